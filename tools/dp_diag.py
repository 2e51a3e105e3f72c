"""Developer probe (not a pytest file): where does the data-parallel step spend its time?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dp_diag.py

Times graph A (forward+backward), the gradient exchange and graph B (Adam) separately with CUDA events, the NCCL
all-reduce of the flat gradient buffer on its own, and probes torch's symmetric-memory plumbing (peer pointers,
signal pads, multicast) that the peer-memory exchange kernel is built on.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    import mopoe_mimic_b200 as P
    from mopoe_mimic_b200.dp import FlatGradAllReduce
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B = int(os.environ.get('DIAG_BATCH', '256'))

    def say(*a):
        if rank == 0:
            print(*a, flush=True)

    # ---- symmetric memory probe -------------------------------------------------------------------------------
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(64 << 20, dtype=torch.float32, device=dev)       # 256 MB
        h = symm.rendezvous(t, group=dist.group.WORLD)
        say('symm: world', h.world_size, 'buffer_ptrs', [hex(p) for p in h.buffer_ptrs], 'signal_pad_size',
            h.signal_pad_size,
            'multicast_ptr', hex(h.multicast_ptr) if h.multicast_ptr else None)
        peer = (rank + 1) % world
        rt = h.get_buffer(peer, (64 << 20,), torch.float32)
        loc = torch.empty_like(t)
        for name, fn in (('p2p read ', lambda: loc.copy_(rt)), ('p2p write', lambda: rt.copy_(loc))):
            fn()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            say('symm %s 256 MB: %.3f ms -> %.0f GB/s' % (name, a.elapsed_time(b) / 5, 0.268435456 / (a.elapsed_time(b) / 5e3)))
        dist.barrier()
    except Exception as e:       # noqa: BLE001
        say('symm probe failed:', repr(e))

    # ---- NCCL all-reduce alone ---------------------------------------------------------------------------------
    fl = P.default_flags(device=dev, batch_size=B, compute_dtype='bf16', distributed=True, world_size=world)
    torch.manual_seed(0)
    exp = P.Experiment(fl)
    exp.set_optimizer()
    vae = exp.mm_vae
    vae.train()
    dist.broadcast(vae.flat_params, 0)
    exp.optimizer.grad_scale = 1.0 / world
    flat = vae.flat_grads
    say('flat grads: %d elements, %.1f MB' % (flat.numel(), flat.numel() * 4 / 1e6))
    for mb in (64, 256, 1024):
        ar = FlatGradAllReduce(bucket_mb=mb)
        ar(flat)
        dist.barrier()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(5):
            ar(flat)
        b.record()
        torch.cuda.synchronize()
        say('nccl all-reduce bucket %4d MB: %.3f ms' % (mb, a.elapsed_time(b) / 5))
    flat.zero_()

    # ---- step segments ------------------------------------------------------------------------------------------
    g = torch.Generator(device='cpu').manual_seed(1 + rank)
    res = {'PA': torch.rand(B, 1, 128, 128, generator=g).to(dev), 'Lateral': torch.rand(B, 1, 128, 128, generator=g).to(dev),
           'text': torch.nn.functional.one_hot(torch.randint(0, 71, (B, 1024), generator=g), 71).float().to(dev)}
    ar = FlatGradAllReduce()
    gs = P.GraphedTrainStep(exp, res, ar)
    for _ in range(3):
        gs(res)
    dist.barrier()
    torch.cuda.synchronize()
    if os.environ.get('DIAG_SEGMENTS', '1') == '0':
        return
    n = 10
    es = [[ev() for _ in range(4)] for _ in range(n)]
    for i in range(n):
        es[i][0].record()
        gs.graph.replay()
        es[i][1].record()
        ar(flat)
        es[i][2].record()
        gs.graph_b.replay()
        es[i][3].record()
    torch.cuda.synchronize()
    for j, name in enumerate(('graph A (fwd+bwd)', 'all-reduce', 'graph B (adam)')):
        say('%-20s %.3f ms' % (name, sum(e[j].elapsed_time(e[j + 1]) for e in es) / n))
    say('whole step           %.3f ms' % (es[0][0].elapsed_time(es[-1][3]) / n))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
