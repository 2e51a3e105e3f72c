#!/usr/bin/env python
"""bench.py — train samples/sec of the MoPoE training step (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU oracle port of the reference step on the host cores

One "step" = forward (3 encoders, fused MoPoE, 3 decoders, likelihoods) + ELBO + backward + gradient exchange and Adam
(N > 1: ONE fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory, or --dp-exchange nccl) on a
synthetic batch: configs[1] of BASELINE.json (PA+Lateral+text, 128 px, 1024x71 char text,
class_dim 128, per-GPU batch 256, bf16 storage / fp32 accumulation).  Weak scaling: per-GPU batch is fixed.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_SAMPLE_TRAIN = 52.09     # 3 x 17.36 GFLOP fwd (2 x 8.681 GMAC) at 128 px tri-modal
WORKLOAD = 'MoPoE PA+Lateral+text 128px char1024x71 class_dim128 joint_elbo'
# BASELINE.json configs.  '2' is the configuration the metric is quoted on (the default, and the only bench line the
# driver reads); the others run the same step at their per-GPU sizes for parity / capacity checks.
CONFIGS = {
    '2': dict(flags={}, batch=256, workload=WORKLOAD, gflop=GFLOP_PER_SAMPLE_TRAIN),
    '4': dict(flags=dict(img_size=256, class_dim=512), batch=64, gflop=167.95,
              workload='MoPoE PA+Lateral+text 256px char1024x71 class_dim512 joint_elbo'),
    '5-joint': dict(flags=dict(mods=('PA', 'text')), batch=128, gflop=29.09,
                    workload='MoPoE PA+text 128px char1024x71 class_dim128 joint_elbo'),
    '5-moe': dict(flags=dict(mods=('PA', 'text'), method='moe'), batch=128, gflop=29.09,
                  workload='MoPoE PA+text 128px char1024x71 class_dim128 moe'),
    '5-poe': dict(flags=dict(mods=('PA', 'text'), method='poe'), batch=128, gflop=58.2,
                  workload='MoPoE PA+text 128px char1024x71 class_dim128 poe (+2 unimodal passes)'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (default: the config\'s)')
    ap.add_argument('--dp-exchange', default='peer', choices=['peer', 'nccl'],
                    help='N>1 gradient exchange: fused reduce-scatter+Adam+all-gather kernel over NVLink peer memory '
                         '(default) or bucketed NCCL all-reduce followed by Adam')
    ap.add_argument('--lr', type=float, default=1e-5,
                    help='Adam learning rate.  The reference default (1e-3) makes the reference model itself diverge to '
                         'inf/NaN on the SECOND step with default init on random inputs (checked with the CPU oracle); '
                         '1e-5 keeps the synthetic run finite.  Throughput does not depend on it')
    ap.add_argument('--text-wire', default='onehot', choices=['onehot', 'uint8'],
                    help='host format of the char text in the e2e leg: the reference\'s fp32 one-hot rows [B,1024,71] '
                         '(default) or one byte per token, expanded on the device (SURVEY N3)')
    ap.add_argument('--image-wire', default='fp32', choices=['fp32', 'uint8'],
                    help='host format of the images in the e2e leg: the reference\'s fp32 in [0,1] (default) or 8-bit pixels, '
                         'ToTensor() evaluated on the device (SURVEY N3)')
    ap.add_argument('--config', default='2', choices=sorted(CONFIGS), help='BASELINE.json configuration (default 2)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--cpu-batch', type=int, default=16)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host instead of replaying the captured CUDA graph')
    ap.add_argument('--profile-kernels', action='store_true', help='print per-op-class device time of one step')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: the per-GPU batch is fixed (default, the config\'s); strong: --global-batch is split over the '
                         'ranks (BASELINE config 3: global 2048 -> 1024 / 512 / 256 per GPU at 2 / 4 / 8 GPUs)')
    ap.add_argument('--global-batch', type=int, default=2048, help='global batch of --scaling strong')
    ap.add_argument('--cpu-kind', default='auto', choices=['auto', 'reference', 'port'],
                    help='CPU baseline: the unmodified reference vendored under baseline/_ref (auto: when present and the '
                         'config is the tri-modal joint_elbo one) or the oracle port')
    ap.add_argument('--no-dp-check', action='store_true', help='N > 1: skip the exchange-vs-all-reduce parity step')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.15)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 7 and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        mx = max((int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit()), default=None)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference(args, steps, warmup, min_seconds=None):
    """The reference's CPU path on the host cores: full step = forward + ELBO + backward + Adam, fp32, all host threads,
    batch `cpu_batch` (a bounded sample of the same workload).  kind "reference": the UNMODIFIED reference modules
    vendored under baseline/_ref (baseline/reference_step.py); kind "port": its algorithm restated in oracle/ (pure torch
    CPU ops, same call sites) — used when the vendored tree is absent or the config needs the oracle's model shim
    (PA+text / poe: the shipped VAEtrimodalMimic cannot run them)."""
    import torch
    cfg_flags = CONFIGS[args.config]['flags']
    native = not any(k in cfg_flags for k in ('mods', 'method'))
    if args.cpu_kind != 'port' and native:
        try:
            from baseline import reference_step as RS
            if RS.available():
                sec, n, cores, loss = RS.timed_run(args.cpu_batch, args.lr, steps, warmup, min_seconds,
                                                   img_size=cfg_flags.get('img_size', 128), class_dim=cfg_flags.get('class_dim', 128))
                return {'value': args.cpu_batch / sec, 'unit': 'samples/s', 'cores': cores, 'kind': 'reference',
                        'sample': '%d timed steps (after %d warm-up) of batch %d, fp32, the unmodified reference '
                                  '(baseline/_ref) on torch CPU, %.2f s/step' % (n, warmup, args.cpu_batch, sec)}, sec
            if args.cpu_kind == 'reference':
                raise RuntimeError('baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists')
        except Exception as e:      # noqa: BLE001
            if args.cpu_kind == 'reference':
                raise
            print('reference arm: vendored reference unavailable (%r); timing the oracle port' % (e,), file=sys.stderr)
    from oracle import mopoe_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fl = O.default_flags(batch_size=args.cpu_batch, **CONFIGS[args.config]['flags'])
    state = O.make_state(fl, 0, torch.float32)
    batch = O.make_batch(fl, 1, torch.float32)
    masks, eps = O.make_noise(fl, 2, torch.float32)
    uni = {m: O.make_noise(fl, 3 + i, torch.float32) for i, m in enumerate(fl.mods)} if fl.method == 'poe' else None
    params = {k: v for k, v in state.items() if v.is_floating_point() and 'running_' not in k}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    times = []
    for it in range(warmup + steps):
        if min_seconds is not None and it >= warmup + 2 and sum(times) >= min_seconds:
            break                       # bounded sample: ~min_seconds of timed CPU work, at least 2 steps
        t0 = time.perf_counter()
        out = O.step_with_grads(state, batch, fl, masks, eps, uni_masks=uni)
        with torch.no_grad():
            O.adam_step(params, out['grads'], m, v, it + 1, lr=args.lr)
            state.update(out['results']['bn_updates'])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {'value': args.cpu_batch / sec, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
            'sample': '%d timed steps (after %d warm-up) of batch %d, fp32, torch CPU ops, %.2f s/step'
                      % (len(times), warmup, args.cpu_batch, sec)}, sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    cb, sec = cpu_reference(args, steps, warmup)
    line = {'metric': 'train samples/sec (3-modality MoPoE, 128px)', 'value': cb['value'], 'unit': 'samples/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': CONFIGS[args.config]['workload'], 'per_gpu_batch': args.cpu_batch,
                       'note': 'the reference step on the host CPU (%s)' % cb['kind']},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_dp_check(exp, px_obj, world, rank, dev):
    """DP parity where the driver can see it (the -m gpu pytest is skipped on a 1-GPU box): ONE more optimizer step on
    the gradients the last timed step left in the flat buffers, done twice — by the fused peer-memory exchange kernel
    on the live buffers, and by the path it replaces (NCCL all-reduce of the gradients, then the flat Adam kernel with
    grad_scale = 1/world) on clones — plus a check that all replicas hold bit-identical parameters afterwards."""
    import torch
    import torch.distributed as dist
    from mopoe_mimic_b200 import _lib as L
    opt = exp.optimizer
    vae = exp.mm_vae
    torch.cuda.synchronize()
    dist.barrier()
    px_obj.gather_moments(opt.m, opt.v)                 # owner-sharded moments -> full tensors on every rank
    eng = vae.rt.engine
    L.call('mopoe_step_advance', L.ptr(eng.rng_step), L.ptr(opt.step_t), L.ptr(opt.coef), float(opt.lr),
           float(opt.betas[0]), float(opt.betas[1]), L.stream_ptr())
    p_ref, m_ref, v_ref = opt.p.clone(), opt.m.clone(), opt.v.clone()
    g_sum = opt.g.clone()
    dist.all_reduce(g_sum)
    L.call('mopoe_adam_flat_dev', L.ptr(p_ref), L.ptr(g_sum), L.ptr(m_ref), L.ptr(v_ref), p_ref.numel(), L.ptr(opt.coef),
           float(opt.betas[0]), float(opt.betas[1]), float(opt.eps), 1.0 / world, L.stream_ptr())
    px_obj.adam_step_all(opt.m, opt.v, opt.coef, opt.betas, opt.eps)
    eng.invalidate_packs()
    torch.cuda.synchronize()
    err = (opt.p - p_ref).abs().max().reshape(1).double()
    scale = p_ref.abs().max().reshape(1).double()
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    hi, lo = opt.p.clone(), opt.p.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    spread = (hi - lo).abs().max().reshape(1).double()
    px_obj.check()
    del p_ref, m_ref, v_ref, g_sum, hi, lo
    return {'what': 'one step: fused peer-memory exchange kernel vs NCCL all-reduce + flat Adam (grad_scale 1/world)',
            'max_abs_err': float(err), 'param_abs_max': float(scale),
            'params_equal_across_ranks': bool(float(spread) == 0.0), 'max_param_spread_across_ranks': float(spread),
            'world': world, 'multicast': bool(getattr(px_obj, 'multicast', False)),
            'buckets': [(b['lo'], b['hi']) for b in px_obj.buckets] if px_obj.buckets else None}


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import mopoe_mimic_b200 as P
    from mopoe_mimic_b200 import _lib as L
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = CONFIGS[args.config]
    B = args.batch or cfg['batch']
    if args.scaling == 'strong':
        if args.global_batch % world:
            raise SystemExit('--global-batch %d does not divide over %d ranks' % (args.global_batch, world))
        B = args.global_batch // world
    fl = P.default_flags(device=dev, batch_size=B, compute_dtype=args.dtype, distributed=world > 1, world_size=world,
                         initial_learning_rate=args.lr, **cfg['flags'])
    torch.manual_seed(0)
    exp = P.Experiment(fl)
    from mopoe_mimic_b200.dp import FlatGradAllReduce, PeerExchange
    peer = world > 1 and args.dp_exchange == 'peer'
    px_obj = None
    if peer:
        # every rank must take the same path: agree on whether the NVLink symmetric-memory plumbing came up
        ok = torch.ones(1, device=dev)
        try:
            px_obj = PeerExchange(dev)
        except Exception as e:       # noqa: BLE001
            print('rank %d: peer-memory exchange unavailable (%r)' % (rank, e), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            peer, px_obj = False, None
            if rank == 0:
                print('falling back to the NCCL all-reduce exchange on all ranks', file=sys.stderr)
    exp.set_optimizer(exchange=px_obj)     # (a PeerExchange broadcasts rank 0's parameters)
    vae = exp.mm_vae
    vae.train()
    # (NCCL path: train_step / GraphedTrainStep attach the all-reduce — rank-0 parameter broadcast + 1/world gradient scale)
    # synthetic inputs of the reference's shapes (dataio/MimicDataset.py:414-428), true one-hot text
    g = torch.Generator(device='cpu').manual_seed(1 + rank)
    px = fl.img_size
    host = {'PA': torch.rand(B, 1, px, px, generator=g).pin_memory(),
            'Lateral': torch.rand(B, 1, px, px, generator=g).pin_memory(),
            'text': torch.nn.functional.one_hot(torch.randint(0, 71, (B, 1024), generator=g), 71).float().pin_memory()}
    host = {k: v for k, v in host.items() if k in fl.mods}
    resident = {k: v.to(dev) for k, v in host.items()}
    if args.text_wire == 'uint8' and 'text' in host:
        if args.no_graph:
            raise SystemExit('--text-wire uint8 needs the graphed step (the expansion kernel fills its static input)')
        host['text'] = host['text'].argmax(-1).to(torch.uint8).pin_memory()
    if args.image_wire == 'uint8':
        if args.no_graph:
            raise SystemExit('--image-wire uint8 needs the graphed step (the expansion kernel fills its static input)')
        for k in ('PA', 'Lateral'):
            if k in host:
                host[k] = (host[k] * 255.0).round().to(torch.uint8).pin_memory()
                resident[k] = (host[k].float() / 255.0).to(dev)         # the same pixel values the device expansion yields
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    ar = FlatGradAllReduce() if (world > 1 and not peer) else None      # bucketed NCCL all-reduce of the flat gradients
    stats_host = torch.empty(16, dtype=torch.float32).pin_memory()
    launches_per_step = None
    if args.no_graph:
        def step_resident():
            return P.train_step(exp, (dict(resident), None), ar)

        def step_e2e():
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            out = P.train_step(exp, (b, None), ar)
            st = P.packed_stats(out)
            stats_host[:st.numel()].copy_(st, non_blocking=True)
            return st.numel() * 4
    else:
        c0 = L.LAUNCHES
        gstep = P.GraphedTrainStep(exp, resident, ar, token_indices=(args.text_wire == 'uint8'))   # 2 eager warm-up steps + 1 captured step
        launches_per_step = (L.LAUNCHES - c0) // 3

        def step_resident():
            return gstep(resident)

        def step_e2e():
            st = gstep(host)                                   # pinned host -> static device inputs, then replay
            stats_host[:st.numel()].copy_(st, non_blocking=True)
            return st.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    if args.profile_kernels and rank == 0:
        # developer aid: device time per kernel name for ONE step (CUPTI via torch.profiler), then exit
        from torch.profiler import ProfilerActivity, profile
        t_cpu0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step_resident()
            torch.cuda.synchronize()
        print('one step wall %.1f ms' % ((time.perf_counter() - t_cpu0) * 1e3))
        rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
        tot = sum(r[2] for r in rows)
        for k, n, t in sorted(rows, key=lambda r: -r[2])[:45]:
            print('%-70s %5d %9.1f us %5.1f%%' % (k[:70], n, t, 100.0 * t / tot))
        print('total device time %.1f ms' % (tot / 1e3))
        flt = os.environ.get('KFILTER')
        if flt:      # per-launch durations of one kernel family, in launch order
            ds = [(e.name[:40], e.device_time_total) for e in prof.events()
                  if flt in e.name and e.device_time_total > 0 and str(e.device_type).endswith('CUDA')]
            print('%d launches of %s:' % (len(ds), flt), ' '.join('%.0f' % d for _, d in ds))
        return
    sampler = ClockSampler(local)
    sampler.start()
    L.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.LAUNCHES if launches_per_step is None else launches_per_step * args.steps
    # end-to-end: pinned host inputs -> device each step, packed stats back
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    d2h = 0
    for _ in range(args.steps):
        d2h = step_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # Rooflines: one instrumented step with a CUDA event pair around EVERY C-ABI call (mopoe_mimic_b200/_lib.call), each
    # annotated by its caller with the algorithmic work of the launch (GEMM flops; bytes of the HBM-bound passes).  The
    # events are captured INTO a copy of the step graph (event-record nodes), so the durations are the kernels' own
    # in-step durations — bracketing eager launches instead also counts the host's enqueue latency between the event and
    # the kernel, which is comparable to these 5-100 us kernels.
    prof, timing = None, None
    dp_check = None
    if world > 1 and peer and not args.no_dp_check:
        dp_check = run_dp_check(exp, px_obj, world, rank, dev)
    # (the instrumented copy stops after backward at N > 1: no bucket of the gradient exchange may be forked from it)
    saved_hook, vae.rt.on_decoders_done = vae.rt.on_decoders_done, None
    if not args.no_graph:
        try:
            L.PROFILE, L.PROFILE_EXTERNAL = [], True
            saved_dataset, fl.dataset = fl.dataset, 'testing'
            g2 = torch.cuda.CUDAGraph()
            # the instrumented copy runs the modality branches on ONE stream: an event pair around a kernel that shares
            # the GPU with another branch's kernels would time the sharing, not the kernel
            saved_env = os.environ.get('MOPOE_BRANCH_STREAMS')
            os.environ['MOPOE_BRANCH_STREAMS'] = '0'
            try:
                with torch.cuda.graph(g2, pool=gstep.graph.pool()):
                    P.forward_backward(exp, (dict(gstep.static), None))
                    if world == 1:
                        exp.optimizer.step()
            finally:
                if saved_env is None:
                    os.environ.pop('MOPOE_BRANCH_STREAMS', None)
                else:
                    os.environ['MOPOE_BRANCH_STREAMS'] = saved_env
            fl.dataset = saved_dataset
            prof = L.PROFILE
            for _ in range(3):
                g2.replay()
            torch.cuda.synchronize()
            _ = [p_['a'].elapsed_time(p_['b']) for p_ in prof[:2]]
            timing = ('CUDA event pairs captured as nodes of a single-stream copy of the step graph (in-step kernel '
                      'durations; the timed step itself overlaps the modality branches on 3 streams)')
        except Exception as e:       # noqa: BLE001
            print('graph-captured event timing unavailable (%r); bracketing eager launches' % (e,), file=sys.stderr)
            prof = None
            torch.cuda.synchronize()
        finally:
            L.PROFILE, L.PROFILE_EXTERNAL = None, False
    if prof is None:
        L.PROFILE = []
        try:
            P.forward_backward(exp, (dict(resident), None))          # eager: includes host enqueue latency per launch
            torch.cuda.synchronize()
        finally:
            prof, L.PROFILE = L.PROFILE, None
        timing = 'CUDA event pairs around eager launches (includes host enqueue latency)'
    vae.rt.on_decoders_done = saved_hook
    for p_ in prof:
        p_['ms'] = p_['a'].elapsed_time(p_['b'])
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        traffic, traffic_note = None, None
        try:       # dram bytes of the step's largest GEMM launch, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_roofline_traffic.json')))
            traffic = tj['dram_bytes_read'] + tj['dram_bytes_written']
            traffic_note = ('NOT measured in this run: ncu --set full dram read+write of ONE launch (%s) from the committed '
                            'capture profiles/r2_ncu_gemm.txt; algorithmic bytes of that launch %.1f MB'
                            % (tj['kernel'], tj['algorithmic_bytes'] / 1e6))
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback'
        gprof = [p_ for p_ in prof if p_.get('flops')]
        gemm_ms = sum(p_['ms'] for p_ in gprof)
        gemm_flop = sum(p_['flops'] for p_ in gprof)
        by_kind = {}
        for p_ in gprof:
            d = by_kind.setdefault(p_['kind'], [0.0, 0.0, 0])
            d[0] += p_['ms']
            d[1] += p_['flops']
            d[2] += 1
        achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # HBM side (north_star items 2-3): algorithmic bytes of every fusion / likelihood / BatchNorm / Adam launch of the
        # step divided by its in-step duration, per kernel class, against the measured copy bandwidth
        peak_hbm = peaks.get('hbm_gbs', 6550.0)
        hbm = {}
        for p_ in prof:
            if p_.get('bytes'):
                d = hbm.setdefault(p_['kind'], [0.0, 0.0, 0])
                d[0] += p_['ms']
                d[1] += p_['bytes']
                d[2] += 1
        roofline_hbm = {'peak': peak_hbm, 'unit': 'GB/s', 'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback',
                        'note': 'achieved = algorithmic bytes (SURVEY.md 8d) / in-step launch time; the fusion kernels move '
                                '<5 MB per launch and are launch-latency-bound by construction',
                        'classes': {k: {'ms': v[0], 'MB': v[1] / 1e6, 'launches': v[2],
                                        'achieved': (v[1] / 1e9) / (v[0] * 1e-3) if v[0] > 0 else 0.0,
                                        'frac': ((v[1] / 1e9) / (v[0] * 1e-3) / peak_hbm) if v[0] > 0 else 0.0}
                                    for k, v in sorted(hbm.items())}}
        tot_ms = sum(v[0] for v in hbm.values())
        tot_b = sum(v[1] for v in hbm.values())
        roofline_hbm['all'] = {'ms': tot_ms, 'MB': tot_b / 1e6, 'achieved': (tot_b / 1e9) / (tot_ms * 1e-3) if tot_ms else 0.0,
                               'frac': (tot_b / 1e9) / (tot_ms * 1e-3) / peak_hbm if tot_ms else 0.0}
        if os.environ.get('MOPOE_BENCH_SHAPES'):       # developer aid: GEMM time per problem shape
            agg = {}
            for p_ in gprof:
                d = agg.setdefault(p_.get('tag'), [0.0, 0.0, 0])
                d[0] += p_['ms']
                d[1] += p_['flops']
                d[2] += 1
            for tag, d in sorted(agg.items(), key=lambda kv: -kv[1][0]):
                print('%-46s n=%3d %8.3f ms %7.1f TF/s' % (tag, d[2], d[0], d[1] / d[0] / 1e9), file=sys.stderr)
            other = {}
            for p_ in prof:
                if not p_.get('flops'):
                    d = other.setdefault(p_['name'], [0.0, 0])
                    d[0] += p_['ms']
                    d[1] += 1
            for name, d in sorted(other.items(), key=lambda kv: -kv[1][0]):
                print('%-46s n=%3d %8.3f ms' % (name, d[1], d[0]), file=sys.stderr)
        # GEMM launches that carry another pass in their epilogue (their time counts as GEMM time, only the GEMM's flops count)
        fused_epi = {}
        for p_ in gprof:
            tag = p_.get('tag') or ''
            label = ('residual_combine' if '+res' in tag else 'bn_backward_sums' if '+bnb' in tag
                     else 'bn_statistics' if '+bn ' in tag else None)
            if label:
                d = fused_epi.setdefault(label, {'launches': 0, 'ms': 0.0})
                d['launches'] += 1
                d['ms'] += p_['ms']
        value = world * B * args.steps / (ms * 1e-3)
        metric = 'train samples/sec (3-modality MoPoE, 128px)'
        if args.config != '2':
            metric = 'train samples/sec (BASELINE config %s)' % args.config
        line = {'metric': metric, 'value': value, 'unit': 'samples/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
                'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': args.dtype,
                'data': 'synthetic',
                'config': {'workload': cfg['workload'], 'per_gpu_batch': B, 'global_batch': B * world,
                           'parallelism': 'dp%d' % world + ('' if world == 1 else (' peer-memory fused exchange' if peer else ' nccl all-reduce')), 'cuda_graph': not args.no_graph, 'lr': args.lr,
                           'branch_streams': os.environ.get('MOPOE_BRANCH_STREAMS', '1') != '0', 'text_wire': args.text_wire, 'image_wire': args.image_wire, 'l2': 'inputs+activations per step >> 126 MB L2 (no flush needed)'},
                'clocks': sampler.summary(),
                'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s',
                        'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h},
                'gpu_launches': launches,
                'last_step': {'total_loss': float(stats_host[0]), 'joint_divergence': float(stats_host[1]),
                              'finite': bool(torch.isfinite(stats_host[:d2h // 4]).all())},
                'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                             'frac': achieved / peak_tf, 'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_src,
                             'timing': timing,
                             'kernel': 'implicit-GEMM conv family (fprop+dgrad+wgrad), %d launches/step' % len(gprof),
                             'gemm_ms_per_step': gemm_ms, 'step_tensor_frac': value / world * cfg['gflop'] / 1e3 / peak_tf,
                             'fused_epilogues': fused_epi,
                             'fused_epilogues_note': 'GEMM launches whose epilogue also does the block\'s residual combine, the '
                                                     'BatchNorm-backward reduction or the BatchNorm statistics (passes that were '
                                                     'separate HBM-bound launches): their whole time is in gemm_ms_per_step, only '
                                                     'the GEMM flops are counted in achieved / frac',

                             'by_kind': {k: {'ms': v[0], 'tflops': (v[1] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else 0.0, 'launches': v[2]}
                                         for k, v in by_kind.items()}},
                'roofline_hbm': roofline_hbm}
        if dp_check is not None:
            line['dp_check'] = dp_check
        if not args.no_cpu_baseline:
            cb, _ = cpu_reference(args, 12, 1, min_seconds=10.0)
            line['cpu_baseline'] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
