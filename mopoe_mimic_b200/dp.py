"""Batch-sharded data parallelism (reference: mimic/main_mimic.py:44-48,65-67, utils/utils.py:179-185,
run_epochs.py:245-247, dataio/utils.py:120-123): one process per GPU, per-rank batch = global // world,
BatchNorm statistics and mixture selection per rank, gradients mean-reduced.

The path shards by samples and has exactly one exchange step — the gradient all-reduce — so that is the only
collective.  All 420 gradients live in ONE flat fp32 buffer: the reduction is a handful of large bucketed
`all_reduce` calls (NCCL over NVLink/NVSwitch on GPUs; the same host logic runs on gloo for the CPU tests), and the
1/world scale is folded into the Adam kernel instead of a separate pass.
"""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """contiguous shard of every tensor in a {name: tensor} batch — what DistributedSampler + per-rank
    batch_size //= world amounts to for one step (dataio/utils.py:120-123, main_mimic.py:48)"""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def bucket_bounds(numel, bucket_elems):
    """[start, end) element ranges covering `numel` in buckets of at most `bucket_elems` (256-B aligned)"""
    bucket_elems = max(64, bucket_elems // 64 * 64)
    return [(s, min(numel, s + bucket_elems)) for s in range(0, numel, bucket_elems)]


class FlatGradAllReduce:
    """Sum-all-reduce of a flat gradient buffer in large buckets.  Callable: `allreduce(flat_grads)`.

    `grad_scale` (= 1 / world) is what the optimizer must multiply gradients with to obtain the DDP mean."""

    def __init__(self, group=None, bucket_mb=256):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.grad_scale = 1.0 / self.world

    def __call__(self, flat):
        if self.world == 1:
            return flat
        for s, e in bucket_bounds(flat.numel(), self.bucket_elems):
            dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=self.group)
        return flat


class PeerExchange:
    """NVLink peer-memory plumbing of the fused exchange kernel (mopoe_dp_adam_exchange, csrc/dp_exchange.cu).

    The flat parameter and gradient buffers are allocated from torch's symmetric-memory pool, so after `connect()`
    every rank holds the address of every peer's buffers; one kernel per step then does reduce-scatter (P2P loads),
    Adam on the own 1/world slice and the parameter all-gather (P2P stores).  No NCCL call is on the step path and
    the whole DP step stays ONE CUDA graph.  Adam moments are owner-sharded (ZeRO-1 style): `gather_moments()`
    assembles the full tensors for a checkpoint."""

    MAX_BUCKETS = 8

    def __init__(self, device, group=None, multicast=None):
        """multicast: None = use NVSwitch multicast (NVLS) when the symmetric allocation has it and world > 2
        (MOPOE_DP_MULTICAST=0/1 overrides); True / False force it."""
        import os
        import torch.distributed._symmetric_memory as symm
        env = os.environ.get('MOPOE_DP_MULTICAST')
        self.want_multicast = multicast if multicast is not None else (None if env is None else env == '1')
        self.symm = symm
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 16:
            raise RuntimeError('PeerExchange supports up to 16 ranks on one NVLink domain')
        self.params = self.grads = self.flags = self.state = self.peers = None
        self.grad_scale = 1.0 / self.world
        # fail HERE (before the model's buffers are moved) if the platform has no symmetric-memory / P2P support
        probe = symm.empty(64, dtype=torch.float32, device=device)
        h = symm.rendezvous(probe, group=self.group)
        if len(h.buffer_ptrs) != self.world or not all(h.buffer_ptrs):
            raise RuntimeError('symmetric-memory rendezvous did not yield a peer address for every rank')

    def alloc(self, total):
        self.params = self.symm.empty(total, dtype=torch.float32, device=self.device)
        self.grads = self.symm.empty(total, dtype=torch.float32, device=self.device)
        self.flags = self.symm.empty(self.MAX_BUCKETS * 32, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        return self.params, self.grads

    def connect(self):
        """exchange buffer addresses with the peers (collective) and broadcast rank 0's parameters"""
        from . import _lib as L
        hp = self.symm.rendezvous(self.params, group=self.group)
        hg = self.symm.rendezvous(self.grads, group=self.group)
        hf = self.symm.rendezvous(self.flags, group=self.group)
        pk = L.DpPeers()
        # the handles describe the symmetric ALLOCATION; our tensors may sit at an offset inside it (same on all ranks)
        og = self.grads.data_ptr() - hg.buffer_ptrs[self.rank]
        op = self.params.data_ptr() - hp.buffer_ptrs[self.rank]
        of = self.flags.data_ptr() - hf.buffer_ptrs[self.rank]
        for r in range(self.world):
            pk.grad[r], pk.param[r], pk.flags[r] = hg.buffer_ptrs[r] + og, hp.buffer_ptrs[r] + op, hf.buffer_ptrs[r] + of
        self.peers, self._handles = pk, (hp, hg, hf)
        have_mc = bool(hp.multicast_ptr) and bool(hg.multicast_ptr)
        use = self.want_multicast if self.want_multicast is not None else self.world > 2
        if use and not have_mc:
            if self.want_multicast:
                raise RuntimeError('NVSwitch multicast was requested but the symmetric allocation has no multicast address')
            use = False
        self.multicast = use
        self.mc_grad, self.mc_param = (hg.multicast_ptr + og, hp.multicast_ptr + op) if use else (0, 0)
        self.state = torch.tensor([1, 0, 0, 0], dtype=torch.int32, device=self.device)   # epoch, arrivals, error, -
        self._bases = ([hg.buffer_ptrs[r] + og for r in range(self.world)], [hp.buffer_ptrs[r] + op for r in range(self.world)],
                       [hf.buffer_ptrs[r] + of for r in range(self.world)])
        self.buckets = None
        dist.broadcast(self.params, 0, group=self.group)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)          # every rank's flags are zeroed and mapped before the first kernel

    def set_buckets(self, bounds):
        """Split the exchange into buckets: [(lo, hi), ...] element ranges (multiples of 4) covering the flat buffers, in the
        order their gradients become final during backward.  Every bucket is exchanged by its own launch on its own
        sub-range (reduce-scatter + Adam + all-gather of THAT range over all ranks), with its own flag slots / epoch state,
        so a bucket can run under the rest of the backward pass (the reference's DDP overlaps its buckets the same way)."""
        from . import _lib as L
        assert self.peers is not None, 'connect() first'
        assert 1 <= len(bounds) <= self.MAX_BUCKETS
        covered = sorted(bounds)
        assert covered[0][0] == 0 and covered[-1][1] == self.params.numel() and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        gb, pb, fb = self._bases
        self.buckets = []
        for k, (lo, hi) in enumerate(bounds):
            assert lo % 4 == 0 and hi % 4 == 0 and hi > lo
            pk = L.DpPeers()
            for r in range(self.world):
                pk.grad[r], pk.param[r], pk.flags[r] = gb[r] + 4 * lo, pb[r] + 4 * lo, fb[r] + 4 * 32 * k
            self.buckets.append(dict(lo=lo, hi=hi, peers=pk,
                                     mc_grad=self.mc_grad + 4 * lo if self.multicast else 0,
                                     mc_param=self.mc_param + 4 * lo if self.multicast else 0,
                                     state=torch.tensor([1, 0, 0, 0], dtype=torch.int32, device=self.device)))

    def adam_step(self, m, v, coef, betas, eps, bucket=None, max_blocks=0):
        """the whole buffer (bucket None; only valid while no buckets are set) or one bucket"""
        import ctypes as C
        from . import _lib as L
        if bucket is None:
            assert self.buckets is None, 'the exchange is bucketed: use adam_step_all() or pass a bucket index'
            L.call('mopoe_dp_adam_exchange', C.byref(self.peers), C.c_void_p(self.mc_grad), C.c_void_p(self.mc_param),
                   L.ptr(m), L.ptr(v), self.params.numel(), self.rank,
                   self.world, L.ptr(self.state), L.ptr(coef), float(betas[0]), float(betas[1]), float(eps),
                   float(self.grad_scale), L.stream_ptr())
            return
        b = self.buckets[bucket]
        lo, hi = b['lo'], b['hi']
        L.call('mopoe_dp_adam_exchange_ex', C.byref(b['peers']), C.c_void_p(b['mc_grad']), C.c_void_p(b['mc_param']),
               L.ptr(m[lo:hi]), L.ptr(v[lo:hi]), hi - lo, self.rank, self.world, L.ptr(b['state']), L.ptr(coef),
               float(betas[0]), float(betas[1]), float(eps), float(self.grad_scale), int(max_blocks), L.stream_ptr())

    def adam_step_all(self, m, v, coef, betas, eps):
        """every bucket (or the whole buffer) back to back on the current stream"""
        if self.buckets is None:
            return self.adam_step(m, v, coef, betas, eps)
        for k in range(len(self.buckets)):
            self.adam_step(m, v, coef, betas, eps, bucket=k)

    def check(self):
        """Raise if a flag barrier of the exchange kernel gave up waiting (one device->host read; call it where the step's
        statistics are read anyway).  All ranks must reach the exchange within MOPOE_DP_TIMEOUT_S (default 600 s, 0 = no
        limit) of each other; put a dist.barrier() after rank-asymmetric work (rank-0 evaluation, checkpointing) so the
        skew is absorbed on the host rather than inside the kernel."""
        errs = [int(self.state[2].item())] + [int(b['state'][2].item()) for b in (self.buckets or [])]
        err = max(errs)
        if err:
            err -= 1
            raise RuntimeError('peer exchange: rank %d never reached barrier %d within MOPOE_DP_TIMEOUT_S on rank %d'
                               % (err % 16, err // 16, self.rank))

    def slice_bounds(self, lo=0, hi=None):
        """owner slices of the range [lo, hi) (default: the whole buffer): rank r owns the r-th run of ceil(n/4/world) float4s"""
        hi = self.params.numel() if hi is None else hi
        n4 = (hi - lo) // 4
        per = (n4 + self.world - 1) // self.world
        return [(lo + min(n4, r * per) * 4, lo + min(n4, (r + 1) * per) * 4) for r in range(self.world)]

    def gather_moments(self, m, v):
        """full Adam moments on every rank (for optimizer checkpoints): each slice comes from its owner"""
        ranges = [(b['lo'], b['hi']) for b in self.buckets] if getattr(self, 'buckets', None) else [(0, self.params.numel())]
        for lo, hi in ranges:
            for r, (s, e) in enumerate(self.slice_bounds(lo, hi)):
                if e > s:
                    dist.broadcast(m[s:e], r, group=self.group)
                    dist.broadcast(v[s:e], r, group=self.group)
        return m, v


def broadcast_flat(flat, src=0, group=None):
    """rank-0 parameters to every rank (what DDP's constructor does)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src, group=group)
    return flat
