"""Host-side engine: channels-last activation descriptors, implicit-GEMM problem builders, weight
packing, and thin wrappers over the C ABI.  No math happens here — every op is a call into
libmopoe_b200.so; torch is used for device memory, streams and (tiny) weight re-layouts.

Layouts
  activation  [B, H+2ph, W+2pw, C]  (1-D text: H = 1, ph = 0), zero border written by the producer
  conv-form weights   Wc [a, (ky, kx, b)]       from a weight tensor Wg[a, b, ky, kx]
  phase-form weights  Wp[py,px] [b, (r, kxi, a)] = Wg[a, b, KY[py][r], KX[px][kxi]],  KY = KX = ((3,1),(2,0))
  full-form weights   Wf [(ky, kx, b), a]
so that  conv k4/s2/p1  = one GEMM over 4-tap row windows of the padded input (conv-form),
         deconv k4/s2/p1 = 2^nd sub-pixel phase GEMMs over 2-tap windows (phase-form),
and each layer's dgrad is the other kind; wgrad always produces conv-form.
"""
import ctypes as C
import os

import torch

from . import _lib as L

KTAPS = ((3, 1), (2, 0))     # phase -> kernel taps in memory order of the 2-wide window


class Act:
    """Channels-last activation: storage tensor `t` [B, H+2ph, W+2pw, C]."""
    __slots__ = ('t', 'B', 'H', 'W', 'C', 'ph', 'pw')

    def __init__(self, t, B, H, W, C_, ph, pw):
        self.t, self.B, self.H, self.W, self.C, self.ph, self.pw = t, B, H, W, C_, ph, pw

    @staticmethod
    def empty(B, H, W, C_, ph, pw, dtype, device):
        return Act(torch.empty((B, H + 2 * ph, W + 2 * pw, C_), dtype=dtype, device=device), B, H, W, C_, ph, pw)

    @staticmethod
    def like(t, B, H, W, C_, ph=0, pw=0):
        assert t.numel() == B * (H + 2 * ph) * (W + 2 * pw) * C_, (t.shape, B, H, W, C_, ph, pw)
        return Act(t, B, H, W, C_, ph, pw)

    @property
    def Ws(self):
        return self.W + 2 * self.pw

    @property
    def Hs(self):
        return self.H + 2 * self.ph

    @property
    def dtype(self):
        return self.t.dtype

    def origin(self):
        """element offset of interior (0,0,0,0) in storage"""
        return (self.ph * self.Ws + self.pw) * self.C

    def view(self):
        isz = self.t.element_size()
        return L.View(self.t.data_ptr() + self.origin() * isz, L.dtype_code(self.t.dtype), self.B, self.H, self.W,
                      self.C, self.ph, self.pw, 0, self.Hs * self.Ws * self.C, self.Ws * self.C, self.C)

    def interior(self):
        return self.t[:, self.ph:self.ph + self.H, self.pw:self.pw + self.W, :]


class Engine:
    """Per-device scratch + launch helpers.  One per model; all calls go to the current CUDA stream."""

    def __init__(self, device, dtype=torch.float32, impl=L.IMPL_AUTO):
        L.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('mopoe_mimic_b200 needs a CUDA device (sm_100a); no CPU fallback exists')
        self.dtype = dtype
        self.impl = impl
        self._ws64 = {}      # scratch per CUDA stream: the modality branches run concurrently on their own streams
        self._wsf = {}
        self.rng_offset = 0
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        # arrival counters for the library's fused-finalize path (last block finalises).  Measured SLOWER on B200 than
        # a separate one-warp-per-channel finalize launch (the lone last block serialises 16 channels), so it stays off.
        self.counters = None
        self._packs = {}
        self._pack_gen, self._packs_stale, self._pack_table = 0, False, None
        import os
        self.batched = os.environ.get('MOPOE_GEMM_BATCHED', '1') != '0'       # phases of a deconv in one launch
        self.persistent = os.environ.get('MOPOE_GEMM_PERSISTENT', '1') != '0'   # persistent kernel for single GEMMs too
        self.fuse_stats = os.environ.get('MOPOE_FUSE_BN_STATS', '1') != '0'     # BatchNorm statistics in the GEMM epilogue
        self.fuse_next_stats = os.environ.get('MOPOE_FUSE_NEXT_BN_STATS', '1') != '0'   # next block's bn1 statistics in combine
        self.fuse_res = os.environ.get('MOPOE_FUSE_RES', '1') != '0'            # residual combine in conv2's GEMM epilogue
        self.fuse_bnb = os.environ.get('MOPOE_FUSE_BNB', '1') != '0'            # bn2-backward sums in conv2's dgrad epilogue
        self.wgrad_streams = os.environ.get('MOPOE_WGRAD_STREAMS', '0') != '0'          # weight gradients on side streams (opt-in: measured +-0.1 ms)
        self._wg_streams, self._wg_used, self._wg_keep = {}, set(), []

    # ---- packed-weight cache ---------------------------------------------------------------------------------
    # Every (weight, form) has a persistent SLOT (destination buffers with fixed addresses).  A slot is valid while its
    # generation matches the engine's and the tensor's torch version is unchanged.  The first step packs slot by slot
    # as the layers run; from then on begin_step() re-packs ALL known slots in one batched launch (prepack), so the
    # forward/backward only ever hit the cache.
    FORMS = {'conv': 0, 'phase': 1, 'full': 2, 'mat': 3, 'matT': 4}

    def _slot_shapes(self, fcode, A, B, KH, KW, bpad):
        if fcode == 0:
            return [(A, KH * KW * bpad)]
        if fcode == 1:
            return [(B, 4 * A)] * 4 if KH > 1 else [(B, 2 * A)] * 2
        if fcode == 2:
            return [(KH * KW * B, A)]
        assert KH == 1 and KW == 1
        # (bpad: zero-padded k-columns of a 'mat' operand / zero-padded rows of a 'matT' operand)
        return [(A, bpad)] if fcode == 3 else [(max(B, bpad), A)]

    def packed(self, Wg, form, bpad=None):
        """GEMM-operand re-layout of an fp32 master weight, cached until the weights change"""
        skey = (Wg.data_ptr(), form, bpad, self.dtype)
        slot = self._packs.get(skey)
        if slot is not None and slot['gen'] == self._pack_gen and slot['version'] == Wg._version:
            return slot['out']
        W = Wg.detach()
        assert W.dtype == torch.float32 and W.is_contiguous()
        if slot is None:
            A, B = W.shape[0], W.shape[1]
            KH, KW = (1, 1) if W.dim() == 2 else ((1, W.shape[2]) if W.dim() == 3 else (W.shape[2], W.shape[3]))
            fcode = self.FORMS[form]
            bp = bpad or B
            alloc = torch.zeros if (fcode == 4 and bp > B) else torch.empty      # (padded matT rows are never written: stay 0)
            dsts = [alloc(sh, dtype=self.dtype, device=self.device) for sh in self._slot_shapes(fcode, A, B, KH, KW, bp)]
            slot = {'W': Wg, 'dsts': dsts, 'out': dsts if fcode == 1 else dsts[0], 'geo': (A, B, KH, KW, fcode, bp),
                    'arr': (C.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts]), 'gen': -1, 'version': -1}
            self._packs[skey] = slot
            self._pack_table = None                      # the batched job table must be rebuilt
        A, B, KH, KW, fcode, bp = slot['geo']
        L.call('mopoe_pack_weight_tiled', L.ptr(W), A, B, KH, KW, fcode, bp, slot['arr'], L.dtype_code(self.dtype),
               L.stream_ptr())
        slot['gen'], slot['version'] = self._pack_gen, Wg._version
        return slot['out']

    def prepack(self):
        """re-pack every known slot in ONE launch (mopoe_pack_weights_batched)"""
        if not self._packs:
            return
        if self._pack_table is None and torch.cuda.is_current_stream_capturing():
            return                       # (the table upload is not capturable: this step packs slot by slot)
        if self._pack_table is None:
            self._build_pack_table()
        tab, njobs, tiles = self._pack_table
        L.call('mopoe_pack_weights_batched', L.ptr(tab), njobs, tiles, L.dtype_code(self.dtype), L.stream_ptr())
        for slot in self._packs.values():
            slot['gen'], slot['version'] = self._pack_gen, slot['W']._version

    def _build_pack_table(self):
        """device-resident job table of mopoe_pack_weights_batched (one descriptor per slot, sorted by first tile)"""
        lib = L.load()
        jobs = (L.PackJob * len(self._packs))()
        tile = 0
        for j, slot in zip(jobs, self._packs.values()):
            A, B, KH, KW, fcode, bp = slot['geo']
            nx = C.c_int(0)
            n = lib.mopoe_pack_job_tiles(A, B, KH, KW, fcode, bp, C.byref(nx))
            j.W = slot['W'].data_ptr()
            for i, d in enumerate(slot['dsts']):
                j.dst[i] = d.data_ptr()
            j.A, j.B, j.KH, j.KW, j.form, j.bpad, j.tile0, j.nx = A, B, KH, KW, fcode, bp, tile, nx.value
            tile += n
        raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
        self._pack_table = (raw.to(self.device), len(jobs), tile)

    def begin_step(self):
        """same within-step Philox offsets every step; the device-side step counter makes the draws differ.  Weights that
        changed behind torch's back (flat Adam) are re-packed here in one launch, and all dropout keep-masks of the step
        (72 at the bench configuration) are drawn by ONE launch once the previous step has shown which are needed."""
        self.rng_offset = 0
        self._masks_begin()
        if self._packs and self._packs_stale:
            self.prepack()
            self._packs_stale = False

    def invalidate_packs(self):
        """call after the parameters changed in place behind torch's back (the flat Adam kernel)"""
        self._pack_gen += 1
        self._packs_stale = True
        if self._pack_table is None and self._packs and not torch.cuda.is_current_stream_capturing():
            self._build_pack_table()

    # ---- scratch ------------------------------------------------------------------------------------
    def ws64(self, n):
        key = torch.cuda.current_stream().cuda_stream
        buf = self._ws64.get(key)
        if buf is None or buf.numel() < n:
            buf = self._ws64[key] = torch.empty(max(n, 1 << 16), dtype=torch.float64, device=self.device)
        return buf

    def wsf(self, nbytes):
        n = (nbytes + 3) // 4
        key = torch.cuda.current_stream().cuda_stream
        buf = self._wsf.get(key)
        if buf is None or buf.numel() < n:
            buf = self._wsf[key] = torch.empty(max(n, 1 << 20), dtype=torch.float32, device=self.device)
        return buf

    @staticmethod
    def nchunk(rows, C_, per=1):
        cg = (C_ + 127) // 128
        want = max(1, (148 * 8 + cg - 1) // cg)
        return int(max(1, min(want, (rows + 31) // 32)))

    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    # ---- elementwise wrappers ---------------------------------------------------------------------------
    def bn_stats(self, x, mask, mode, rmean=None, rvar=None, eps=1e-5, momentum=0.1):
        rows = x.B * x.H * x.W
        nc = self.nchunk(rows, x.C)
        ws = self.ws64(2 * nc * x.C)
        stats = self.f32(2, x.C)
        self._bytes('bn_stats', x, 1)
        L.call('mopoe_bn_stats', C.byref(x.view()), L.ptr(mask), mode, L.ptr(ws), nc, eps, momentum,
               L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(rmean), L.ptr(rvar), L.ptr(self.counters), L.stream_ptr())
        return stats

    def bn_apply(self, x, mask, mode, stats, gamma, beta, relu, out):
        self._bytes('bn_apply', x, 2)
        L.call('mopoe_bn_apply', C.byref(x.view()), L.ptr(mask), mode, L.ptr(stats[0]), L.ptr(stats[1]),
               L.ptr(gamma), L.ptr(beta), int(relu), C.byref(out.view()), L.stream_ptr())
        return out

    def combine(self, r, stats, gamma, beta, c, mask, mode, a, b, out, bn=None, eps=1e-5, momentum=0.1):
        """out = a * BN(r) + b * (c * 2mask).  bn = (running_mean, running_var) of the BatchNorm that reads `out` next (the
        following block's bn1): also return its training-mode statistics [2, C], produced in the same pass"""
        self._bytes('combine', r, 3)
        if bn is None:
            L.call('mopoe_combine', C.byref(r.view()), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(beta),
                   C.byref(c.view()), L.ptr(mask), mode, float(a), float(b), C.byref(out.view()), L.stream_ptr())
            return out
        nc = self.nchunk(r.B * r.H * r.W, r.C)
        ws = self.ws64(2 * nc * r.C)
        st = self.f32(2, r.C)
        L.call('mopoe_combine_bn', C.byref(r.view()), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(beta),
               C.byref(c.view()), L.ptr(mask), mode, float(a), float(b), C.byref(out.view()), L.ptr(ws), nc, eps, momentum,
               L.ptr(st[0]), L.ptr(st[1]), L.ptr(bn[0]), L.ptr(bn[1]), L.stream_ptr())
        return out, st

    def bn_bwd(self, dy, gate, gscale, x, mask, mode, stats, gamma, dgamma, dbeta, addend, out, accumulate=False, beta=None,
               sums=None):
        """Both halves of the BN(+ReLU, +dropout) backward; returns `out` = d/d(x).  gate = the saved post-ReLU
        activation (None: no ReLU).  beta: the BatchNorm's bias when `gate` is this BatchNorm's own relu output — lets both
        passes recompute the gate from x (bit-identical) instead of reading the activation.  sums: the reduction's result
        when the GEMM that produced dy already delivered it (_gemm_bnbwd; dgamma / dbeta are then written too)."""
        gv = C.byref(gate.view()) if gate is not None else None
        recomp = gate is not None and beta is not None
        gg, gb = (L.ptr(gamma), L.ptr(beta)) if recomp else (None, None)
        if sums is None:
            rows = x.B * x.H * x.W
            nc = self.nchunk(rows, x.C)
            ws = self.ws64(2 * nc * x.C)
            sums = self.f32(2, x.C)
            self._bytes('bn_bwd_reduce', x, (2 if recomp else 3) if gate is not None else 2)
            L.call('mopoe_bn_bwd_reduce', C.byref(dy.view()), gv, float(gscale), C.byref(x.view()), L.ptr(mask), mode,
                   L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(ws), nc, L.ptr(dgamma), L.ptr(dbeta), int(accumulate), L.ptr(sums),
                   gg, gb, L.ptr(self.counters), L.stream_ptr())
        else:
            assert recomp and gscale == 1.0
        av = C.byref(addend.view()) if addend is not None else None
        self._bytes('bn_bwd_apply', x, 3 + (gate is not None and not recomp) + (addend is not None))
        L.call('mopoe_bn_bwd_apply', C.byref(dy.view()), gv, float(gscale), C.byref(x.view()), L.ptr(mask), mode,
               L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(sums), av, C.byref(out.view()), gb, L.stream_ptr())
        return out

    def combine_bwd(self, dy, a, r, stats, gamma, dgamma, dbeta, mask2, mode2, b, dr, dc, accumulate=False):
        """backward of y = a*BN(r) + b*(c*2mask2): BN reduction over (dy, r), then ONE pass writing dr and dc"""
        rows = r.B * r.H * r.W
        nc = self.nchunk(rows, r.C)
        ws = self.ws64(2 * nc * r.C)
        sums = self.f32(2, r.C)
        self._bytes('bn_bwd_reduce', r, 2)
        L.call('mopoe_bn_bwd_reduce', C.byref(dy.view()), None, float(a), C.byref(r.view()), None, L.MASK_NONE,
               L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(ws), nc, L.ptr(dgamma), L.ptr(dbeta), int(accumulate), L.ptr(sums),
               None, None, L.ptr(self.counters), L.stream_ptr())
        self._bytes('combine_bwd_apply', r, 4)
        L.call('mopoe_combine_bwd_apply', C.byref(dy.view()), float(a), C.byref(r.view()), L.ptr(stats[0]),
               L.ptr(stats[1]), L.ptr(gamma), L.ptr(sums), L.ptr(mask2), mode2, float(b), C.byref(dr.view()),
               C.byref(dc.view()), L.stream_ptr())
        return dr, dc

    def scale_mask(self, dy, mask, mode, scale, out):
        L.call('mopoe_scale_mask', C.byref(dy.view()), L.ptr(mask), mode, float(scale), C.byref(out.view()),
               L.stream_ptr())
        return out

    def colsum(self, v, out=None, accumulate=False):
        rows = v.B * v.H * v.W
        nc = self.nchunk(rows, v.C)
        ws = self.ws64(2 * nc * v.C)
        if out is None:
            out = self.f32(v.C)
        L.call('mopoe_colsum', C.byref(v.view()), L.ptr(out), int(accumulate), L.ptr(ws), nc, L.ptr(self.counters),
               L.stream_ptr())
        return out

    def convert(self, src_view, nchw, dst):
        L.call('mopoe_convert', C.byref(src_view), int(nchw), C.byref(dst.view()), L.stream_ptr())
        return dst

    # ---- weight-gradient side streams ------------------------------------------------------------------------
    # A block's weight gradients feed nothing but the optimizer, while its input gradients sit on the critical chain
    # dgrad -> BN sums -> finalize -> BN apply -> dgrad ...  In the deep stages that chain is a string of 5-us launches on a
    # handful of SMs and the weight-gradient GEMMs (large weights there) are the bulk of the work: they go to a side stream
    # of the branch stream they were issued from and overlap the chain.  The side streams are joined before anything
    # reads the flat gradient (FlatAdam.step / the decoders' bucket), and the operands of the deferred kernels are kept
    # alive until then (the caching allocator would otherwise hand their memory to the branch stream's next tensors).
    def wgrad_side(self):
        """the side stream paired with the current stream, forked at this point (None: disabled / gradients not flat)"""
        if not self.wgrad_streams or os.environ.get('MOPOE_BRANCH_STREAMS', '1') == '0':      # (bench.py flips it at run time)
            return None
        cur = torch.cuda.current_stream()
        key = cur.cuda_stream
        st = self._wg_streams.get(key)
        if st is None:
            st = self._wg_streams[key] = torch.cuda.Stream()
        st.wait_stream(cur)
        self._wg_used.add(key)
        return st

    def wgrad_keep(self, *tensors):
        self._wg_keep.extend(tensors)

    def join_wgrad_sides(self, final=True):
        """make the current stream wait for every weight-gradient side stream used since the last final join.  final:
        the caller is the step's last join (everything later is ordered behind the current stream): the operands kept
        alive for the deferred kernels may go.  A non-final join (the decoders' optimizer bucket, itself on a side stream)
        must keep them: the branch streams do not wait here and would reuse the memory."""
        cur = torch.cuda.current_stream()
        for key in sorted(self._wg_used):
            cur.wait_stream(self._wg_streams[key])
        if final:
            self._wg_used = set()
            self._wg_keep = []

    # ---- dropout keep-masks --------------------------------------------------------------------------------
    # Mask k of a step covers the Philox counter blocks [off_k, off_k + ceil(n_k / 128)) (128 mask bytes per counter), so
    # the masks of a whole step are ONE contiguous stream: a single launch over a buffer that lays them out at 128-byte
    # boundaries draws exactly the bytes the per-mask launches would.  The sizes are learnt from the previous step; a
    # step that asks for anything else (eval pass, another fusion method) falls back to per-mask launches and re-learns.
    def _masks_begin(self):
        rec = getattr(self, '_mask_rec', None)
        if rec:                                   # what the last step drew becomes the plan
            self._mask_plan = (self._mask_seed, tuple(rec))
        self._mask_rec, self._mask_seed = [], None
        self._mask_buf, self._mask_pos = None, 0
        plan = getattr(self, '_mask_plan', None)
        if plan is None or os.environ.get('MOPOE_BATCHED_MASKS', '1') == '0':
            return
        seed, sizes = plan
        total = sum((n + 127) // 128 for n in sizes) * 128
        self._mask_buf = torch.empty(total, dtype=torch.uint8, device=self.device)
        L.call('mopoe_dropout_mask', L.ptr(self._mask_buf), total, int(seed) & (2 ** 64 - 1), 0, L.ptr(self.rng_step),
               L.stream_ptr())

    def dropout_mask(self, n, seed):
        if getattr(self, '_mask_rec', None) is None:
            self._mask_rec, self._mask_seed, self._mask_buf, self._mask_pos = [], None, None, 0
        if self._mask_seed is None:
            self._mask_seed = seed
        self._mask_rec.append(n)
        blocks = (n + 127) // 128
        buf = self._mask_buf
        if buf is not None:
            pseed, sizes = self._mask_plan
            k = self._mask_pos
            if k < len(sizes) and sizes[k] == n and pseed == seed and self.rng_offset * 128 + n <= buf.numel():
                m = buf[self.rng_offset * 128:self.rng_offset * 128 + n]
                self._mask_pos += 1
                self.rng_offset += blocks
                return m
            self._mask_buf = None                 # the step departs from the plan: per-mask launches from here on
        m = torch.empty(n, dtype=torch.uint8, device=self.device)
        L.call('mopoe_dropout_mask', L.ptr(m), n, int(seed) & (2 ** 64 - 1), self.rng_offset, L.ptr(self.rng_step),
               L.stream_ptr())
        self.rng_offset += blocks
        return m

    # ---- implicit-GEMM problem builders -------------------------------------------------------------------
    @staticmethod
    def _timed(kind, flops, fn, tag=None):
        """annotate the C call fn() makes with its algorithmic work (bench.py's per-call timing lives in _lib.call)"""
        L.annotate(kind=kind, flops=flops, tag=tag)
        return fn()

    @staticmethod
    def _bytes(kind, v, passes, extra=0):
        """annotate the next call as an HBM pass: `passes` reads/writes of the interior of activation `v`"""
        if L.PROFILE is not None:
            L.annotate(kind=kind, bytes=passes * v.B * v.H * v.W * v.C * v.t.element_size() + extra)

    def _gemm(self, win, wp, bias, rows, bn=None, out=None, res=None, bnb=None):
        if res is not None:
            return self._gemm_res([win], [wp], bias, [rows], out, res, [self.rows_of(res['r'])])
        if bnb is not None:
            return self._gemm_bnbwd([win], [wp], bias, [rows], out, bnb, [self.rows_of(bnb['x'])])
        return self._gemm_batched([win], [wp], bias, [rows], bn, out)

    def _gemm_bnbwd(self, wins, wps, bias, rows_list, out, bnb, x_rows):
        """an input-gradient GEMM whose result `out` feeds a BatchNorm(+ReLU, +dropout) backward: the BatchNorm-backward
        sums come out of the GEMM epilogue (mopoe_conv_gemm_bnbwd) instead of a reduction pass over (out, x).  bnb:
        dict(x, mask, mode, stats, gamma, beta, dgamma, dbeta, accumulate).  Always runs the GEMM; returns the sums [2, C]
        (for bn_bwd(..., sums=)) or None when the fused epilogue does not apply."""
        n = len(wins)
        x = bnb['x']
        ok = (bias is None and wins[0].a_dtype == L.BF16 and self.impl != L.IMPL_SIMT and self.persistent and self.fuse_bnb
              and x.dtype == torch.bfloat16 and (x.B, x.H, x.W, x.C) == (out.B, out.H, out.W, out.C))
        if ok:
            req = L.BnBwdReq()
            XR = (L.Rows * n)(*x_rows)
            req.x = XR
            mask = bnb['mask']
            req.mask, req.mask_mode = (mask.data_ptr() if mask is not None else None), bnb['mode']
            req.accumulate = int(bool(bnb['accumulate']))
            st = bnb['stats']
            req.mean, req.invstd = st[0].data_ptr(), st[1].data_ptr()
            req.gamma, req.beta = bnb['gamma'].data_ptr(), bnb['beta'].data_ptr()
            ws = self.ws64(8 * 160 * out.C)
            req.ws, req.ws_doubles = ws.data_ptr(), ws.numel()
            sums = self.f32(2, out.C)
            req.dgamma, req.dbeta, req.sums = bnb['dgamma'].data_ptr(), bnb['dbeta'].data_ptr(), sums.data_ptr()
            WA = (L.Window * n)(*wins)
            RA = (L.Rows * n)(*rows_list)
            ok = bool(L.load().mopoe_conv_gemm_bnbwd_eligible(n, WA, RA, self.impl, C.byref(req)))
        if not ok:
            self._gemm_batched(wins, wps, bias, rows_list, None, out)
            return None
        PA = (C.c_void_p * n)(*[wp.data_ptr() for wp in wps])
        w0 = wins[0]
        flops = sum(2.0 * w.E0 * w.E1 * w.E2 * rr.N * w.R * w.KW for w, rr in zip(wins, rows_list))
        self._timed('fprop/dgrad', flops, lambda: L.call('mopoe_conv_gemm_bnbwd', n, WA, PA, RA, self.impl, C.byref(req),
                                                         L.stream_ptr()),
                    'x%d+bnb M=%dx%dx%d N=%d K=%dx%d' % (n, w0.E2, w0.E1, w0.E0, rows_list[0].N, w0.R, w0.KW))
        return sums

    def _gemm_res(self, wins, wps, bias, rows_list, out, res, res_rows):
        """conv2 of a residual block with the block's combine in the epilogue (mopoe_conv_gemm_res): `out` receives
        a * BN(r) + b * dropout(conv2) directly, zero border included.  res: dict(r, stats, gamma, beta, a, b, mask, mode, next_bn).  Returns None
        when the fused epilogue does not apply (nothing launched; the caller runs GEMM + combine), else the statistics of
        the next block's bn1 (res['next_bn'] = its running buffers) or True."""
        n = len(wins)
        if wins[0].a_dtype != L.BF16 or self.impl == L.IMPL_SIMT or not self.persistent:
            return None
        for wp in wps:
            assert wp.is_contiguous() and wp.dtype == torch.bfloat16
        r = res['r']
        assert r.ph == 0 and r.pw == 0 and r.t.is_contiguous() and (r.B, r.H, r.W, r.C) == (out.B, out.H, out.W, out.C)
        req = L.ResReq()
        RR = (L.Rows * n)(*res_rows)
        req.r = RR
        st = res['stats']
        req.mean, req.invstd = st[0].data_ptr(), st[1].data_ptr()
        req.gamma, req.beta = res['gamma'].data_ptr(), res['beta'].data_ptr()
        req.a, req.b = float(res['a']), float(res['b'])
        mask = res['mask']
        req.mask, req.mask_mode = (mask.data_ptr() if mask is not None else None), res['mode']
        oview = out.view()
        req.out = C.pointer(oview)               # the launch also writes the zero border of `out`
        bnreq = stats = keep = None
        if res.get('next_bn') is not None:
            bnreq, stats, keep = self._bn_request(out, (None, L.MASK_NONE) + tuple(res['next_bn']))
        bnp = C.byref(bnreq) if bnreq is not None else None
        WA = (L.Window * n)(*wins)
        RA = (L.Rows * n)(*rows_list)
        if not L.load().mopoe_conv_gemm_res_eligible(n, WA, L.ptr(bias), RA, self.impl, C.byref(req), bnp):
            return None
        PA = (C.c_void_p * n)(*[wp.data_ptr() for wp in wps])
        w0 = wins[0]
        flops = sum(2.0 * w.E0 * w.E1 * w.E2 * rr.N * w.R * w.KW for w, rr in zip(wins, rows_list))
        self._timed('fprop/dgrad', flops, lambda: L.call('mopoe_conv_gemm_res', n, WA, PA, L.ptr(bias), RA, self.impl,
                                                         C.byref(req), bnp, L.stream_ptr()),
                    'x%d+res M=%dx%dx%d N=%d K=%dx%d' % (n, w0.E2, w0.E1, w0.E0, rows_list[0].N, w0.R, w0.KW))
        return stats if stats is not None else True

    def _bn_request(self, out, bn):
        """mopoe_bn_req_t for a GEMM whose output `out` feeds a training-mode BatchNorm; bn = (mask, mode, rmean, rvar)"""
        mask, mode, rmean, rvar = bn
        rows = out.B * out.H * out.W
        nc = self.nchunk(rows, out.C)
        nd = max(2 * nc, 8 * 160) * out.C
        ws = self.ws64(nd)
        stats = self.f32(2, out.C)
        req = L.BnReq()
        req.out = out.view()
        req.mask, req.mask_mode, req.nchunk = (mask.data_ptr() if mask is not None else None), mode, nc
        req.ws, req.ws_doubles = ws.data_ptr(), ws.numel()
        req.eps, req.momentum = 1e-5, 0.1
        req.mean, req.invstd = stats[0].data_ptr(), stats[1].data_ptr()
        req.running_mean = rmean.data_ptr() if rmean is not None else None
        req.running_var = rvar.data_ptr() if rvar is not None else None
        return req, stats, (ws, mask, rmean, rvar)

    def _gemm_batched(self, wins, wps, bias, rows_list, bn=None, out=None):
        """up to 4 same-shape problems in ONE launch (persistent tcgen05 kernel when eligible).  bn = (mask, mode,
        running_mean, running_var): also produce the training-mode BatchNorm statistics of the output `out` (fused into
        the GEMM epilogue where the library can); returns them as a [2, C] tensor (mean, 1/sqrt(var + eps))."""
        n = len(wins)
        for wp in wps:
            assert wp.is_contiguous() and wp.dtype == self._win_dtype(wins[0])
        flops = sum(2.0 * w.E0 * w.E1 * w.E2 * r.N * w.R * w.KW for w, r in zip(wins, rows_list))
        if n == 1 and self.impl != L.IMPL_SIMT and wins[0].a_dtype == L.BF16:
            # weight-bound problem (few tiles, long reduction): split-K over the idle SMs
            nbytes = L.load().mopoe_conv_gemm_splitk_ws(C.byref(wins[0]), C.byref(rows_list[0]), self.impl)
            if nbytes:
                ws = self.wsf(nbytes)
                w0 = wins[0]
                self._timed('fprop/dgrad', flops, lambda: L.call('mopoe_conv_gemm_splitk', C.byref(w0), L.ptr(wps[0]), L.ptr(bias),
                                                                 C.byref(rows_list[0]), L.ptr(ws), nbytes, self.impl,
                                                                 L.stream_ptr()),
                            'sk M=%dx%dx%d N=%d K=%dx%d' % (w0.E2, w0.E1, w0.E0, rows_list[0].N, w0.R, w0.KW))
                return self.bn_stats(out, bn[0], bn[1], bn[2], bn[3]) if bn is not None else None
        if bn is not None and not self.fuse_stats:
            self._gemm_batched(wins, wps, bias, rows_list)
            return self.bn_stats(out, bn[0], bn[1], bn[2], bn[3])
        if bn is not None:
            req, stats, keep = self._bn_request(out, bn)
            WA = (L.Window * n)(*wins)
            RA = (L.Rows * n)(*rows_list)
            PA = (C.c_void_p * n)(*[wp.data_ptr() for wp in wps])
            w0 = wins[0]
            self._timed('fprop/dgrad', flops, lambda: L.call('mopoe_conv_gemm_bn', n, WA, PA, L.ptr(bias), RA, self.impl,
                                                             C.byref(req), L.stream_ptr()),
                        'x%d+bn M=%dx%dx%d N=%d K=%dx%d' % (n, w0.E2, w0.E1, w0.E0, rows_list[0].N, w0.R, w0.KW))
            return stats
        if self.batched and (n > 1 or self.persistent):
            WA = (L.Window * n)(*wins)
            RA = (L.Rows * n)(*rows_list)
            PA = (C.c_void_p * n)(*[wp.data_ptr() for wp in wps])
            w0 = wins[0]
            self._timed('fprop/dgrad', flops, lambda: L.call('mopoe_conv_gemm_batched', n, WA, PA, L.ptr(bias), RA,
                                                             self.impl, L.stream_ptr()),
                        'x%d M=%dx%dx%d N=%d K=%dx%d' % (n, w0.E2, w0.E1, w0.E0, rows_list[0].N, w0.R, w0.KW))
        else:
            for w, wp, r in zip(wins, wps, rows_list):
                f1 = 2.0 * w.E0 * w.E1 * w.E2 * r.N * w.R * w.KW
                self._timed('fprop/dgrad', f1, lambda w=w, wp=wp, r=r: L.call(
                    'mopoe_conv_gemm', C.byref(w), L.ptr(wp), L.ptr(bias), C.byref(r), self.impl, L.stream_ptr()))

    @staticmethod
    def _win_dtype(win):
        return torch.float32 if win.a_dtype == L.F32 else torch.bfloat16

    def _wgrad(self, win, rows, N, K):
        out = self.f32(N, K)
        nbytes = L.load().mopoe_conv_wgrad_ws(C.byref(win), C.byref(rows), self.impl)
        ws = self.wsf(nbytes) if nbytes else None
        flops = 2.0 * win.E0 * win.E1 * win.E2 * N * K
        self._timed('wgrad', flops, lambda: L.call('mopoe_conv_wgrad', C.byref(win), C.byref(rows), L.ptr(out), 0,
                                                   L.ptr(ws), nbytes, self.impl, L.stream_ptr()),
                    'wg0 M=%dx%dx%d N=%d K=%dx%d' % (win.E2, win.E1, win.E0, N, win.R, win.KW))
        return out

    def _wgrad_param(self, win, rows, param, taps, bpad):
        """weight gradient accumulated straight into param.grad (parameter layout [a, b, *taps]); False if the
        parameter has no contiguous .grad to accumulate into (caller then uses the packed-gradient path)"""
        g = param.grad
        if g is None or not g.is_contiguous() or g.dtype != torch.float32:
            return False
        A, B = param.shape[0], param.shape[1]
        lib = L.load()
        nbytes = lib.mopoe_conv_wgrad_param_ws(C.byref(win), C.byref(rows), self.impl)
        ws = self.wsf(nbytes)
        flops = 2.0 * win.E0 * win.E1 * win.E2 * rows.N * win.R * win.KW
        self._timed('wgrad', flops, lambda: L.call('mopoe_conv_wgrad_param', C.byref(win), C.byref(rows), L.ptr(g), A, B,
                                                   taps, bpad, 1, L.ptr(ws), nbytes, self.impl, L.stream_ptr()),
                    'wg M=%dx%dx%d N=%d K=%dx%d' % (win.E2, win.E1, win.E0, rows.N, win.R, win.KW))
        return True

    def wgrad_down_param(self, xwin, k, s, p, yrows, param, bpad=None):
        win, OH, OW = self.win_down(xwin, k, s, p)
        assert (yrows.H, yrows.W, yrows.B) == (OH, OW, xwin.B), ((yrows.H, yrows.W), (OH, OW))
        taps = param[0, 0].numel()
        return self._wgrad_param(win, self.rows_of(yrows), param, taps, bpad or param.shape[1])

    @staticmethod
    def win_down(x, k, s, p):
        """windows of a stride-s, kernel-k (k taps contiguous along W) conv over padded `x`"""
        assert x.pw >= p and (x.H == 1 or x.ph >= p)
        Cc = x.C
        if x.H == 1:
            OW = (x.W + 2 * p - k) // s + 1
            return L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), OW, 1, x.B, 1, k * Cc, 0,
                            (x.pw - p) * Cc, s * Cc, 0, x.Ws * Cc, 0), 1, OW
        OH = (x.H + 2 * p - k) // s + 1
        OW = (x.W + 2 * p - k) // s + 1
        a_off = ((x.ph - p) * x.Ws + (x.pw - p)) * Cc
        return L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), OW, OH, x.B, k, k * Cc, 0,
                        a_off, s * Cc, s * x.Ws * Cc, x.Hs * x.Ws * Cc, x.Ws * Cc), OH, OW

    @staticmethod
    def rows_of(act, N=None):
        """row addressing over the interior pixels of `act` in (w, h, b) order"""
        return L.Rows(act.t.data_ptr(), L.dtype_code(act.dtype), N or act.C, act.origin(), act.C,
                      act.Ws * act.C, act.Hs * act.Ws * act.C)

    def gemm_down(self, x, wc, bias, k, s, p, n, out_dtype=None, out=None, bn=None, res=None, bnb=None):
        """bnb: see _gemm_bnbwd -> returns (out, sums or None)"""
        win, OH, OW = self.win_down(x, k, s, p)
        if out is None:
            out = Act.empty(x.B, OH, OW, n, 0, 0, out_dtype or x.dtype, self.device)
        assert (out.B, out.H, out.W, out.C) == (x.B, OH, OW, n)
        if res is not None:
            return self._gemm(win, wc, bias, self.rows_of(out), None, out, res)
        if bnb is not None:
            return out, self._gemm(win, wc, bias, self.rows_of(out), None, out, None, bnb)
        st = self._gemm(win, wc, bias, self.rows_of(out), bn, out)
        return out if bn is None else (out, st)

    @staticmethod
    def phase_rows(act, n, py, px):
        """rows of sub-pixel phase (py, px) of a stride-2 deconv's output `act` (any border): pixel (2y+py, 2x+px)"""
        if act.H == 1:
            return L.Rows(act.t.data_ptr(), L.dtype_code(act.dtype), n, act.origin() + px * n, 2 * n, 0, act.Ws * n)
        return L.Rows(act.t.data_ptr(), L.dtype_code(act.dtype), n, act.origin() + (py * act.Ws + px) * n, 2 * n,
                      2 * act.Ws * n, act.Hs * act.Ws * n)

    def wgrad_down(self, xwin, k, s, p, yrows):
        win, OH, OW = self.win_down(xwin, k, s, p)
        assert (yrows.H, yrows.W, yrows.B) == (OH, OW, xwin.B), ((yrows.H, yrows.W), (OH, OW))
        return self._wgrad(win, self.rows_of(yrows), yrows.C, win.R * win.KW)

    def gemm_up(self, x, wph, bias, n, out_dtype=None, bn=None, out=None, res=None, bnb=None):
        """stride-2 k4 p1 transposed conv as 2^nd sub-pixel phase GEMMs (x must carry a border >= 1).  out: an existing
        (possibly bordered) activation to write the interior of.  res: see _gemm_res (returns its result)."""
        Cc = x.C
        assert x.pw >= 1 and (x.H == 1 or x.ph >= 1)
        if out is None:
            out = Act.empty(x.B, 1 if x.H == 1 else 2 * x.H, 2 * x.W, n, 0, 0, out_dtype or x.dtype, self.device)
        assert (out.B, out.H, out.W, out.C) == (x.B, 1 if x.H == 1 else 2 * x.H, 2 * x.W, n)
        wins, rows_l, res_rows = [], [], []
        if x.H == 1:
            for px in range(2):
                wins.append(L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.W, 1, x.B, 1, 2 * Cc, 0,
                                     (x.pw - 1 + px) * Cc, Cc, 0, x.Ws * Cc, 0))
                rows_l.append(self.phase_rows(out, n, 0, px))
                if res is not None or bnb is not None:
                    res_rows.append(self.phase_rows(res['r'] if res is not None else bnb['x'], n, 0, px))
        else:
            for py in range(2):
                for px in range(2):
                    a_off = ((x.ph - 1 + py) * x.Ws + (x.pw - 1 + px)) * Cc
                    wins.append(L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.W, x.H, x.B, 2, 2 * Cc, 0,
                                         a_off, Cc, x.Ws * Cc, x.Hs * x.Ws * Cc, x.Ws * Cc))
                    rows_l.append(self.phase_rows(out, n, py, px))
                    if res is not None or bnb is not None:
                        res_rows.append(self.phase_rows(res['r'] if res is not None else bnb['x'], n, py, px))
        if res is not None:
            return self._gemm_res(wins, list(wph), bias, rows_l, out, res, res_rows)
        if bnb is not None:
            return out, self._gemm_bnbwd(wins, list(wph), bias, rows_l, out, bnb, res_rows)
        st = self._gemm_batched(wins, list(wph), bias, rows_l, bn, out)
        return out if bn is None else (out, st)

    def gemm_unfold(self, dy, wfull, n, H, W, k, p):
        """input gradient of a conv whose stride equals its kernel (the 256-px stage: k4 s4 p1, FeatureExtractorImg.py
        :52-59): windows tile the input without overlap, so every dy pixel scatters ONE k x k x n patch and dgrad is a
        plain GEMM.  One problem per patch row ky (same A, weight slice and output origin differ): row (b,oy,ox) of
        problem ky is the k*n contiguous elements dX[b, k*oy-p+ky, k*ox-p .. +k, :].  The result carries a border of p
        that receives the (unused) gradients of the padding; input pixels no window covers stay zero."""
        Cc = dy.C
        assert wfull.shape == (k * k * n, Cc) and p <= 1
        out = Act(torch.zeros(dy.B, H + 2 * p, W + 2 * p, n, dtype=dy.dtype, device=self.device), dy.B, H, W, n, p, p)
        assert k * dy.H - 1 < out.Hs and k * dy.W <= out.Ws
        win = L.Window(dy.t.data_ptr(), L.dtype_code(dy.dtype), dy.W, dy.H, dy.B, 1, Cc, 0, dy.origin(), Cc,
                       dy.Ws * Cc, dy.Hs * dy.Ws * Cc, 0)
        wins, wps, rows_l = [], [], []
        for ky in range(k):
            wins.append(win)
            wps.append(wfull[ky * k * n:(ky + 1) * k * n])
            rows_l.append(L.Rows(out.t.data_ptr(), L.dtype_code(out.dtype), k * n, ky * out.Ws * n, k * n,
                                 k * out.Ws * n, out.Hs * out.Ws * n))
        self._gemm_batched(wins, wps, None, rows_l)
        return out

    def gemm_rows(self, x, w, bias, n, out_shape=None, out_dtype=None, bn=None, out=None):
        """pointwise GEMM over the interior pixels of x: out[m, n] = sum_c x[m, c] w[n, c] (+bias).  out: an existing
        (possibly bordered) activation to write the interior of."""
        Cc = x.C
        flat = x.ph == 0 and x.pw == 0 and (out is None or (out.ph == 0 and out.pw == 0))
        if flat:
            win = L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.B * x.H * x.W, 1, 1, 1, Cc, 0, 0, Cc, 0, 0, 0)
        else:
            win = L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.W, x.H, x.B, 1, Cc, 0, x.origin(), Cc,
                           x.Ws * Cc, x.Hs * x.Ws * Cc, 0)
        B, H, W = out_shape[:3] if out_shape else (x.B, x.H, x.W)
        nn_ = out_shape[3] if out_shape else n
        if out is None:
            out = Act.empty(B, H, W, nn_, 0, 0, out_dtype or x.dtype, self.device)
        else:
            assert out_shape is None and (out.B, out.H, out.W, out.C) == (x.B, x.H, x.W, n)
        if flat:
            rows = L.Rows(out.t.data_ptr(), L.dtype_code(out.dtype), n, 0, n, 0, 0)
        elif out.ph or out.pw:
            rows = self.rows_of(out)
        else:
            rows = L.Rows(out.t.data_ptr(), L.dtype_code(out.dtype), n, 0, n, x.W * n, x.H * x.W * n)
        assert bn is None or out_shape is None, 'fused statistics need GEMM columns == output channels'
        st = self._gemm(win, w, bias, rows, bn, out)
        return out if bn is None else (out, st)

    def wgrad_rows_param(self, x, dy, param):
        """param.grad[n, c] += sum_m dy[m, n] x[m, c] (a 1x1 conv / linear weight [n, c, 1..]); False if the parameter
        has no flat fp32 .grad to accumulate into"""
        Cc = x.C
        win = L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.W, x.H, x.B, 1, Cc, 0, x.origin(), Cc,
                       x.Ws * Cc, x.Hs * x.Ws * Cc, 0)
        assert (x.B, x.H, x.W) == (dy.B, dy.H, dy.W) and tuple(param.shape[:2]) == (dy.C, Cc)
        return self._wgrad_param(win, self.rows_of(dy), param, 1, Cc)

    def wgrad_rows(self, x, dy, n=None):
        """dW[n, c] = sum_m dy[m, n] x[m, c] over interior pixels (both any padding); n: only the first n columns of dy"""
        Cc = x.C
        win = L.Window(x.t.data_ptr(), L.dtype_code(x.dtype), x.W, x.H, x.B, 1, Cc, 0, x.origin(), Cc,
                       x.Ws * Cc, x.Hs * x.Ws * Cc, 0)
        assert (x.B, x.H, x.W) == (dy.B, dy.H, dy.W)
        return self._wgrad(win, self.rows_of(dy, N=n), n or dy.C, Cc)

    def zero_border(self, act):
        L.call('mopoe_zero_border', C.byref(act.view()), L.stream_ptr())
        return act


# ---- weight packing (tiny torch re-layouts of the fp32 master weights) ------------------------------------
def conv_form(Wg, dtype):
    a, b = Wg.shape[:2]
    if Wg.dim() == 3:
        return Wg.permute(0, 2, 1).reshape(a, -1).to(dtype).contiguous()
    return Wg.permute(0, 2, 3, 1).reshape(a, -1).to(dtype).contiguous()


def conv_form_grad(g, shape):
    """inverse of conv_form for a gradient [a, taps*b] -> Wg layout"""
    a, b = shape[:2]
    if len(shape) == 3:
        return g.view(a, shape[2], b).permute(0, 2, 1)
    return g.view(a, shape[2], shape[3], b).permute(0, 3, 1, 2)


def phase_form(Wg, dtype):
    a, b = Wg.shape[:2]
    out = []
    if Wg.dim() == 3:
        for px in range(2):
            w = Wg[:, :, list(KTAPS[px])]                      # [a, b, 2]
            out.append(w.permute(1, 2, 0).reshape(b, -1).to(dtype).contiguous())
        return out
    for py in range(2):
        for px in range(2):
            w = Wg[:, :, list(KTAPS[py]), :][:, :, :, list(KTAPS[px])]   # [a, b, 2, 2]
            out.append(w.permute(1, 2, 3, 0).reshape(b, -1).to(dtype).contiguous())
    return out


def full_form(Wg, dtype):
    a, b = Wg.shape[:2]
    if Wg.dim() == 3:
        return Wg.permute(2, 1, 0).reshape(-1, a).to(dtype).contiguous()
    return Wg.permute(2, 3, 1, 0).reshape(-1, a).to(dtype).contiguous()
