// conv_direct.cu — the two single-channel image layers.  K = 9 (first conv) and N = 1 (last deconv)
// are not tensor-core shapes, so they are CUDA-core kernels bound by the 128-channel
// side of their traffic.
//   first:  nn.Conv2d(1, C, 3, stride=2, padding=1, bias=False)          FeatureExtractorImg.py:29-34
//   last:   nn.ConvTranspose2d(C, 1, 3, stride=2, padding=1, output_padding=1)   DataGeneratorImg.py:84-90
// The 3x3xC filter is staged once per CTA in shared memory as [tap][C] (or kept in registers), each thread
// owns 8 consecutive channels of one pixel, so the per-pixel overhead (index math, 9 scalar taps) is
// amortised over a 16-byte (bf16) / 32-byte (fp32) channel vector.
#include "common.cuh"

constexpr int CV8 = 8;

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&o)[8]) {
    float a[4], b[4];
    ldv<4>(p, a);
    ldv<4>(p + 4, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i] = a[i]; o[4 + i] = b[i]; }
}
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float (&o)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&o)[8]) {
    float a[4] = {o[0], o[1], o[2], o[3]}, b[4] = {o[4], o[5], o[6], o[7]};
    stv<4>(p, a);
    stv<4>(p + 4, b);
}
template <>
__device__ __forceinline__ void st8<bf16>(bf16* p, const float (&o)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
}

// The 9 taps of a stride-2, pad-1, 3x3 window around source pixel (2*oy, 2*ox): rows 2oy-1 .. 2oy+1, cols 2ox-1 .. 2ox+1.
// With an even source extent only the FIRST row / column can fall outside (oy == 0 / ox == 0), so the whole bounds
// logic is two predicates and the addressing is three row pointers with immediate offsets — the per-tap 64-bit index
// math of the naive form was 2/3 of this kernel family's instruction stream (ncu: 250 integer ops per 72 FMAs).
__device__ __forceinline__ void load_taps(const float* __restrict__ img, int SW, int oy, int ox, float (&t)[9]) {
    const float* pc = img + (2 * oy) * SW + 2 * ox;
    const float* pu = pc - SW;
    const float* pd = pc + SW;
    const bool top = oy > 0, left = ox > 0;
    t[0] = (top && left) ? __ldg(pu - 1) : 0.f;
    t[1] = top ? __ldg(pu) : 0.f;
    t[2] = top ? __ldg(pu + 1) : 0.f;
    t[3] = left ? __ldg(pc - 1) : 0.f;
    t[4] = __ldg(pc);
    t[5] = __ldg(pc + 1);
    t[6] = left ? __ldg(pd - 1) : 0.f;
    t[7] = __ldg(pd);
    t[8] = __ldg(pd + 1);
}

// o[i] += sum_t tap[t] * w[t][i] with the packed fp32 FMA of sm_100 (FFMA2: two IEEE fmas per issue slot — same bits
// as scalar fmaf, half the instructions; these kernels are issue-bound)
__device__ __forceinline__ void fma_taps(const float (&tap)[9], const float (&w)[9][8], float (&o)[8]) {
    float2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float2(o[2 * j], o[2 * j + 1]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float2 tv = make_float2(tap[t], tap[t]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(tv, make_float2(w[t][2 * j], w[t][2 * j + 1]), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = acc[j].x; o[2 * j + 1] = acc[j].y; }
}

// 8 consecutive channels kept packed between the load and the use
template <typename T>
struct Packed8;
template <>
struct Packed8<bf16> {
    uint4 v;
    __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
};
template <>
struct Packed8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = reinterpret_cast<const float4*>(p)[0];
        b = reinterpret_cast<const float4*>(p)[1];
    }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
};

// stage w[C][9] -> smem [9][C]
__device__ __forceinline__ void stage_filter(const float* __restrict__ w, float* ws, int C) {
    for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < 9 * C; i += blockDim.x * blockDim.y) {
        int c = i / 9, t = i - c * 9;
        ws[t * C + c] = w[i];
    }
    __syncthreads();
}

// ---- first conv forward: thread = (storage position of out, 8 channels) -------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv3x3s2_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               int H, int W, DView<T> out, long long total) {
    extern __shared__ float wsm[];
    stage_filter(w, wsm, out.C);                      // once per CTA; the CTA then strides over many pixels
    const int C = out.C;
    const unsigned CG = (unsigned)C / CV8;
    // the 9 x 8 filter taps of this thread's channel octet live in REGISTERS: when C/8 divides the grid stride (every
    // power-of-two width) the octet never changes, so the inner loop is 9 broadcast loads + 72 FMAs + one 16-byte store
    // instead of 18 shared-memory vector loads per pixel (which bounded the kernel at ~1 TB/s of output)
    float wr[9][CV8];
    int cprev = -1;
    // each CTA owns a CONTIGUOUS run of pixels (not a grid-stride comb): successive iterations walk along image rows, so
    // the 3x3 tap loads of neighbouring pixels hit L1 (a comb jumps ~9 images per iteration: 36 % L1 hits, 7 long-
    // scoreboard stalls per issue in ncu)
    const unsigned per_cta = (unsigned)(((total + gridDim.x - 1) / gridDim.x + 255) / 256) * 256u;
    const unsigned cta_begin = blockIdx.x * per_cta;
    const unsigned cta_end = (unsigned)min((long long)cta_begin + per_cta, total);
    for (unsigned idx = cta_begin + threadIdx.x; idx < cta_end; idx += 256u) {
        const int c = (int)(idx % CG) * CV8;
        if (c != cprev) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 w0 = *reinterpret_cast<const float4*>(wsm + t * C + c);
                const float4 w1 = *reinterpret_cast<const float4*>(wsm + t * C + c + 4);
                wr[t][0] = w0.x; wr[t][1] = w0.y; wr[t][2] = w0.z; wr[t][3] = w0.w;
                wr[t][4] = w1.x; wr[t][5] = w1.y; wr[t][6] = w1.z; wr[t][7] = w1.w;
            }
            cprev = c;
        }
        unsigned pos = out.fCV8.div(idx), ws, hs, q, bb;
        out.fWs.divmod(pos, q, ws);
        out.fHs.divmod(q, bb, hs);
        const int b = (int)bb, oy = (int)hs - out.ph, ox = (int)ws - out.pw;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        if (oy >= 0 && oy < out.H && ox >= 0 && ox < out.W) {
            float xv[9];
            load_taps(x + (long long)b * H * W, W, oy, ox, xv);
            fma_taps(xv, wr, o);
        }
        st8<T>(out.p + (long long)b * out.sB + (long long)oy * out.sH + (long long)ox * out.sW + c, o);
    }
}
extern "C" int mopoe_conv3x3s2_c1_fwd(const float* x, const float* w, int B, int H, int W, const mopoe_view_t* out,
                                      void* stream) {
    MOPOE_REQUIRE(out->B == B && out->H == H / 2 && out->W == W / 2 && out->C % CV8 == 0, "conv3x3s2_c1_fwd: bad out view");
    long long total = (long long)B * (out->H + 2 * out->ph) * (out->W + 2 * out->pw) * (out->C / CV8);
    MOPOE_DISPATCH_T(out->dtype, T, {
        conv3x3s2_c1_fwd_kernel<T><<<(unsigned)min((long long)ceil_div64(total, 256), 148ll * 16), 256, 9 * out->C * sizeof(float), (cudaStream_t)stream>>>(
            x, w, H, W, make_dview<T>(out), total);
    });
    MOPOE_CHECK_LAUNCH("conv3x3s2_c1_fwd");
    return 0;
}

// ---- tap-gradient reduction shared by both layers ------------------------------------------------------------
// acc[c, tap] = sum_rows v[row, c] * s[row, tap]   where v is a channel vector (dy or x) and s the 9 scalar taps
// block = (16 channel-octet lanes, 16 row lanes); grid = (ceil(C/128), nchunk); ws layout [chunk][9][C]
template <typename T, bool FIRST>
__global__ void __launch_bounds__(256, 2) tap_grad_kernel(DView<const T> v, const float* __restrict__ s, int SH, int SW,
                                                       double* __restrict__ ws, int nchunk) {
    // FIRST: v = dy [B,OH,OW,C], s = image x [B,SH,SW], tap (ky,kx) reads s[2*oy-1+ky, 2*ox-1+kx]
    // else : v = x  [B,H,W,C],   s = dout [B,SH=2H,SW=2W],   tap (ky,kx) reads s[2*t-1+ky, 2*s-1+kx]  (same form)
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = (blockIdx.x * 16 + tx) * CV8;
    const bool cvalid = c < v.C;
    const unsigned rows = (unsigned)v.B * v.H * v.W;
    // block y owns a CONTIGUOUS run of rows (16-row groups walk along image rows: the scalar taps hit L1)
    const unsigned rpc = ((rows + nchunk - 1) / nchunk + 15) / 16 * 16;
    const unsigned row_begin = blockIdx.y * rpc, row_end = min(rows, row_begin + rpc);
    const unsigned gstride = 16;
    float acc[9][CV8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < CV8; ++i) acc[t][i] = 0.f;
    if (cvalid) {
        // The pass is latency-bound: what matters is bytes in flight per SM.  The activation octets of U rows are
        // prefetched PACKED (4 registers each for bf16) before the first use; the 9 scalar taps of a row are L1 hits and
        // are loaded just in time, so they do not occupy registers while the activation loads are outstanding.
        constexpr int U = sizeof(T) == 2 ? 8 : 4;
        for (unsigned r0 = row_begin + ty; r0 < row_end; r0 += U * gstride) {
            Packed8<T> g[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned r = r0 + u * gstride;
                if (r < row_end) {
                    unsigned px, t2, py, b;
                    v.fW.divmod(r, t2, px);
                    v.fH.divmod(t2, b, py);
                    g[u].load(v.p + (long long)b * v.sB + (long long)py * v.sH + (long long)px * v.sW + c);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned r = r0 + u * gstride;
                if (r < row_end) {
                    unsigned px, t2, py, b;
                    v.fW.divmod(r, t2, px);
                    v.fH.divmod(t2, b, py);
                    float sv[9], gf[CV8];
                    load_taps(s + (long long)b * SH * SW, SW, (int)py, (int)px, sv);
                    g[u].unpack(gf);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const float2 tv = make_float2(sv[t], sv[t]);
#pragma unroll
                        for (int j = 0; j < CV8 / 2; ++j) {
                            const float2 rr = __ffma2_rn(make_float2(gf[2 * j], gf[2 * j + 1]), tv,
                                                         make_float2(acc[t][2 * j], acc[t][2 * j + 1]));
                            acc[t][2 * j] = rr.x;
                            acc[t][2 * j + 1] = rr.y;
                        }
                    }
                }
            }
        }
    }
    // rows per thread <= rpc/16 (a few hundred): fp32 strip sums, combined across the 16 row lanes in fp64
    __shared__ double sm[16][16][CV8 + 1];
    for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int i = 0; i < CV8; ++i) sm[ty][tx][i] = (double)acc[t][i];
        __syncthreads();
        if (ty < CV8 && cvalid) {        // thread (tx, ty=i) reduces channel c+i over the 16 row lanes
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j) a += sm[j][tx][ty];
            ws[((long long)blockIdx.y * 9 + t) * v.C + c + ty] = a;
        }
        __syncthreads();
    }
}
// out[c*ntap + t] (+)= sum_k ws[(k*ntap + t)*C + c]   (one warp per output, fixed lane-strided order)
__global__ void __launch_bounds__(256) tap_finalize_kernel(const double* ws, int nchunk, int ntap, int C, float* out,
                                                           int accumulate) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= C * ntap) return;
    const int c = i / ntap, t = i % ntap, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int k = lane; k < nchunk; k += 32) s += ws[((long long)k * ntap + t) * C + c];
    s = warp_sum(s);
    if (lane == 0) out[i] = (accumulate ? out[i] : 0.f) + (float)s;
}

extern "C" int mopoe_conv3x3s2_c1_wgrad(const float* x, const mopoe_view_t* dy, int B, int H, int W, float* dw,
                                        int accumulate, double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(dy->B == B && dy->H == H / 2 && dy->W == W / 2 && dy->C % CV8 == 0, "conv3x3s2_c1_wgrad: bad dy view");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 block(16, 16), grid((dy->C + 127) / 128, nchunk);
    MOPOE_DISPATCH_T(dy->dtype, T, {
        tap_grad_kernel<T, true><<<grid, block, 0, st>>>(make_dview<const T>(dy), x, H, W, ws, nchunk);
    });
    MOPOE_CHECK_LAUNCH("conv3x3s2_c1_wgrad");
    tap_finalize_kernel<<<(dy->C * 9 + 7) / 8, 256, 0, st>>>(ws, nchunk, 9, dy->C, dw, accumulate);
    MOPOE_CHECK_LAUNCH("tap_finalize");
    return 0;
}

// ---- last deconv forward: one warp per 2x2 output quad, filter taps in registers ----------------------------------
// out(2t,2s)     = x(t,s).w11
// out(2t,2s+1)   = x(t,s).w12 + x(t,s+1).w10
// out(2t+1,2s)   = x(t,s).w21 + x(t+1,s).w01
// out(2t+1,2s+1) = x(t,s).w22 + x(t,s+1).w20 + x(t+1,s).w02 + x(t+1,s+1).w00      (w[ky][kx], oy = 2iy-1+ky)
template <typename T>
__global__ void __launch_bounds__(256) deconv3x3s2_c1_fwd_kernel(DView<const T> x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                                 long long quads) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const float bb = bias[0];
    for (int cb = 0; cb < x.C; cb += 128) {            // one pass per 128-channel slab (one pass when C <= 128)
        const int c = cb + lane * 4;
        const bool cvalid = c < x.C;
        float wr[4][9];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int t = 0; t < 9; ++t) wr[i][t] = cvalid ? __ldg(w + (c + i) * 9 + t) : 0.f;
        for (long long q = warp0; q < quads; q += nwarps) {
            int s = (int)(q % x.W);
            long long t2 = q / x.W;
            int t = (int)(t2 % x.H);
            int b = (int)(t2 / x.H);
            const bool rs = s + 1 < x.W, dn = t + 1 < x.H;
            float a[4] = {0, 0, 0, 0}, ar[4] = {0, 0, 0, 0}, ad[4] = {0, 0, 0, 0}, adr[4] = {0, 0, 0, 0};
            if (cvalid) {
                const T* p = x.p + (long long)b * x.sB + (long long)t * x.sH + (long long)s * x.sW + c;
                ldv<4>(p, a);
                if (rs) ldv<4>(p + x.sW, ar);
                if (dn) ldv<4>(p + x.sH, ad);
                if (rs && dn) ldv<4>(p + x.sH + x.sW, adr);
            }
            float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o00 += a[i] * wr[i][4];
                o01 += a[i] * wr[i][5] + ar[i] * wr[i][3];
                o10 += a[i] * wr[i][7] + ad[i] * wr[i][1];
                o11 += a[i] * wr[i][8] + ar[i] * wr[i][6] + ad[i] * wr[i][2] + adr[i] * wr[i][0];
            }
            // transposing butterfly: 4 values x 32 lanes -> 4 totals with 2 + 1 + 3 shuffles instead of 4 x 5
            {
                const bool hi = lane & 16;
                float sa = hi ? o00 : o10, sb = hi ? o01 : o11;          // what this lane gives away
                float ka = hi ? o10 : o00, kb = hi ? o11 : o01;          // what it keeps (lo: row 0, hi: row 1)
                ka += __shfl_xor_sync(0xffffffffu, sa, 16);
                kb += __shfl_xor_sync(0xffffffffu, sb, 16);
                const bool h8 = lane & 8;
                float give = h8 ? ka : kb, keep = h8 ? kb : ka;          // lo8: col 0, hi8: col 1
                keep += __shfl_xor_sync(0xffffffffu, give, 8);
                keep += __shfl_xor_sync(0xffffffffu, keep, 4);
                keep += __shfl_xor_sync(0xffffffffu, keep, 2);
                keep += __shfl_xor_sync(0xffffffffu, keep, 1);
                // lanes 0 / 8 / 16 / 24 now hold o00 / o01 / o10 / o11
                o00 = __shfl_sync(0xffffffffu, keep, 0);
                o01 = __shfl_sync(0xffffffffu, keep, 8);
                o10 = __shfl_sync(0xffffffffu, keep, 16);
                o11 = __shfl_sync(0xffffffffu, keep, 24);
            }
            if (lane == 0) {
                const int OW = 2 * x.W;
                float* ob = out + ((long long)b * 2 * x.H + 2 * t) * OW + 2 * s;
                if (cb == 0) {
                    *reinterpret_cast<float2*>(ob) = make_float2(o00 + bb, o01 + bb);
                    *reinterpret_cast<float2*>(ob + OW) = make_float2(o10 + bb, o11 + bb);
                } else {
                    ob[0] += o00; ob[1] += o01; ob[OW] += o10; ob[OW + 1] += o11;
                }
            }
        }
    }
}
// Two-phase form of the same layer (the default): phase A computes, for every INPUT pixel, the 9 channel dot products
//   p[k] = sum_c x[b,t,s,c] * w[c,k]      (one thread per pixel walking its 256 contiguous bytes; the [C][9] filter is
// read from shared memory as warp-wide broadcasts; 2 pixels per thread share each filter read; packed FFMA2)
// and phase B assembles each 2x2 output quad from the taps of its (up to) 4 source pixels.  Every activation byte is
// read exactly once and there is no cross-lane reduction: the warp-per-quad kernel above spends its time in a 7-deep
// shuffle chain and 4x neighbour re-reads (0.47 ms for a 268 MB input; HBM floor 0.05 ms).
constexpr int TAPS_STRIDE = 12;      // 9 taps padded to 3 float4
template <typename T>
__global__ void __launch_bounds__(128) deconv_taps_kernel(DView<const T> x, const float* __restrict__ w,
                                                          float* __restrict__ taps, long long pixels) {
    extern __shared__ float wsm[];                   // [C][12]
    const int C = x.C;
    for (int i = threadIdx.x; i < C * TAPS_STRIDE; i += 128) {
        const int c = i / TAPS_STRIDE, k = i - c * TAPS_STRIDE;
        wsm[i] = k < 9 ? w[c * 9 + k] : 0.f;
    }
    __syncthreads();
    const long long half = (pixels + 1) / 2;
    for (long long p0 = (long long)blockIdx.x * 128 + threadIdx.x; p0 < half; p0 += (long long)gridDim.x * 128) {
        const long long p1 = p0 + half;
        const bool has1 = p1 < pixels;
        const T* xp[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long pp = u == 0 ? p0 : (has1 ? p1 : p0);
            const int s_ = (int)(pp % x.W);
            const long long t2 = pp / x.W;
            const int t_ = (int)(t2 % x.H), b_ = (int)(t2 / x.H);
            xp[u] = x.p + (long long)b_ * x.sB + (long long)t_ * x.sH + (long long)s_ * x.sW;
        }
        float2 acc[2][5];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 5; ++k) acc[u][k] = make_float2(0.f, 0.f);
        for (int c0 = 0; c0 < C; c0 += CV8) {
            float xv[2][CV8];
            ld8<T>(xp[0] + c0, xv[0]);
            ld8<T>(xp[1] + c0, xv[1]);
#pragma unroll
            for (int i = 0; i < CV8; ++i) {
                const float4 wa = *reinterpret_cast<const float4*>(wsm + (c0 + i) * TAPS_STRIDE);
                const float4 wb = *reinterpret_cast<const float4*>(wsm + (c0 + i) * TAPS_STRIDE + 4);
                const float4 wc = *reinterpret_cast<const float4*>(wsm + (c0 + i) * TAPS_STRIDE + 8);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float2 xx = make_float2(xv[u][i], xv[u][i]);
                    acc[u][0] = __ffma2_rn(xx, make_float2(wa.x, wa.y), acc[u][0]);
                    acc[u][1] = __ffma2_rn(xx, make_float2(wa.z, wa.w), acc[u][1]);
                    acc[u][2] = __ffma2_rn(xx, make_float2(wb.x, wb.y), acc[u][2]);
                    acc[u][3] = __ffma2_rn(xx, make_float2(wb.z, wb.w), acc[u][3]);
                    acc[u][4] = __ffma2_rn(xx, make_float2(wc.x, wc.y), acc[u][4]);     // wc.y is the zero pad
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !has1) break;
            float4* o = reinterpret_cast<float4*>(taps + (u == 0 ? p0 : p1) * TAPS_STRIDE);
            o[0] = make_float4(acc[u][0].x, acc[u][0].y, acc[u][1].x, acc[u][1].y);
            o[1] = make_float4(acc[u][2].x, acc[u][2].y, acc[u][3].x, acc[u][3].y);
            o[2] = make_float4(acc[u][4].x, 0.f, 0.f, 0.f);
        }
    }
}
// out(2t,2s) = p11(t,s);  out(2t,2s+1) = p12(t,s) + p10(t,s+1);  out(2t+1,2s) = p21(t,s) + p01(t+1,s);
// out(2t+1,2s+1) = p22(t,s) + p20(t,s+1) + p02(t+1,s) + p00(t+1,s+1)        (p_kykx = taps[ky*3+kx])
__global__ void __launch_bounds__(256) deconv_assemble_kernel(const float* __restrict__ taps, const float* __restrict__ bias,
                                                              float* __restrict__ out, int H, int W, long long quads,
                                                              int TS) {                  // TS: floats per pixel in `taps`
    const float bb = bias[0];
    for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < quads; q += (long long)gridDim.x * 256) {
        const int s = (int)(q % W);
        const long long t2 = q / W;
        const int t = (int)(t2 % H);
        const long long b = t2 / H;
        const bool rs = s + 1 < W, dn = t + 1 < H;
        const float* p = taps + q * TS;
        const float* pr = p + TS;
        const float* pd = p + (long long)W * TS;
        const float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
        const float a8 = p[8];
        float o00 = a1.x, o01 = a1.y, o10 = a1.w, o11 = a8;        // p11, p12, p21, p22
        if (rs) { o01 += pr[3]; o11 += pr[6]; }                    // p10, p20 of (t, s+1)
        if (dn) { o10 += pd[1]; o11 += pd[2]; }                    // p01, p02 of (t+1, s)
        if (rs && dn) o11 += pd[TS];                               // p00 of (t+1, s+1)
        (void)a0;
        const int OW = 2 * W;
        float* ob = out + (b * 2 * H + 2 * t) * OW + 2 * s;
        *reinterpret_cast<float2*>(ob) = make_float2(o00 + bb, o01 + bb);
        *reinterpret_cast<float2*>(ob + OW) = make_float2(o10 + bb, o11 + bb);
    }
}
extern "C" size_t mopoe_deconv3x3s2_c1_fwd_ws(const mopoe_view_t* x) {
    return (size_t)x->B * x->H * x->W * TAPS_STRIDE * sizeof(float);
}
extern "C" int mopoe_deconv3x3s2_c1_fwd(const mopoe_view_t* x, const float* w, const float* bias, float* out, void* ws,
                                        size_t ws_bytes, void* stream) {
    MOPOE_REQUIRE(x->C % CV8 == 0, "deconv3x3s2_c1_fwd: C=%d", x->C);
    const long long quads = (long long)x->B * x->H * x->W;
    cudaStream_t st = (cudaStream_t)stream;
    if (ws == nullptr) {                 // no workspace: single-kernel warp-per-quad form
        long long blocks = ceil_div64(quads, 8 * 8);
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (blocks < 1) blocks = 1;
        MOPOE_DISPATCH_T(x->dtype, T, {
            deconv3x3s2_c1_fwd_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(make_dview<const T>(x), w, bias, out, quads);
        });
        MOPOE_CHECK_LAUNCH("deconv3x3s2_c1_fwd");
        return 0;
    }
    MOPOE_REQUIRE(ws_bytes >= mopoe_deconv3x3s2_c1_fwd_ws(x), "deconv3x3s2_c1_fwd: workspace %zu too small", ws_bytes);
    MOPOE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "deconv3x3s2_c1_fwd: unaligned workspace");
    long long blocks = ceil_div64((quads + 1) / 2, 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)x->C * TAPS_STRIDE * sizeof(float);
    MOPOE_REQUIRE(smem <= 48 * 1024, "deconv3x3s2_c1_fwd: C=%d too large for the filter stage", x->C);
    MOPOE_DISPATCH_T(x->dtype, T, {
        deconv_taps_kernel<T><<<(unsigned)blocks, 128, smem, st>>>(make_dview<const T>(x), w, (float*)ws, quads);
    });
    MOPOE_CHECK_LAUNCH("deconv_taps");
    long long b2 = ceil_div64(quads, 256);
    if (b2 > 148 * 16) b2 = 148 * 16;
    deconv_assemble_kernel<<<(unsigned)b2, 256, 0, st>>>((const float*)ws, bias, out, x->H, x->W, quads, TAPS_STRIDE);
    MOPOE_CHECK_LAUNCH("deconv_assemble");
    return 0;
}

// the assembly half alone: the 9 tap products per input pixel come from a tensor-core GEMM ([M, C] x [C, 16] -> taps
// [M, stride] fp32, mopoe_conv_gemm with the 16-row zero-padded filter) instead of deconv_taps_kernel
extern "C" int mopoe_deconv3x3s2_c1_assemble(const float* taps, int stride, const float* bias, float* out, int B, int H, int W,
                                             void* stream) {
    MOPOE_REQUIRE(stride >= 9 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(taps) & 15) == 0, "deconv_assemble: stride=%d", stride);
    const long long quads = (long long)B * H * W;
    long long b2 = ceil_div64(quads, 256);
    if (b2 > 148 * 16) b2 = 148 * 16;
    deconv_assemble_kernel<<<(unsigned)b2, 256, 0, (cudaStream_t)stream>>>(taps, bias, out, H, W, quads, stride);
    MOPOE_CHECK_LAUNCH("deconv_assemble");
    return 0;
}

// ---- last deconv backward ------------------------------------------------------------------------------------
// dx[b,t,s,c] = sum_{ky,kx} dout[b, 2t-1+ky, 2s-1+kx] * w[c,ky,kx]
template <typename T>
__global__ void __launch_bounds__(256) deconv3x3s2_c1_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                                DView<T> dx, long long total) {
    extern __shared__ float wsm[];
    stage_filter(w, wsm, dx.C);
    const int C = dx.C;
    const unsigned CG = (unsigned)C / CV8;
    const int OH = 2 * dx.H, OW = 2 * dx.W;
    float wr[9][CV8];                                  // register-resident taps (see conv3x3s2_c1_fwd_kernel)
    int cprev = -1;
    // each CTA owns a CONTIGUOUS run of pixels (not a grid-stride comb): successive iterations walk along image rows, so
    // the 3x3 tap loads of neighbouring pixels hit L1 (a comb jumps ~9 images per iteration: 36 % L1 hits, 7 long-
    // scoreboard stalls per issue in ncu)
    const unsigned per_cta = (unsigned)(((total + gridDim.x - 1) / gridDim.x + 255) / 256) * 256u;
    const unsigned cta_begin = blockIdx.x * per_cta;
    const unsigned cta_end = (unsigned)min((long long)cta_begin + per_cta, total);
    for (unsigned idx = cta_begin + threadIdx.x; idx < cta_end; idx += 256u) {
        const int c = (int)(idx % CG) * CV8;
        if (c != cprev) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 w0 = *reinterpret_cast<const float4*>(wsm + t * C + c);
                const float4 w1 = *reinterpret_cast<const float4*>(wsm + t * C + c + 4);
                wr[t][0] = w0.x; wr[t][1] = w0.y; wr[t][2] = w0.z; wr[t][3] = w0.w;
                wr[t][4] = w1.x; wr[t][5] = w1.y; wr[t][6] = w1.z; wr[t][7] = w1.w;
            }
            cprev = c;
        }
        unsigned pos = dx.fCV8.div(idx), ws, hs, q, bb;
        dx.fWs.divmod(pos, q, ws);
        dx.fHs.divmod(q, bb, hs);
        const int b = (int)bb, t = (int)hs - dx.ph, s = (int)ws - dx.pw;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        if (t >= 0 && t < dx.H && s >= 0 && s < dx.W) {
            float gv[9];
            load_taps(dout + (long long)b * OH * OW, OW, t, s, gv);
            fma_taps(gv, wr, o);
        }
        st8<T>(dx.p + (long long)b * dx.sB + (long long)t * dx.sH + (long long)s * dx.sW + c, o);
    }
}
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ v, long long n, double* __restrict__ part) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc += (double)v[i];
    __shared__ double sm[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += sm[i];
        part[blockIdx.x] = s;
    }
}
__global__ void sum_final_kernel(const double* part, int n, float* out, int accumulate) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)s;
}
extern "C" int mopoe_deconv3x3s2_c1_bwd(const mopoe_view_t* x, const float* w, const float* dout, const mopoe_view_t* dx,
                                        float* dw, float* dbias, int accumulate, double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(x->C % CV8 == 0 && (!dx || (dx->C == x->C && dx->B == x->B && dx->H == x->H && dx->W == x->W &&
                                              dx->dtype == x->dtype)), "deconv3x3s2_c1_bwd: bad views");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = dx ? (long long)dx->B * (dx->H + 2 * dx->ph) * (dx->W + 2 * dx->pw) * (dx->C / CV8) : 0;
    dim3 block(16, 16), grid((x->C + 127) / 128, nchunk);
    MOPOE_DISPATCH_T(x->dtype, T, {
        // (dx == NULL: the caller forms it as a tensor-core GEMM over mopoe_im2col3x3s2 patches of dout)
        if (dx) deconv3x3s2_c1_dx_kernel<T><<<(unsigned)min((long long)ceil_div64(total, 256), 148ll * 16), 256, 9 * x->C * sizeof(float), st>>>(
            dout, w, make_dview<T>(dx), total);
        if (dw) tap_grad_kernel<T, false><<<grid, block, 0, st>>>(make_dview<const T>(x), dout, 2 * x->H, 2 * x->W, ws, nchunk);
    });
    MOPOE_CHECK_LAUNCH("deconv3x3s2_c1_bwd");
    if (dw) {          // (NULL: the caller forms the tap gradient as a tcgen05 weight-gradient GEMM over mopoe_im2col3x3s2)
        tap_finalize_kernel<<<(x->C * 9 + 7) / 8, 256, 0, st>>>(ws, nchunk, 9, x->C, dw, accumulate);
        MOPOE_CHECK_LAUNCH("tap_finalize");
    }
    double* part = ws + (long long)nchunk * 9 * x->C;
    long long n = (long long)x->B * 4 * x->H * x->W;
    sum_partial_kernel<<<nchunk, 256, 0, st>>>(dout, n, part);
    sum_final_kernel<<<1, 32, 0, st>>>(part, nchunk, dbias, accumulate);
    MOPOE_CHECK_LAUNCH("dbias");
    return 0;
}

// ---- 3x3 / stride-2 / pad-1 patches of a single-channel image as a 16-column bf16 matrix ------------------------------
// out[(b, oy, ox)][t = ky*3 + kx] = src[b, 2*oy - 1 + ky, 2*ox - 1 + kx]  (0 outside the image, 0 for t >= 9).
// With it the two single-channel weight gradients are ordinary weight-gradient GEMMs on the tensor cores:
//   first conv  (FeatureExtractorImg.py:29-34):  dW[c, t] = sum_m dY[m, c] * patches(x)[m, t]
//   last deconv (DataGeneratorImg.py:84-90):      dW[c, t] = sum_m  X[m, c] * patches(dOut)[m, t]
// i.e. mopoe_conv_wgrad with the 128-channel activation as the window operand and the patches as the 16-wide row operand —
// one streaming read of the activation at tensor-core speed instead of the register-blocked CUDA-core reduction
// (tap_grad_kernel: 235 us per launch against a 45 us HBM floor).
__global__ void __launch_bounds__(256) im2col3x3s2_kernel(const float* __restrict__ src, int SH, int SW, long long pixels,
                                                          int cols, bf16* __restrict__ out) {
    const int OW = SW / 2, OH = SH / 2;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < pixels; p += (long long)gridDim.x * 256) {
        const int ox = (int)(p % OW);
        const long long t2 = p / OW;
        const int oy = (int)(t2 % OH);
        const long long b = t2 / OH;
        float t[9];
        load_taps(src + b * SH * SW, SW, oy, ox, t);
        uint4 lo, hi;
        __nv_bfloat162 h;
        h = __floats2bfloat162_rn(t[0], t[1]); lo.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(t[2], t[3]); lo.y = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(t[4], t[5]); lo.z = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(t[6], t[7]); lo.w = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(t[8], 0.f); hi.x = *reinterpret_cast<uint32_t*>(&h);
        hi.y = hi.z = hi.w = 0u;
        uint4* o = reinterpret_cast<uint4*>(out + p * cols);
        o[0] = lo;
        o[1] = hi;
        for (int k = 2; k < cols / 8; ++k) o[k] = make_uint4(0u, 0u, 0u, 0u);      // K padded to the GEMM's 64-element k-block
    }
}
extern "C" int mopoe_im2col3x3s2(const float* src, int B, int SH, int SW, int cols, void* out_bf16, void* stream) {
    MOPOE_REQUIRE(B > 0 && SH > 0 && SW > 0 && SH % 2 == 0 && SW % 2 == 0, "im2col3x3s2: bad shape [%d,%d,%d]", B, SH, SW);
    MOPOE_REQUIRE(cols >= 16 && cols % 8 == 0, "im2col3x3s2: cols=%d (>= 16, multiple of 8)", cols);
    MOPOE_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0, "im2col3x3s2: unaligned output");
    const long long pixels = (long long)B * (SH / 2) * (SW / 2);
    long long blocks = ceil_div64(pixels, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    im2col3x3s2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, SH, SW, pixels, cols, reinterpret_cast<bf16*>(out_bf16));
    MOPOE_CHECK_LAUNCH("im2col3x3s2");
    return 0;
}
