// elementwise.cu — HBM-bound passes of the residual blocks: training-mode BatchNorm statistics,
// fused BN+ReLU(+dropout) apply, residual combine, their backward halves, layout conversion,
// dropout-mask generation, weight re-layout and the flat Adam update.
//
// All streaming kernels address activations through channels-last views; one thread owns 8 consecutive
// channels of one pixel (16-byte bf16 / 2x16-byte fp32 vector accesses, a warp covers 256 channels-pixels
// contiguously), index math is 32-bit, and every reduction accumulates short fp32 strips into fp64 with a
// fixed two-stage order (per-chunk partials, then one warp per channel sums the chunks lane-strided) —
// run-to-run deterministic, no atomics.
#include <stdarg.h>

#include "common.cuh"

// ---- error string ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void mopoe_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* mopoe_last_error(void) { return g_err; }
extern "C" int mopoe_version(void) { return 101; }

constexpr int VEC = 8;
constexpr int EW_THREADS = 256;

// shared-memory-staged persistent variants (stream.cu): 0 launched, 1 error, -1 shape not eligible
int mopoe_staged_bn_apply(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, int relu, const mopoe_view_t* out, cudaStream_t st);
int mopoe_staged_combine(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         const mopoe_view_t* c, const uint8_t* mask, int mask_mode, float a, float b, const mopoe_view_t* out,
                         double* ws, int nchunk_cap, int* nchunk_used, cudaStream_t st);
int mopoe_staged_bn_bwd_apply(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale, const mopoe_view_t* x,
                              const uint8_t* mask, int mask_mode, const float* mean, const float* invstd, const float* gamma,
                              const float* sums, const mopoe_view_t* addend, const mopoe_view_t* out, const mopoe_view_t* out2,
                              const uint8_t* mask2, int mask2_mode, float scale2, const float* gate_beta, cudaStream_t st);
int mopoe_staged_reduce(int mode, const mopoe_view_t* x, const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                        const uint8_t* mask, int mask_mode, const float* mean, const float* invstd, double* ws, int nchunk_cap,
                        int* nchunk_used, const float* gate_gamma, const float* gate_beta, cudaStream_t st);
static int ew_staged() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_EW_STAGED");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static int check_same(const mopoe_view_t* a, const mopoe_view_t* b, const char* what) {
    if (a->B != b->B || a->H != b->H || a->W != b->W || a->C != b->C || a->dtype != b->dtype)
        MOPOE_FAIL("%s: view mismatch [%d,%d,%d,%d]/%d vs [%d,%d,%d,%d]/%d", what, a->B, a->H, a->W, a->C,
                   a->dtype, b->B, b->H, b->W, b->C, b->dtype);
    return 0;
}

template <typename T>
__device__ __forceinline__ void ld8v(const T* p, float (&o)[8]) {
    if constexpr (sizeof(T) == 2) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
    } else {
        float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
}
template <typename T>
__device__ __forceinline__ void st8v(T* p, const float (&o)[8]) {
    if constexpr (sizeof(T) == 2) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    } else {
        reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}
__device__ __forceinline__ void ld8f(const float* p, float (&o)[8]) { ld8v<float>(p, o); }
// 8 consecutive channels held PACKED (4 registers for bf16) between the load and the use
template <typename T>
struct Raw8;
template <>
struct Raw8<bf16> {
    uint4 v;
    __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
    }
};
template <>
struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = reinterpret_cast<const float4*>(p)[0];
        b = reinterpret_cast<const float4*>(p)[1];
    }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
};
// keep-mask bytes -> multiplier 2 (keep) / 0 (drop); bc_idx / el_idx are element indices into the mask
__device__ __forceinline__ void ldmask8(const uint8_t* m, int mode, int bc_idx, int el_idx, float (&o)[8]) {
    if (mode == MOPOE_MASK_NONE) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 1.f;
    } else {
        const uint2 t = *reinterpret_cast<const uint2*>(m + (mode == MOPOE_MASK_BC ? bc_idx : el_idx));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = ((t.x >> (8 * i)) & 0xffu) ? 2.f : 0.f;
            o[4 + i] = ((t.y >> (8 * i)) & 0xffu) ? 2.f : 0.f;
        }
    }
}

__device__ __forceinline__ void mask8(const uint2& t, float (&o)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o[i] = ((t.x >> (8 * i)) & 0xffu) ? 2.f : 0.f;
        o[4 + i] = ((t.y >> (8 * i)) & 0xffu) ? 2.f : 0.f;
    }
}

// y = v * sc + sh  ==  gamma * (v - mean) * invstd + beta.  ONE definition shared by the forward apply kernel and
// the backward kernels that RECOMPUTE the ReLU gate from x instead of re-reading the activation: identical
// instruction sequence -> bit-identical sign decisions.
__device__ __forceinline__ void bn_affine(const float (&mu)[8], const float (&is)[8], const float (&ga)[8],
                                          const float (&be)[8], float (&sc)[8], float (&sh)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {                      // explicit intrinsics: identical in every kernel (stream.cu affine8)
        sc[i] = __fmul_rn(is[i], ga[i]);
        sh[i] = __fmaf_rn(-mu[i], sc[i], be[i]);
    }
}
__device__ __forceinline__ float bn_eval(float v, float sc, float sh) { return __fmaf_rn(v, sc, sh); }

// position of a flat thread index over the STORAGE (interior + zero border) of an output view
struct Pos {
    int b, h, w, c;
    bool interior;
};
template <typename T>
__device__ __forceinline__ Pos decode_storage(const DView<T>& o, unsigned idx) {
    const unsigned CV = (unsigned)o.C / VEC, Ws = (unsigned)(o.W + 2 * o.pw), Hs = (unsigned)(o.H + 2 * o.ph);
    Pos p;
    p.c = (int)(idx % CV) * VEC;
    unsigned pos = idx / CV;
    const unsigned ws = pos % Ws;
    pos /= Ws;
    const unsigned hs = pos % Hs;
    p.b = (int)(pos / Hs);
    p.h = (int)hs - o.ph;
    p.w = (int)ws - o.pw;
    p.interior = p.h >= 0 && p.h < o.H && p.w >= 0 && p.w < o.W;
    return p;
}
template <typename T>
static long long storage_threads(const DView<T>& o) {
    return (long long)o.B * (o.H + 2 * o.ph) * (o.W + 2 * o.pw) * (o.C / VEC);
}
// element offset relative to the interior origin; 32-bit (hosts check numel < 2^31), may be negative on the border
template <typename T>
__device__ __forceinline__ int vaddr(const DView<T>& v, int b, int h, int w, int c) {
    return b * (int)v.sB + h * (int)v.sH + w * (int)v.sW + c;
}
// Launch shape of the apply kernels: every thread keeps ONE channel octet (its per-channel parameters stay in
// registers) and strides over ~4 pixels, so the total thread count must be a multiple of C/8.
static unsigned gcd_u(unsigned a, unsigned b) { while (b) { unsigned t = a % b; a = b; b = t; } return a; }
static int apply_grid(long long total, int C, unsigned& grid, unsigned& stride) {
    MOPOE_REQUIRE(total > 0 && total < (1ll << 31), "elementwise: %lld work items do not fit 32-bit indexing", total);
    const unsigned CV = (unsigned)C / VEC;
    const unsigned m = CV / gcd_u(CV, EW_THREADS);          // blocks must be a multiple of m
    unsigned blocks = (unsigned)ceil_div64(ceil_div64(total, 4), EW_THREADS);
    blocks = (blocks + m - 1) / m * m;
    grid = blocks;
    stride = blocks * EW_THREADS;
    return 0;
}
template <typename T>
__device__ __forceinline__ Pos decode_pixel(const DView<T>& o, unsigned pos, int c) {
    Pos p;
    p.c = c;
    unsigned ws, hs, q, bb;
    o.fWs.divmod(pos, q, ws);
    o.fHs.divmod(q, bb, hs);
    p.b = (int)bb;
    p.h = (int)hs - o.ph;
    p.w = (int)ws - o.pw;
    p.interior = p.h >= 0 && p.h < o.H && p.w >= 0 && p.w < o.W;
    return p;
}

// ---- per-channel reductions (BN stats, BN backward sums, bias gradient) ------------------------------
// block = (16 channel-octet lanes, 16 row lanes); grid = (ceil(C/128), nchunk)
enum { RED_STATS = 0, RED_BNBWD = 1, RED_COLSUM = 2 };
constexpr int RED_ROWS = 16;
#ifndef RED_U_BWD
#define RED_U_BWD 4
#endif
#ifndef RED_U_ONE
#define RED_U_ONE 4
#endif

// Fused finalize: the LAST block of a channel group to finish (a self-resetting arrival counter) sums the per-chunk
// partials in a FIXED lane-strided order — so which block happens to be last does not change a single bit — and
// writes mean / invstd / running stats (BN statistics) or dbeta / dgamma / sums (BN backward, bias gradient).
struct FinArgs {
    int* counter;            // one int per channel group (grid.x); NULL -> separate finalize launch
    double count;
    float eps, momentum;
    float *mean, *invstd, *rmean, *rvar;     // RED_STATS
    float *out0, *out1, *sums;               // RED_BNBWD / RED_COLSUM
    int accumulate;
};
__device__ __forceinline__ void finalize_channel(int mode, const FinArgs& f, int C, int c, double s, double q) {
    if (mode == RED_STATS) {
        double m = s / f.count;
        double var = q / f.count - m * m;
        if (var < 0.0) var = 0.0;
        f.mean[c] = (float)m;
        f.invstd[c] = (float)(1.0 / sqrt(var + (double)f.eps));
        if (f.rmean) {
            double unb = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
            f.rmean[c] = (float)((1.0 - f.momentum) * (double)f.rmean[c] + f.momentum * m);
            f.rvar[c] = (float)((1.0 - f.momentum) * (double)f.rvar[c] + f.momentum * unb);
        }
    } else {
        if (f.out0) f.out0[c] = (f.accumulate ? f.out0[c] : 0.f) + (float)s;   // dbeta / colsum
        if (f.out1) f.out1[c] = (f.accumulate ? f.out1[c] : 0.f) + (float)q;   // dgamma
        if (f.sums) {
            f.sums[c] = (float)s;
            f.sums[C + c] = (float)q;
        }
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) reduce_rows_kernel(DView<const T> x, DView<const T> dy, DView<const T> gate,
                                                          int has_gate, float gscale, const uint8_t* mask,
                                                          int mask_mode, const float* mean, const float* invstd,
                                                          const float* ggamma, const float* gbeta, double* ws,
                                                          int nchunk, FinArgs fin) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = (blockIdx.x * 16 + tx) * VEC;
    const bool cvalid = c < x.C;
    // rows are dealt to the blocks in interleaved groups of 16 (block y takes groups y, y+nchunk, ...): at any moment
    // the whole grid streams one compact window of the tensor (DRAM-page / TLB friendly), unlike a blocked split
    const unsigned rows = (unsigned)x.B * x.H * x.W;
    const unsigned r1 = rows;
    const unsigned gstride = (unsigned)nchunk * RED_ROWS;
    float f0[VEC], f1[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) f0[i] = f1[i] = 0.f;
    float mu[VEC], is[VEC];
    if (MODE == RED_BNBWD && cvalid) {
        ld8f(mean + c, mu);
        ld8f(invstd + c, is);
    }
    if (cvalid) {
        constexpr int U = MODE == RED_BNBWD ? RED_U_BWD : RED_U_ONE;      // row groups in flight per thread
        for (unsigned rb = blockIdx.y * RED_ROWS + ty; rb < r1; rb += U * gstride) {
            // issue all loads of the U row groups first (kept PACKED: 4 registers per bf16 octet), unpack at use
            Raw8<T> xr[U], gr[U], tr[U];
            uint2 mr[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned r = rb + u * gstride;
                ok[u] = r < r1;
                if (ok[u]) {
                    unsigned w, t, h, b;
                    x.fW.divmod(r, t, w);
                    x.fH.divmod(t, b, h);
                    xr[u].load(x.p + vaddr(x, b, h, w, c));
                    if (mask_mode != MOPOE_MASK_NONE)
                        mr[u] = *reinterpret_cast<const uint2*>(mask + (mask_mode == MOPOE_MASK_BC ? (int)(b * x.C + c)
                                                                                                  : (int)(r * x.C + c)));
                    if (MODE == RED_BNBWD) {
                        gr[u].load(dy.p + vaddr(dy, b, h, w, c));
                        if (has_gate == 1) tr[u].load(gate.p + vaddr(gate, b, h, w, c));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    float xv[VEC], mk[VEC];
                    xr[u].unpack(xv);
                    if (mask_mode != MOPOE_MASK_NONE) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            mk[i] = ((mr[u].x >> (8 * i)) & 0xffu) ? 2.f : 0.f;
                            mk[4 + i] = ((mr[u].y >> (8 * i)) & 0xffu) ? 2.f : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) mk[i] = 1.f;
                    }
                    if (MODE == RED_STATS) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float v = xv[i] * mk[i];
                            f0[i] += v;
                            f1[i] = fmaf(v, v, f1[i]);
                        }
                    } else if (MODE == RED_COLSUM) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) f0[i] += xv[i];
                    } else {
                        float g[VEC], gt[VEC];
                        gr[u].unpack(g);
                        if (has_gate == 1) tr[u].unpack(gt);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            float gg = gscale * g[i];
                            if (has_gate == 1 && !(gt[i] > 0.f)) gg = 0.f;
                            const float xh = (xv[i] * mk[i] - mu[i]) * is[i];
                            f0[i] += gg;
                            f1[i] = fmaf(gg, xh, f1[i]);
                        }
                    }
                }
            }
        }
    }
    // each thread summed <= rows / (16 * nchunk) (~100) values in fp32; everything above that level is fp64
    __shared__ double sm[2][RED_ROWS][16][VEC + 1];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        sm[0][ty][tx][i] = (double)f0[i];
        sm[1][ty][tx][i] = (double)f1[i];
    }
    __syncthreads();
    // thread (tx, ty): ty < 8 reduces sum-0 of channel c+ty, ty >= 8 reduces sum-1 of channel c+ty-8
    if (cvalid) {
        const int which = ty >> 3, ch = ty & 7;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < RED_ROWS; ++j) a += sm[which][j][tx][ch];
        ws[((long long)blockIdx.y * 2 + which) * x.C + c + ch] = a;
    }
    if (fin.counter == nullptr) return;
    // ---- fused finalize by the last-arriving block of this channel group ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    const int tid = ty * 16 + tx;
    if (tid == 0) s_last = (atomicAdd(fin.counter + blockIdx.x, 1) == nchunk - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = 0; j < 16; ++j) {
        const int ch = blockIdx.x * 128 + warp * 16 + j;          // 8 warps x 16 channels = this group's 128
        if (ch >= x.C) break;
        double a = 0.0, b = 0.0;
        for (int k = lane; k < nchunk; k += 32) {
            a += __ldcg(ws + ((long long)k * 2 + 0) * x.C + ch);
            b += __ldcg(ws + ((long long)k * 2 + 1) * x.C + ch);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) finalize_channel(MODE, fin, x.C, ch, a, b);
    }
    if (tid == 0) fin.counter[blockIdx.x] = 0;                    // self-resetting: ready for the next launch
}

// finalize: FIN_SPLIT warps per channel (block = 8 warps = FIN_CH channels).  Warp j of a channel sums chunks
// j*32 + lane, + 32*FIN_SPLIT, ... then a shuffle tree, and the FIN_SPLIT warp totals are added in warp order — a fixed
// order, so the result is run-to-run deterministic.  ~9 dependent L2 round trips per lane instead of ~37: these
// launches are pure latency (the partials are a few MB in L2), and a training step has ~260 of them.
constexpr int FIN_SPLIT = 4, FIN_CH = 8 / FIN_SPLIT;
__device__ __forceinline__ bool chunk_sums(const double* ws, int nchunk, int C, int& c, double& s, double& q) {
    __shared__ double part[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int local = warp / FIN_SPLIT, j = warp - local * FIN_SPLIT;
    c = blockIdx.x * FIN_CH + local;
    double a = 0.0, b = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int k = j * 32 + lane; k < nchunk; k += 32 * FIN_SPLIT) {
            a += ws[((long long)k * 2 + 0) * C + c];
            b += ws[((long long)k * 2 + 1) * C + c];
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
        part[0][warp] = a;
        part[1][warp] = b;
    }
    __syncthreads();
    if (c >= C || j != 0 || lane != 0) return false;
    s = q = 0.0;
#pragma unroll
    for (int t = 0; t < FIN_SPLIT; ++t) {
        s += part[0][warp + t];
        q += part[1][warp + t];
    }
    return true;
}
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* ws, int nchunk, int C, double count, float eps,
                                                          float momentum, float* mean, float* invstd, float* rmean,
                                                          float* rvar) {
    int c;
    double s, q;
    if (!chunk_sums(ws, nchunk, C, c, s, q)) return;
    double m = s / count;
    double var = q / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) {
        double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * m);
        rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unb);
    }
}

__global__ void __launch_bounds__(256) sums_finalize_kernel(const double* ws, int nchunk, int C, float* out0, float* out1,
                                                            int accumulate, float* sums) {
    int c;
    double s, q;
    if (!chunk_sums(ws, nchunk, C, c, s, q)) return;
    if (out0) out0[c] = (accumulate ? out0[c] : 0.f) + (float)s;   // dbeta / colsum
    if (out1) out1[c] = (accumulate ? out1[c] : 0.f) + (float)q;   // dgamma
    if (sums) {
        sums[c] = (float)s;
        sums[C + c] = (float)q;
    }
}

template <int MODE>
static int launch_reduce(const mopoe_view_t* x, const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                         const uint8_t* mask, int mask_mode, const float* mean, const float* invstd, double* ws,
                         int nchunk, cudaStream_t st, const FinArgs& fin, const float* ggamma = nullptr,
                         const float* gbeta = nullptr) {
    MOPOE_REQUIRE(x->C % VEC == 0, "reduce: C=%d not a multiple of %d", x->C, VEC);
    MOPOE_REQUIRE(nchunk >= 1 && ws, "reduce: bad workspace");
    MOPOE_REQUIRE((long long)x->B * x->H * x->W < (1ll << 31), "reduce: too many rows");
    dim3 block(16, RED_ROWS), grid((x->C + 127) / 128, nchunk);
    MOPOE_DISPATCH_T(x->dtype, T, {
        DView<const T> xv = make_dview<const T>(x);
        DView<const T> dv = dy ? make_dview<const T>(dy) : xv;
        DView<const T> gv = gate ? make_dview<const T>(gate) : xv;
        MOPOE_REQUIRE(!(ggamma || gbeta) || gate, "bn_bwd_reduce: gate recompute is not compiled in; pass the gate view");
        const int gmode = gate ? 1 : 0;
        reduce_rows_kernel<T, MODE><<<grid, block, 0, st>>>(xv, dv, gv, gmode, gscale, mask, mask_mode, mean, invstd,
                                                           ggamma, gbeta, ws, nchunk, fin);
    });
    MOPOE_CHECK_LAUNCH("reduce_rows");
    return 0;
}

// staged reduction when eligible (and no fused-finalize counters were asked for), else the register-staged kernel
template <int MODE>
static int reduce_any(const mopoe_view_t* x, const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale, const uint8_t* mask,
                      int mask_mode, const float* mean, const float* invstd, double* ws, int nchunk, int* used, cudaStream_t st,
                      const FinArgs& fin, const float* ggamma = nullptr, const float* gbeta = nullptr) {
    *used = nchunk;
    if (ew_staged() && !fin.counter && x->C % VEC == 0) {
        const int r = mopoe_staged_reduce(MODE, x, dy, gate, gscale, mask, mask_mode, mean, invstd, ws, nchunk, used, ggamma, gbeta, st);
        if (r >= 0) return r;
        *used = nchunk;
    }
    return launch_reduce<MODE>(x, dy, gate, gscale, mask, mask_mode, mean, invstd, ws, nchunk, st, fin);   // (reads x: exact)
}

extern "C" int mopoe_bn_stats(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, double* ws, int nchunk,
                              float eps, float momentum, float* mean, float* invstd, float* running_mean,
                              float* running_var, int* counters, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FinArgs fin = {};
    fin.counter = counters;
    fin.count = (double)x->B * x->H * x->W;
    fin.eps = eps; fin.momentum = momentum;
    fin.mean = mean; fin.invstd = invstd; fin.rmean = running_mean; fin.rvar = running_var;
    if (reduce_any<RED_STATS>(x, nullptr, nullptr, 1.f, mask, mask_mode, nullptr, nullptr, ws, nchunk, &nchunk, st, fin)) return 1;
    if (!counters) {
        bn_finalize_kernel<<<(x->C + FIN_CH - 1) / FIN_CH, 256, 0, st>>>(ws, nchunk, x->C, fin.count, eps, momentum, mean, invstd,
                                                          running_mean, running_var);
        MOPOE_CHECK_LAUNCH("bn_finalize");
    }
    return 0;
}

int mopoe_bn_finalize_launch(const double* ws, int nchunk, int C, double count, float eps, float momentum, float* mean,
                             float* invstd, float* rmean, float* rvar, void* stream) {
    bn_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, 256, 0, (cudaStream_t)stream>>>(ws, nchunk, C, count, eps, momentum, mean,
                                                                                   invstd, rmean, rvar);
    MOPOE_CHECK_LAUNCH("bn_finalize");
    return 0;
}

// finalize of BatchNorm-backward partial sums produced elsewhere (the GEMM epilogue, gemm_api.cu): ws = [nchunk][2][C]
int mopoe_sums_finalize_launch(const double* ws, int nchunk, int C, float* dbeta, float* dgamma, int accumulate, float* sums,
                               void* stream) {
    sums_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, 256, 0, (cudaStream_t)stream>>>(ws, nchunk, C, dbeta, dgamma, accumulate, sums);
    MOPOE_CHECK_LAUNCH("bn_bwd_finalize");
    return 0;
}

extern "C" int mopoe_colsum(const mopoe_view_t* v, float* out, int accumulate, double* ws, int nchunk, int* counters,
                            void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FinArgs fin = {};
    fin.counter = counters;
    fin.out0 = out; fin.accumulate = accumulate;
    if (reduce_any<RED_COLSUM>(v, nullptr, nullptr, 1.f, nullptr, MOPOE_MASK_NONE, nullptr, nullptr, ws, nchunk, &nchunk, st, fin))
        return 1;
    if (!counters) {
        sums_finalize_kernel<<<(v->C + FIN_CH - 1) / FIN_CH, 256, 0, st>>>(ws, nchunk, v->C, out, nullptr, accumulate, nullptr);
        MOPOE_CHECK_LAUNCH("colsum_finalize");
    }
    return 0;
}

extern "C" int mopoe_bn_bwd_reduce(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                                   const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean,
                                   const float* invstd, double* ws, int nchunk, float* dgamma, float* dbeta,
                                   int accumulate, float* sums, const float* gate_gamma, const float* gate_beta,
                                   int* counters, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (check_same(x, dy, "bn_bwd_reduce(dy)")) return 1;
    if (gate && check_same(x, gate, "bn_bwd_reduce(gate)")) return 1;
    FinArgs fin = {};
    fin.counter = counters;
    fin.out0 = dbeta; fin.out1 = dgamma; fin.sums = sums; fin.accumulate = accumulate;
    if (reduce_any<RED_BNBWD>(x, dy, gate, gscale, mask, mask_mode, mean, invstd, ws, nchunk, &nchunk, st, fin, gate_gamma, gate_beta))
        return 1;
    if (!counters) {
        sums_finalize_kernel<<<(x->C + FIN_CH - 1) / FIN_CH, 256, 0, st>>>(ws, nchunk, x->C, dbeta, dgamma, accumulate, sums);
        MOPOE_CHECK_LAUNCH("bn_bwd_finalize");
    }
    return 0;
}

// =====================================================================================================================
// Row-wise streaming passes (default; MOPOE_EW_ROWS=0 selects the per-octet kernels below).
//
// ncu of the per-octet kernels (profiles/r2_ncu_elementwise.txt): ~125 instructions and ONE 16-byte load in flight per
// thread and item, 1.2 eligible warps per scheduler, 66 % of the cycles stalled on that load -> 4.3 TB/s where a plain
// copy moves 6.2.  Here a thread owns one octet COLUMN (8 channels of one pixel column of the storage row) of U
// consecutive storage rows of the output: the (batch, row) decode is one division per row, the column / channel decode
// one per thread, the per-channel coefficients are loaded once per thread, and all U x (operands) 16-byte loads are
// issued before the first use.
struct RowGeo {
    unsigned NRG, RS, RO;          // row groups, storage rows (B * Hs), octets per storage row (Ws * C/8)
    int Hs, H, ph, o_lo, o_hi, W;  // interior octets of a row: [o_lo, o_hi)
    FastDiv fRO, fHs, fCV;
};
static int make_rowgeo(const mopoe_view_t* out, int U, RowGeo& g, unsigned& blocks) {
    const int CV = out->C / VEC;
    g.Hs = out->H + 2 * out->ph;
    g.H = out->H; g.ph = out->ph; g.W = out->W;
    g.RS = (unsigned)out->B * g.Hs;
    g.RO = (unsigned)(out->W + 2 * out->pw) * CV;
    g.o_lo = out->pw * CV;
    g.o_hi = (out->pw + out->W) * CV;
    g.NRG = (g.RS + U - 1) / U;
    g.fRO = FastDiv(g.RO); g.fHs = FastDiv((unsigned)g.Hs); g.fCV = FastDiv((unsigned)CV);
    const long long threads = (long long)g.NRG * g.RO;
    MOPOE_REQUIRE(threads > 0 && threads < (1ll << 31), "elementwise: %lld work items do not fit 32-bit indexing", threads);
    blocks = (unsigned)ceil_div64(threads, EW_THREADS);
    return 0;
}
static bool rows_ok(const mopoe_view_t* v) { return v->sW == v->C && v->C % VEC == 0; }
static int ew_rows() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_EW_ROWS");
        v = e ? atoi(e) : 1;
    }
    return v;
}
// per-thread decode shared by the row kernels
struct RowThread {
    unsigned rg, o;
    int c, coff;
    bool col_in;
};
__device__ __forceinline__ bool row_thread(const RowGeo& g, RowThread& t) {
    const unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    g.fRO.divmod(idx, t.rg, t.o);
    if (t.rg >= g.NRG) return false;
    unsigned wq, cv;
    g.fCV.divmod(t.o, wq, cv);
    t.c = (int)cv * VEC;
    t.col_in = (int)t.o >= g.o_lo && (int)t.o < g.o_hi;
    t.coff = ((int)t.o - g.o_lo) * VEC;
    return true;
}
// storage row rs -> (b, h); returns "row exists"
__device__ __forceinline__ bool row_decode(const RowGeo& g, unsigned rs, int& b, int& h) {
    unsigned bb, hs;
    g.fHs.divmod(rs, bb, hs);
    b = (int)bb;
    h = (int)hs - g.ph;
    return rs < g.RS;
}
template <typename T>
__device__ __forceinline__ void store_raw(T* p, const float (&o)[8]) { st8v<T>(p, o); }

template <typename T, int U>
__global__ void __launch_bounds__(EW_THREADS) bn_apply_rows_kernel(const RowGeo g, DView<const T> x, const uint8_t* mask,
                                                                   int mask_mode, const float* mean, const float* invstd,
                                                                   const float* gamma, const float* beta, int relu,
                                                                   DView<T> out) {
    RowThread t;
    if (!row_thread(g, t)) return;
    Raw8<T> xr[U];
    uint2 mr[U];
    bool valid[U], in[U];
    int ooff[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int b, h;
        valid[u] = row_decode(g, t.rg * U + u, b, h);
        in[u] = valid[u] && t.col_in && h >= 0 && h < g.H;
        ooff[u] = b * (int)out.sB + h * (int)out.sH + t.coff;
        if (in[u]) {
            xr[u].load(x.p + (b * (int)x.sB + h * (int)x.sH + t.coff));
            if (mask_mode != MOPOE_MASK_NONE)
                mr[u] = *reinterpret_cast<const uint2*>(
                    mask + (mask_mode == MOPOE_MASK_BC ? b * x.C + t.c : (b * g.H + h) * (g.W * x.C) + t.coff));
        }
    }
    float sc[VEC], sh[VEC];
    {
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + t.c, mu); ld8f(invstd + t.c, is); ld8f(gamma + t.c, ga); ld8f(beta + t.c, be);
        bn_affine(mu, is, ga, be, sc, sh);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (in[u]) {
            float xv[VEC];
            xr[u].unpack(xv);
            if (mask_mode != MOPOE_MASK_NONE) {
                float mk[VEC];
                mask8(mr[u], mk);
#pragma unroll
                for (int i = 0; i < VEC; ++i) xv[i] *= mk[i];
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) xv[i] *= 1.f;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float y = bn_eval(xv[i], sc[i], sh[i]);
                o[i] = (relu && y < 0.f) ? 0.f : y;
            }
        }
        st8v<T>(out.p + ooff[u], o);
    }
}

template <typename T, int U>
__global__ void __launch_bounds__(EW_THREADS) combine_rows_kernel(const RowGeo g, DView<const T> r, const float* mean,
                                                                  const float* invstd, const float* gamma, const float* beta,
                                                                  DView<const T> cc, const uint8_t* mask, int mask_mode, float a,
                                                                  float bcoef, DView<T> out) {
    RowThread t;
    if (!row_thread(g, t)) return;
    Raw8<T> rr[U], cr[U];
    uint2 mr[U];
    bool valid[U], in[U];
    int ooff[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int b, h;
        valid[u] = row_decode(g, t.rg * U + u, b, h);
        in[u] = valid[u] && t.col_in && h >= 0 && h < g.H;
        ooff[u] = b * (int)out.sB + h * (int)out.sH + t.coff;
        if (in[u]) {
            rr[u].load(r.p + (b * (int)r.sB + h * (int)r.sH + t.coff));
            cr[u].load(cc.p + (b * (int)cc.sB + h * (int)cc.sH + t.coff));
            if (mask_mode != MOPOE_MASK_NONE)
                mr[u] = *reinterpret_cast<const uint2*>(
                    mask + (mask_mode == MOPOE_MASK_BC ? b * r.C + t.c : (b * g.H + h) * (g.W * r.C) + t.coff));
        }
    }
    float sc[VEC], sh[VEC];                    // a * BN(r) = r * sc + sh
    {
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + t.c, mu); ld8f(invstd + t.c, is); ld8f(gamma + t.c, ga); ld8f(beta + t.c, be);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sc[i] = a * is[i] * ga[i]; sh[i] = a * be[i] - mu[i] * sc[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (in[u]) {
            float rv[VEC], cv[VEC], mk[VEC];
            rr[u].unpack(rv);
            cr[u].unpack(cv);
            if (mask_mode != MOPOE_MASK_NONE) {
                mask8(mr[u], mk);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) mk[i] = 1.f;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = fmaf(rv[i], sc[i], sh[i]) + bcoef * (cv[i] * mk[i]);
        }
        st8v<T>(out.p + ooff[u], o);
    }
}

// BN(+ReLU gate, +dropout, +addend, + second output) backward apply, row-wise.  Same arithmetic, operation for operation,
// as bn_bwd_apply_oneshot_kernel (coefficients k1 / ca / cb, fmaf nesting) -> bit-identical results.
template <typename T, int U, bool GATE, bool ADD, bool OUT2>
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_rows_kernel(
    const RowGeo g, DView<const T> dy, DView<const T> gate, float gscale, DView<const T> x, const uint8_t* mask, int mask_mode,
    const float* mean, const float* invstd, const float* gamma, const float* sums, float inv_cnt, DView<const T> addend,
    DView<T> out, DView<T> out2, const uint8_t* mask2, int mask2_mode, float scale2) {
    RowThread t;
    if (!row_thread(g, t)) return;
    const int C = x.C;
    Raw8<T> gR[U], tR[U], xR[U], aR[U];
    uint2 mR[U], m2R[U];
    bool valid[U], in[U];
    int ooff[U], o2off[U];
    const bool masked = mask_mode != MOPOE_MASK_NONE, masked2 = OUT2 && mask2_mode != MOPOE_MASK_NONE;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int b, h;
        valid[u] = row_decode(g, t.rg * U + u, b, h);
        in[u] = valid[u] && t.col_in && h >= 0 && h < g.H;
        ooff[u] = b * (int)out.sB + h * (int)out.sH + t.coff;
        if (OUT2) o2off[u] = b * (int)out2.sB + h * (int)out2.sH + t.coff;
        if (in[u]) {
            gR[u].load(dy.p + (b * (int)dy.sB + h * (int)dy.sH + t.coff));
            if (GATE) tR[u].load(gate.p + (b * (int)gate.sB + h * (int)gate.sH + t.coff));
            xR[u].load(x.p + (b * (int)x.sB + h * (int)x.sH + t.coff));
            if (ADD) aR[u].load(addend.p + (b * (int)addend.sB + h * (int)addend.sH + t.coff));
            const int bc = b * C + t.c, el = (b * g.H + h) * (g.W * C) + t.coff;
            if (masked) mR[u] = *reinterpret_cast<const uint2*>(mask + (mask_mode == MOPOE_MASK_BC ? bc : el));
            if (masked2) m2R[u] = *reinterpret_cast<const uint2*>(mask2 + (mask2_mode == MOPOE_MASK_BC ? bc : el));
        }
    }
    float k1[VEC], ca[VEC], cb[VEC];
    {
        float mu[VEC], is[VEC], ga[VEC], sg[VEC], sgx[VEC];
        ld8f(mean + t.c, mu); ld8f(invstd + t.c, is); ld8f(gamma + t.c, ga); ld8f(sums + t.c, sg); ld8f(sums + C + t.c, sgx);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            k1[i] = ga[i] * is[i];
            const float mg = sg[i] * inv_cnt, mgx = sgx[i] * inv_cnt;
            ca[i] = -k1[i] * is[i] * mgx;
            cb[i] = k1[i] * (mu[i] * is[i] * mgx - mg);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (!in[u]) {
            st8v<T>(out.p + ooff[u], o);
            if (OUT2) st8v<T>(out2.p + o2off[u], o);
            continue;
        }
        float gv[VEC], v[VEC];
        gR[u].unpack(gv);
        xR[u].unpack(v);
        if (OUT2) {
            float o2[VEC], mk2[VEC];
            if (masked2) {
                mask8(m2R[u], mk2);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) mk2[i] = 1.f;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) o2[i] = scale2 * gv[i] * mk2[i];
            st8v<T>(out2.p + o2off[u], o2);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) gv[i] *= gscale;
        if (GATE) {
            float gt[VEC];
            tR[u].unpack(gt);
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                if (!(gt[i] > 0.f)) gv[i] = 0.f;
        }
        if (masked) {
            float mk[VEC];
            mask8(mR[u], mk);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], gv[i], fmaf(ca[i], v[i] * mk[i], cb[i])) * mk[i];
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], gv[i], fmaf(ca[i], v[i], cb[i]));
        }
        if (ADD) {
            float ad[VEC];
            aR[u].unpack(ad);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] += ad[i];
        }
        st8v<T>(out.p + ooff[u], o);
    }
}
#ifndef BWD_ROWS_U
#define BWD_ROWS_U 2
#endif
#ifndef FWD_ROWS_U
#define FWD_ROWS_U 4
#endif
template <typename T>
static int launch_bwd_rows(const mopoe_view_t* outv, cudaStream_t st, DView<const T> dy, DView<const T> gate, bool has_gate,
                           float gscale, DView<const T> x, const uint8_t* mask, int mask_mode, const float* mean,
                           const float* invstd, const float* gamma, const float* sums, float inv_cnt, DView<const T> addend,
                           bool has_add, DView<T> out, DView<T> out2, bool has_out2, const uint8_t* mask2, int mask2_mode,
                           float scale2) {
    RowGeo g;
    unsigned blocks;
    constexpr int U = BWD_ROWS_U;
    if (make_rowgeo(outv, U, g, blocks)) return 1;
#define MOPOE_RW(G, A, O)                                                                                                      \
    bn_bwd_apply_rows_kernel<T, U, G, A, O><<<blocks, EW_THREADS, 0, st>>>(g, dy, gate, gscale, x, mask, mask_mode, mean, invstd, \
                                                                          gamma, sums, inv_cnt, addend, out, out2, mask2,        \
                                                                          mask2_mode, scale2)
    if (has_out2) {
        if (has_gate) { if (has_add) MOPOE_RW(true, true, true); else MOPOE_RW(true, false, true); }
        else { if (has_add) MOPOE_RW(false, true, true); else MOPOE_RW(false, false, true); }
    } else {
        if (has_gate) { if (has_add) MOPOE_RW(true, true, false); else MOPOE_RW(true, false, false); }
        else { if (has_add) MOPOE_RW(false, true, false); else MOPOE_RW(false, false, false); }
    }
#undef MOPOE_RW
    return 0;
}

// ---- forward apply kernels ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) bn_apply_kernel(DView<const T> x, const uint8_t* mask, int mask_mode,
                                                              const float* mean, const float* invstd,
                                                              const float* gamma, const float* beta, int relu,
                                                              DView<T> out, unsigned total, unsigned stride) {
    unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= total) return;
    const unsigned CV = (unsigned)out.C / VEC;
    const int c = (int)(idx % CV) * VEC;
    float sc[VEC], sh[VEC];                    // y = v * sc + sh
    {
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(beta + c, be);
        bn_affine(mu, is, ga, be, sc, sh);
    }
#pragma unroll 2
    for (; idx < total; idx += stride) {
        const Pos p = decode_pixel(out, out.fCV8.div(idx), c);
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (p.interior) {
            float xv[VEC], mk[VEC];
            ld8v<T>(x.p + vaddr(x, p.b, p.h, p.w, p.c), xv);
            ldmask8(mask, mask_mode, p.b * x.C + p.c, ((p.b * x.H + p.h) * x.W + p.w) * x.C + p.c, mk);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float y = bn_eval(xv[i] * mk[i], sc[i], sh[i]);
                o[i] = (relu && y < 0.f) ? 0.f : y;
            }
        }
        st8v<T>(out.p + vaddr(out, p.b, p.h, p.w, p.c), o);
    }
}

extern "C" int mopoe_bn_apply(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean,
                              const float* invstd, const float* gamma, const float* beta, int relu,
                              const mopoe_view_t* out, void* stream) {
    if (check_same(x, out, "bn_apply")) return 1;
    MOPOE_REQUIRE(x->C % VEC == 0, "bn_apply: C=%d", x->C);
    if (ew_staged()) {
        const int r = mopoe_staged_bn_apply(x, mask, mask_mode, mean, invstd, gamma, beta, relu, out, (cudaStream_t)stream);
        if (r >= 0) return r;
    }
    if (ew_rows() && rows_ok(x) && rows_ok(out)) {
        RowGeo g;
        unsigned blocks;
        if (make_rowgeo(out, FWD_ROWS_U, g, blocks)) return 1;
        MOPOE_DISPATCH_T(x->dtype, T, {
            bn_apply_rows_kernel<T, FWD_ROWS_U><<<blocks, EW_THREADS, 0, (cudaStream_t)stream>>>(
                g, make_dview<const T>(x), mask, mask_mode, mean, invstd, gamma, beta, relu, make_dview<T>(out));
        });
        MOPOE_CHECK_LAUNCH("bn_apply_rows");
        return 0;
    }
    MOPOE_DISPATCH_T(x->dtype, T, {
        DView<T> ov = make_dview<T>(out);
        unsigned grid, stride;
        const long long total = storage_threads(ov);
        if (apply_grid(total, x->C, grid, stride)) return 1;
        bn_apply_kernel<T><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(make_dview<const T>(x), mask, mask_mode, mean,
                                                                         invstd, gamma, beta, relu, ov, (unsigned)total,
                                                                         stride);
    });
    MOPOE_CHECK_LAUNCH("bn_apply");
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) combine_kernel(DView<const T> r, const float* mean, const float* invstd,
                                                             const float* gamma, const float* beta, DView<const T> cc,
                                                             const uint8_t* mask, int mask_mode, float a, float bcoef,
                                                             DView<T> out, unsigned total, unsigned stride) {
    unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= total) return;
    const unsigned CV = (unsigned)out.C / VEC;
    const int c = (int)(idx % CV) * VEC;
    float sc[VEC], sh[VEC];                    // a * BN(r) = r * sc + sh
    {
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(beta + c, be);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sc[i] = a * is[i] * ga[i]; sh[i] = a * be[i] - mu[i] * sc[i]; }
    }
#pragma unroll 2
    for (; idx < total; idx += stride) {
        const Pos p = decode_pixel(out, out.fCV8.div(idx), c);
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (p.interior) {
            float rv[VEC], cv[VEC], mk[VEC];
            ld8v<T>(r.p + vaddr(r, p.b, p.h, p.w, p.c), rv);
            ld8v<T>(cc.p + vaddr(cc, p.b, p.h, p.w, p.c), cv);
            ldmask8(mask, mask_mode, p.b * r.C + p.c, ((p.b * r.H + p.h) * r.W + p.w) * r.C + p.c, mk);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = fmaf(rv[i], sc[i], sh[i]) + bcoef * (cv[i] * mk[i]);
        }
        st8v<T>(out.p + vaddr(out, p.b, p.h, p.w, p.c), o);
    }
}

extern "C" int mopoe_combine(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, const mopoe_view_t* c, const uint8_t* mask, int mask_mode, float a,
                             float b, const mopoe_view_t* out, void* stream) {
    if (check_same(r, out, "combine(out)") || check_same(r, c, "combine(c)")) return 1;
    MOPOE_REQUIRE(r->C % VEC == 0, "combine: C=%d", r->C);
    if (ew_staged()) {
        const int rc = mopoe_staged_combine(r, mean, invstd, gamma, beta, c, mask, mask_mode, a, b, out, nullptr, 0, nullptr,
                                            (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    if (ew_rows() && rows_ok(r) && rows_ok(c) && rows_ok(out)) {
        RowGeo g;
        unsigned blocks;
        if (make_rowgeo(out, FWD_ROWS_U, g, blocks)) return 1;
        MOPOE_DISPATCH_T(r->dtype, T, {
            combine_rows_kernel<T, FWD_ROWS_U><<<blocks, EW_THREADS, 0, (cudaStream_t)stream>>>(
                g, make_dview<const T>(r), mean, invstd, gamma, beta, make_dview<const T>(c), mask, mask_mode, a, b,
                make_dview<T>(out));
        });
        MOPOE_CHECK_LAUNCH("combine_rows");
        return 0;
    }
    MOPOE_DISPATCH_T(r->dtype, T, {
        DView<T> ov = make_dview<T>(out);
        unsigned grid, stride;
        const long long total = storage_threads(ov);
        if (apply_grid(total, r->C, grid, stride)) return 1;
        combine_kernel<T><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(make_dview<const T>(r), mean, invstd, gamma, beta,
                                                                        make_dview<const T>(c), mask, mask_mode, a, b, ov,
                                                                        (unsigned)total, stride);
    });
    MOPOE_CHECK_LAUNCH("combine");
    return 0;
}

// combine + the training-mode BatchNorm statistics of its output (the NEXT block's bn1: ResidualBlocks.py:84-86 applied to
// the previous block's `out`), in the same pass over the data when the staged kernel applies; else combine, then mopoe_bn_stats.
extern "C" int mopoe_combine_bn(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, const mopoe_view_t* c, const uint8_t* mask, int mask_mode, float a, float b,
                                const mopoe_view_t* out, double* ws, int nchunk, float eps, float momentum, float* out_mean,
                                float* out_invstd, float* running_mean, float* running_var, void* stream) {
    if (check_same(r, out, "combine_bn(out)") || check_same(r, c, "combine_bn(c)")) return 1;
    MOPOE_REQUIRE(r->C % VEC == 0 && ws && nchunk >= 1, "combine_bn: C=%d / workspace", r->C);
    cudaStream_t st = (cudaStream_t)stream;
    if (ew_staged()) {
        int used = 0;
        const int rc = mopoe_staged_combine(r, mean, invstd, gamma, beta, c, mask, mask_mode, a, b, out, ws, nchunk, &used, st);
        if (rc > 0) return 1;
        if (rc == 0) {
            const double count = (double)r->B * r->H * r->W;
            bn_finalize_kernel<<<(r->C + FIN_CH - 1) / FIN_CH, 256, 0, st>>>(ws, used, r->C, count, eps, momentum, out_mean, out_invstd,
                                                                           running_mean, running_var);
            MOPOE_CHECK_LAUNCH("combine_bn_finalize");
            return 0;
        }
    }
    if (mopoe_combine(r, mean, invstd, gamma, beta, c, mask, mask_mode, a, b, out, stream)) return 1;
    return mopoe_bn_stats(out, nullptr, MOPOE_MASK_NONE, ws, nchunk, eps, momentum, out_mean, out_invstd, running_mean, running_var,
                          nullptr, stream);
}

// ---- backward apply kernels ---------------------------------------------------------------------------
// out  = gamma*invstd*(g - sums_g/cnt - xhat*sums_gx/cnt) * 2mask + addend,  g = gscale * dy * [gate > 0]
// out2 = scale2 * dy * 2mask2   (optional second output sharing the read of dy: the dropout2 branch of a block)
// These passes are LATENCY-bound, not issue-bound (ncu: 39 % issue slots, long-scoreboard stalls, 115 registers -> 2
// blocks/SM): what sets their speed is the number of bytes in flight per SM.  So: operands stay PACKED between load and
// use (4 registers per bf16 octet), the loads of U pixels are issued before the first use, the per-channel terms are
// folded into 3 coefficients (dv = k1*g + ca*v + cb) and the register budget is capped for 3 blocks/SM.
#ifndef BWD_U
#define BWD_U 2
#endif
#ifndef BWD_MINB
#define BWD_MINB 2
#endif
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, BWD_MINB) bn_bwd_apply_kernel(DView<const T> dy, DView<const T> gate, int has_gate,
                                                                     float gscale, DView<const T> x, const uint8_t* mask,
                                                                     int mask_mode, const float* mean, const float* invstd,
                                                                     const float* gamma, const float* sums, float inv_cnt,
                                                                     DView<const T> addend, int has_add, DView<T> out,
                                                                     DView<T> out2, int has_out2, const uint8_t* mask2,
                                                                     int mask2_mode, float scale2, const float* gbeta,
                                                                     unsigned total, unsigned stride) {
    unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int C = x.C;
    const unsigned CV = (unsigned)C / VEC;
    const int c = (int)(idx % CV) * VEC;
    // dv = k1 * (g - m_g - xhat * m_gx),  xhat = v*is - mu*is   ==   k1*g + ca*v + cb
    float k1[VEC], ca[VEC], cb[VEC];
    {
        float mu[VEC], is[VEC], ga[VEC], sg[VEC], sgx[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(sums + c, sg); ld8f(sums + C + c, sgx);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            k1[i] = ga[i] * is[i];
            const float mg = sg[i] * inv_cnt, mgx = sgx[i] * inv_cnt;
            ca[i] = -k1[i] * is[i] * mgx;
            cb[i] = k1[i] * (mu[i] * is[i] * mgx - mg);
        }
    }
    constexpr int U = BWD_U;
    for (; idx < total; idx += U * stride) {
        Pos p[U];
        bool valid[U], in[U];
        Raw8<T> gR[U], tR[U], xR[U], aR[U];
        uint2 mR[U], m2R[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned id = idx + u * stride;
            valid[u] = id < total;
            in[u] = false;
            if (valid[u]) {
                p[u] = decode_pixel(out, out.fCV8.div(id), c);
                in[u] = p[u].interior;
            }
            if (in[u]) {
                const Pos& q = p[u];
                gR[u].load(dy.p + vaddr(dy, q.b, q.h, q.w, q.c));
                if (has_gate == 1) tR[u].load(gate.p + vaddr(gate, q.b, q.h, q.w, q.c));
                xR[u].load(x.p + vaddr(x, q.b, q.h, q.w, q.c));
                if (has_add) aR[u].load(addend.p + vaddr(addend, q.b, q.h, q.w, q.c));
                const int bc = q.b * C + q.c, el = ((q.b * x.H + q.h) * x.W + q.w) * C + q.c;
                if (mask_mode != MOPOE_MASK_NONE)
                    mR[u] = *reinterpret_cast<const uint2*>(mask + (mask_mode == MOPOE_MASK_BC ? bc : el));
                if (has_out2 && mask2_mode != MOPOE_MASK_NONE)
                    m2R[u] = *reinterpret_cast<const uint2*>(mask2 + (mask2_mode == MOPOE_MASK_BC ? bc : el));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            const Pos& q = p[u];
            float o[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = 0.f;
            if (in[u]) {
                float g[VEC], v[VEC];
                gR[u].unpack(g);
                xR[u].unpack(v);
                if (has_out2) {
                    float o2[VEC], mk2[VEC];
                    if (mask2_mode != MOPOE_MASK_NONE) {
                        mask8(m2R[u], mk2);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) mk2[i] = 1.f;
                    }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o2[i] = scale2 * g[i] * mk2[i];
                    st8v<T>(out2.p + vaddr(out2, q.b, q.h, q.w, q.c), o2);
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) g[i] *= gscale;
                if (has_gate == 1) {
                    float gt[VEC];
                    tR[u].unpack(gt);
#pragma unroll
                    for (int i = 0; i < VEC; ++i)
                        if (!(gt[i] > 0.f)) g[i] = 0.f;
                }
                if (mask_mode != MOPOE_MASK_NONE) {
                    float mk[VEC];
                    mask8(mR[u], mk);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], g[i], fmaf(ca[i], v[i] * mk[i], cb[i])) * mk[i];
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], g[i], fmaf(ca[i], v[i], cb[i]));
                }
                if (has_add) {
                    float ad[VEC];
                    aR[u].unpack(ad);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] += ad[i];
                }
            } else if (has_out2) {
                st8v<T>(out2.p + vaddr(out2, q.b, q.h, q.w, q.c), o);      // zero border of the second output
            }
            st8v<T>(out.p + vaddr(out, q.b, q.h, q.w, q.c), o);
        }
    }
}

// Variant B of the same pass (MOPOE_EW_ONESHOT=0 falls back to the kernel above): no grid-stride loop — one channel
// octet per thread, the per-channel coefficients live in shared memory (computed once per block), the optional operands
// are template flags and the arithmetic runs pair by pair straight from the packed registers.  A thread then carries
// ~40 registers instead of ~110: occupancy, not per-thread unrolling, is what feeds HBM here (measured).
template <typename T>
__device__ __forceinline__ float2 pair_of(const Raw8<T>& r, int i);
template <>
__device__ __forceinline__ float2 pair_of<bf16>(const Raw8<bf16>& r, int i) {
    const uint32_t w = i == 0 ? r.v.x : (i == 1 ? r.v.y : (i == 2 ? r.v.z : r.v.w));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 pair_of<float>(const Raw8<float>& r, int i) {
    return i == 0 ? make_float2(r.a.x, r.a.y) : (i == 1 ? make_float2(r.a.z, r.a.w) : (i == 2 ? make_float2(r.b.x, r.b.y) : make_float2(r.b.z, r.b.w)));
}
__device__ __forceinline__ float2 mask_pair(const uint2& t, int i) {
    const uint32_t w = (i < 2 ? t.x : t.y) >> (16 * (i & 1));
    return make_float2((w & 0xffu) ? 2.f : 0.f, (w & 0xff00u) ? 2.f : 0.f);
}
template <typename T>
__device__ __forceinline__ void store_pairs(T* p, const float2 (&o)[4]);
template <>
__device__ __forceinline__ void store_pairs<bf16>(bf16* p, const float2 (&o)[4]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[i].x, o[i].y);
    *reinterpret_cast<uint4*>(p) = t;
}
template <>
__device__ __forceinline__ void store_pairs<float>(float* p, const float2 (&o)[4]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
    reinterpret_cast<float4*>(p)[1] = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
}

template <typename T, bool GATE, bool ADD, bool OUT2>
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_oneshot_kernel(
    DView<const T> dy, DView<const T> gate, float gscale, DView<const T> x, const uint8_t* mask, int mask_mode,
    const float* mean, const float* invstd, const float* gamma, const float* sums, float inv_cnt, DView<const T> addend,
    DView<T> out, DView<T> out2, const uint8_t* mask2, int mask2_mode, float scale2, unsigned total) {
    extern __shared__ float coef[];          // k1[C] | ca[C] | cb[C]
    const int C = x.C;
    const unsigned id = blockIdx.x * EW_THREADS + threadIdx.x;
    const bool valid = id < total;
    unsigned pix, cv;
    out.fCV8.divmod(valid ? id : 0u, pix, cv);
    const Pos q = decode_pixel(out, pix, (int)cv * VEC);
    const bool in = valid && q.interior;
    // the operand loads go out FIRST: they are in flight while the block computes its coefficient table
    Raw8<T> gR, tR, xR, aR;
    uint2 mR = make_uint2(0, 0), m2R = make_uint2(0, 0);
    const bool masked = mask_mode != MOPOE_MASK_NONE, masked2 = OUT2 && mask2_mode != MOPOE_MASK_NONE;
    if (in) {
        gR.load(dy.p + vaddr(dy, q.b, q.h, q.w, q.c));
        if (GATE) tR.load(gate.p + vaddr(gate, q.b, q.h, q.w, q.c));
        xR.load(x.p + vaddr(x, q.b, q.h, q.w, q.c));
        if (ADD) aR.load(addend.p + vaddr(addend, q.b, q.h, q.w, q.c));
        const int bc = q.b * C + q.c, el = ((q.b * x.H + q.h) * x.W + q.w) * C + q.c;
        if (masked) mR = *reinterpret_cast<const uint2*>(mask + (mask_mode == MOPOE_MASK_BC ? bc : el));
        if (masked2) m2R = *reinterpret_cast<const uint2*>(mask2 + (mask2_mode == MOPOE_MASK_BC ? bc : el));
    }
    for (int ch = threadIdx.x; ch < C; ch += EW_THREADS) {
        const float is = invstd[ch], k1 = gamma[ch] * is;
        const float mg = sums[ch] * inv_cnt, mgx = sums[C + ch] * inv_cnt;
        coef[ch] = k1;
        coef[C + ch] = -k1 * is * mgx;
        coef[2 * C + ch] = k1 * (mean[ch] * is * mgx - mg);
    }
    __syncthreads();
    if (!valid) return;
    T* const op = out.p + vaddr(out, q.b, q.h, q.w, q.c);
    float2 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_float2(0.f, 0.f);
    if (!in) {                                               // zero border of the output(s)
        store_pairs<T>(op, o);
        if (OUT2) store_pairs<T>(out2.p + vaddr(out2, q.b, q.h, q.w, q.c), o);
        return;
    }
    if (OUT2) {
        float2 o2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 g = pair_of<T>(gR, i);
            const float2 mk2 = masked2 ? mask_pair(m2R, i) : make_float2(1.f, 1.f);
            o2[i] = make_float2(scale2 * g.x * mk2.x, scale2 * g.y * mk2.y);
        }
        store_pairs<T>(out2.p + vaddr(out2, q.b, q.h, q.w, q.c), o2);
    }
    const float* ck = coef + q.c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 g = pair_of<T>(gR, i);
        const float2 v = pair_of<T>(xR, i);
        const float2 k1 = *reinterpret_cast<const float2*>(ck + 2 * i);
        const float2 ca = *reinterpret_cast<const float2*>(ck + C + 2 * i);
        const float2 cb = *reinterpret_cast<const float2*>(ck + 2 * C + 2 * i);
        g.x *= gscale;
        g.y *= gscale;
        if (GATE) {
            const float2 gt = pair_of<T>(tR, i);
            if (!(gt.x > 0.f)) g.x = 0.f;
            if (!(gt.y > 0.f)) g.y = 0.f;
        }
        float2 r;
        if (masked) {
            const float2 mk = mask_pair(mR, i);
            r.x = fmaf(k1.x, g.x, fmaf(ca.x, v.x * mk.x, cb.x)) * mk.x;
            r.y = fmaf(k1.y, g.y, fmaf(ca.y, v.y * mk.y, cb.y)) * mk.y;
        } else {
            r.x = fmaf(k1.x, g.x, fmaf(ca.x, v.x, cb.x));
            r.y = fmaf(k1.y, g.y, fmaf(ca.y, v.y, cb.y));
        }
        if (ADD) {
            const float2 ad = pair_of<T>(aR, i);
            r.x += ad.x;
            r.y += ad.y;
        }
        o[i] = r;
    }
    store_pairs<T>(op, o);
}
template <typename T>
static void launch_oneshot(long long total, cudaStream_t st, DView<const T> dy, DView<const T> gate, bool has_gate, float gscale,
                           DView<const T> x, const uint8_t* mask, int mask_mode, const float* mean, const float* invstd,
                           const float* gamma, const float* sums, float inv_cnt, DView<const T> addend, bool has_add,
                           DView<T> out, DView<T> out2, bool has_out2, const uint8_t* mask2, int mask2_mode, float scale2) {
    const size_t sm = (size_t)3 * x.C * sizeof(float);
    const unsigned grid = (unsigned)ceil_div64(total, EW_THREADS);
#define MOPOE_OS(G, A, O)                                                                                                        \
    bn_bwd_apply_oneshot_kernel<T, G, A, O><<<grid, EW_THREADS, sm, st>>>(dy, gate, gscale, x, mask, mask_mode, mean, invstd, gamma, \
                                                                          sums, inv_cnt, addend, out, out2, mask2, mask2_mode,     \
                                                                          scale2, (unsigned)total)
    if (has_out2) {
        if (has_gate) { if (has_add) MOPOE_OS(true, true, true); else MOPOE_OS(true, false, true); }
        else { if (has_add) MOPOE_OS(false, true, true); else MOPOE_OS(false, false, true); }
    } else {
        if (has_gate) { if (has_add) MOPOE_OS(true, true, false); else MOPOE_OS(true, false, false); }
        else { if (has_add) MOPOE_OS(false, true, false); else MOPOE_OS(false, false, false); }
    }
#undef MOPOE_OS
}
static int ew_oneshot() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_EW_ONESHOT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static int bn_bwd_apply_impl(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale, const mopoe_view_t* x,
                             const uint8_t* mask, int mask_mode, const float* mean, const float* invstd,
                             const float* gamma, const float* sums, const mopoe_view_t* addend, const mopoe_view_t* out,
                             const mopoe_view_t* out2, const uint8_t* mask2, int mask2_mode, float scale2, void* stream,
                             const float* gate_beta = nullptr) {
    if (check_same(x, dy, "bn_bwd_apply(dy)") || check_same(x, out, "bn_bwd_apply(out)")) return 1;
    MOPOE_REQUIRE(gate_beta == nullptr || gate, "bn_bwd_apply: pass the gate view as well (the register-staged kernels read it)");
    if (gate && check_same(x, gate, "bn_bwd_apply(gate)")) return 1;
    if (addend && check_same(x, addend, "bn_bwd_apply(addend)")) return 1;
    if (out2) {
        if (check_same(x, out2, "bn_bwd_apply(out2)")) return 1;
        MOPOE_REQUIRE(out2->ph == out->ph && out2->pw == out->pw, "bn_bwd_apply: out and out2 must share their border");
    }
    MOPOE_REQUIRE(x->C % VEC == 0, "bn_bwd_apply: C=%d", x->C);
    float inv_cnt = 1.f / ((float)x->B * (float)x->H * (float)x->W);
    if (ew_staged()) {
        const int r = mopoe_staged_bn_bwd_apply(dy, gate, gscale, x, mask, mask_mode, mean, invstd, gamma, sums, addend, out, out2,
                                                mask2, mask2_mode, scale2, gate_beta, (cudaStream_t)stream);
        if (r >= 0) return r;
    }
    MOPOE_DISPATCH_T(x->dtype, T, {
        DView<T> ov = make_dview<T>(out);
        DView<const T> xv = make_dview<const T>(x);
        unsigned grid, stride;
        const long long total = storage_threads(ov);
        if (apply_grid(total, x->C, grid, stride)) return 1;
        const int os = ew_oneshot();
        const bool rw = ew_rows() && rows_ok(x) && rows_ok(dy) && rows_ok(out) && (!gate || rows_ok(gate)) &&
                        (!addend || rows_ok(addend)) && (!out2 || rows_ok(out2));
        if (rw) {
            if (launch_bwd_rows<T>(out, (cudaStream_t)stream, make_dview<const T>(dy), gate ? make_dview<const T>(gate) : xv,
                                   gate != nullptr, gscale, xv, mask, mask_mode, mean, invstd, gamma, sums, inv_cnt,
                                   addend ? make_dview<const T>(addend) : xv, addend != nullptr, ov,
                                   out2 ? make_dview<T>(out2) : ov, out2 != nullptr, mask2, mask2_mode, scale2))
                return 1;
        } else if (os) {
            launch_oneshot<T>(total, (cudaStream_t)stream, make_dview<const T>(dy), gate ? make_dview<const T>(gate) : xv,
                              gate != nullptr, gscale, xv, mask, mask_mode, mean, invstd, gamma, sums, inv_cnt,
                              addend ? make_dview<const T>(addend) : xv, addend != nullptr, ov,
                              out2 ? make_dview<T>(out2) : ov, out2 != nullptr, mask2, mask2_mode, scale2);
        } else
        bn_bwd_apply_kernel<T><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(
            make_dview<const T>(dy), gate ? make_dview<const T>(gate) : xv, gate ? 1 : 0, gscale, xv, mask, mask_mode,
            mean, invstd, gamma, sums, inv_cnt, addend ? make_dview<const T>(addend) : xv, addend != nullptr, ov,
            out2 ? make_dview<T>(out2) : ov, out2 != nullptr, mask2, mask2_mode, scale2, nullptr, (unsigned)total, stride);
    });
    MOPOE_CHECK_LAUNCH("bn_bwd_apply");
    return 0;
}
extern "C" int mopoe_bn_bwd_apply(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                                  const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean,
                                  const float* invstd, const float* gamma, const float* sums,
                                  const mopoe_view_t* addend, const mopoe_view_t* out, const float* gate_beta,
                                  void* stream) {
    return bn_bwd_apply_impl(dy, gate, gscale, x, mask, mask_mode, mean, invstd, gamma, sums, addend, out, nullptr, nullptr,
                             MOPOE_MASK_NONE, 0.f, stream, gate_beta);
}
// backward of `y = a*BN(r) + b*(c*2mask2)` in ONE pass over dy:  dr = BN-backward(a*dy),  dc = b * dy * 2mask2
extern "C" int mopoe_combine_bwd_apply(const mopoe_view_t* dy, float a, const mopoe_view_t* r, const float* mean,
                                       const float* invstd, const float* gamma, const float* sums,
                                       const uint8_t* mask2, int mask2_mode, float b, const mopoe_view_t* dr,
                                       const mopoe_view_t* dc, void* stream) {
    return bn_bwd_apply_impl(dy, nullptr, a, r, nullptr, MOPOE_MASK_NONE, mean, invstd, gamma, sums, nullptr, dr, dc, mask2,
                             mask2_mode, b, stream);
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) scale_mask_kernel(DView<const T> dy, const uint8_t* mask, int mask_mode,
                                                                float scale, DView<T> out, unsigned total, unsigned stride) {
    unsigned idx = blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= total) return;
    const unsigned CV = (unsigned)out.C / VEC;
    const int c = (int)(idx % CV) * VEC;
#pragma unroll 2
    for (; idx < total; idx += stride) {
        const Pos p = decode_pixel(out, out.fCV8.div(idx), c);
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 0.f;
        if (p.interior) {
            float g[VEC], mk[VEC];
            ld8v<T>(dy.p + vaddr(dy, p.b, p.h, p.w, p.c), g);
            ldmask8(mask, mask_mode, p.b * dy.C + p.c, ((p.b * dy.H + p.h) * dy.W + p.w) * dy.C + p.c, mk);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = scale * g[i] * mk[i];
        }
        st8v<T>(out.p + vaddr(out, p.b, p.h, p.w, p.c), o);
    }
}

extern "C" int mopoe_scale_mask(const mopoe_view_t* dy, const uint8_t* mask, int mask_mode, float scale,
                                const mopoe_view_t* out, void* stream) {
    if (check_same(dy, out, "scale_mask")) return 1;
    MOPOE_REQUIRE(dy->C % VEC == 0, "scale_mask: C=%d", dy->C);
    MOPOE_DISPATCH_T(dy->dtype, T, {
        DView<T> ov = make_dview<T>(out);
        unsigned grid, stride;
        const long long total = storage_threads(ov);
        if (apply_grid(total, dy->C, grid, stride)) return 1;
        scale_mask_kernel<T><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(make_dview<const T>(dy), mask, mask_mode, scale,
                                                                           ov, (unsigned)total, stride);
    });
    MOPOE_CHECK_LAUNCH("scale_mask");
    return 0;
}

// ---- zero border of a bordered activation (producers that only write the interior: GEMM outputs) ---------------------
template <typename T>
__global__ void __launch_bounds__(256) zero_border_kernel(DView<T> o, long long total) {
    const int Ws = o.W + 2 * o.pw, Hs = o.H + 2 * o.ph;
    const int CV = o.C / VEC;
    // border pixels per image: the ph top / bottom rows (full width) + the pw left / right columns of the H middle rows
    const int per = 2 * o.ph * Ws + 2 * o.pw * o.H;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cv = (int)(i % CV);
        const long long q = i / CV;
        const int k = (int)(q % per);
        const long long b = q / per;
        int hs, ws;
        if (k < o.ph * Ws) { hs = k / Ws; ws = k - hs * Ws; }
        else if (k < 2 * o.ph * Ws) { const int k2 = k - o.ph * Ws; hs = o.ph + o.H + k2 / Ws; ws = k2 % Ws; }
        else { const int k2 = k - 2 * o.ph * Ws; hs = o.ph + k2 / (2 * o.pw); const int j = k2 % (2 * o.pw); ws = j < o.pw ? j : o.W + j; }
        const float z[VEC] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        st8v<T>(o.p + b * o.sB + (long long)(hs - o.ph) * o.sH + (long long)(ws - o.pw) * o.sW + cv * VEC, z);
        (void)Hs;
    }
}
extern "C" int mopoe_zero_border(const mopoe_view_t* v, void* stream) {
    MOPOE_REQUIRE(v->C % VEC == 0, "zero_border: C=%d", v->C);
    if (v->ph == 0 && v->pw == 0) return 0;
    const long long per = 2ll * v->ph * (v->W + 2 * v->pw) + 2ll * v->pw * v->H;
    const long long total = (long long)v->B * per * (v->C / VEC);
    long long blocks = ceil_div64(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    MOPOE_DISPATCH_T(v->dtype, T, {
        zero_border_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(make_dview<T>(v), total);
    });
    MOPOE_CHECK_LAUNCH("zero_border");
    return 0;
}

// ---- layout / dtype conversion ------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void __launch_bounds__(EW_THREADS) convert_kernel(DView<const TS> s, int src_nchw, DView<TD> d,
                                                            long long total) {
    long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int Ws = d.W + 2 * d.pw, Hs = d.H + 2 * d.ph;
    int c = (int)(idx % d.C);
    long long pos = idx / d.C;
    int ws = (int)(pos % Ws);
    pos /= Ws;
    int hs = (int)(pos % Hs);
    int b = (int)(pos / Hs);
    int h = hs - d.ph, w = ws - d.pw;
    float v = 0.f;
    if (h >= 0 && h < d.H && w >= 0 && w < d.W && c < s.C) {
        long long a = src_nchw ? (((long long)b * s.C + c) * s.H + h) * s.W + w : vaddr(s, b, h, w, c);
        if constexpr (sizeof(TS) == 4) v = s.p[a]; else v = __bfloat162float(s.p[a]);
    }
    long long o = vaddr(d, b, h, w, c);
    if constexpr (sizeof(TD) == 4) d.p[o] = v; else d.p[o] = __float2bfloat16_rn(v);
}

extern "C" int mopoe_convert(const mopoe_view_t* src, int src_nchw, const mopoe_view_t* dst, void* stream) {
    MOPOE_REQUIRE(src->B == dst->B && src->H == dst->H && src->W == dst->W && src->C <= dst->C,
                  "convert: shape mismatch");
    long long total = (long long)dst->B * (dst->H + 2 * dst->ph) * (dst->W + 2 * dst->pw) * dst->C;
    unsigned grid = (unsigned)ceil_div64(total, EW_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    MOPOE_DISPATCH_T(src->dtype, TS, {
        MOPOE_DISPATCH_T(dst->dtype, TD, {
            convert_kernel<TS, TD><<<grid, EW_THREADS, 0, st>>>(make_dview<const TS>(src), src_nchw,
                                                                 make_dview<TD>(dst), total);
        });
    });
    MOPOE_CHECK_LAUNCH("convert");
    return 0;
}

// ---- dropout masks: Philox-4x32-10, one 128-bit block -> 128 keep bytes per thread -------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__global__ void dropout_mask_kernel(uint8_t* mask, long long n, uint64_t seed, uint64_t offset,
                                    const unsigned long long* __restrict__ step_ptr) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long base = t * 128;
    if (base >= n) return;
    uint64_t cidx = (uint64_t)t + offset;
    // the training-step index lives in device memory so that a captured CUDA graph draws fresh masks on every replay
    const unsigned long long step = step_ptr ? step_ptr[0] : 0ull;
    uint32_t ctr[4] = {(uint32_t)cidx, (uint32_t)(cidx >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0x6d6f7065u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (base + 128 <= n && ((uintptr_t)(mask + base) & 15) == 0) {
#pragma unroll
        for (int wd = 0; wd < 4; ++wd) {
            uint32_t bits = ctr[wd];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t nib = (bits >> (q * 16 + j * 4)) & 0xFu;
                    o[j] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                }
                *reinterpret_cast<uint4*>(mask + base + wd * 32 + q * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    } else {
        for (int i = 0; i < 128 && base + i < n; ++i) mask[base + i] = (ctr[i >> 5] >> (i & 31)) & 1u;
    }
}
extern "C" int mopoe_dropout_mask(uint8_t* mask, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_ptr,
                                  void* stream) {
    if (n <= 0) return 0;
    long long threads = ceil_div64(n, 128);
    dropout_mask_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        mask, n, seed, offset, (const unsigned long long*)step_ptr);
    MOPOE_CHECK_LAUNCH("dropout_mask");
    return 0;
}

// ---- token indices -> one-hot rows (word-encoded text) ------------------------------------------------------------------
// nn.Embedding as a GEMM operand (word_encoding/mmvae_text_enc.py:27-28,69): out[row, v] = (v == idx[row]), rows of Vp
// (>= V, multiple of 8) elements in the activation dtype; idx_out receives the integer indices.
template <typename T>
__global__ void __launch_bounds__(256) onehot_kernel(const float* __restrict__ idx, long long rows, int V, int Vp,
                                                     T* __restrict__ out, int* __restrict__ idx_out) {
    const long long total = rows * (Vp / 8);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long row = i / (Vp / 8);
        const int v0 = (int)(i - row * (Vp / 8)) * 8;
        int id = (int)idx[row];
        id = id < 0 ? 0 : (id >= V ? V - 1 : id);
        if (v0 == 0 && idx_out) idx_out[row] = id;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v0 + k == id) ? 1.f : 0.f;
        st8v<T>(out + row * Vp + v0, o);
    }
}
extern "C" int mopoe_onehot(const float* idx, int64_t rows, int V, int Vp, void* out, int out_dtype, int32_t* idx_out,
                            void* stream) {
    MOPOE_REQUIRE(rows > 0 && V >= 1 && Vp >= V && Vp % 8 == 0, "onehot: rows=%lld V=%d Vp=%d", (long long)rows, V, Vp);
    long long blocks = ceil_div64(rows * (Vp / 8), 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (out_dtype == MOPOE_F32)
        onehot_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, rows, V, Vp, (float*)out, idx_out);
    else
        onehot_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, rows, V, Vp, (bf16*)out, idx_out);
    MOPOE_CHECK_LAUNCH("onehot");
    return 0;
}

// Character indices shipped as one byte per token (SURVEY N3 wire format: 1 KB instead of 291 KB per report) -> the fp32
// one-hot rows [rows, V] the text encoder consumes (dataio/MimicDataset.py:92-96 builds them on the host in the reference).
__global__ void __launch_bounds__(256) onehot_u8_kernel(const uint8_t* __restrict__ idx, long long rows, int V,
                                                        float* __restrict__ out) {
    const long long total = rows * V;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long row = i / V;
        out[i] = ((int)(i - row * V) == (int)idx[row]) ? 1.f : 0.f;
    }
}
extern "C" int mopoe_onehot_u8(const uint8_t* idx, int64_t rows, int V, float* out, void* stream) {
    MOPOE_REQUIRE(rows > 0 && V >= 1 && V <= 256, "onehot_u8: rows=%lld V=%d", (long long)rows, V);
    long long blocks = ceil_div64(rows * V, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    onehot_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, rows, V, out);
    MOPOE_CHECK_LAUNCH("onehot_u8");
    return 0;
}

// 8-bit images on the wire: out = float(u8) / 255 — exactly torchvision's ToTensor() (the reference's loader,
// dataio/MimicDataset.py) evaluated on the device, so the host ships 1 byte per pixel instead of 4
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t* __restrict__ src, long long n, float* __restrict__ dst) {
    const long long n16 = n >> 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            reinterpret_cast<float4*>(dst)[i * 4 + k] =
                make_float4((float)(w[k] & 0xffu) / 255.f, (float)((w[k] >> 8) & 0xffu) / 255.f,
                            (float)((w[k] >> 16) & 0xffu) / 255.f, (float)(w[k] >> 24) / 255.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n16 << 4; i < n; ++i) dst[i] = (float)src[i] / 255.f;
}
extern "C" int mopoe_u8_to_unit(const uint8_t* src, int64_t n, float* dst, void* stream) {
    if (n <= 0) return 0;
    MOPOE_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "u8_to_unit: unaligned buffers");
    long long blocks = ceil_div64(n >> 4, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    u8_to_unit_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, n, dst);
    MOPOE_CHECK_LAUNCH("u8_to_unit");
    return 0;
}

// ---- flat Adam -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n4,
                                                   long long n, float lr_c, float b1, float b2, float eps, float inv_sqrt_bc2,
                                                   float gscale, const float* __restrict__ coef) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (coef) {            // device-side bias corrections (graph-captured step)
        lr_c = coef[0];
        inv_sqrt_bc2 = coef[1];
    }
    if (i < n4) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gk = G[k] * gscale;
            M[k] = b1 * M[k] + (1.f - b1) * gk;
            V[k] = b2 * V[k] + (1.f - b2) * gk * gk;
            P[k] -= lr_c * M[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    } else if (i == n4) {
        for (long long k = n4 * 4; k < n; ++k) {
            float gk = g[k] * gscale;
            m[k] = b1 * m[k] + (1.f - b1) * gk;
            v[k] = b2 * v[k] + (1.f - b2) * gk * gk;
            p[k] -= lr_c * m[k] / (sqrtf(v[k]) * inv_sqrt_bc2 + eps);
        }
    }
}
extern "C" int mopoe_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                               float beta2, float eps, int step, float grad_scale, void* stream) {
    if (n <= 0) return 0;
    MOPOE_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam: unaligned buffers");
    double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long n4 = n / 4;
    adam_kernel<<<(unsigned)ceil_div64(n4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
        p, g, m, v, n4, n, (float)(lr / bc1), beta1, beta2, eps, (float)(1.0 / sqrt(bc2)), grad_scale, nullptr);
    MOPOE_CHECK_LAUNCH("adam");
    return 0;
}
// same update, coefficients {lr/(1-b1^t), 1/sqrt(1-b2^t)} read from device memory (written by mopoe_step_advance)
extern "C" int mopoe_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* coef, float beta1,
                                   float beta2, float eps, float grad_scale, void* stream) {
    if (n <= 0) return 0;
    MOPOE_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam: unaligned buffers");
    long long n4 = n / 4;
    adam_kernel<<<(unsigned)ceil_div64(n4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n4, n, 0.f, beta1, beta2, eps,
                                                                                    1.f, grad_scale, coef);
    MOPOE_CHECK_LAUNCH("adam");
    return 0;
}

// ---- weight re-layout: fp32 master weights W[A][B][KH][KW] -> packed GEMM operand (bf16 or fp32) ------------------
//   form 0 conv : dst[a][ky][kx][b']            (b' < bpad; zero for b' >= B)      conv-form  [A, KH*KW*bpad]
//   form 1 phase: dst[b][r][kxi][a] = W[a][b][KT[py][r]][KT[px][kxi]]               phase-form [B, (KH>1?2:1)*2*A]
//   form 2 full : dst[ky][kx][b][a]                                                  full-form  [KH*KW*B, A]
//   form 3 mat  : dst[a][b]        form 4 matT : dst[b][a]                           (KH = KW = 1)
template <typename TD>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ W, int A, int B, int KH, int KW, int form,
                                                          int py, int px, int bpad, TD* __restrict__ dst, long long total) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int KT[2][2] = {{3, 1}, {2, 0}};
    int a, b, ky, kx;
    bool zero = false;
    if (form == 0) {
        b = (int)(i % bpad);
        long long t = i / bpad;
        kx = (int)(t % KW); t /= KW;
        ky = (int)(t % KH);
        a = (int)(t / KH);
        zero = b >= B;
    } else if (form == 1) {
        a = (int)(i % A);
        long long t = i / A;
        int kxi = (int)(t % 2); t /= 2;
        int r = 0;
        if (KH > 1) { r = (int)(t % 2); t /= 2; }
        b = (int)t;
        kx = KT[px][kxi];
        ky = KH > 1 ? KT[py][r] : 0;
    } else if (form == 2) {
        a = (int)(i % A);
        long long t = i / A;
        b = (int)(t % B); t /= B;
        kx = (int)(t % KW);
        ky = (int)(t / KW);
    } else if (form == 3) {
        b = (int)(i % B); a = (int)(i / B); ky = kx = 0;
    } else {
        a = (int)(i % A); b = (int)(i / A); ky = kx = 0;
    }
    float v = zero ? 0.f : W[(((long long)a * B + b) * KH + ky) * KW + kx];
    if constexpr (sizeof(TD) == 4) dst[i] = v; else dst[i] = __float2bfloat16_rn(v);
}
extern "C" int mopoe_pack_weight(const float* W, int A, int B, int KH, int KW, int form, int py, int px, int bpad,
                                 void* dst, int dst_dtype, void* stream) {
    long long total;
    if (form == 0) total = (long long)A * KH * KW * bpad;
    else if (form == 1) total = (long long)B * (KH > 1 ? 2 : 1) * 2 * A;
    else if (form == 2) total = (long long)KH * KW * B * A;
    else if (form == 3 || form == 4) total = (long long)A * B;
    else MOPOE_FAIL("pack_weight: bad form %d", form);
    MOPOE_REQUIRE(form != 0 || bpad >= B, "pack_weight: bpad < B");
    MOPOE_REQUIRE(form != 1 || KW == 4, "pack_weight: phase form needs a 4-tap kernel");
    unsigned grid = (unsigned)ceil_div64(total, 256);
    if (dst_dtype == MOPOE_F32)
        pack_weight_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(W, A, B, KH, KW, form, py, px, bpad, (float*)dst, total);
    else
        pack_weight_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(W, A, B, KH, KW, form, py, px, bpad, (bf16*)dst, total);
    MOPOE_CHECK_LAUNCH("pack_weight");
    return 0;
}

// ---- device-side step state: keeps the whole training step capturable in a CUDA graph ------------------------------
// state[0] = step count (as float bits in an int), adam_coef[0] = lr / (1 - b1^t), adam_coef[1] = 1 / sqrt(1 - b2^t)
__global__ void step_advance_kernel(unsigned long long* rng_step, int* adam_step, float* adam_coef, float lr, float b1, float b2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (rng_step) rng_step[0] += 1ull;
        if (adam_step) {
            int t = adam_step[0] + 1;
            adam_step[0] = t;
            double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
            adam_coef[0] = (float)((double)lr / bc1);
            adam_coef[1] = (float)(1.0 / sqrt(bc2));
        }
    }
}
extern "C" int mopoe_step_advance(uint64_t* rng_step, int32_t* adam_step, float* adam_coef, float lr, float beta1,
                                  float beta2, void* stream) {
    step_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long*)rng_step, adam_step, adam_coef, lr, beta1, beta2);
    MOPOE_CHECK_LAUNCH("step_advance");
    return 0;
}
