// layout.cu — weight re-layout kernels, sector-efficient in BOTH directions through a shared-memory tile.
//
// The fp32 master weights stay in the reference's layouts W[A][B][taps] (Conv: A = out, B = in; ConvTranspose:
// A = in, B = out; taps = KH*KW, fastest).  The GEMM kernels want K-major bf16 operands:
//   conv-form   dst[a][t][b']           (b' < bpad, zero beyond B)      block = (a, 64 b's)
//   phase-form  dst_p[b][(r,kxi)][a]    tap = KT[py][r], KT[px][kxi]    block = (b, 64 a's)
//   full-form   dst[t][b][a]                                            block = (b, 64 a's)
// A naive thread-per-output kernel reads 4 useful bytes of every 32-byte sector (the taps are the fastest source
// dim, never the fastest destination dim); here each block loads a [64][taps] tile with fully used sectors,
// transposes it in shared memory, and writes 128-byte contiguous runs.
// The inverse (weight GRADIENT, conv-form partial sums -> parameter layout) is fused with the split-K reduction and
// the accumulation into the flat gradient buffer: mopoe_wgrad_finish.
#include "common.cuh"

constexpr int LT = 64;          // tile extent along the transposed dim
constexpr int MAXT = 16;        // taps

template <typename TD>
__device__ __forceinline__ void put(TD* p, float v) {
    if constexpr (sizeof(TD) == 4) *p = v; else *p = __float2bfloat16_rn(v);
}

// conv-form: tile grid (ceil(bpad/64), A)
template <typename TD>
__device__ __forceinline__ void pack_conv_tile(float (&s)[LT][MAXT + 1], int bx, int by, const float* __restrict__ W, int A,
                                               int B, int T, int bpad, TD* __restrict__ dst) {
    const int a = by, b0 = bx * LT;
    const float* src = W + ((long long)a * B + b0) * T;
    const int nb = min(LT, B - b0);                        // may be <= 0 for the zero-padded tail tile
    for (int i = threadIdx.x; i < LT * T; i += 256) {
        const int bi = i / T, t = i - bi * T;
        s[bi][t] = bi < nb ? src[i] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LT * T; i += 256) {
        const int t = i / LT, bi = i - t * LT;
        if (b0 + bi < bpad) put(dst + ((long long)a * T + t) * bpad + b0 + bi, s[bi][t]);
    }
}
template <typename TD>
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ W, int A, int B, int T, int bpad,
                                                        TD* __restrict__ dst) {
    __shared__ float s[LT][MAXT + 1];
    pack_conv_tile<TD>(s, blockIdx.x, blockIdx.y, W, A, B, T, bpad, dst);
}

struct PackDst {
    void* p[4];
};
// phase-form (form 1) / full-form (form 2) / transposed 1x1 (form 2 with T = 1): tile grid (ceil(A/64), B)
template <typename TD>
__device__ __forceinline__ void pack_inner_a_tile(float (&s)[LT][MAXT + 1], int bx, int by, const float* __restrict__ W, int A,
                                                  int B, int KH, int KW, int form, const PackDst& dst) {
    const int T = KH * KW;
    const int b = by, a0 = bx * LT;
    const int na = min(LT, A - a0);
    for (int i = threadIdx.x; i < LT * T; i += 256) {
        const int ai = i / T, t = i - ai * T;
        s[ai][t] = ai < na ? W[((long long)(a0 + ai) * B + b) * T + t] : 0.f;
    }
    __syncthreads();
    const int KT[2][2] = {{3, 1}, {2, 0}};
    for (int i = threadIdx.x; i < LT * T; i += 256) {
        const int t = i / LT, ai = i - t * LT;
        if (ai >= na) continue;
        const int ky = t / KW, kx = t - ky * KW;
        if (form == 2) {
            put(reinterpret_cast<TD*>(dst.p[0]) + ((long long)t * B + b) * A + a0 + ai, s[ai][t]);
        } else {
            // tap (ky,kx) belongs to phase (py,px) at window slot (r,kxi):  KT[p][slot] = tap
            int py = 0, r = 0, px, kxi;
            if (KH > 1) { py = (ky == 3 || ky == 1) ? 0 : 1; r = (ky == KT[py][0]) ? 0 : 1; }
            px = (kx == 3 || kx == 1) ? 0 : 1;
            kxi = (kx == KT[px][0]) ? 0 : 1;
            const int nslot = KH > 1 ? 4 : 2;
            const int slot = KH > 1 ? r * 2 + kxi : kxi;
            const int ph = KH > 1 ? py * 2 + px : px;
            put(reinterpret_cast<TD*>(dst.p[ph]) + ((long long)b * nslot + slot) * A + a0 + ai, s[ai][t]);
        }
    }
}
template <typename TD>
__global__ void __launch_bounds__(256) pack_inner_a_kernel(const float* __restrict__ W, int A, int B, int KH, int KW, int form,
                                                           PackDst dst) {
    __shared__ float s[LT][MAXT + 1];
    pack_inner_a_tile<TD>(s, blockIdx.x, blockIdx.y, W, A, B, KH, KW, form, dst);
}

// Every re-layout of a training step in ONE launch.  Thread-per-item, no shared memory, no barriers: an item is
//   conv / mat  (form 0, 3): (a, b-octet, tap)      -> 8 scalar loads W[a][b0..b0+7][t]  -> ONE 16-byte store dst[a][t][b0..]
//   phase / full / matT (1, 2, 4): (a-hexadecad, b, tap) -> 16 scalar loads W[a0..a0+15][b][t] -> 32 contiguous bytes of dst
// with the tap index fastest across the lanes, so that every load instruction of a warp reads whole 32-byte sectors
// (phase / full: 128 contiguous bytes) and every store writes whole sectors.  All 8 / 16 loads of a thread are
// independent: the kernel is limited by bytes in flight, not by a load -> barrier -> store round trip per 4-KB tile as
// the shared-memory version was (1.58 ms per step for 1.2 GB = 0.12 of the HBM roofline).
constexpr int PACK_ITEMS_PER_TILE = 256;

template <typename TD>
__device__ __forceinline__ void store8(TD* p, const float (&v)[8]) {
    if constexpr (sizeof(TD) == 2) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    } else {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <typename TD>
__device__ __forceinline__ void pack_item(const mopoe_pack_job_t& j, long long item) {
    const int A = j.A, B = j.B, T = j.KH * j.KW;
    const float* __restrict__ W = reinterpret_cast<const float*>(j.W);
    if (j.form == 0 || j.form == 3) {
        const int OB = (j.bpad + 7) / 8;
        const long long total = (long long)A * OB * T;
        if (item >= total) return;
        const int t = (int)(item % T);
        const long long q = item / T;
        const int ob = (int)(q % OB), a = (int)(q / OB);
        const int b0 = ob * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (b0 + i < B) ? __ldg(W + ((long long)a * B + b0 + i) * T + t) : 0.f;
        TD* d = reinterpret_cast<TD*>(j.dst[0]) + ((long long)a * T + t) * j.bpad + b0;
        if (b0 + 8 <= j.bpad && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            store8<TD>(d, v);
        } else {
            for (int i = 0; i < 8 && b0 + i < j.bpad; ++i) put(d + i, v[i]);
        }
        return;
    }
    const int OA = (A + 15) / 16;
    const long long total = (long long)OA * B * T;
    if (item >= total) return;
    const int t = (int)(item % T);
    const long long q = item / T;
    const int b = (int)(q % B), a0 = (int)(q / B) * 16;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (a0 + i < A) ? __ldg(W + ((long long)(a0 + i) * B + b) * T + t) : 0.f;
    TD* d;
    if (j.form == 1) {
        // tap (ky,kx) belongs to phase (py,px) at window slot (r,kxi):  KT[p][slot] = tap,  KT = ((3,1),(2,0))
        const int ky = t / j.KW, kx = t - ky * j.KW;
        const int px = (kx == 3 || kx == 1) ? 0 : 1, kxi = (kx == 3 || kx == 2) ? 0 : 1;
        int ph = px, slot = kxi, nslot = 2;
        if (j.KH > 1) {
            const int py = (ky == 3 || ky == 1) ? 0 : 1, r = (ky == 3 || ky == 2) ? 0 : 1;
            ph = py * 2 + px;
            slot = r * 2 + kxi;
            nslot = 4;
        }
        d = reinterpret_cast<TD*>(j.dst[ph]) + ((long long)b * nslot + slot) * A + a0;
    } else {
        d = reinterpret_cast<TD*>(j.dst[0]) + ((long long)t * B + b) * A + a0;
    }
    if (a0 + 16 <= A && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        float lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { lo[i] = v[i]; hi[i] = v[8 + i]; }
        store8<TD>(d, lo);
        store8<TD>(d + 8, hi);
    } else {
        for (int i = 0; i < 16 && a0 + i < A; ++i) put(d + i, v[i]);
    }
}

template <typename TD>
__global__ void __launch_bounds__(PACK_ITEMS_PER_TILE) pack_batched_kernel(const mopoe_pack_job_t* __restrict__ jobs, int njobs,
                                                                           int total_tiles) {
    __shared__ mopoe_pack_job_t sj;
    const int tile = (int)blockIdx.x;
    if (threadIdx.x == 0) {                       // jobs are sorted by their first tile
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
        }
        sj = jobs[lo];
    }
    __syncthreads();
    pack_item<TD>(sj, (long long)(tile - sj.tile0) * PACK_ITEMS_PER_TILE + threadIdx.x);
}
template <typename TD>
__global__ void __launch_bounds__(PACK_ITEMS_PER_TILE) pack_single_kernel(const mopoe_pack_job_t j) {
    pack_item<TD>(j, (long long)blockIdx.x * PACK_ITEMS_PER_TILE + threadIdx.x);
}
static long long pack_job_items(int A, int B, int KH, int KW, int form, int bpad) {
    const long long T = (long long)KH * KW;
    if (form == 0 || form == 3) {
        if (bpad < B) bpad = B;
        return (long long)A * ((bpad + 7) / 8) * T;
    }
    return (long long)((A + 15) / 16) * B * T;
}
extern "C" int mopoe_pack_job_tiles(int A, int B, int KH, int KW, int form, int bpad, int* nx) {
    if (nx) *nx = 0;                              // (tiles are flat runs of 256 items; kept for ABI compatibility)
    return (int)((pack_job_items(A, B, KH, KW, form, bpad) + PACK_ITEMS_PER_TILE - 1) / PACK_ITEMS_PER_TILE);
}
extern "C" int mopoe_pack_weights_batched(const mopoe_pack_job_t* jobs_dev, int njobs, int total_tiles, int dst_dtype,
                                          void* stream) {
    if (njobs <= 0 || total_tiles <= 0) return 0;
    MOPOE_REQUIRE(jobs_dev != nullptr, "pack_weights_batched: null job table");
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == MOPOE_F32) pack_batched_kernel<float><<<(unsigned)total_tiles, PACK_ITEMS_PER_TILE, 0, st>>>(jobs_dev, njobs, total_tiles);
    else pack_batched_kernel<bf16><<<(unsigned)total_tiles, PACK_ITEMS_PER_TILE, 0, st>>>(jobs_dev, njobs, total_tiles);
    MOPOE_CHECK_LAUNCH("pack_weights_batched");
    return 0;
}

// form: 0 conv (dsts[0]), 1 phase (dsts[0..3] 2-D / dsts[0..1] 1-D), 2 full (dsts[0]), 3 mat = conv with T=1, 4 matT = full with T=1
extern "C" int mopoe_pack_weight_tiled(const float* W, int A, int B, int KH, int KW, int form, int bpad, void* const* dsts,
                                       int dst_dtype, void* stream) {
    const int T = KH * KW;
    MOPOE_REQUIRE(T >= 1 && T <= MAXT, "pack_weight: taps=%d", T);
    MOPOE_REQUIRE(form >= 0 && form <= 4, "pack_weight: bad form %d", form);
    MOPOE_REQUIRE(form != 1 || KW == 4, "pack_weight: phase form needs a 4-tap kernel");
    MOPOE_REQUIRE(!(form == 3 || form == 4) || T == 1, "pack_weight: mat forms need a 1x1 kernel");
    mopoe_pack_job_t j = {};
    j.W = W;
    const int n = form == 1 ? (KH > 1 ? 4 : 2) : 1;
    for (int i = 0; i < n; ++i) j.dst[i] = dsts[i];
    j.A = A; j.B = B; j.KH = KH; j.KW = KW;
    j.form = form == 4 ? 2 : form;                 // matT = full-form with one tap
    j.bpad = bpad < B ? B : bpad;
    const long long items = pack_job_items(A, B, KH, KW, form, bpad);
    const unsigned grid = (unsigned)((items + PACK_ITEMS_PER_TILE - 1) / PACK_ITEMS_PER_TILE);
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == MOPOE_F32) pack_single_kernel<float><<<grid, PACK_ITEMS_PER_TILE, 0, st>>>(j);
    else pack_single_kernel<bf16><<<grid, PACK_ITEMS_PER_TILE, 0, st>>>(j);
    MOPOE_CHECK_LAUNCH("pack_weight_tiled");
    return 0;
}

// grad[a][b][t] (+)= sum_z part[z][a][t*bpad + b]      grid (ceil(B/64), A)
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ part, int Z, int A, int B, int T, int bpad,
                                                           float* __restrict__ grad, int accumulate) {
    __shared__ float s[LT][MAXT + 1];
    const int a = blockIdx.y, b0 = blockIdx.x * LT;
    const long long zstride = (long long)A * T * bpad;
    for (int i = threadIdx.x; i < LT * T; i += 256) {
        const int t = i / LT, bi = i - t * LT;
        float v = 0.f;
        if (b0 + bi < B) {
            const float* p = part + ((long long)a * T + t) * bpad + b0 + bi;
            int z = 0;
            for (; z + 4 <= Z; z += 4) {                               // 4 loads in flight; summation order stays fixed
                const float v0 = __ldcs(p + z * zstride), v1 = __ldcs(p + (z + 1) * zstride);
                const float v2 = __ldcs(p + (z + 2) * zstride), v3 = __ldcs(p + (z + 3) * zstride);
                v += v0; v += v1; v += v2; v += v3;
            }
            for (; z < Z; ++z) v += __ldcs(p + z * zstride);
        }
        s[bi][t] = v;
    }
    __syncthreads();
    const int nb = min(LT, B - b0);
    float* g = grad + ((long long)a * B + b0) * T;
    for (int i = threadIdx.x; i < nb * T; i += 256) {
        const int bi = i / T, t = i - bi * T;
        g[i] = (accumulate ? g[i] : 0.f) + s[bi][t];
    }
}
// Same result bit for bit (the z-order of the sum is unchanged), 16-byte accesses in both directions: one thread owns 4
// consecutive b of one tap (a float4 of every partial, all Z of them issued back to back), the transposed tile leaves as
// float4 runs of the parameter layout.  The scalar kernel above kept 4 x 4 B in flight per thread: 1.4 ms / step for
// 1.8 GB (0.2 of HBM).  Needs bpad % 4 == 0, (T * B) % 4 == 0 and 16-byte aligned bases.
__global__ void __launch_bounds__(256) wgrad_finish_v4_kernel(const float* __restrict__ part, int Z, int A, int B, int T, int bpad,
                                                              float* __restrict__ grad, int accumulate) {
    __shared__ float s[LT][MAXT + 1];
    const int a = blockIdx.y, b0 = blockIdx.x * LT;
    const long long zstride = (long long)A * T * bpad;
    constexpr int Q = LT / 4;
    for (int i = threadIdx.x; i < Q * T; i += 256) {
        const int t = i / Q, q = i - t * Q;
        const int b = b0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < bpad) {
            const float4* p = reinterpret_cast<const float4*>(part + ((long long)a * T + t) * bpad + b);
            const long long zs4 = zstride / 4;
            int z = 0;
            for (; z + 4 <= Z; z += 4) {
                const float4 v0 = __ldcs(p + z * zs4), v1 = __ldcs(p + (z + 1) * zs4);
                const float4 v2 = __ldcs(p + (z + 2) * zs4), v3 = __ldcs(p + (z + 3) * zs4);
                v.x += v0.x; v.y += v0.y; v.z += v0.z; v.w += v0.w;
                v.x += v1.x; v.y += v1.y; v.z += v1.z; v.w += v1.w;
                v.x += v2.x; v.y += v2.y; v.z += v2.z; v.w += v2.w;
                v.x += v3.x; v.y += v3.y; v.z += v3.z; v.w += v3.w;
            }
            for (; z < Z; ++z) {
                const float4 v0 = __ldcs(p + z * zs4);
                v.x += v0.x; v.y += v0.y; v.z += v0.z; v.w += v0.w;
            }
        }
        s[4 * q][t] = v.x; s[4 * q + 1][t] = v.y; s[4 * q + 2][t] = v.z; s[4 * q + 3][t] = v.w;
    }
    __syncthreads();
    const int nb = min(LT, B - b0);
    float* g = grad + ((long long)a * B + b0) * T;
    const int n = nb * T;                       // multiple of 4 (host check), g 16-byte aligned
    for (int i = 4 * threadIdx.x; i < n; i += 4 * 256) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int bi = (i + j) / T, t = (i + j) - bi * T;
            o[j] = s[bi][t];
        }
        float4* gp = reinterpret_cast<float4*>(g + i);
        if (accumulate) {
            const float4 e = *gp;
            o[0] += e.x; o[1] += e.y; o[2] += e.z; o[3] += e.w;
        }
        *gp = make_float4(o[0], o[1], o[2], o[3]);
    }
}
// v5: 256 channels per block and FOUR float4 items per thread, the loads of a split issued together (v4 kept 16 bytes per
// thread in flight: 1.26 ms / step where the traffic — one partial + gradient read + write — is worth 0.3 ms).
constexpr int LT5 = 256;
__global__ void __launch_bounds__(256) wgrad_finish_v5_kernel(const float* __restrict__ part, int Z, int A, int B, int T, int bpad,
                                                              float* __restrict__ grad, int accumulate) {
    __shared__ float s[LT5][MAXT + 1];
    const int a = blockIdx.y, b0 = blockIdx.x * LT5;
    const long long zs4 = ((long long)A * T * bpad) / 4;
    constexpr int Q = LT5 / 4;
    const int items = Q * T;                                   // <= 1024
    const float4* p[4];
    bool ok[4];
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + 256 * k;
        const int t = i / Q, q = i - t * Q;
        const int b = b0 + 4 * q;
        ok[k] = i < items && b < bpad;
        p[k] = reinterpret_cast<const float4*>(part + ((long long)a * T + (ok[k] ? t : 0)) * bpad + (ok[k] ? b : 0));
        acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int z = 0; z < Z; ++z) {                              // fixed z order: the sum is bit-identical to v4 / the scalar kernel
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = ok[k] ? __ldcs(p[k] + z * zs4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k].x += v[k].x; acc[k].y += v[k].y; acc[k].z += v[k].z; acc[k].w += v[k].w; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + 256 * k;
        if (i < items) {
            const int t = i / Q, q = i - t * Q;
            s[4 * q][t] = acc[k].x; s[4 * q + 1][t] = acc[k].y; s[4 * q + 2][t] = acc[k].z; s[4 * q + 3][t] = acc[k].w;
        }
    }
    __syncthreads();
    const int nb = min(LT5, B - b0);
    float* g = grad + ((long long)a * B + b0) * T;
    const int n = nb * T;                       // multiple of 4 (host check), g 16-byte aligned
    float4 e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * (threadIdx.x + 256 * k);
        e[k] = (accumulate && i < n) ? *reinterpret_cast<const float4*>(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * (threadIdx.x + 256 * k);
        if (i >= n) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int bi = (i + j) / T, t = (i + j) - bi * T;
            o[j] = s[bi][t];
        }
        *reinterpret_cast<float4*>(g + i) = make_float4(o[0] + e[k].x, o[1] + e[k].y, o[2] + e[k].z, o[3] + e[k].w);
    }
}
void mopoe_wgrad_finish_launch(const float* part, int Z, int A, int B, int T, int bpad, float* grad, int accumulate,
                               cudaStream_t st) {
    static int v5 = -1;
    if (v5 < 0) {
        const char* e = getenv("MOPOE_WGRAD_FINISH_V5");
        v5 = (e && e[0] == '0') ? 0 : 1;
    }
    {
        const bool ok5 = v5 && bpad % 4 == 0 && T <= MAXT && ((long long)B * T) % 4 == 0 && ((B % LT5) * T) % 4 == 0 &&
                         ((long long)A * T * bpad) % 4 == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(grad) & 15) == 0 && T >= 4;      // (1x1 layers: 64 items per block — the v4 tile fits them better)
        if (ok5) {
            dim3 grid5((B + LT5 - 1) / LT5, A);
            wgrad_finish_v5_kernel<<<grid5, 256, 0, st>>>(part, Z, A, B, T, bpad, grad, accumulate);
            return;
        }
    }
    dim3 grid((B + LT - 1) / LT, A);
    static int v4 = -1;
    if (v4 < 0) {
        const char* e = getenv("MOPOE_WGRAD_FINISH_V4");
        v4 = (e && e[0] == '0') ? 0 : 1;
    }
    // every tile's slice of the gradient row must start 16-byte aligned and hold a multiple of 4 floats
    const bool ok = v4 && bpad % 4 == 0 && T <= MAXT && (LT * T) % 4 == 0 && ((long long)B * T) % 4 == 0 &&
                    ((min(LT, B - (B / LT) * LT) * T) % 4 == 0) && ((long long)A * T * bpad) % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(part) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
    if (ok)
        wgrad_finish_v4_kernel<<<grid, 256, 0, st>>>(part, Z, A, B, T, bpad, grad, accumulate);
    else
        wgrad_finish_kernel<<<grid, 256, 0, st>>>(part, Z, A, B, T, bpad, grad, accumulate);
}
