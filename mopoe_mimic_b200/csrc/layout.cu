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

// Every re-layout of a training step in ONE launch: block -> (job, tile) through the jobs' tile prefix (binary search).
// A step re-packs ~240 weight tensors after the optimizer update; as separate launches they cost ~1.8 ms of launch
// latency for ~0.9 GB of traffic.
constexpr int PACK_TPB = 8;      // consecutive tiles per block: one job look-up (a chain of L2 round trips) per 8 tiles
template <typename TD>
__global__ void __launch_bounds__(256) pack_batched_kernel(const mopoe_pack_job_t* __restrict__ jobs, int njobs, int total_tiles) {
    __shared__ float s[LT][MAXT + 1];
    __shared__ mopoe_pack_job_t sj;
    __shared__ int s_idx;
    const int first = (int)blockIdx.x * PACK_TPB;
    if (threadIdx.x == 0) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].tile0 <= first) lo = mid; else hi = mid - 1;
        }
        s_idx = lo;
        sj = jobs[lo];
    }
    __syncthreads();
    for (int k = 0; k < PACK_TPB; ++k) {
        const int tile = first + k;
        if (tile >= total_tiles) break;
        // jobs are sorted by tile0: crossing into the next job is a single step
        const int next0 = s_idx + 1 < njobs ? jobs[s_idx + 1].tile0 : total_tiles;
        if (tile >= next0) {
            __syncthreads();
            if (threadIdx.x == 0) {
                s_idx += 1;
                sj = jobs[s_idx];
            }
            __syncthreads();
        }
        const int local = tile - sj.tile0, bx = local % sj.nx, by = local / sj.nx;
        const int T = sj.KH * sj.KW;
        if (sj.form == 0 || sj.form == 3) {
            pack_conv_tile<TD>(s, bx, by, sj.W, sj.A, sj.B, T, sj.bpad, reinterpret_cast<TD*>(sj.dst[0]));
        } else {
            PackDst d;
#pragma unroll
            for (int i = 0; i < 4; ++i) d.p[i] = sj.dst[i];
            pack_inner_a_tile<TD>(s, bx, by, sj.W, sj.A, sj.B, sj.KH, sj.KW, sj.form == 1 ? 1 : 2, d);
        }
        __syncthreads();                 // the tile buffer is reused by the next iteration
    }
}
extern "C" int mopoe_pack_job_tiles(int A, int B, int KH, int KW, int form, int bpad, int* nx) {
    if (form == 0 || form == 3) {
        if (bpad < B) bpad = B;
        *nx = (bpad + LT - 1) / LT;
        return *nx * A;
    }
    *nx = (A + LT - 1) / LT;
    return *nx * B;
}
extern "C" int mopoe_pack_weights_batched(const mopoe_pack_job_t* jobs_dev, int njobs, int total_tiles, int dst_dtype,
                                          void* stream) {
    if (njobs <= 0 || total_tiles <= 0) return 0;
    MOPOE_REQUIRE(jobs_dev != nullptr, "pack_weights_batched: null job table");
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == MOPOE_F32) pack_batched_kernel<float><<<(unsigned)((total_tiles + PACK_TPB - 1) / PACK_TPB), 256, 0, st>>>(jobs_dev, njobs, total_tiles);
    else pack_batched_kernel<bf16><<<(unsigned)((total_tiles + PACK_TPB - 1) / PACK_TPB), 256, 0, st>>>(jobs_dev, njobs, total_tiles);
    MOPOE_CHECK_LAUNCH("pack_weights_batched");
    return 0;
}

// form: 0 conv (dsts[0]), 1 phase (dsts[0..3] 2-D / dsts[0..1] 1-D), 2 full (dsts[0]), 3 mat = conv with T=1, 4 matT = full with T=1
extern "C" int mopoe_pack_weight_tiled(const float* W, int A, int B, int KH, int KW, int form, int bpad, void* const* dsts,
                                       int dst_dtype, void* stream) {
    const int T = KH * KW;
    MOPOE_REQUIRE(T >= 1 && T <= MAXT, "pack_weight: taps=%d", T);
    cudaStream_t st = (cudaStream_t)stream;
    if (form == 0 || form == 3) {
        if (bpad < B) bpad = B;
        dim3 grid((bpad + LT - 1) / LT, A);
        if (dst_dtype == MOPOE_F32) pack_conv_kernel<float><<<grid, 256, 0, st>>>(W, A, B, T, bpad, (float*)dsts[0]);
        else pack_conv_kernel<bf16><<<grid, 256, 0, st>>>(W, A, B, T, bpad, (bf16*)dsts[0]);
    } else if (form == 1 || form == 2 || form == 4) {
        MOPOE_REQUIRE(form != 1 || KW == 4, "pack_weight: phase form needs a 4-tap kernel");
        PackDst d = {};
        const int n = form == 1 ? (KH > 1 ? 4 : 2) : 1;
        for (int i = 0; i < n; ++i) d.p[i] = dsts[i];
        dim3 grid((A + LT - 1) / LT, B);
        const int f = form == 1 ? 1 : 2;
        if (dst_dtype == MOPOE_F32) pack_inner_a_kernel<float><<<grid, 256, 0, st>>>(W, A, B, KH, KW, f, d);
        else pack_inner_a_kernel<bf16><<<grid, 256, 0, st>>>(W, A, B, KH, KW, f, d);
    } else {
        MOPOE_FAIL("pack_weight: bad form %d", form);
    }
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
void mopoe_wgrad_finish_launch(const float* part, int Z, int A, int B, int T, int bpad, float* grad, int accumulate,
                               cudaStream_t st) {
    dim3 grid((B + LT - 1) / LT, A);
    wgrad_finish_kernel<<<grid, 256, 0, st>>>(part, Z, A, B, T, bpad, grad, accumulate);
}
