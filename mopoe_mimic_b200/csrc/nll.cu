// nll.cu — fused reconstruction log-likelihood reductions (HBM-bound; north_star item 3).
//
//   Laplace:      sum_i -(log(2b) + |x_i - loc_i| / b)                      modalities/Modality.py:25-30
//   Categorical:  log_softmax over the vocabulary + gather at argmax(target) char_encoding/DataGeneratorText.py:51,75
//                                                                            modalities/utils.py:7-8
// Reductions: per-thread fp32 over a short strip -> block fp64 -> per-block partial -> one warp sums
// the partials in a fixed order (deterministic).
#include "common.cuh"

__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double sm[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < 8 ? sm[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__global__ void final_sum_kernel(const double* part, int n, float* out, float scale) {
    double acc = 0.0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + threadIdx.x;
        acc += warp_sum(i < n ? part[i] : 0.0);
    }
    if (threadIdx.x == 0) out[0] = (float)(acc * (double)scale);
}

// ---- Laplace ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) laplace_fwd_kernel(const float* __restrict__ loc, const float* __restrict__ x,
                                                          long long n, float inv_b, double* __restrict__ part) {
    const long long n4 = n >> 2;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 a = __ldg(reinterpret_cast<const float4*>(loc) + i);
        float4 b = __ldg(reinterpret_cast<const float4*>(x) + i);
        acc += fabsf(b.x - a.x) + fabsf(b.y - a.y) + fabsf(b.z - a.z) + fabsf(b.w - a.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 << 2; i < n; ++i) acc += fabsf(x[i] - loc[i]);
    double s = block_sum_256((double)acc * (double)inv_b);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void laplace_final_kernel(const double* part, int nchunk, double n, float log2b, float* out) {
    double acc = 0.0;
    for (int i0 = 0; i0 < nchunk; i0 += 32) {
        int i = i0 + threadIdx.x;
        acc += warp_sum(i < nchunk ? part[i] : 0.0);
    }
    if (threadIdx.x == 0) out[0] = (float)(-(n * (double)log2b) - acc);
}
extern "C" int mopoe_laplace_logprob_sum(const float* loc, const float* x, int64_t n, float scale, float* out,
                                         double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(nchunk >= 1 && nchunk <= 65535, "laplace: nchunk=%d", nchunk);
    MOPOE_REQUIRE((((uintptr_t)loc | (uintptr_t)x) & 15) == 0, "laplace: unaligned");
    cudaStream_t st = (cudaStream_t)stream;
    laplace_fwd_kernel<<<nchunk, 256, 0, st>>>(loc, x, n, 1.f / scale, ws);
    MOPOE_CHECK_LAUNCH("laplace_fwd");
    // the reference evaluates log(2*scale) on an fp32 tensor (torch.tensor(0.75), ConvNetworksImgMimic.py:54)
    laplace_final_kernel<<<1, 32, 0, st>>>(ws, nchunk, (double)n, logf(2.f * scale), out);
    MOPOE_CHECK_LAUNCH("laplace_final");
    return 0;
}
__global__ void __launch_bounds__(256) laplace_bwd_kernel(const float* __restrict__ loc, const float* __restrict__ x,
                                                          long long n, float inv_b, const float* __restrict__ gout,
                                                          float* __restrict__ dloc) {
    const float g = gout[0] * inv_b;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 a = __ldg(reinterpret_cast<const float4*>(loc) + i);
        float4 b = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 o;   // d/dloc of -|x - loc|/b = sign(x - loc)/b
        o.x = b.x > a.x ? g : (b.x < a.x ? -g : 0.f);
        o.y = b.y > a.y ? g : (b.y < a.y ? -g : 0.f);
        o.z = b.z > a.z ? g : (b.z < a.z ? -g : 0.f);
        o.w = b.w > a.w ? g : (b.w < a.w ? -g : 0.f);
        reinterpret_cast<float4*>(dloc)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 << 2; i < n; ++i) dloc[i] = x[i] > loc[i] ? g : (x[i] < loc[i] ? -g : 0.f);
}
extern "C" int mopoe_laplace_logprob_bwd(const float* loc, const float* x, int64_t n, float scale, const float* gout,
                                         float* dloc, void* stream) {
    long long blocks = ceil_div64(n >> 2, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    laplace_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(loc, x, n, 1.f / scale, gout, dloc);
    MOPOE_CHECK_LAUNCH("laplace_bwd");
    return 0;
}

// elementwise log-density (evaluation callers: importance-sampled likelihoods reduce it per sample, utils/likelihood.py:120)
__global__ void __launch_bounds__(256) laplace_elem_kernel(const float* __restrict__ loc, const float* __restrict__ x,
                                                           long long n, float b, float log2b, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        out[i] = -log2b - fabsf(x[i] - loc[i]) / b;          // torch.distributions.Laplace.log_prob's own operation order
}
extern "C" int mopoe_laplace_logprob_elem(const float* loc, const float* x, int64_t n, float scale, float* out, void* stream) {
    if (n <= 0) return 0;
    long long blocks = ceil_div64(n, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    laplace_elem_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(loc, x, n, scale, logf(2.f * scale), out);
    MOPOE_CHECK_LAUNCH("laplace_elem");
    return 0;
}

// ---- Categorical: one warp per (b, l) row of V logits -------------------------------------------------
constexpr int CAT_MAXV_PER_LANE = 8;   // V <= 256

__global__ void __launch_bounds__(256) categorical_fwd_kernel(const float* __restrict__ y, const float* __restrict__ target,
                                                              const int* __restrict__ idx_in, long long rows, int V,
                                                              float* __restrict__ logits_out, int* __restrict__ idx_out,
                                                              double* __restrict__ part) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double acc = 0.0;
    for (long long row = (long long)blockIdx.x * 8 + wib; row < rows; row += (long long)gridDim.x * 8) {
        const float* yr = y + row * V;
        float v[CAT_MAXV_PER_LANE];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
            int k = lane + 32 * j;
            v[j] = k < V ? yr[k] : -INFINITY;
            mx = fmaxf(mx, v[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) se += (lane + 32 * j < V) ? expf(v[j] - mx) : 0.f;
        se = warp_sum(se);
        const float lse = mx + logf(se);
        int id;
        if (target) {   // argmax of the target row, first maximum wins (torch .max(-1)[1] semantics)
            const float* tr = target + row * V;
            float best = -INFINITY;
            int bi = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
                int k = lane + 32 * j;
                if (k < V) {
                    float t = tr[k];
                    if (t > best) { best = t; bi = k; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            id = bi;
        } else {
            id = idx_in[row];
        }
        if (idx_out && lane == 0) idx_out[row] = id;
        if (logits_out) {
#pragma unroll
            for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
                int k = lane + 32 * j;
                if (k < V) logits_out[row * V + k] = v[j] - lse;
            }
        }
        // OneHotCategorical(logits=l).log_prob renormalises l again (idempotent up to rounding): emulate it
        float l2[CAT_MAXV_PER_LANE], mx2 = -INFINITY;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
            l2[j] = (lane + 32 * j < V) ? v[j] - lse : -INFINITY;
            mx2 = fmaxf(mx2, l2[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
        float se2 = 0.f;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) se2 += (lane + 32 * j < V) ? expf(l2[j] - mx2) : 0.f;
        se2 = warp_sum(se2);
        const float lse2 = mx2 + logf(se2);
        float pick = 0.f;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j)
            if (lane + 32 * j == id) pick = l2[j] - lse2;
        pick = warp_sum(pick);
        if (lane == 0) acc += (double)pick;
    }
    __shared__ double sm[8];
    if (lane == 0) sm[wib] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i];
        part[blockIdx.x] = s;
    }
}
// Vocabulary-sized rows (word-encoded text, V in the thousands): the same arithmetic with the row re-read from L1/L2 in
// lane-strided passes instead of held in registers (same per-lane summation order as the register version).
__global__ void __launch_bounds__(256) categorical_fwd_big_kernel(const float* __restrict__ y, const float* __restrict__ target,
                                                                  const int* __restrict__ idx_in, long long rows, int V,
                                                                  float* __restrict__ logits_out, int* __restrict__ idx_out,
                                                                  double* __restrict__ part) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double acc = 0.0;
    for (long long row = (long long)blockIdx.x * 8 + wib; row < rows; row += (long long)gridDim.x * 8) {
        const float* yr = y + row * V;
        float mx = -INFINITY;
        for (int k = lane; k < V; k += 32) mx = fmaxf(mx, yr[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int k = lane; k < V; k += 32) se += expf(yr[k] - mx);
        se = warp_sum(se);
        const float lse = mx + logf(se);
        int id;
        if (target) {
            const float* tr = target + row * V;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int k = lane; k < V; k += 32) {
                const float t = tr[k];
                if (t > best) { best = t; bi = k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            id = bi;
        } else {
            id = idx_in[row];
        }
        if (idx_out && lane == 0) idx_out[row] = id;
        if (logits_out)
            for (int k = lane; k < V; k += 32) logits_out[row * V + k] = yr[k] - lse;
        // OneHotCategorical(logits=l).log_prob renormalises l again (idempotent up to rounding): emulate it
        const float mx2 = mx - lse;
        float se2 = 0.f;
        for (int k = lane; k < V; k += 32) se2 += expf((yr[k] - lse) - mx2);
        se2 = warp_sum(se2);
        const float lse2 = mx2 + logf(se2);
        if (lane == 0) acc += (double)((yr[id] - lse) - lse2);
    }
    __shared__ double sm[8];
    if (lane == 0) sm[wib] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i];
        part[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) categorical_bwd_big_kernel(const float* __restrict__ y, const int* __restrict__ idx,
                                                                  long long rows, int V, const float* __restrict__ gout,
                                                                  float* __restrict__ dy) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const float g = gout[0];
    for (long long row = (long long)blockIdx.x * 8 + wib; row < rows; row += (long long)gridDim.x * 8) {
        const float* yr = y + row * V;
        float mx = -INFINITY;
        for (int k = lane; k < V; k += 32) mx = fmaxf(mx, yr[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int k = lane; k < V; k += 32) se += expf(yr[k] - mx);
        se = warp_sum(se);
        const float inv = 1.f / se;
        const int id = idx[row];
        for (int k = lane; k < V; k += 32) dy[row * V + k] = g * ((k == id ? 1.f : 0.f) - expf(yr[k] - mx) * inv);
    }
}

// ---- Categorical, V <= 256: staged rows ------------------------------------------------------------------------------
// The char-text hot path (V = 71, 262,144 rows at B = 256): a block stages CAT_RB consecutive rows of the scores (and of the
// one-hot target) in shared memory with 16-byte cp.async copies — the rows are contiguous in memory, so the whole chunk is
// one dense stream — and then ONE THREAD PER ROW walks its row out of shared memory (row stride V words: conflict-free for
// odd V, at worst 2-way otherwise).  No shuffles, ~12 instructions per element; the warp-per-row kernel above spent ~35
// shuffles and 6 dependent global loads per row and reached 0.15 of the HBM roofline.  The forward also stores the row's
// logsumexp, so the backward is a flat elementwise pass (one read of the scores, one write of the gradient).
constexpr int CAT_RB_THREADS = 64;      // rows (= threads) per block: 36 KB of staging at V = 71 -> 6 blocks per SM
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(CAT_RB_THREADS)
categorical_fwd_staged_kernel(const float* __restrict__ y, const float* __restrict__ target, const int* __restrict__ idx_in,
                              long long rows, int V, int RB, float* __restrict__ logits_out, int* __restrict__ idx_out,
                              float* __restrict__ lse_out, double* __restrict__ part) {
    extern __shared__ __align__(16) float cat_sm[];
    float* ys = cat_sm;                                   // [RB * V]
    float* ts = cat_sm + (size_t)RB * V;                  // [RB * V] (one-hot target; unused with index targets)
    __shared__ float lse_s[CAT_RB_THREADS];
    __shared__ double red[CAT_RB_THREADS / 32];
    const int tid = threadIdx.x;
    const long long nchunks = (rows + RB - 1) / RB;
    double acc = 0.0;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long row0 = ch * RB;
        const int nrows = (int)(rows - row0 < RB ? rows - row0 : RB);
        const int n = nrows * V, n4 = n >> 2;
        const float* yg = y + row0 * V;                   // 16-byte aligned: RB % 4 == 0 and the base is
        const float* tg = target ? target + row0 * V : nullptr;
        for (int i = tid; i < n4; i += CAT_RB_THREADS) {
            cp_async16(ys + 4 * i, yg + 4 * i);
            if (tg) cp_async16(ts + 4 * i, tg + 4 * i);
        }
        for (int i = (n4 << 2) + tid; i < n; i += CAT_RB_THREADS) {
            ys[i] = yg[i];
            if (tg) ts[i] = tg[i];
        }
        cp_async_wait_all();
        __syncthreads();
        float pick = 0.f;
        if (tid < nrows) {
            const float* yr = ys + tid * V;
            float mx = -INFINITY;
            for (int k = 0; k < V; ++k) mx = fmaxf(mx, yr[k]);
            float se = 0.f;
            for (int k = 0; k < V; ++k) se += expf(yr[k] - mx);
            const float lse = mx + logf(se);
            int id;
            if (tg) {            // argmax of the target row, first maximum wins (torch .max(-1)[1] semantics)
                const float* tr = ts + tid * V;
                float best = -INFINITY;
                id = 0;
                for (int k = 0; k < V; ++k) {
                    const float t = tr[k];
                    if (t > best) { best = t; id = k; }
                }
            } else {
                id = idx_in[row0 + tid];
            }
            // OneHotCategorical(logits=l).log_prob renormalises l again (idempotent up to rounding): emulate it
            const float mx2 = mx - lse;
            float se2 = 0.f;
            for (int k = 0; k < V; ++k) se2 += expf((yr[k] - lse) - mx2);
            const float lse2 = mx2 + logf(se2);
            pick = (yr[id] - lse) - lse2;
            if (idx_out) idx_out[row0 + tid] = id;
            if (lse_out) lse_out[row0 + tid] = lse;
            lse_s[tid] = lse;
        }
        acc += (double)pick;
        if (logits_out) {
            __syncthreads();
            float* og = logits_out + row0 * V;
            for (int i = tid; i < n; i += CAT_RB_THREADS) og[i] = ys[i] - lse_s[i / V];
        }
        __syncthreads();                                  // the chunk buffers are reused by the next iteration
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < CAT_RB_THREADS / 32; ++i) s += red[i];
        part[blockIdx.x] = s;
    }
}

// d/dy of sum_rows log_softmax(y)[idx] * g = g * (onehot - softmax(y)), softmax from the saved row logsumexp
__global__ void __launch_bounds__(256) categorical_bwd_flat_kernel(const float* __restrict__ y, const int* __restrict__ idx,
                                                                   const float* __restrict__ lse, long long rows, int V,
                                                                   const float* __restrict__ gout, float* __restrict__ dy) {
    const float g = gout[0];
    const long long n = rows * V, n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(y) + i);
        long long row = (i << 2) / V;
        int col = (int)((i << 2) - row * V);
        float l = __ldg(lse + row);
        int id = __ldg(idx + row);
        const float in[4] = {v.x, v.y, v.z, v.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (col == V) {
                col = 0;
                ++row;
                l = __ldg(lse + row);
                id = __ldg(idx + row);
            }
            o[j] = g * ((col == id ? 1.f : 0.f) - expf(in[j] - l));
            ++col;
        }
        reinterpret_cast<float4*>(dy)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long e = n4 << 2; e < n; ++e) {
            const long long row = e / V;
            const int col = (int)(e - row * V);
            dy[e] = g * ((col == idx[row] ? 1.f : 0.f) - expf(y[e] - lse[row]));
        }
}

static int cat_rows_per_block(int V) {
    int rb = (96 * 1024 / 8) / V;                  // scores + target chunk <= 96 KB
    if (rb > CAT_RB_THREADS) rb = CAT_RB_THREADS;
    return rb / 4 * 4;
}
extern "C" int mopoe_categorical_logprob_sum(const float* y, const float* target, const int32_t* idx, int64_t rows,
                                             int V, float* logits_out, int32_t* idx_out, float* lse_out, float* out,
                                             double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(V >= 1, "categorical: V=%d", V);
    MOPOE_REQUIRE(target || idx, "categorical: need target or idx");
    MOPOE_REQUIRE(nchunk >= 1 && nchunk <= 65535, "categorical: nchunk=%d", nchunk);
    cudaStream_t st = (cudaStream_t)stream;
    int nparts = nchunk;
    const int RB = cat_rows_per_block(V);
    const bool aligned = ((((uintptr_t)y | (uintptr_t)target) & 15) == 0);
    if (V <= 32 * CAT_MAXV_PER_LANE && RB >= 4 && aligned) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(categorical_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            if (e != cudaSuccess) MOPOE_FAIL("categorical: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            attr_set = true;
        }
        const long long nchunks = (rows + RB - 1) / RB;
        const size_t smem = (size_t)RB * V * sizeof(float) * (target ? 2 : 1);
        int bps = (int)((200 * 1024) / (smem + 1024));
        if (bps < 1) bps = 1;
        if (bps > 12) bps = 12;
        long long grid = (long long)148 * bps;
        if (grid > nchunks) grid = nchunks;
        if (grid > nchunk) grid = nchunk;
        nparts = (int)grid;
        categorical_fwd_staged_kernel<<<(unsigned)grid, CAT_RB_THREADS, smem, st>>>(y, target, idx, rows, V, RB, logits_out, idx_out,
                                                                                 lse_out, ws);
    } else {
        if (V <= 32 * CAT_MAXV_PER_LANE)
            categorical_fwd_kernel<<<nchunk, 256, 0, st>>>(y, target, idx, rows, V, logits_out, idx_out, ws);
        else
            categorical_fwd_big_kernel<<<nchunk, 256, 0, st>>>(y, target, idx, rows, V, logits_out, idx_out, ws);
        if (lse_out) {      // these kernels do not produce the row logsumexp: mark it unusable for the flat backward
            cudaError_t e = cudaMemsetAsync(lse_out, 0xff, (size_t)rows * sizeof(float), st);       // NaN pattern
            if (e != cudaSuccess) MOPOE_FAIL("categorical: memset: %s", cudaGetErrorString(e));
        }
    }
    MOPOE_CHECK_LAUNCH("categorical_fwd");
    final_sum_kernel<<<1, 32, 0, st>>>(ws, nparts, out, 1.f);
    MOPOE_CHECK_LAUNCH("categorical_final");
    return 0;
}
/* 1 if mopoe_categorical_logprob_sum fills lse_out for this (V, alignment) — i.e. the flat backward may be used */
extern "C" int mopoe_categorical_has_lse(const float* y, const float* target, int V) {
    return V <= 32 * CAT_MAXV_PER_LANE && cat_rows_per_block(V) >= 4 && ((((uintptr_t)y | (uintptr_t)target) & 15) == 0);
}
__global__ void __launch_bounds__(256) categorical_bwd_kernel(const float* __restrict__ y, const int* __restrict__ idx,
                                                              long long rows, int V, const float* __restrict__ gout,
                                                              float* __restrict__ dy) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const float g = gout[0];
    for (long long row = (long long)blockIdx.x * 8 + wib; row < rows; row += (long long)gridDim.x * 8) {
        const float* yr = y + row * V;
        float v[CAT_MAXV_PER_LANE];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
            int k = lane + 32 * j;
            v[j] = k < V ? yr[k] : -INFINITY;
            mx = fmaxf(mx, v[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) se += (lane + 32 * j < V) ? expf(v[j] - mx) : 0.f;
        se = warp_sum(se);
        const float inv = 1.f / se;
        const int id = idx[row];
#pragma unroll
        for (int j = 0; j < CAT_MAXV_PER_LANE; ++j) {
            int k = lane + 32 * j;
            if (k < V) dy[row * V + k] = g * ((k == id ? 1.f : 0.f) - expf(v[j] - mx) * inv);
        }
    }
}
extern "C" int mopoe_categorical_logprob_bwd(const float* y, const int32_t* idx, const float* lse, int64_t rows, int V,
                                             const float* gout, float* dy, void* stream) {
    MOPOE_REQUIRE(V >= 1, "categorical: V=%d", V);
    if (lse && ((((uintptr_t)y | (uintptr_t)dy) & 15) == 0)) {
        long long blocks = ceil_div64((rows * V) >> 2, 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        categorical_bwd_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, idx, lse, rows, V, gout, dy);
        MOPOE_CHECK_LAUNCH("categorical_bwd_flat");
        return 0;
    }
    long long blocks = ceil_div64(rows, 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    if (V <= 32 * CAT_MAXV_PER_LANE)
        categorical_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, idx, rows, V, gout, dy);
    else
        categorical_bwd_big_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, idx, rows, V, gout, dy);
    MOPOE_CHECK_LAUNCH("categorical_bwd");
    return 0;
}
