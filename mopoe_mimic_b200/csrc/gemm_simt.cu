// gemm_simt.cu — CUDA-core implicit-GEMM convolution (fp32 accumulate).  This is the fp32
// VALIDATION path (tcgen05 kind::tf32 keeps only 10 mantissa bits, so rtol 1e-5 parity needs FFMA) and
// the catch-all for shapes the tcgen05 kernels do not take (N = 71, tiny K).  Same operand description
// (mopoe_window_t / mopoe_rows_t) as the tensor-core kernels, so both paths share all host-side layout logic.
#include "common.cuh"

struct WinDev {
    const void* a;
    long long a_off, sA0, sA1, sA2, sAr;
    int E0, E1, E2, R, KW;
};
struct RowsDev {
    void* d;
    long long d_off, s0, s1, s2;
    int N;
};
static WinDev to_dev(const mopoe_window_t* A) {
    WinDev w;
    w.a = A->a; w.a_off = A->a_off; w.sA0 = A->sA0; w.sA1 = A->sA1; w.sA2 = A->sA2; w.sAr = A->sAr;
    w.E0 = A->E0; w.E1 = A->E1; w.E2 = A->E2; w.R = A->R; w.KW = A->KW;
    return w;
}
static RowsDev to_dev(const mopoe_rows_t* D) {
    RowsDev r;
    r.d = D->d; r.d_off = D->d_off; r.s0 = D->s0; r.s1 = D->s1; r.s2 = D->s2; r.N = D->N;
    return r;
}

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&o)[4]) { ldv<4>(p, o); }
template <typename T>
__device__ __forceinline__ float ld1(const T* p) {
    if constexpr (sizeof(T) == 4) return *p; else return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void st1(T* p, float v) {
    if constexpr (sizeof(T) == 4) *p = v; else *p = __float2bfloat16_rn(v);
}

constexpr int BM = 64, BN = 64, BK = 8;

// ---- D[m,n] = sum_k A[m,k] Wp[n,k] + bias[n] ---------------------------------------------------------------
template <typename TA, typename TD>
__global__ void __launch_bounds__(256) gemm_fwd_simt_kernel(WinDev A, const TA* __restrict__ Wp, const float* __restrict__ bias,
                                                            RowsDev D, int M, int K) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m_base = blockIdx.x * BM, n_base = blockIdx.y * BN;
    const TA* a = (const TA*)A.a;
    // loader role: threads 0..127 fetch A (row = l/2, 4 k's), threads 128..255 fetch B
    const int l = tid & 127;
    const int lrow = l >> 1, lk = (l & 1) * 4;
    long long lbase = -1;
    if (tid < 128) {
        int m = m_base + lrow;
        if (m < M) {
            int m0 = m % A.E0;
            int t = m / A.E0;
            int m1 = t % A.E1, m2 = t / A.E1;
            lbase = A.a_off + m0 * A.sA0 + m1 * A.sA1 + m2 * A.sA2;
        }
    } else {
        int n = n_base + lrow;
        if (n < D.N) lbase = (long long)n * K;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (lbase >= 0) {
            if (tid < 128) {
                int kk = k0 + lk;
                int r = kk / A.KW, kw = kk - r * A.KW;
                ld4(a + lbase + r * A.sAr + kw, v);
            } else {
                ld4(Wp + lbase + k0 + lk, v);
            }
        }
        float(*S)[BM + 4] = tid < 128 ? As : Bs;
#pragma unroll
        for (int i = 0; i < 4; ++i) S[lk + i][lrow] = v[i];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[4], bv[4];
            *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    TD* d = (TD*)D.d;
    const int n0 = n_base + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias)
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n0 + j < D.N) bv[j] = bias[n0 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m_base + ty * 4 + i;
        if (m >= M) continue;
        int m0 = m % A.E0;
        int t = m / A.E0;
        int m1 = t % A.E1, m2 = t / A.E1;
        long long o = D.d_off + m0 * D.s0 + m1 * D.s1 + m2 * D.s2 + n0;
        float ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ov[j] = acc[i][j] + bv[j];
        if (n0 + 3 < D.N && ((o & 3) == 0)) {
            stv<4>(d + o, ov);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n0 + j < D.N) st1(d + o + j, ov[j]);
        }
    }
}

int mopoe_conv_gemm_simt(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* stream) {
    const long long Ml = (long long)A->E0 * A->E1 * A->E2;
    const int K = A->R * A->KW;
    MOPOE_REQUIRE(Ml > 0 && Ml < (1ll << 31), "conv_gemm: M=%lld", Ml);
    MOPOE_REQUIRE(A->KW % 8 == 0, "conv_gemm: KW=%d must be a multiple of 8", A->KW);
    MOPOE_REQUIRE((A->a_off | A->sA0 | A->sA1 | A->sA2 | A->sAr) % 4 == 0, "conv_gemm: window strides must be multiples of 4 elements");
    dim3 grid((unsigned)ceil_div64(Ml, BM), (unsigned)ceil_div64(D->N, BN));
    WinDev w = to_dev(A);
    RowsDev r = to_dev(D);
    cudaStream_t st = (cudaStream_t)stream;
    if (A->a_dtype == MOPOE_F32 && D->d_dtype == MOPOE_F32)
        gemm_fwd_simt_kernel<float, float><<<grid, 256, 0, st>>>(w, (const float*)Wp, bias, r, (int)Ml, K);
    else if (A->a_dtype == MOPOE_BF16 && D->d_dtype == MOPOE_BF16)
        gemm_fwd_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>(w, (const bf16*)Wp, bias, r, (int)Ml, K);
    else if (A->a_dtype == MOPOE_BF16 && D->d_dtype == MOPOE_F32)
        gemm_fwd_simt_kernel<bf16, float><<<grid, 256, 0, st>>>(w, (const bf16*)Wp, bias, r, (int)Ml, K);
    else
        MOPOE_FAIL("conv_gemm: unsupported dtype pair %d -> %d", A->a_dtype, D->d_dtype);
    MOPOE_CHECK_LAUNCH("gemm_fwd_simt");
    return 0;
}

// ---- dWp[n, j] = sum_m dY[m, n] * A[m, j]   (reduction over pixels, split across blockIdx.z) ----------------
template <typename T>
__global__ void __launch_bounds__(256) gemm_wgrad_simt_kernel(WinDev A, RowsDev Y, float* __restrict__ out, int M, int K,
                                                              int m_per_split, int direct, int accumulate) {
    __shared__ float Ys[BK][BN + 4];
    __shared__ float As[BK][BM + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n_base = blockIdx.y * BN, j_base = blockIdx.x * BM;
    const int z = blockIdx.z;
    const int m_begin = z * m_per_split, m_end = min(M, m_begin + m_per_split);
    const T* a = (const T*)A.a;
    const T* y = (const T*)Y.d;
    const int l = tid & 127;
    const int lrow = l >> 4, lc = (l & 15) * 4;      // 8 rows (m) x 16 quads
    long long coff = -1;                             // column offset for this loader thread
    bool yvec = false;
    if (tid < 128) {
        int n = n_base + lc;
        if (n < Y.N) coff = n;
        yvec = (n + 3 < Y.N) && ((Y.N & 3) == 0);
    } else {
        int j = j_base + lc;
        if (j < K) {
            int r = j / A.KW;
            coff = r * A.sAr + (j - r * A.KW);
        }
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
    for (int mb = m_begin; mb < m_end; mb += BK) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        int m = mb + lrow;
        if (m < m_end && coff >= 0) {
            int m0 = m % A.E0;
            int t = m / A.E0;
            int m1 = t % A.E1, m2 = t / A.E1;
            if (tid < 128) {
                long long o = Y.d_off + m0 * Y.s0 + m1 * Y.s1 + m2 * Y.s2 + coff;
                if (yvec && ((o & 3) == 0)) ld4(y + o, v);
                else
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (coff + i < Y.N) v[i] = ld1(y + o + i);
            } else {
                ld4(a + A.a_off + m0 * A.sA0 + m1 * A.sA1 + m2 * A.sA2 + coff, v);
            }
        }
        float(*S)[BN + 4] = tid < 128 ? Ys : As;
        *reinterpret_cast<float4*>(&S[lrow][lc]) = make_float4(v[0], v[1], v[2], v[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float yv[4], av[4];
            *reinterpret_cast<float4*>(yv) = *reinterpret_cast<const float4*>(&Ys[k][ty * 4]);
            *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(yv[i], av[jj], acc[i][jj]);
        }
        __syncthreads();
    }
    float* o = direct ? out : out + (long long)z * Y.N * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n_base + ty * 4 + i;
        if (n >= Y.N) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            int j = j_base + tx * 4 + jj;
            if (j < K) {
                long long p = (long long)n * K + j;
                o[p] = (direct && accumulate ? o[p] : 0.f) + acc[i][jj];
            }
        }
    }
}
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ ws, int Z, long long n,
                                                           float* __restrict__ out, int accumulate) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < Z; ++z) s += ws[(long long)z * n + i];
    out[i] = (accumulate ? out[i] : 0.f) + s;
}

void mopoe_split_reduce_launch(const float* ws, int Z, long long n, float* out, int accumulate, cudaStream_t st) {
    split_reduce_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(ws, Z, n, out, accumulate);
}

static int wgrad_splits(long long M, int N, int K) {
    long long tiles = ceil_div64(N, BN) * ceil_div64(K, BM);
    long long z = ceil_div64(148 * 4, tiles);
    long long zmax = ceil_div64(M, 256);
    if (z > zmax) z = zmax;
    if (z < 1) z = 1;
    if (z > 512) z = 512;
    return (int)z;
}
size_t mopoe_conv_wgrad_ws_simt(const mopoe_window_t* A, const mopoe_rows_t* dY) {
    long long M = (long long)A->E0 * A->E1 * A->E2;
    int K = A->R * A->KW;
    int Z = wgrad_splits(M, dY->N, K);
    return Z > 1 ? (size_t)Z * dY->N * K * sizeof(float) : 0;
}
void mopoe_wgrad_finish_launch(const float* part, int Z, int A, int B, int T, int bpad, float* grad, int accumulate,
                               cudaStream_t st);
size_t mopoe_conv_wgrad_ws_simt_fin(const mopoe_window_t* A, const mopoe_rows_t* dY) {
    long long M = (long long)A->E0 * A->E1 * A->E2;
    int K = A->R * A->KW;
    return (size_t)wgrad_splits(M, dY->N, K) * dY->N * K * sizeof(float);
}
int mopoe_conv_wgrad_simt(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate, void* ws,
                          size_t ws_bytes, void* stream, const int* fin) {
    const long long Ml = (long long)A->E0 * A->E1 * A->E2;
    const int K = A->R * A->KW;
    MOPOE_REQUIRE(Ml > 0 && Ml < (1ll << 31), "conv_wgrad: M=%lld", Ml);
    MOPOE_REQUIRE(A->KW % 8 == 0, "conv_wgrad: KW=%d must be a multiple of 8", A->KW);
    MOPOE_REQUIRE(A->a_dtype == dY->d_dtype, "conv_wgrad: A and dY dtypes differ");
    int Z = wgrad_splits(Ml, dY->N, K);
    size_t need = (Z > 1 || fin) ? (size_t)Z * dY->N * K * sizeof(float) : 0;
    MOPOE_REQUIRE(ws_bytes >= need, "conv_wgrad: workspace %zu < %zu", ws_bytes, need);
    int mps = (int)(ceil_div64(ceil_div64(Ml, Z), BK) * BK);
    dim3 grid((unsigned)ceil_div64(K, BM), (unsigned)ceil_div64(dY->N, BN), Z);
    cudaStream_t st = (cudaStream_t)stream;
    WinDev w = to_dev(A);
    RowsDev r = to_dev(dY);
    float* out = (Z > 1 || fin) ? (float*)ws : dWp;
    MOPOE_DISPATCH_T(A->a_dtype, T, {
        gemm_wgrad_simt_kernel<T><<<grid, 256, 0, st>>>(w, r, out, (int)Ml, K, mps, Z == 1, fin ? 0 : accumulate);
    });
    MOPOE_CHECK_LAUNCH("gemm_wgrad_simt");
    if (fin) {
        MOPOE_REQUIRE(fin[0] == dY->N && fin[2] * fin[3] == K, "conv_wgrad: finish spec does not match N/K");
        mopoe_wgrad_finish_launch((const float*)ws, Z, fin[0], fin[1], fin[2], fin[3], dWp, accumulate, st);
        MOPOE_CHECK_LAUNCH("wgrad_finish");
    } else if (Z > 1) {
        long long n = (long long)dY->N * K;
        split_reduce_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>((const float*)ws, Z, n, dWp, accumulate);
        MOPOE_CHECK_LAUNCH("split_reduce");
    }
    return 0;
}
