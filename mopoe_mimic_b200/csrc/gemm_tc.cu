// gemm_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution kernels (bf16 operands, fp32 accumulate).
//
// fprop / dgrad  (conv_gemm_tc_kernel):   D[m, n] = sum_{r,k} A[m, r, k] * Wp[n, r*KW + k] + bias[n]
//   * A is never materialised: a 5-D TMA tensor map (k, m0, m1, r, m2) with OVERLAPPING strides describes
//     the conv windows of the bordered channels-last activation directly (a 4-tap row of a k4/s2 conv is
//     4C contiguous elements); one box = 128 output pixels x 64 window elements, landed in shared memory
//     in the 128-byte-swizzled K-major layout tcgen05.mma consumes.  No im2col buffer, no padding logic:
//     the zero border lives in the tensor, tile overhang is TMA zero fill.
//   * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, accumulator
//     128 x BN fp32 in TMEM), warps 2-5 = epilogue (tcgen05.ld -> +bias -> bf16/fp32 -> global rows).
//   * smem ring of 4-8 stages (A 16 KB + B BN*128 B each), mbarrier full/empty pairs, tcgen05.commit
//     releases a stage as soon as the MMAs reading it retire.
//
// wgrad (conv_wgrad_tc_kernel):   dWp[n, r*KW + j] = sum_m dY[m, n] * A[m, r, j]
//   * the reduction runs over pixels, so both operands are MN-major: the same 5-D window map (box 64 window
//     elements x 64 pixels) and a 4-D row map of dY, UMMA descriptors with the MN-major bit set;
//   * split over pixel ranges across CTAs, fp32 partial tiles to a workspace, fixed-order reduction
//     (deterministic — no atomics).
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

// ---- driver entry point for tensor-map encoding (no libcuda link) ------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_tc_state = -1;   // -1 unknown, 0 unavailable, 1 ok

static int tc_init() {
    if (g_tc_state >= 0) return g_tc_state;
    g_tc_state = 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
        return 0;
    g_encode = (PFN_encodeTiled)fn;
    g_tc_state = 1;
    return 1;
}
extern "C" int mopoe_tc_available(void) { return tc_init(); }

static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, const char* what) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 1; i < rank; ++i) {
        uint64_t s = strides_elems[i] * 2;           // bf16 bytes
        if (dims[i] == 1 && s == 0) s = 16;          // a singleton dim never advances; any legal stride will do
        if (s % 16 != 0) MOPOE_FAIL("%s: stride[%d]=%llu bytes not a multiple of 16", what, i, (unsigned long long)s);
        gstr[i - 1] = s;
    }
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        MOPOE_FAIL("%s: cuTensorMapEncodeTiled failed (CUresult %d) dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]", what,
                   (int)r, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                   (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                   rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    }
    return 0;
}

static int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
// split a `rows`-row tile over (m0, m1, m2): BX * BY * NB == rows
static void tile_split(int E0, int E1, int rows, int& BX, int& BY, int& NB) {
    BX = E0 >= rows ? rows : pow2_ceil(E0);
    int rem = rows / BX;
    BY = rem < pow2_ceil(E1) ? rem : pow2_ceil(E1);
    NB = rem / BY;
}

constexpr int TC_THREADS = 192;
constexpr int SMEM_LIMIT = 232448;   // 227 KB
constexpr int SMEM_HEADER = 1024;    // barriers + tmem pointer + bias staging lives after the stages

struct TcFwdParams {
    int E0, E1, E2, BX, BY, NB, T0, T1, T2;
    int R, KW, N, BN, stages, tmem_cols;
    long long d_off, s0, s1, s2;
    void* d;
    int d_is_bf16;
    const float* bias;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-B alignment
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * 128;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t hdr = base + (uint32_t)p.stages * stage_bytes;  // header after the ring
    // header: full[stages] | empty[stages] | tmem_full | tmem_ptr | bias[BN]
    const uint32_t full0 = hdr, empty0 = hdr + 8u * p.stages, tmem_full = hdr + 16u * p.stages, tmem_slot = tmem_full + 8;
    uint8_t* const hdr_gen = gen + (size_t)p.stages * stage_bytes;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(hdr_gen + 16 * p.stages + 8);
    float* bias_s = reinterpret_cast<float*>(hdr_gen + 16 * p.stages + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int t0 = tile % p.T0, t1 = (tile / p.T0) % p.T1, t2 = tile / (p.T0 * p.T1);
    const int n0 = blockIdx.y * p.BN;
    const int kpw = p.KW >> 6;                                    // 64-element k-blocks per window row
    const int nkb = p.R * kpw;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < p.BN; j += 128) bias_s[j] = (p.bias && n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nkb; ++it) {
                const int r = it / kpw, kc = it - r * kpw;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
                mbar_expect_tx(full0 + 8 * stage, stage_bytes);
                tma_load_5d(sa, &mapA, full0 + 8 * stage, kc * 64, t0 * p.BX, t1 * p.BY, r, t2 * p.NB);
                tma_load_2d(sb, &mapB, full0 + 8 * stage, r * p.KW + kc * 64, n0);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nkb; ++it) {
                mbar_wait(full0 + 8 * stage, phase);
                fence_after();
                const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
                const uint64_t da = smem_desc_sw128(sa, 0, 1024), db = smem_desc_sw128(sb, 0, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)                       // 4 x (K = 16) per 64-element block: +32 B per step
                    umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
                umma_commit(empty0 + 8 * stage);                  // frees the smem stage when these MMAs retire
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full);
        }
    } else {
        // ===== epilogue: 4 warps, warp q reads TMEM lanes [32q, 32q+32) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int i1 = row % p.BX, i2 = (row / p.BX) % p.BY, i4 = row / (p.BX * p.BY);
        const int m0 = t0 * p.BX + i1, m1 = t1 * p.BY + i2, m2 = t2 * p.NB + i4;
        const bool rvalid = m0 < p.E0 && m1 < p.E1 && m2 < p.E2;
        const long long o = p.d_off + (long long)m0 * p.s0 + (long long)m1 * p.s1 + (long long)m2 * p.s2 + n0;
        mbar_wait(tmem_full, 0);
        fence_after();
        for (int c = 0; c < p.BN; c += 16) {
            float v[16];
            __syncwarp();
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (rvalid && n0 + c < p.N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += bias_s[c + j];
            const bool full16 = n0 + c + 16 <= p.N;
            if (p.d_is_bf16) {
                bf16* dp = reinterpret_cast<bf16*>(p.d) + o + c;
                if (full16 && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        w[j] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    reinterpret_cast<uint4*>(dp)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    reinterpret_cast<uint4*>(dp)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                } else {
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c + j < p.N) dp[j] = __float2bfloat16_rn(v[j]);
                }
            } else {
                float* dp = reinterpret_cast<float*>(p.d) + o + c;
                if (full16 && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<float4*>(dp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c + j < p.N) dp[j] = v[j];
                }
            }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static int pick_bn(int N) {
    if (N <= 256) return (N + 15) / 16 * 16;
    const int cands[] = {256, 192, 160, 128};
    int best = 128, best_tiles = 1 << 30, best_waste = 1 << 30;
    for (int c : cands) {
        int tiles = (N + c - 1) / c, waste = tiles * c - N;
        if (waste < best_waste || (waste == best_waste && tiles < best_tiles)) {
            best = c; best_tiles = tiles; best_waste = waste;
        }
    }
    return best;
}

int mopoe_tc_fwd_eligible(const mopoe_window_t* A, const mopoe_rows_t* D) {
    if (!tc_init()) return 0;
    if (A->a_dtype != MOPOE_BF16) return 0;
    if (A->KW % 64 != 0) return 0;
    if (D->N < 16) return 0;
    if ((A->sA0 % 8) || (A->sA1 % 8) || (A->sA2 % 8) || (A->sAr % 8) || (A->a_off % 8)) return 0;
    if ((reinterpret_cast<uintptr_t>(A->a) & 15) != 0) return 0;
    if ((long long)A->E0 * A->E1 * A->E2 >= (1ll << 31)) return 0;
    return 1;
}

static bool g_fwd_attr_set = false;

int mopoe_conv_gemm_tc(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* stream) {
    TcFwdParams p;
    p.E0 = A->E0; p.E1 = A->E1; p.E2 = A->E2; p.R = A->R; p.KW = A->KW; p.N = D->N;
    tile_split(p.E0, p.E1, 128, p.BX, p.BY, p.NB);
    p.T0 = (p.E0 + p.BX - 1) / p.BX; p.T1 = (p.E1 + p.BY - 1) / p.BY; p.T2 = (p.E2 + p.NB - 1) / p.NB;
    p.BN = pick_bn(p.N);
    p.tmem_cols = pow2_ceil(p.BN < 32 ? 32 : p.BN);
    const int stage_bytes = 128 * 128 + p.BN * 128;
    const int hdr_bytes = 16 * 8 + 16 + 4 * p.BN + 64;
    int stages = (SMEM_LIMIT - 1024 - hdr_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    const int nkb = p.R * (p.KW / 64);
    if (stages > nkb) stages = nkb;
    MOPOE_REQUIRE(stages >= 1, "conv_gemm_tc: no room for a stage (BN=%d)", p.BN);
    p.stages = stages;
    p.d_off = D->d_off; p.s0 = D->s0; p.s1 = D->s1; p.s2 = D->s2; p.d = D->d;
    p.d_is_bf16 = D->d_dtype == MOPOE_BF16;
    p.bias = bias;
    CUtensorMap mapA, mapB;
    {
        const uint64_t dims[5] = {(uint64_t)A->KW, (uint64_t)A->E0, (uint64_t)A->E1, (uint64_t)A->R, (uint64_t)A->E2};
        const uint64_t str[5] = {1, (uint64_t)A->sA0, (uint64_t)A->sA1, (uint64_t)A->sAr, (uint64_t)A->sA2};
        const uint32_t box[5] = {64, (uint32_t)p.BX, (uint32_t)p.BY, 1, (uint32_t)p.NB};
        const bf16* basep = reinterpret_cast<const bf16*>(A->a) + A->a_off;
        if (encode_map(&mapA, basep, 5, dims, str, box, "conv_gemm_tc(A)")) return 1;
    }
    {
        const uint64_t K = (uint64_t)A->R * A->KW;
        const uint64_t dims[2] = {K, (uint64_t)p.N};
        const uint64_t str[2] = {1, K};
        const uint32_t box[2] = {64, (uint32_t)p.BN};
        if (encode_map(&mapB, Wp, 2, dims, str, box, "conv_gemm_tc(B)")) return 1;
    }
    const int smem = 1024 + stages * stage_bytes + hdr_bytes;
    if (!g_fwd_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        g_fwd_attr_set = true;
    }
    dim3 grid((unsigned)(p.T0 * p.T1 * p.T2), (unsigned)((p.N + p.BN - 1) / p.BN));
    conv_gemm_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(mapA, mapB, p);
    MOPOE_CHECK_LAUNCH("conv_gemm_tc");
    return 0;
}

// =====================================================================================================================
// wgrad:  dWp[n, r*KW + j] = sum_m dY[m, n] * A[m, r, j]
// =====================================================================================================================
constexpr int WG_BKM = 64;            // pixels (reduction rows) per pipeline stage

struct TcWgParams {
    int E0, E1, E2, bx, by, nb, T0, T1, T2, TM;
    int R, KW, N, BNJ, JT, stages, tmem_cols;
    int Z, m_per_split;
    long long K;                      // R*KW
    float* out;                       // dWp (Z == 1) or workspace [Z][N][K]
    int accumulate;
    int NTN, tiles_out, items;        // persistent schedule: n-tiles (pair: n-tile PAIRS), output tiles, tiles_out * Z work items
    int pair;                         // host only: CTA-pair kernel
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapY, const TcWgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t y_bytes = 2 * WG_BKM * 128, a_bytes = (uint32_t)(p.BNJ / 64) * WG_BKM * 128;
    const uint32_t stage_bytes = y_bytes + a_bytes;
    const uint32_t hdr = base + (uint32_t)p.stages * stage_bytes;
    const uint32_t full0 = hdr, empty0 = hdr + 8u * p.stages, tmem_full = hdr + 16u * p.stages, tmem_slot = tmem_full + 8;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(gen + (size_t)p.stages * stage_bytes + 16 * p.stages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jt = blockIdx.x;                       // tile over (r, j-block)
    const int r = jt / p.JT, j0 = (jt - r * p.JT) * p.BNJ;
    const int n0 = blockIdx.y * 128;
    const int z = blockIdx.z;
    const int mt_begin = z * p.m_per_split;
    const int mt_end = min(p.TM, mt_begin + p.m_per_split);
    const int nkb = mt_end - mt_begin;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapY);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = mt_begin; mt < mt_end; ++mt) {
                const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t sy = base + stage * stage_bytes, sa = sy + y_bytes;
                const uint32_t bar = full0 + 8 * stage;
                mbar_expect_tx(bar, stage_bytes);
                tma_load_4d(sy, &mapY, bar, n0, t0 * p.bx, t1 * p.by, t2 * p.nb);
                tma_load_4d(sy + WG_BKM * 128, &mapY, bar, n0 + 64, t0 * p.bx, t1 * p.by, t2 * p.nb);
                for (int qd = 0; qd < p.BNJ / 64; ++qd)
                    tma_load_5d(sa + qd * WG_BKM * 128, &mapA, bar, j0 + 64 * qd, t0 * p.bx, t1 * p.by, r, t2 * p.nb);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.BNJ, 1, 1);      // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nkb; ++it) {
                mbar_wait(full0 + 8 * stage, phase);
                fence_after();
                const uint32_t sy = base + stage * stage_bytes, sa = sy + y_bytes;
                // MN-major SW128: 64-wide MN chunks are LBO = BKM*128 B apart, 8 reduction rows are SBO = 1024 B apart
                const uint64_t dy = smem_desc_sw128(sy, WG_BKM * 128, 1024), da = smem_desc_sw128(sa, WG_BKM * 128, 1024);
#pragma unroll
                for (int k = 0; k < WG_BKM / 16; ++k)                      // 16 reduction rows = 2048 B per step
                    umma_bf16(tmem_base, dy + 128 * k, da + 128 * k, idesc, (it | k) != 0);
                umma_commit(empty0 + 8 * stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        const int n = n0 + q * 32 + lane;
        float* orow = p.out + ((long long)z * p.N + n) * p.K + (long long)r * p.KW + j0;
        const bool direct_acc = p.Z == 1 && p.accumulate;
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            fence_after();
        }
        for (int c = 0; c < p.BNJ; c += 16) {
            float v[16];
            if (nkb > 0) {
                __syncwarp();
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (n < p.N && j0 + c < p.KW) {
                float* dp = orow + c;
                if (j0 + c + 16 <= p.KW && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (direct_acc) {
                            float4 e = reinterpret_cast<float4*>(dp)[j];
                            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                        }
                        reinterpret_cast<float4*>(dp)[j] = o;
                    }
                } else {
                    for (int j = 0; j < 16; ++j)
                        if (j0 + c + j < p.KW) dp[j] = (direct_acc ? dp[j] : 0.f) + v[j];
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// PERSISTENT variant: one CTA per SM loops over work items (output tile x pixel split); the accumulator is
// double-buffered in TMEM (2 x BNJ columns), so draining item i (tcgen05.ld -> fp32 partial tile in global) overlaps
// the MMAs of item i+1 and the TMEM allocation / barrier set-up is paid once per SM instead of once per tile.
// Items are ordered split-major: the CTAs running side by side work on the SAME pixel range, so the dY / window
// tiles they share come out of L2.
// PAIR: clusters of 2 CTAs run one 256 x BNJ tcgen05.mma.cta_group::2 per step: CTA r holds the dY tile of ITS n-tile
// (2*np + r) and window columns [r*BNJ/2, (r+1)*BNJ/2) — 32 KB per CTA and k-step instead of 48 (ncu: the single-CTA kernel
// sits at 51 B/clk/SM of operand fill with the tensor pipe 58 % active).
template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_persist_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapY,
                             const TcWgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int item_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nqd = PAIR ? p.BNJ / 128 : p.BNJ / 64;          // 64-column window chunks this CTA loads per stage
    const uint32_t y_bytes = 2 * WG_BKM * 128, a_bytes = (uint32_t)nqd * WG_BKM * 128;
    const uint32_t stage_bytes = y_bytes + a_bytes;
    const uint32_t hdr = base + (uint32_t)p.stages * stage_bytes;
    // header: full[stages] | empty[stages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
    const uint32_t full0 = hdr, empty0 = hdr + 8u * p.stages, tfull0 = hdr + 16u * p.stages, tempty0 = tfull0 + 16,
                   tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(gen + (size_t)p.stages * stage_bytes + 16 * p.stages + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapY);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, PAIR ? 2 : 128);           // pair: one (remote) arrival per CTA, at the leader
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_2cta(tmem_slot, (uint32_t)p.tmem_cols);
        else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        {
            // (whole warp converged, one elected lane issues: see tc::elect_one)
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader0 = PAIR ? mapa_shared(full0, 0) : full0;
            for (int item = item_first; item < p.items; item += item_step) {
                const int z = item / p.tiles_out, tile = item - z * p.tiles_out;
                const int jt = tile / p.NTN, nt = PAIR ? 2 * (tile - jt * p.NTN) + (int)rank : tile - jt * p.NTN;
                const int r = jt / p.JT, j0 = (jt - r * p.JT) * p.BNJ, n0 = nt * 128;
                const int mt_begin = z * p.m_per_split, mt_end = min(p.TM, mt_begin + p.m_per_split);
                for (int mt = mt_begin; mt < mt_end; ++mt) {
                    const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sy = base + stage * stage_bytes, sa = sy + y_bytes;
                    if (!elect_one()) {
                    } else if (PAIR) {
                        const uint32_t bar = full_leader0 + 8 * stage;
                        if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * stage_bytes);
                        tma_load_4d_2cta(sy, &mapY, bar, n0, t0 * p.bx, t1 * p.by, t2 * p.nb);
                        tma_load_4d_2cta(sy + WG_BKM * 128, &mapY, bar, n0 + 64, t0 * p.bx, t1 * p.by, t2 * p.nb);
                        for (int qd = 0; qd < nqd; ++qd)
                            tma_load_5d_2cta(sa + qd * WG_BKM * 128, &mapA, bar, j0 + (int)rank * (p.BNJ >> 1) + 64 * qd, t0 * p.bx,
                                             t1 * p.by, r, t2 * p.nb);
                    } else {
                        const uint32_t bar = full0 + 8 * stage;
                        mbar_expect_tx(bar, stage_bytes);
                        tma_load_4d(sy, &mapY, bar, n0, t0 * p.bx, t1 * p.by, t2 * p.nb);
                        tma_load_4d(sy + WG_BKM * 128, &mapY, bar, n0 + 64, t0 * p.bx, t1 * p.by, t2 * p.nb);
                        for (int qd = 0; qd < nqd; ++qd)
                            tma_load_5d(sa + qd * WG_BKM * 128, &mapA, bar, j0 + 64 * qd, t0 * p.bx, t1 * p.by, r, t2 * p.nb);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, p.BNJ, 1, 1);      // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int item = item_first; item < p.items; item += item_step, ++iter) {
                const int z = item / p.tiles_out;
                const int mt_begin = z * p.m_per_split, mt_end = min(p.TM, mt_begin + p.m_per_split);
                const int nkb = mt_end - mt_begin;
                const int acc = iter & 1;
                const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);              // epilogue has drained this accumulator
                fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BNJ);
                for (int it = 0; it < nkb; ++it) {
                    mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const uint32_t sy = base + stage * stage_bytes, sa = sy + y_bytes;
                    const uint64_t dy = smem_desc_sw128(sy, WG_BKM * 128, 1024), da = smem_desc_sw128(sa, WG_BKM * 128, 1024);
                    if (!elect_one()) {
                    } else if (PAIR) {
#pragma unroll
                        for (int k = 0; k < WG_BKM / 16; ++k)
                            umma_bf16_2cta(d_tmem, dy + 128 * k, da + 128 * k, idesc, (it | k) != 0);
                        umma_commit_2cta(empty0 + 8 * stage, 3);
                    } else {
#pragma unroll
                        for (int k = 0; k < WG_BKM / 16; ++k)
                            umma_bf16(d_tmem, dy + 128 * k, da + 128 * k, idesc, (it | k) != 0);
                        umma_commit(empty0 + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) {
                    if (PAIR) umma_commit_2cta(tfull0 + 8 * acc, 3);
                    else umma_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const bool direct_acc = p.Z == 1 && p.accumulate;
        const uint32_t tempty_leader0 = PAIR ? mapa_shared(tempty0, 0) : tempty0;
        int iter = 0;
        for (int item = item_first; item < p.items; item += item_step, ++iter) {
            const int z = item / p.tiles_out, tile = item - z * p.tiles_out;
            const int jt = tile / p.NTN, nt = PAIR ? 2 * (tile - jt * p.NTN) + (int)rank : tile - jt * p.NTN;
            const int r = jt / p.JT, j0 = (jt - r * p.JT) * p.BNJ;
            const int n = nt * 128 + q * 32 + lane;
            const int acc = iter & 1;
            const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
            float* orow = p.out + ((long long)z * p.N + n) * p.K + (long long)r * p.KW + j0;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BNJ);
            for (int c = 0; c < p.BNJ; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(t_addr + (uint32_t)c, v);
                if (n < p.N && j0 + c < p.KW) {
                    float* dp = orow + c;
                    if (j0 + c + 16 <= p.KW && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            if (direct_acc) {
                                float4 e = reinterpret_cast<float4*>(dp)[j];
                                o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                            }
                            reinterpret_cast<float4*>(dp)[j] = o;
                        }
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (j0 + c + j < p.KW) dp[j] = (direct_acc ? dp[j] : 0.f) + v[j];
                    }
                }
            }
            fence_before();
            if (PAIR) {
                named_bar_sync(1, 128);                                  // all 4 epilogue warps of this CTA have drained
                if (threadIdx.x == 64) mbar_arrive_cluster(tempty_leader0 + 8 * acc);
            } else {
                mbar_arrive(tempty0 + 8 * acc);
            }
        }
    }
    fence_before();
    if (PAIR) {
        cluster_sync_all();
        if (warp == 1) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

void mopoe_split_reduce_launch(const float* ws, int Z, long long n, float* out, int accumulate, cudaStream_t st);
void mopoe_wgrad_finish_launch(const float* part, int Z, int A, int B, int T, int bpad, float* grad, int accumulate,
                               cudaStream_t st);

extern "C" int mopoe_tc_wgrad_built(void) { return 1; }

int mopoe_tc_wgrad_eligible(const mopoe_window_t* A, const mopoe_rows_t* dY) {
    if (!tc_init()) return 0;
    if (A->a_dtype != MOPOE_BF16 || dY->d_dtype != MOPOE_BF16) return 0;
    if (A->KW % 64 != 0 || dY->N < 16 || dY->N % 8 != 0) return 0;     // (narrow dY: TMA zero-fills the rest of the 128-row tile)
    if ((A->sA0 % 8) || (A->sA1 % 8) || (A->sA2 % 8) || (A->sAr % 8) || (A->a_off % 8)) return 0;
    if ((dY->s0 % 8) || (dY->s1 % 8) || (dY->s2 % 8) || (dY->d_off % 8)) return 0;
    if ((reinterpret_cast<uintptr_t>(A->a) & 15) || (reinterpret_cast<uintptr_t>(dY->d) & 15)) return 0;
    return 1;
}

static int wg_persistent() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_WGRAD_PERSISTENT");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

static void wg_plan(const mopoe_window_t* A, const mopoe_rows_t* dY, TcWgParams& p) {
    p.E0 = A->E0; p.E1 = A->E1; p.E2 = A->E2; p.R = A->R; p.KW = A->KW; p.N = dY->N;
    p.K = (long long)A->R * A->KW;
    tile_split(p.E0, p.E1, WG_BKM, p.bx, p.by, p.nb);
    p.T0 = (p.E0 + p.bx - 1) / p.bx; p.T1 = (p.E1 + p.by - 1) / p.by; p.T2 = (p.E2 + p.nb - 1) / p.nb;
    p.TM = p.T0 * p.T1 * p.T2;
    const int cands[] = {256, 192, 128, 64};
    int best = 64, best_cost = 1 << 30;
    for (int c : cands) {
        int tiles = (p.KW + c - 1) / c;
        int cost = tiles * c;                        // MMA columns issued per window row (waste included)
        if (cost < best_cost) { best = c; best_cost = cost; }
    }
    p.BNJ = best;
    p.JT = (p.KW + p.BNJ - 1) / p.BNJ;
    p.tmem_cols = pow2_ceil(p.BNJ < 32 ? 32 : p.BNJ);
    // CTA pairs: two n-tiles of one window tile per cluster.  Only for an EVEN number of n-tiles (an odd count would spend a
    // whole tile of MMAs on out-of-range rows: N = 384 -> 4 tiles for 3) and 128-column halves of the window tile.
    const int ntn_single = (p.N + 127) / 128;
    static int wg_pair = -1;
    if (wg_pair < 0) {
        const char* e = getenv("MOPOE_WGRAD_PAIR");
        wg_pair = (e && e[0] == '0') ? 0 : 1;
    }
    p.pair = wg_pair && wg_persistent() && ntn_single % 2 == 0 && p.BNJ % 128 == 0;
    const int units = p.pair ? 74 : 148;               // concurrent work items: SMs, or SM pairs
    const int tiles_out = p.R * p.JT * (p.pair ? ntn_single / 2 : ntn_single);
    // split count: fill the 148 SMs in whole waves (a 2.05-wave grid wastes a third of its time in the tail) while
    // keeping >= 8 pixel blocks per CTA and the partials workspace bounded
    const long long ws_cap = 1ll << 29;              // 512 MB of partials at most
    int Z = 1;
    double best_eff = -1.0;
    for (int z = 1; z <= 64 && z <= p.TM; ++z) {
        if (z > 1 && (p.TM / z < 8 || (long long)z * p.N * p.K * 4 > ws_cap)) break;
        const long long items = (long long)tiles_out * z;
        const long long waves = (items + units - 1) / units;
        if (waves > 4 && z > 1) break;
        const double eff = (double)items / (double)(waves * units);
        if (eff > best_eff + 0.02) { best_eff = eff; Z = z; }
    }
    if (wg_persistent()) {
        // persistent schedule: no per-item set-up cost, so the split only trades tail waste against partial-tile traffic.
        // estimated time (us) = ceil(items / 148) * (k-blocks per item * t_kb + t_item) + Z * N * K * 8 B / ~5 TB/s
        const double t_kb = 0.14 * (double)(p.BNJ / 64) + 0.05, t_item = 0.6;
        double best_t = 1e30;
        Z = 1;
        // few output tiles (the 16 x 128 tap gradients of the single-channel layers): allow one split per SM
        const int zmax = tiles_out <= 2 ? units : 64;
        for (int z = 1; z <= zmax && z <= p.TM; ++z) {
            const int mps = (p.TM + z - 1) / z;
            const int zz = (p.TM + mps - 1) / mps;
            if (zz != z) continue;
            if (z > 1 && (mps < 4 || (long long)z * p.N * p.K * 4 > ws_cap)) break;
            const long long items = (long long)tiles_out * z;
            const double t = (double)((items + units - 1) / units) * (mps * t_kb + t_item) + (z > 1 ? (double)z : 0.5) * p.N * (double)p.K * 8.0 / 5.0e6;
            if (t < best_t * 0.98) { best_t = t; Z = z; }
        }
    }
    p.m_per_split = (p.TM + Z - 1) / Z;
    p.Z = (p.TM + p.m_per_split - 1) / p.m_per_split;
    p.NTN = p.pair ? ntn_single / 2 : ntn_single;
    p.tiles_out = tiles_out;
    p.items = tiles_out * p.Z;
    if (p.pair && p.items < units / 2) {          // too little work for 74 pairs: plan again for single CTAs
        static thread_local bool again = false;
        if (!again) {
            again = true;
            const int saved = wg_pair;
            wg_pair = 0;
            wg_plan(A, dY, p);
            wg_pair = saved;
            again = false;
            return;
        }
    }
    if (wg_persistent()) p.tmem_cols = pow2_ceil(2 * p.BNJ < 32 ? 32 : 2 * p.BNJ);
    const int stage_bytes = 2 * WG_BKM * 128 + (p.pair ? p.BNJ / 128 : p.BNJ / 64) * WG_BKM * 128;
    int stages = (SMEM_LIMIT - 1024 - 256) / stage_bytes;
    if (stages > 8) stages = 8;
    p.stages = stages;
}

size_t mopoe_conv_wgrad_ws_tc(const mopoe_window_t* A, const mopoe_rows_t* dY) {
    TcWgParams p;
    wg_plan(A, dY, p);
    return p.Z > 1 ? (size_t)p.Z * p.N * p.K * sizeof(float) : 0;
}

static bool g_wg_attr_set = false;

// fin != nullptr: partial sums always go to `ws`, then ONE kernel reduces the splits, re-lays the conv-form gradient
// out into the parameter's own layout and accumulates into the (flat) gradient buffer: fin = {A, B, taps, bpad}
int mopoe_conv_wgrad_tc(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate, void* ws,
                        size_t ws_bytes, void* stream, const int* fin) {
    TcWgParams p;
    wg_plan(A, dY, p);
    const size_t need = (p.Z > 1 || fin) ? (size_t)p.Z * p.N * p.K * sizeof(float) : 0;
    MOPOE_REQUIRE(ws_bytes >= need, "conv_wgrad_tc: workspace %zu < %zu", ws_bytes, need);
    p.out = (p.Z > 1 || fin) ? (float*)ws : dWp;
    p.accumulate = fin ? 0 : accumulate;
    if (fin) p.Z = p.Z;   // (Z == 1 with fin: the kernel's direct path writes the single partial into ws[0])
    CUtensorMap mapA, mapY;
    {
        const uint64_t dims[5] = {(uint64_t)A->KW, (uint64_t)A->E0, (uint64_t)A->E1, (uint64_t)A->R, (uint64_t)A->E2};
        const uint64_t str[5] = {1, (uint64_t)A->sA0, (uint64_t)A->sA1, (uint64_t)A->sAr, (uint64_t)A->sA2};
        const uint32_t box[5] = {64, (uint32_t)p.bx, (uint32_t)p.by, 1, (uint32_t)p.nb};
        if (encode_map(&mapA, reinterpret_cast<const bf16*>(A->a) + A->a_off, 5, dims, str, box, "conv_wgrad_tc(A)")) return 1;
    }
    {
        const uint64_t dims[4] = {(uint64_t)dY->N, (uint64_t)A->E0, (uint64_t)A->E1, (uint64_t)A->E2};
        const uint64_t str[4] = {1, (uint64_t)dY->s0, (uint64_t)dY->s1, (uint64_t)dY->s2};
        const uint32_t box[4] = {64, (uint32_t)p.bx, (uint32_t)p.by, (uint32_t)p.nb};
        if (encode_map(&mapY, reinterpret_cast<const bf16*>(dY->d) + dY->d_off, 4, dims, str, box, "conv_wgrad_tc(dY)")) return 1;
    }
    const int stage_bytes = 2 * WG_BKM * 128 + (p.pair ? p.BNJ / 128 : p.BNJ / 64) * WG_BKM * 128;
    const int smem = 1024 + p.stages * stage_bytes + 256;
    if (!g_wg_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) MOPOE_FAIL("conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        g_wg_attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (wg_persistent()) {
        static bool attr_p = false;
        static int num_sms = 148;
        if (!attr_p) {
            cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(conv_wgrad_tc_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
            if (e != cudaSuccess) MOPOE_FAIL("conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
            if (const char* e2 = getenv("MOPOE_GEMM_SMS")) {
                const int v = atoi(e2);
                if (v >= 2 && v <= num_sms) num_sms = v & ~1;
            }
            attr_p = true;
        }
        if (p.pair) {
            const int clusters = p.items < num_sms / 2 ? p.items : num_sms / 2;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(2 * clusters));
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = (size_t)smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, conv_wgrad_tc_persist_kernel<true>, mapA, mapY, p);
            if (e != cudaSuccess) MOPOE_FAIL("conv_wgrad_tc_pair: launch: %s", cudaGetErrorString(e));
        } else {
            const int grid = p.items < num_sms ? p.items : num_sms;
            conv_wgrad_tc_persist_kernel<false><<<grid, TC_THREADS, smem, st>>>(mapA, mapY, p);
        }
        MOPOE_CHECK_LAUNCH("conv_wgrad_tc_persist");
    } else {
        dim3 grid((unsigned)(p.R * p.JT), (unsigned)((p.N + 127) / 128), (unsigned)p.Z);
        conv_wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(mapA, mapY, p);
        MOPOE_CHECK_LAUNCH("conv_wgrad_tc");
    }
    if (fin) {
        MOPOE_REQUIRE(fin[0] == p.N && (long long)fin[2] * fin[3] == p.K, "conv_wgrad_tc: finish spec does not match N/K");
        mopoe_wgrad_finish_launch((const float*)ws, p.Z, fin[0], fin[1], fin[2], fin[3], dWp, accumulate, st);
        MOPOE_CHECK_LAUNCH("wgrad_finish");
    } else if (p.Z > 1) {
        mopoe_split_reduce_launch((const float*)ws, p.Z, (long long)p.N * p.K, dWp, accumulate, st);
        MOPOE_CHECK_LAUNCH("split_reduce");
    }
    return 0;
}
size_t mopoe_conv_wgrad_ws_tc_fin(const mopoe_window_t* A, const mopoe_rows_t* dY) {
    TcWgParams p;
    wg_plan(A, dY, p);
    return (size_t)p.Z * p.N * p.K * sizeof(float);
}

// ---- helpers shared with gemm_tc_persist.cu ---------------------------------------------------------------------------
int mopoe_tc_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box, const char* what) {
    return encode_map(map, base, rank, dims, strides_elems, box, what);
}
int mopoe_tc_init_state() { return tc_init(); }
void mopoe_tc_tile_split(int E0, int E1, int rows, int& BX, int& BY, int& NB) { tile_split(E0, E1, rows, BX, BY, NB); }
int mopoe_tc_pick_bn(int N) { return pick_bn(N); }
