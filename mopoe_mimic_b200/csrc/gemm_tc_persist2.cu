// gemm_tc_persist2.cu — EXPERIMENTAL (off unless MOPOE_GEMM_BM256=1; not yet validated on hardware): the persistent
// fprop/dgrad kernel of gemm_tc_persist.cu with a 256 x BN CTA tile for BN <= 128.
//
// Why: ncu shows the N = 128 layers L2-bound, not tensor-bound (46 % tensor-pipe active vs 75 % at N = 256): a 128 x 128
// tile fetches 32 KB of operands per 2.1 MFLOP.  Two M sub-tiles per CTA share every B stage (48 KB per 4.2 MFLOP — the
// N = 256 ratio); the two fp32 accumulators and their double buffers use 4 x BN <= 512 TMEM columns.
//   stage = A0 (128 rows, 16 KB) | A1 (next m-tile, 16 KB) | B (BN rows);   per k-block: 4 MMAs into acc0, 4 into acc1.
// Schedule, barriers and epilogue are those of conv_gemm_tc_persist_kernel; a work item is (m-tile PAIR, problem, n-tile).
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

int mopoe_tc_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box, const char* what);
int mopoe_tc_init_state();
void mopoe_tc_tile_split(int E0, int E1, int rows, int& BX, int& BY, int& NB);
int mopoe_tc_pick_bn(int N);

constexpr int TC2_THREADS = 192;
constexpr int TC2_SMEM_LIMIT = 232448;
constexpr int TC2_MAXP = 4;

struct Tc2Maps {
    CUtensorMap a[TC2_MAXP];
    CUtensorMap b[TC2_MAXP];
};
struct Tc2Params {
    int E0, E1, E2, BX, BY, NB, T0, T1, T2, MT, MTP;      // MT m-tiles, MTP = ceil(MT / 2) pairs
    int R, KW, N, BN, NT, stages, tmem_cols;
    int nprob, total_items;
    long long d_off[TC2_MAXP];
    long long s0, s1, s2;
    void* d;
    int d_is_bf16;
    const float* bias;
};

__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_gemm_tc_persist_bm256_kernel(const __grid_constant__ Tc2Maps maps, const Tc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * 128;
    const uint32_t stage_bytes = 2 * a_bytes + b_bytes;
    const uint32_t hdr = base + (uint32_t)p.stages * stage_bytes;
    const uint32_t full0 = hdr, empty0 = hdr + 8u * p.stages, tfull0 = hdr + 16u * p.stages, tempty0 = tfull0 + 16,
                   tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(gen + (size_t)p.stages * stage_bytes + 16 * p.stages + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpw = p.KW >> 6;
    const int nkb = p.R * kpw;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) {
            prefetch_tmap(&maps.a[i]);
            prefetch_tmap(&maps.b[i]);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int nt = item % p.NT;
                const int tq = item / p.NT;
                const int prob = tq % p.nprob, mtp = tq / p.nprob;
                const int mtA = 2 * mtp, mtB = min(2 * mtp + 1, p.MT - 1);     // an odd tail pair repeats its first tile
                const int a0 = mtA % p.T0, a1 = (mtA / p.T0) % p.T1, a2 = mtA / (p.T0 * p.T1);
                const int b0 = mtB % p.T0, b1 = (mtB / p.T0) % p.T1, b2 = mtB / (p.T0 * p.T1);
                const int n0 = nt * p.BN;
                const CUtensorMap* ma = &maps.a[prob];
                const CUtensorMap* mb = &maps.b[prob];
                for (int it = 0; it < nkb; ++it) {
                    const int r = it / kpw, kc = it - r * kpw;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sa0 = base + stage * stage_bytes, sa1 = sa0 + a_bytes, sb = sa1 + a_bytes;
                    mbar_expect_tx(full0 + 8 * stage, stage_bytes);
                    tma_load_5d(sa0, ma, full0 + 8 * stage, kc * 64, a0 * p.BX, a1 * p.BY, r, a2 * p.NB);
                    tma_load_5d(sa1, ma, full0 + 8 * stage, kc * 64, b0 * p.BX, b1 * p.BY, r, b2 * p.NB);
                    tma_load_2d(sb, mb, full0 + 8 * stage, r * p.KW + kc * 64, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * p.BN), d1 = d0 + (uint32_t)p.BN;
                for (int it = 0; it < nkb; ++it) {
                    mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const uint32_t sa0 = base + stage * stage_bytes, sa1 = sa0 + a_bytes, sb = sa1 + a_bytes;
                    const uint64_t da0 = smem_desc_sw128(sa0, 0, 1024), da1 = smem_desc_sw128(sa1, 0, 1024),
                                   db = smem_desc_sw128(sb, 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d0, da0 + 2 * k, db + 2 * k, idesc, (it | k) != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d1, da1 + 2 * k, db + 2 * k, idesc, (it | k) != 0);
                    umma_commit(empty0 + 8 * stage);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull0 + 8 * acc);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int i1 = row % p.BX, i2 = (row / p.BX) % p.BY, i4 = row / (p.BX * p.BY);
        int iter = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
            const int nt = item % p.NT;
            const int tq = item / p.NT;
            const int prob = tq % p.nprob, mtp = tq / p.nprob;
            const int n0 = nt * p.BN;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
            for (int sub = 0; sub < 2; ++sub) {
                const int mt = 2 * mtp + sub;
                if (mt >= p.MT) break;                         // odd tail: the second accumulator holds a repeat, drop it
                const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
                const int m0 = t0 * p.BX + i1, m1 = t1 * p.BY + i2, m2 = t2 * p.NB + i4;
                const bool rvalid = m0 < p.E0 && m1 < p.E1 && m2 < p.E2;
                const long long o = p.d_off[prob] + (long long)m0 * p.s0 + (long long)m1 * p.s1 + (long long)m2 * p.s2 + n0;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.BN + sub * p.BN);
                for (int c = 0; c < p.BN; c += 16) {
                    float v[16];
                    __syncwarp();
                    tmem_ld16(t_addr + (uint32_t)c, v);
                    if (rvalid && n0 + c < p.N) {
                        if (p.bias) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (n0 + c + j < p.N) v[j] += __ldg(p.bias + n0 + c + j);
                        }
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
            mbar_arrive(tempty0 + 8 * acc);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static bool g_p2_attr_set = false;
static int g_p2_sms = 0;

// 0 = not applicable (caller uses the regular persistent kernel), 1 = error, 2 = launched
int mopoe_conv_gemm_tc_batched_bm256(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                     const mopoe_rows_t* D, void* stream) {
    if (nprob < 1 || nprob > TC2_MAXP || !mopoe_tc_init_state()) return 0;
    Tc2Params p;
    p.E0 = A[0].E0; p.E1 = A[0].E1; p.E2 = A[0].E2; p.R = A[0].R; p.KW = A[0].KW; p.N = D[0].N;
    p.BN = mopoe_tc_pick_bn(p.N);
    if (p.BN > 128) return 0;
    mopoe_tc_tile_split(p.E0, p.E1, 128, p.BX, p.BY, p.NB);
    p.T0 = (p.E0 + p.BX - 1) / p.BX; p.T1 = (p.E1 + p.BY - 1) / p.BY; p.T2 = (p.E2 + p.NB - 1) / p.NB;
    p.MT = p.T0 * p.T1 * p.T2;
    if (p.MT < 2) return 0;
    p.MTP = (p.MT + 1) / 2;
    p.NT = (p.N + p.BN - 1) / p.BN;
    int cols = 4 * p.BN, pc = 32;
    while (pc < cols) pc <<= 1;
    if (pc > 512) return 0;
    p.tmem_cols = pc;
    const int stage_bytes = 2 * 128 * 128 + p.BN * 128;
    const int hdr_bytes = 16 * 8 + 48 + 64;
    int stages = (TC2_SMEM_LIMIT - 1024 - hdr_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return 0;
    p.stages = stages;
    p.nprob = nprob;
    p.total_items = p.MTP * p.NT * nprob;
    p.s0 = D[0].s0; p.s1 = D[0].s1; p.s2 = D[0].s2; p.d = D[0].d;
    p.d_is_bf16 = D[0].d_dtype == MOPOE_BF16;
    p.bias = bias;
    Tc2Maps maps;
    for (int i = 0; i < nprob; ++i) {
        p.d_off[i] = D[i].d_off;
        const uint64_t dims[5] = {(uint64_t)A[i].KW, (uint64_t)A[i].E0, (uint64_t)A[i].E1, (uint64_t)A[i].R, (uint64_t)A[i].E2};
        const uint64_t str[5] = {1, (uint64_t)A[i].sA0, (uint64_t)A[i].sA1, (uint64_t)A[i].sAr, (uint64_t)A[i].sA2};
        const uint32_t box[5] = {64, (uint32_t)p.BX, (uint32_t)p.BY, 1, (uint32_t)p.NB};
        if (mopoe_tc_encode(&maps.a[i], reinterpret_cast<const bf16*>(A[i].a) + A[i].a_off, 5, dims, str, box, "conv_gemm_tc2(A)"))
            return 1;
        const uint64_t K = (uint64_t)A[i].R * A[i].KW;
        const uint64_t dimsb[2] = {K, (uint64_t)p.N};
        const uint64_t strb[2] = {1, K};
        const uint32_t boxb[2] = {64, (uint32_t)p.BN};
        if (mopoe_tc_encode(&maps.b[i], Wp[i], 2, dimsb, strb, boxb, "conv_gemm_tc2(B)")) return 1;
    }
    for (int i = nprob; i < TC2_MAXP; ++i) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; }
    if (!g_p2_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_persist_bm256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TC2_SMEM_LIMIT);
        if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_p2_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_p2_sms <= 0) g_p2_sms = 148;
        g_p2_attr_set = true;
    }
    const int smem = 1024 + stages * stage_bytes + hdr_bytes;
    const int grid = p.total_items < g_p2_sms ? p.total_items : g_p2_sms;
    conv_gemm_tc_persist_bm256_kernel<<<grid, TC2_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_tc_persist_bm256 launch: %s", cudaGetErrorString(e));
    return 2;
}
