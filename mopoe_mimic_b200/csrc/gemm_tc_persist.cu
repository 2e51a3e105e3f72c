// gemm_tc_persist.cu — PERSISTENT, batched tcgen05 implicit-GEMM (fprop / dgrad family).
//
// Same math and operand maps as conv_gemm_tc_kernel (gemm_tc.cu), different schedule:
//   * grid = min(#tiles, #SMs): every CTA loops over tiles (static round-robin in (m-tile, problem, n-tile) order, so
//     CTAs running side by side read the same activation rows out of L2 — across n-tiles AND across phases);
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns): while the 4 epilogue warps drain tile i
//     (tcgen05.ld -> +bias -> bf16 -> global), the MMA warp already accumulates tile i+1 — the epilogue and the
//     pipeline fill/drain no longer sit on the tensor pipe's critical path;
//   * up to 4 problems of identical shape in ONE launch: the sub-pixel phases of a stride-2 transposed conv
//     (and of a conv's dgrad) share (M, N, K) and differ only in window origin, weights and output origin,
//     so they become one grid instead of four quarter-size launches.
//   * bf16 outputs leave through a TMA-STORE epilogue (template flag TMA_EPI): 64-column groups of the accumulator are read
//     with one batch of tcgen05.ld (no wait per 16 columns), cast, written into a 128-byte-swizzled shared-memory staging
//     tile and stored by ONE cp.async.bulk.tensor per group (two staging tiles alternate).  The register epilogue wrote
//     32 B per lane to 32 different rows per instruction (4x the LSU wavefronts of a row-contiguous store) and was the
//     bottleneck of every short-K layer (K = 512: 9,000 cycles per tile against 2,048 of MMA).
//   * fused BatchNorm statistics (TcStats): while a group sits in the staging tile, each epilogue warp sums its own 32 rows
//     column-wise (per-channel sum and sum of squares of the bf16-ROUNDED, dropout-masked values — what the consuming BN
//     reads), accumulates per CTA in shared memory and writes fixed-order partials for bn_finalize_kernel: the separate
//     statistics pass over the activation (reduce_rows_kernel<.,0>: 1.35 ms / step) disappears for conv1 -> bn2 and
//     shortcut conv -> BN.  Deterministic: fixed tile -> CTA assignment, no atomics.
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

int mopoe_tc_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box, const char* what);
int mopoe_tc_init_state();
void mopoe_tc_tile_split(int E0, int E1, int rows, int& BX, int& BY, int& NB);
int mopoe_tc_pick_bn(int N);

// TMA warp, MMA warp, 4 epilogue warps, 4 statistics warps.  (Measured and dropped: 384 threads with the statistics warps on
// warps 6, 7, 10, 11 — off the two schedulers that issue the TMA and MMA warps — was 0.15 ms/step SLOWER.)
constexpr int TCP_THREADS = 320;
constexpr int TCP_SMEM_LIMIT = 232448;
constexpr int TCP_MAXP = 4;

struct TcMaps {
    CUtensorMap a[TCP_MAXP];
    CUtensorMap b[TCP_MAXP];
    CUtensorMap d[TCP_MAXP];      // output maps (n, m0, m1, m2) of the TMA-store epilogue
    CUtensorMap r[TCP_MAXP];      // RES: the shortcut branch's rows, same geometry as d
};
struct TcStats {
    double* ws;                   // [gridDim.x * 4][2][N] partial (sum, sum of squares); NULL: no statistics
    const uint8_t* mask;          // dropout keep-mask applied between this GEMM and the BatchNorm (x * 2 * mask)
    int mask_mode;                // MOPOE_MASK_NONE / _BC ([B, N]) / _ELEM ([rows, N])
    int rows_per_b;               // flat output rows per sample (MASK_BC: sample = flat row / rows_per_b)
};
// RES: the residual combine of the block fused into the epilogue (ResidualBlocks.py:29-33):
//   stored = a * BN(r) + b * ((acc + bias) * mask_scale * mask),  acc the fp32 accumulator (never rounded to bf16 in between)
struct TcRes {
    const float *mean, *invstd, *gamma, *beta;     // the shortcut's BatchNorm
    float a, b;
    const uint8_t* mask;          // dropout keep-mask on the GEMM result; byte address = mask + moff[prob] + m0*ms0 + m1*ms1 + m2*ms2 + n
    int mask_mode;
    long long moff[TCP_MAXP], ms0, ms1, ms2;
    // zero border of the output activation, written by the otherwise idle statistics warps (zb == NULL: none)
    bf16* zb;                     // interior element (0, 0, 0, 0)
    int zB, zH, zW, zph, zpw;
    long long zsB, zsH, zsW;
};
struct TcPersistParams {
    int E0, E1, E2, BX, BY, NB, T0, T1, T2;
    int R, KW, N, BN, NT, stages, tmem_cols;
    int nprob, tiles_per_prob, total_tiles;
    long long d_off[TCP_MAXP];
    long long s0, s1, s2;
    void* d;
    int d_is_bf16;
    const float* bias;
    int nacc;                     // NT * BN: columns of the per-CTA statistics accumulators
    TcStats st;
    // split-K (register epilogue only): work item = (tile, split); split s reduces k-steps [s*kb_per_split, ...) and
    // writes an fp32 partial tile at d + s * split_stride.  ksplit == 1: off.
    int ksplit, kb_per_split;
    long long split_stride;
    TcRes rs;
};

constexpr int EPI_BAR = 1;        // named barrier of the 4 epilogue warps
constexpr int STG_FULL_BAR = 2;   // +buffer: staging tile written (128 epilogue threads arrive, 128 statistics threads wait)
constexpr int STG_FREE_BAR = 4;   // +buffer: statistics warps are done reading it (they arrive, the epilogue threads wait)
constexpr int STAT_BAR = 6;       // the 4 statistics warps among themselves
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
constexpr uint32_t STG_BYTES = 128 * 128;   // one staging tile: 128 rows x 64 bf16, SWIZZLE_128B

// PAIR: the kernel runs as clusters of 2 CTAs (one TPC) that execute ONE 256 x BN tcgen05.mma.cta_group::2 per k-step: CTA r
// holds the 128 rows of ITS m-tile (2*mp + r) and columns [r*BN/2, (r+1)*BN/2) of the weight tile, the leader (r = 0)
// issues the MMAs, each CTA's TMEM receives the 128 x BN accumulator of its own rows.  Why: the single-CTA kernel is bound by
// the SM's operand fill — ncu: 65-67 B/clk/SM of L2->shared traffic on the 128 x 256 layers (75 % tensor pipe), 57 B/clk on
// the 128 x 128 layers (47 %): a CTA moves 48 KB per 128x256x64 k-step (85 flop/B).  A pair moves 32 KB per CTA for the same
// math (128 flop/B); N = 128 layers 24 KB instead of 32 (85 instead of 64 flop/B).
// RES (with TMA_EPI): the epilogue also reads the shortcut branch's tile (TMA into ONE buffer, refilled one 64-column group ahead)
// and stores a * BN(r) + b * dropout(acc) — the block's residual combine without the round trip of the conv2 output through
// HBM (one write + one read of the activation) and without the combine launch.  The statistics warps then sum the COMBINED
// tile: the statistics of the next block's bn1.
// BNB (with TMA_EPI): the GEMM is the input gradient that feeds a BatchNorm(+ReLU, +dropout) backward; the statistics warps
// read, next to every staged output tile dy, the matching tile of the BatchNorm's INPUT x (maps.r; their own one-tile TMA
// stream, refilled one group ahead) and accumulate the two per-channel sums of the BatchNorm backward, sum(g) and sum(g * xhat) with
// g = dy * [relu gate recomputed from x] — the reduction pass over (dy, x) that otherwise follows the GEMM.
template <bool TMA_EPI, bool PAIR, bool RES = false, bool BNB = false>
__global__ void __launch_bounds__(TCP_THREADS, 1)
conv_gemm_tc_persist_kernel(const __grid_constant__ TcMaps maps, const TcPersistParams p) {
    static_assert(!RES || TMA_EPI, "the residual epilogue is a TMA-store epilogue");
    static_assert(!BNB || (TMA_EPI && !RES), "the BatchNorm-backward sums ride on the TMA-store epilogue's statistics warps");
    constexpr bool AUX = RES || BNB;                               // a second tile stream (maps.r) next to the output's
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * (PAIR ? 64 : 128);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    // ring | [TMA_EPI: 2 staging tiles (1024-B aligned: stage_bytes is a multiple of 1024) | RES / BNB: 1 auxiliary tile |
    //         statistics accumulators | bias row | RES: coefficient rows (scale, shift)] | header
    const uint32_t stg0 = base + (uint32_t)p.stages * stage_bytes;
    const uint32_t tiles_bytes = (AUX ? 3u : 2u) * STG_BYTES;
    const uint32_t res0 = stg0 + 2 * STG_BYTES;
    const uint32_t acc_bytes = (TMA_EPI && p.st.ws) ? (uint32_t)(8 * p.nacc) * 4u : 0u;
    const uint32_t bias_bytes = (TMA_EPI && p.bias) ? (uint32_t)p.nacc * 4u : 0u;
    const uint32_t coef_bytes = RES ? (uint32_t)p.nacc * 8u : BNB ? (uint32_t)p.nacc * 16u : 0u;
    const uint32_t extra = TMA_EPI ? tiles_bytes + acc_bytes + bias_bytes + coef_bytes : 0u;
    float* const sacc = reinterpret_cast<float*>(gen + (size_t)p.stages * stage_bytes + tiles_bytes);
    const float* const sbias = reinterpret_cast<const float*>(gen + (size_t)p.stages * stage_bytes + tiles_bytes + acc_bytes);
    float* const scoef = reinterpret_cast<float*>(gen + (size_t)p.stages * stage_bytes + tiles_bytes + acc_bytes + bias_bytes);
    const uint32_t hdr = stg0 + extra;
    // header: full[stages] | empty[stages] | tmem_full[2] | tmem_empty[2] | tmem_ptr | res_full[2]
    const uint32_t full0 = hdr, empty0 = hdr + 8u * p.stages, tfull0 = hdr + 16u * p.stages, tempty0 = tfull0 + 16,
                   tmem_slot = tempty0 + 16, resfull0 = tmem_slot + 8, aux_dummy = resfull0 + 8;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(gen + (size_t)p.stages * stage_bytes + extra + 16 * p.stages + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kpw = p.KW >> 6;
    const int nkb = p.R * kpw;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) {
            prefetch_tmap(&maps.a[i]);
            prefetch_tmap(&maps.b[i]);
            if (TMA_EPI) prefetch_tmap(&maps.d[i]);
            if (AUX) prefetch_tmap(&maps.r[i]);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, PAIR ? 2 : 128);           // pair: one (remote) arrival per CTA, at the leader
        }
        if (AUX) mbar_init(resfull0, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_2cta(tmem_slot, (uint32_t)p.tmem_cols);
        else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    fence_before();
    if (PAIR) cluster_sync_all();                                  // the peer's barriers exist before anything signals them
    else __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        {
            // ===== TMA producer (whole warp converged; one elected lane issues) =====
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader0 = PAIR ? mapa_shared(full0, 0) : full0;     // pair: bytes are counted at the leader
            for (int tile = tile_first; tile < p.total_tiles; tile += tile_step) {
                // tile order (m-tile, problem, n-tile): the sub-pixel phases of one pixel block run side by side, so
                // their (overlapping) activation windows are fetched from DRAM once — problem-major order re-read the
                // whole input per phase (ncu: 568 MB read for a 151 MB operand)
                const int sp = tile % p.ksplit, tl = tile / p.ksplit;
                const int nt = tl % p.NT;
                const int tq = tl / p.NT;
                const int prob = tq % p.nprob, mt = PAIR ? 2 * (tq / p.nprob) + (int)rank : tq / p.nprob;
                const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
                const int n0 = nt * p.BN;
                const CUtensorMap* ma = &maps.a[prob];
                const CUtensorMap* mb = &maps.b[prob];
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                for (int it = kb0; it < kb1; ++it) {
                    const int r = it / kpw, kc = it - r * kpw;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
                    if (elect_one()) {
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * stage_bytes);
                            tma_load_5d_2cta(sa, ma, full_leader0 + 8 * stage, kc * 64, t0 * p.BX, t1 * p.BY, r, t2 * p.NB);
                            tma_load_2d_2cta(sb, mb, full_leader0 + 8 * stage, r * p.KW + kc * 64, n0 + (int)rank * (p.BN >> 1));
                        } else {
                            mbar_expect_tx(full0 + 8 * stage, stage_bytes);
                            tma_load_5d(sa, ma, full0 + 8 * stage, kc * 64, t0 * p.BX, t1 * p.BY, r, t2 * p.NB);
                            tma_load_2d(sb, mb, full0 + 8 * stage, r * p.KW + kc * 64, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (pair: the leader CTA only); whole warp converged, one elected lane issues =====
            const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, p.BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = tile_first; tile < p.total_tiles; tile += tile_step, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);      // epilogue has drained this accumulator buffer
                fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                const int sp = tile % p.ksplit;
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                for (int it = kb0; it < kb1; ++it) {
                    mbar_wait(full0 + 8 * stage, phase);
                    fence_after();
                    const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
                    const uint64_t da = smem_desc_sw128(sa, 0, 1024), db = smem_desc_sw128(sb, 0, 1024);
                    if (elect_one()) {
                        if (PAIR) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, ((it - kb0) | k) != 0);
                            umma_commit_2cta(empty0 + 8 * stage, 3);             // frees the stage in BOTH CTAs
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, ((it - kb0) | k) != 0);
                            umma_commit(empty0 + 8 * stage);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) {
                    if (PAIR) umma_commit_2cta(tfull0 + 8 * acc, 3);             // both CTAs' epilogues
                    else umma_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
            }
        }
    } else if (TMA_EPI && warp < 6) {
        // ===== TMA-store epilogue (bf16 output, BN % 64 == 0): 4 warps, warp q owns TMEM lanes / tile rows [32q, 32q+32) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool leader = threadIdx.x == 64;                     // issues the bulk stores of this CTA
        const bool stats = p.st.ws != nullptr;
        if (p.bias)                                                // the bias row, zero-padded to the tile grid
            for (int i = threadIdx.x - 64; i < p.nacc; i += 128) const_cast<float*>(sbias)[i] = i < p.N ? __ldg(p.bias + i) : 0.f;
        if (RES) {
            // per-column scale / shift of a * BN(r), the expressions of the stand-alone combine pass (stream.cu CombineBody)
            for (int i = threadIdx.x - 64; i < p.nacc; i += 128) {
                float sc = 0.f, sh = 0.f;
                if (i < p.N) {
                    sc = p.rs.a * __ldg(p.rs.invstd + i) * __ldg(p.rs.gamma + i);
                    sh = p.rs.a * __ldg(p.rs.beta + i) - __ldg(p.rs.mean + i) * sc;
                }
                scoef[i] = sc;
                scoef[p.nacc + i] = sh;
            }
        }
        if (p.bias || RES) named_bar_sync(EPI_BAR, 128);
        const uint32_t swz = (uint32_t)(row & 7);
        int iter = 0;
        uint32_t sbuf = 0;
        const uint32_t tempty_leader0 = PAIR ? mapa_shared(tempty0, 0) : tempty0;
        // RES: ONE residual tile buffer.  A group starts by pulling the thread's residual row (64 bf16) into registers; once all
        // 128 rows are out, the leader thread refills the buffer with the next group's tile, which lands under this group's
        // TMEM loads, math and stores.  (Two buffers cost a stage of the operand ring: -17 % on the fill-bound 128-column layers.)
        const int ri1 = row % p.BX, ri2 = (row / p.BX) % p.BY, ri4 = row / (p.BX * p.BY);
        uint32_t gcount = 0;
        auto res_issue = [&](int tile_, int g_) {
            const int nt_ = tile_ % p.NT, tq_ = tile_ / p.NT;
            const int prob_ = tq_ % p.nprob, mt_ = PAIR ? 2 * (tq_ / p.nprob) + (int)rank : tq_ / p.nprob;
            const int u0 = mt_ % p.T0, u1 = (mt_ / p.T0) % p.T1, u2 = mt_ / (p.T0 * p.T1);
            mbar_expect_tx(resfull0, STG_BYTES);
            tma_load_4d(res0, &maps.r[prob_], resfull0, nt_ * p.BN + g_ * 64, u0 * p.BX, u1 * p.BY, u2 * p.NB);
        };
        if (RES && leader && tile_first < p.total_tiles) res_issue(tile_first, 0);
        for (int tile = tile_first; tile < p.total_tiles; tile += tile_step, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
            const int nt = tile % p.NT;
            const int tq = tile / p.NT;
            const int prob = tq % p.nprob, mt = PAIR ? 2 * (tq / p.nprob) + (int)rank : tq / p.nprob;
            const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
            const int n0 = nt * p.BN;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            const int ngroups = p.BN >> 6;
            for (int g = 0; g < ngroups; ++g) {
                const int c0 = n0 + g * 64;
                uint32_t w[32];                                     // 64 bf16, packed (RES: the residual row first, then the result)
                if (RES) {
                    mbar_wait(resfull0, gcount & 1u);
                    const uint32_t rrow = res0 + (uint32_t)row * 128u;
                    uint32_t chk = 0u;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t rw[4];
                        ld_shared_v4(rrow + ((((uint32_t)ch) ^ swz) << 4), rw);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { w[4 * ch + i] = rw[i]; chk ^= rw[i]; }
                    }
                    // the refill below may only be issued once these loads have RETURNED (a barrier arrival alone is not ordered
                    // behind outstanding shared-memory loads): a store that depends on every loaded word sits before the barrier
                    if (chk == 0x9e3779b9u) st_shared_b32(aux_dummy, chk);
                    named_bar_sync(EPI_BAR, 128);
                    if (leader) {
                        int ntile = tile, ng = g + 1;
                        if (ng == ngroups) { ntile = tile + tile_step; ng = 0; }
                        if (ntile < p.total_tiles) res_issue(ntile, ng);
                    }
                    ++gcount;
                }
                uint32_t r0[32], r1[32];
                tmem_ld32_nowait(t_addr + (uint32_t)(g * 64), r0);
                tmem_ld32_nowait(t_addr + (uint32_t)(g * 64 + 32), r1);
                tmem_ld_wait();
                if (g == ngroups - 1) {
                    // every TMEM read of this accumulator buffer is complete: hand it back before the stores
                    fence_before();
                    if (!PAIR) mbar_arrive(tempty0 + 8 * acc);
                }
                if (RES) {
                    const int m0 = t0 * p.BX + ri1, m1 = t1 * p.BY + ri2, m2 = t2 * p.NB + ri4;
                    const bool rvalid = m0 < p.E0 && m1 < p.E1 && m2 < p.E2;
                    const uint8_t* mrow = nullptr;                  // this row's 64 keep-mask bytes
                    if (p.rs.mask_mode != MOPOE_MASK_NONE && rvalid)
                        mrow = p.rs.mask + p.rs.moff[prob] + (long long)m0 * p.rs.ms0 + (long long)m1 * p.rs.ms1 +
                               (long long)m2 * p.rs.ms2 + c0;
                    const float mscale = p.rs.mask_mode == MOPOE_MASK_NONE ? 1.f : 2.f;
                    uint2 mk[8];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        mk[ch] = mrow ? __ldg(reinterpret_cast<const uint2*>(mrow) + ch) : make_uint2(0x01010101u, 0x01010101u);
                    const float4* s4 = reinterpret_cast<const float4*>(scoef + c0);             // broadcast reads
                    const float4* h4 = reinterpret_cast<const float4*>(scoef + p.nacc + c0);
                    const float4* b4 = reinterpret_cast<const float4*>(sbias + c0);
                    const bool has_bias = p.bias != nullptr;
                    const float bco = p.rs.b;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint32_t rw[4] = {w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]};
                        const float4 sA = s4[2 * ch], sB = s4[2 * ch + 1], hA = h4[2 * ch], hB = h4[2 * ch + 1];
                        const float scv[8] = {sA.x, sA.y, sA.z, sA.w, sB.x, sB.y, sB.z, sB.w};
                        const float shv[8] = {hA.x, hA.y, hA.z, hA.w, hB.x, hB.y, hB.z, hB.w};
                        float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (has_bias) {
                            const float4 bA = b4[2 * ch], bB = b4[2 * ch + 1];
                            bv[0] = bA.x; bv[1] = bA.y; bv[2] = bA.z; bv[3] = bA.w;
                            bv[4] = bB.x; bv[5] = bB.y; bv[6] = bB.z; bv[7] = bB.w;
                        }
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t rwv = rw[i >> 1];
                            const float rv = __uint_as_float((i & 1) ? (rwv & 0xffff0000u) : (rwv << 16));
                            const float cv = __uint_as_float(ch < 4 ? r0[8 * (ch & 3) + i] : r1[8 * (ch & 3) + i]) + bv[i];
                            const uint32_t mword = i < 4 ? mk[ch].x : mk[ch].y;
                            const float mf = ((mword >> (8 * (i & 3))) & 0xffu) ? mscale : 0.f;
                            o[i] = fmaf(rv, scv[i], shv[i]) + bco * (cv * mf);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                            w[4 * ch + i] = *reinterpret_cast<uint32_t*>(&h);
                        }
                    }
                } else if (p.bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(sbias + c0);     // broadcast reads
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 ba = b4[j], bb = b4[8 + j];
                        __nv_bfloat162 h;
                        h = __floats2bfloat162_rn(__uint_as_float(r0[4 * j]) + ba.x, __uint_as_float(r0[4 * j + 1]) + ba.y);
                        w[2 * j] = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(r0[4 * j + 2]) + ba.z, __uint_as_float(r0[4 * j + 3]) + ba.w);
                        w[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(r1[4 * j]) + bb.x, __uint_as_float(r1[4 * j + 1]) + bb.y);
                        w[16 + 2 * j] = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(r1[4 * j + 2]) + bb.z, __uint_as_float(r1[4 * j + 3]) + bb.w);
                        w[16 + 2 * j + 1] = *reinterpret_cast<uint32_t*>(&h);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        __nv_bfloat162 ha = __floats2bfloat162_rn(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                        __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                        w[j] = *reinterpret_cast<uint32_t*>(&ha);
                        w[16 + j] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                }
                // the staging tile we are about to overwrite: its previous bulk store must have finished READING it, and the
                // statistics warps must be done with it
                if (leader) bulk_wait_read<1>();
                if (stats) named_bar_sync(STG_FREE_BAR + (int)sbuf, 256);
                named_bar_sync(EPI_BAR, 128);
                // pair: all 128 threads of this CTA are past their last TMEM load: ONE arrival at the leader's barrier
                if (PAIR && g == ngroups - 1 && leader) mbar_arrive_cluster(tempty_leader0 + 8 * acc);
                const uint32_t stg = stg0 + sbuf * STG_BYTES;
                const uint32_t rbase = stg + (uint32_t)row * 128u;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)                      // 16-byte chunk ch of the row -> swizzled slot
                    st_shared_v4(rbase + ((((uint32_t)ch) ^ swz) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
                fence_proxy_async_smem();
                if (stats) named_bar_arrive(STG_FULL_BAR + (int)sbuf, 256);   // the statistics warps may read the tile
                named_bar_sync(EPI_BAR, 128);                       // all 128 rows are in the staging tile and fenced
                if (leader) {
                    tma_store_4d(&maps.d[prob], stg, c0, t0 * p.BX, t1 * p.BY, t2 * p.NB);
                    bulk_commit();
                }
                sbuf ^= 1u;
            }
        }
        if (leader) bulk_wait<0>();                                 // the last stores have landed before the CTA exits
        if (stats) {                                                // drain: pairs with the statistics warps' last arrivals
            named_bar_sync(STG_FREE_BAR, 256);
            named_bar_sync(STG_FREE_BAR + 1, 256);
        }
    } else if (TMA_EPI) {
        // ===== statistics warps (fused BatchNorm statistics): warp q sums rows [32q, 32q+32) of every staged 128 x 64 tile
        // column-wise while the epilogue warps already convert the next group.  (Round 2, first cut: the epilogue warps did
        // this themselves — on short-K layers, where the epilogue is the critical path, it doubled the kernel: 158 us against
        // 87 for the M = 1M 1x1 layer.)
        const bool stats = p.st.ws != nullptr;
        const int sq = warp - 6;
        if (RES && p.rs.zb) {
            // border pixels per image: the ph top / bottom rows (full width) + the pw left / right columns of the H middle
            // rows (same enumeration as zero_border_kernel, elementwise.cu); items = (pixel, 16-byte chunk), dealt to all CTAs
            const int Ws = p.rs.zW + 2 * p.rs.zpw;
            const int per = 2 * p.rs.zph * Ws + 2 * p.rs.zpw * p.rs.zH;
            const int cv8 = p.N >> 3;
            const long long total = (long long)p.rs.zB * per * cv8;
            for (long long i = (long long)blockIdx.x * 128 + (sq * 32 + lane); i < total; i += (long long)gridDim.x * 128) {
                const int cv = (int)(i % cv8);
                const long long qq = i / cv8;
                const int k = (int)(qq % per);
                const long long b = qq / per;
                int hs, ws;
                if (k < p.rs.zph * Ws) { hs = k / Ws; ws = k - hs * Ws; }
                else if (k < 2 * p.rs.zph * Ws) { const int k2 = k - p.rs.zph * Ws; hs = p.rs.zph + p.rs.zH + k2 / Ws; ws = k2 % Ws; }
                else {
                    const int k2 = k - 2 * p.rs.zph * Ws;
                    hs = p.rs.zph + k2 / (2 * p.rs.zpw);
                    const int j = k2 % (2 * p.rs.zpw);
                    ws = j < p.rs.zpw ? j : p.rs.zW + j;
                }
                bf16* dst = p.rs.zb + b * p.rs.zsB + (long long)(hs - p.rs.zph) * p.rs.zsH + (long long)(ws - p.rs.zpw) * p.rs.zsW + cv * 8;
                *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (stats) {
            const int q = sq;
            const int row = q * 32 + lane;
            const int st_tid = q * 32 + lane;
            const int i1 = row % p.BX, i2 = (row / p.BX) % p.BY, i4 = row / (p.BX * p.BY);
            for (int i = st_tid; i < 8 * p.nacc; i += 128) sacc[i] = 0.f;
            const int ngroups_c = p.BN >> 6;
            // BNB: the x tile of linear group G (= tile iteration * groups per tile + group); ONE buffer, see below
            auto x_issue = [&](int G) {
                const int it_ = G / ngroups_c, g_ = G - it_ * ngroups_c;
                const long long tile_l = (long long)tile_first + (long long)it_ * tile_step;
                if (tile_l >= p.total_tiles) return;
                const int tile_ = (int)tile_l;
                const int nt_ = tile_ % p.NT, tq_ = tile_ / p.NT;
                const int prob_ = tq_ % p.nprob, mt_ = PAIR ? 2 * (tq_ / p.nprob) + (int)rank : tq_ / p.nprob;
                const int u0 = mt_ % p.T0, u1 = (mt_ / p.T0) % p.T1, u2 = mt_ / (p.T0 * p.T1);
                mbar_expect_tx(resfull0, STG_BYTES);
                tma_load_4d(res0, &maps.r[prob_], resfull0, nt_ * p.BN + g_ * 64, u0 * p.BX, u1 * p.BY, u2 * p.NB);
            };
            if (BNB) {
                // per-column mean, 1/std and the forward's affine (sc, sh) — the SAME instruction sequence as the forward apply
                // pass (stream.cu affine8), so the recomputed gate is the stored activation's sign bit for bit
                for (int i = st_tid; i < p.nacc; i += 128) {
                    float mu = 0.f, is = 0.f, sc = 0.f, sh = 0.f;
                    if (i < p.N) {
                        mu = __ldg(p.rs.mean + i);
                        is = __ldg(p.rs.invstd + i);
                        sc = __fmul_rn(is, __ldg(p.rs.gamma + i));
                        sh = __fmaf_rn(-mu, sc, __ldg(p.rs.beta + i));
                    }
                    scoef[i] = mu;
                    scoef[p.nacc + i] = is;
                    scoef[2 * p.nacc + i] = sc;
                    scoef[3 * p.nacc + i] = sh;
                }
                if (st_tid == 0) x_issue(0);
            }
            named_bar_sync(STAT_BAR, 128);
            named_bar_arrive(STG_FREE_BAR, 256);                    // both staging tiles start out free
            named_bar_arrive(STG_FREE_BAR + 1, 256);
            uint32_t sbuf = 0;
            int gcount = 0;
            for (int tile = tile_first; tile < p.total_tiles; tile += tile_step) {
                const int nt = tile % p.NT;
                const int tq = tile / p.NT;
                const int mt = PAIR ? 2 * (tq / p.nprob) + (int)rank : tq / p.nprob;
                const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
                const int n0 = nt * p.BN;
                const int m0 = t0 * p.BX + i1, m1 = t1 * p.BY + i2, m2 = t2 * p.NB + i4;
                const bool rvalid = m0 < p.E0 && m1 < p.E1 && m2 < p.E2;
                const unsigned flat = (unsigned)((m2 * p.E1 + m1) * p.E0 + m0);        // flat output row (host checks < 2^31)
                const unsigned vmask = __ballot_sync(0xffffffffu, rvalid);
                // row key of the dropout mask: sample index (Dropout2d mask [B, N]) or flat row (elementwise mask [rows, N])
                const unsigned mykey = p.st.mask_mode == MOPOE_MASK_BC ? flat / (unsigned)p.st.rows_per_b : flat;
                const unsigned key_lo = __shfl_sync(0xffffffffu, mykey, 0), key_hi = __shfl_sync(0xffffffffu, mykey, 31);
                // Dropout2d mask, all 32 rows of the warp in one sample: ONE mask value per column
                const bool warp_mask = p.st.mask_mode == MOPOE_MASK_BC && key_lo == key_hi;
                unsigned long long mkw = 0ull;                        // 4 groups x 16 bits
                if (warp_mask) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int cn = n0 + g * 64 + 2 * lane;
                        if (g * 64 < p.BN && cn < p.N)
                            mkw |= (unsigned long long)*reinterpret_cast<const unsigned short*>(p.st.mask + (size_t)key_lo * p.N + cn) << (16 * g);
                    }
                }
                const int ngroups = p.BN >> 6;
                for (int g = 0; g < ngroups; ++g) {
                    const int c0 = n0 + g * 64;
                    const uint32_t stg = stg0 + sbuf * STG_BYTES;
                    // Lane l owns columns c0 + 2l, c0 + 2l + 1: one 128-byte row per load, conflict-free.
                    const int cn = c0 + 2 * lane;
                    const bool cvalid = cn < p.N;
                    const uint32_t jchunk = (uint32_t)lane >> 2, wsel = ((uint32_t)lane & 3u) << 2;
                    const uint32_t lbase = stg + (uint32_t)(q * 32) * 128u + wsel;
                    // dropout mask between this GEMM and the BatchNorm: per row (elementwise mask, or a Dropout2d mask when
                    // the warp's rows span several samples) — global loads issued ahead of the wait for the tile
                    const bool row_masks = p.st.mask_mode == MOPOE_MASK_ELEM || (p.st.mask_mode == MOPOE_MASK_BC && key_lo != key_hi);
                    unsigned short mk[32];
                    if (row_masks) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const unsigned key = __shfl_sync(0xffffffffu, mykey, i);
                            mk[i] = (cvalid && ((vmask >> i) & 1u))
                                        ? *reinterpret_cast<const unsigned short*>(p.st.mask + (size_t)key * p.N + cn) : (unsigned short)0;
                        }
                    }
                    uint32_t xw[BNB ? 32 : 1];
                    if (BNB) {
                        // pull this warp's 32 x 2 x values into registers ahead of the output tile; once every statistics thread has
                        // them the single x buffer is refilled with the next group's tile (lands under this group's math)
                        mbar_wait(resfull0, (uint32_t)gcount & 1u);
                        const uint32_t xbase = res0 + (uint32_t)(q * 32) * 128u + wsel;
                        uint32_t chk = 0u;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            xw[i] = ld_shared_b32(xbase + (uint32_t)i * 128u + ((jchunk ^ (uint32_t)(i & 7)) << 4));
                            chk ^= xw[i];
                        }
                        // (a store that depends on every loaded word, before the barrier: the loads have returned — see RES)
                        if (chk == 0x9e3779b9u) st_shared_b32(aux_dummy, chk);
                        named_bar_sync(STAT_BAR, 128);
                        if (st_tid == 0) x_issue(gcount + 1);
                        ++gcount;
                    }
                    named_bar_sync(STG_FULL_BAR + (int)sbuf, 256);      // the epilogue warps have written (and fenced) the tile
                    if (BNB) {
                        const float2 mu2 = *reinterpret_cast<const float2*>(scoef + cn);
                        const float2 is2 = *reinterpret_cast<const float2*>(scoef + p.nacc + cn);
                        const float2 sc2 = *reinterpret_cast<const float2*>(scoef + 2 * p.nacc + cn);
                        const float2 sh2 = *reinterpret_cast<const float2*>(scoef + 3 * p.nacc + cn);
                        // dropout on x: one factor per column (Dropout2d, warp inside one sample), per row, or none
                        float wf0 = 1.f, wf1 = 1.f;
                        if (warp_mask) {
                            const unsigned mkv = (unsigned)(mkw >> (16 * (g & 3))) & 0xffffu;
                            wf0 = (mkv & 0xffu) ? 2.f : 0.f;
                            wf1 = (mkv & 0xff00u) ? 2.f : 0.f;
                        }
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            uint32_t dv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                dv[i] = ld_shared_b32(lbase + (uint32_t)(hlf * 16 + i) * 128u + ((jchunk ^ (uint32_t)(i & 7)) << 4));
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int r_ = hlf * 16 + i;
                                float d0 = __uint_as_float(dv[i] << 16), d1 = __uint_as_float(dv[i] & 0xffff0000u);
                                float x0 = __uint_as_float(xw[r_] << 16), x1 = __uint_as_float(xw[r_] & 0xffff0000u);
                                if (row_masks) {
                                    x0 *= (mk[r_] & 0xffu) ? 2.f : 0.f;           // exact: the forward's masked input
                                    x1 *= (mk[r_] & 0xff00u) ? 2.f : 0.f;
                                } else if (warp_mask) {
                                    x0 *= wf0;
                                    x1 *= wf1;
                                }
                                // relu gate: the stored activation is > 0  <=>  y rounds to a non-zero bf16 (stream.cu gate_open);
                                // tile overhang rows contribute nothing
                                const bool rv_ = vmask == 0xffffffffu || ((vmask >> r_) & 1u);
                                if (rv_ && __fmaf_rn(x0, sc2.x, sh2.x) > 0x1p-134f) { s0 += d0; q0 = fmaf(d0, x0 - mu2.x, q0); }
                                if (rv_ && __fmaf_rn(x1, sc2.y, sh2.y) > 0x1p-134f) { s1 += d1; q1 = fmaf(d1, x1 - mu2.y, q1); }
                            }
                        }
                        q0 *= is2.x;                                    // sum g * xhat = invstd * sum g * (x - mean)
                        q1 *= is2.y;
                        named_bar_arrive(STG_FREE_BAR + (int)sbuf, 256);
                        if (cvalid) {
                            float* a = sacc + (size_t)(q * 2) * p.nacc + cn;           // exclusive owner of these entries
                            a[0] += s0;
                            a[1] += s1;
                            a[p.nacc] += q0;
                            a[p.nacc + 1] += q1;
                        }
                        sbuf ^= 1u;
                        continue;
                    }
                    uint32_t wv[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) wv[i] = ld_shared_b32(lbase + (uint32_t)i * 128u + ((jchunk ^ (uint32_t)(i & 7)) << 4));
                    float sa0 = 0.f, sa1 = 0.f, qa0 = 0.f, qa1 = 0.f, sb0 = 0.f, sb1 = 0.f, qb0 = 0.f, qb1 = 0.f;
                    if (row_masks) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float x0 = __uint_as_float(wv[i] << 16), x1 = __uint_as_float(wv[i] & 0xffff0000u);
                            x0 = (mk[i] & 0xffu) ? 2.f * x0 : 0.f;
                            x1 = (mk[i] & 0xff00u) ? 2.f * x1 : 0.f;
                            if (i & 1) { sb0 += x0; sb1 += x1; qb0 = fmaf(x0, x0, qb0); qb1 = fmaf(x1, x1, qb1); }
                            else { sa0 += x0; sa1 += x1; qa0 = fmaf(x0, x0, qa0); qa1 = fmaf(x1, x1, qa1); }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float x0 = __uint_as_float(wv[i] << 16), x1 = __uint_as_float(wv[i] & 0xffff0000u);
                            if (vmask != 0xffffffffu && !((vmask >> i) & 1u)) x0 = x1 = 0.f;       // tile overhang rows
                            if (i & 1) { sb0 += x0; sb1 += x1; qb0 = fmaf(x0, x0, qb0); qb1 = fmaf(x1, x1, qb1); }
                            else { sa0 += x0; sa1 += x1; qa0 = fmaf(x0, x0, qa0); qa1 = fmaf(x1, x1, qa1); }
                        }
                    }
                    // (the values are in registers: the tile may be overwritten)
                    named_bar_arrive(STG_FREE_BAR + (int)sbuf, 256);
                    float s0 = sa0 + sb0, s1 = sa1 + sb1, q0 = qa0 + qb0, q1 = qa1 + qb1;
                    if (warp_mask) {
                        const unsigned mkv = (unsigned)(mkw >> (16 * (g & 3))) & 0xffffu;
                        const float f0 = (mkv & 0xffu) ? 2.f : 0.f, f1 = (mkv & 0xff00u) ? 2.f : 0.f;
                        s0 *= f0; q0 *= f0 * f0;
                        s1 *= f1; q1 *= f1 * f1;
                    }
                    if (cvalid) {
                        float* a = sacc + (size_t)(q * 2) * p.nacc + cn;           // exclusive owner of these entries
                        a[0] += s0;
                        a[1] += s1;
                        a[p.nacc] += q0;
                        a[p.nacc + 1] += q1;
                    }
                    sbuf ^= 1u;
                }
            }
            named_bar_sync(STAT_BAR, 128);
            for (int i = st_tid; i < 8 * p.N; i += 128) {
                const int n = i % p.N, k = i / p.N;                 // k = q * 2 + which
                p.st.ws[((size_t)blockIdx.x * 8 + k) * p.N + n] = (double)sacc[(size_t)k * p.nacc + n];
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: 4 warps, warp q owns TMEM lanes [32q, 32q+32) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int i1 = row % p.BX, i2 = (row / p.BX) % p.BY, i4 = row / (p.BX * p.BY);
        int iter = 0;
        for (int tile = tile_first; tile < p.total_tiles; tile += tile_step, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (uint32_t)(iter >> 1) & 1u;
            const int sp = tile % p.ksplit, tl = tile / p.ksplit;
            const int nt = tl % p.NT;
            const int tq = tl / p.NT;
            const int prob = tq % p.nprob, mt = tq / p.nprob;
            const int t0 = mt % p.T0, t1 = (mt / p.T0) % p.T1, t2 = mt / (p.T0 * p.T1);
            const int n0 = nt * p.BN;
            const int m0 = t0 * p.BX + i1, m1 = t1 * p.BY + i2, m2 = t2 * p.NB + i4;
            const bool rvalid = m0 < p.E0 && m1 < p.E1 && m2 < p.E2;
            const long long o = p.d_off[prob] + (long long)m0 * p.s0 + (long long)m1 * p.s1 + (long long)m2 * p.s2 + n0 +
                                (long long)sp * p.split_stride;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
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
            // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld16): hand it back
            fence_before();
            mbar_arrive(tempty0 + 8 * acc);
        }
        }
    fence_before();
    if (PAIR) {
        cluster_sync_all();                    // no remote arrival / pair MMA may still target a CTA that tears down
        if (warp == 1) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

static bool g_persist_attr_set = false;
static int g_num_sms = 0;

static int pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_GEMM_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}
static int tma_epi_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_GEMM_TMA_EPI");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

// Fused BatchNorm statistics request of mopoe_conv_gemm_tc_batched_ex (host side of TcStats)
struct TcStatsReq {
    double* ws;              // >= 8 * #SMs * N doubles
    size_t ws_doubles;
    const uint8_t* mask;
    int mask_mode;
    int rows_per_b;
    int* nchunk_out;         // number of [2][N] partial slabs written (for bn_finalize_kernel)
};

// Residual combine request (host side of TcRes): D is the block's OUTPUT, R the shortcut branch's rows
struct TcResReq {
    const mopoe_rows_t* R;   // nprob row addressings of r (bf16, N columns), like D
    const float *mean, *invstd, *gamma, *beta;
    float a, b;
    const uint8_t* mask;     // keep-mask on the GEMM result: MASK_BC [E2, N]; MASK_ELEM: one byte per element of r, laid out like r
    int mask_mode;
    int dry_run;             // only answer whether the fused epilogue applies (0) or not (2): nothing is launched
    const mopoe_view_t* out; // may be NULL; the activation D addresses: the launch also writes its zero border
    int kind;                // 0: residual combine (above).  1: BatchNorm-backward sums — R addresses the BatchNorm's input x,
                             // mean / invstd / gamma / beta are that BatchNorm's, the dropout mask on x travels in the
                             // statistics request (TcStatsReq), which is mandatory; a, b, mask, out are unused
};
static int bnb_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_GEMM_BNB");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}
static int res_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_GEMM_RES");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

// nprob problems of identical (E0,E1,E2,R,KW,N, output strides): window base / weights / output origin differ.
// stats != NULL: fuse the output's per-channel statistics when the TMA-store epilogue applies; *stats->nchunk_out = 0
// tells the caller that they were NOT produced (it then runs the separate statistics pass).
// res != NULL: fuse the residual combine; returns 2 (nothing launched) when that epilogue does not apply to the problem.
int mopoe_conv_gemm_tc_batched_ex(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                  const mopoe_rows_t* D, const TcStatsReq* stats, void* stream, const TcResReq* res) {
    MOPOE_REQUIRE(nprob >= 1 && nprob <= TCP_MAXP, "conv_gemm_tc_batched: nprob=%d", nprob);
    if (!mopoe_tc_init_state()) MOPOE_FAIL("conv_gemm_tc_batched: tcgen05 path unavailable on this device");
    if (stats && stats->nchunk_out) *stats->nchunk_out = 0;
    TcPersistParams p;
    p.E0 = A[0].E0; p.E1 = A[0].E1; p.E2 = A[0].E2; p.R = A[0].R; p.KW = A[0].KW; p.N = D[0].N;
    for (int i = 1; i < nprob; ++i) {
        MOPOE_REQUIRE(A[i].E0 == p.E0 && A[i].E1 == p.E1 && A[i].E2 == p.E2 && A[i].R == p.R && A[i].KW == p.KW &&
                          D[i].N == p.N && D[i].s0 == D[0].s0 && D[i].s1 == D[0].s1 && D[i].s2 == D[0].s2 &&
                          D[i].d == D[0].d && D[i].d_dtype == D[0].d_dtype,
                      "conv_gemm_tc_batched: problems differ in shape");
    }
    mopoe_tc_tile_split(p.E0, p.E1, 128, p.BX, p.BY, p.NB);
    p.T0 = (p.E0 + p.BX - 1) / p.BX; p.T1 = (p.E1 + p.BY - 1) / p.BY; p.T2 = (p.E2 + p.NB - 1) / p.NB;
    p.BN = mopoe_tc_pick_bn(p.N);
    // wide layers whose best tile is not a multiple of 64 columns (N = 640 -> 160): take 128 so that the TMA-store epilogue
    // (64-column groups) and the fused statistics apply
    if (tma_epi_enabled() && D[0].d_dtype == MOPOE_BF16 && p.N > 256 && p.BN % 64 != 0) {
        // N = 640: choose among the 64-column multiples by a wave-aware cost: tiles are dealt to 148 SMs in whole waves, a
        // tile's k-step costs max(MMA cycles, operand fill at ~55 B/clk/SM) (profiles/r2_ncu_gemm.txt).  M = 4096 rows
        // (32 m-tiles): 128 columns -> 160 tiles = 2 waves; 192 -> 128 tiles = 1 wave, 37 % less time despite 17 % padding.
        const long long mt = (long long)p.T0 * p.T1 * p.T2 * nprob;
        int best = 128;
        double best_cost = 1e30;
        for (int bn : {128, 192, 256}) {
            const long long tiles = mt * ((p.N + bn - 1) / bn);
            const double kstep = fmax(2.0 * bn, (16384.0 + 128.0 * bn) / 55.0);
            const double cost = (double)((tiles + 147) / 148) * kstep;
            if (cost < best_cost * 0.97) { best_cost = cost; best = bn; }
        }
        p.BN = best;
    }
    p.NT = (p.N + p.BN - 1) / p.BN;
    p.nacc = p.NT * p.BN;
    int cols = 2 * p.BN, pc = 32;
    while (pc < cols) pc <<= 1;
    MOPOE_REQUIRE(pc <= 512, "conv_gemm_tc_batched: BN=%d needs %d TMEM columns", p.BN, pc);
    p.tmem_cols = pc;
    p.nprob = nprob;
    p.tiles_per_prob = p.T0 * p.T1 * p.T2 * p.NT;
    p.total_tiles = p.tiles_per_prob * nprob;
    p.s0 = D[0].s0; p.s1 = D[0].s1; p.s2 = D[0].s2; p.d = D[0].d;
    p.d_is_bf16 = D[0].d_dtype == MOPOE_BF16;
    p.bias = bias;
    p.st.ws = nullptr; p.st.mask = nullptr; p.st.mask_mode = MOPOE_MASK_NONE; p.st.rows_per_b = 1;
    p.ksplit = 1; p.kb_per_split = p.R * (p.KW >> 6); p.split_stride = 0;
    if (!g_persist_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_tc_batched: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
        if (const char* e = getenv("MOPOE_GEMM_SMS")) {       // experiment: leave SMs to the kernels of the other branches
            const int v = atoi(e);
            if (v >= 2 && v <= g_num_sms) g_num_sms = v & ~1;
        }
        g_persist_attr_set = true;
    }
    int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
    // TMA-store epilogue: bf16 rows whose strides / origins are 16-byte aligned, 64-column groups
    bool tma_epi = tma_epi_enabled() && p.d_is_bf16 && p.BN % 64 == 0 && p.N % 8 == 0 && p.s0 % 8 == 0 && p.s1 % 8 == 0 &&
                   p.s2 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.d) & 15) == 0 &&
                   (long long)p.E0 * p.E1 * p.E2 < (1ll << 31);
    for (int i = 0; i < nprob && tma_epi; ++i) tma_epi = D[i].d_off % 8 == 0;
    bool fuse_stats = false;
    const bool bnb = res && res->kind == 1;
    // (BatchNorm-backward sums: a Dropout2d mask is keyed by the sample = E2 index, which holds for every phase problem)
    const bool bc_by_sample = bnb && stats && stats->mask_mode == MOPOE_MASK_BC && stats->mask && p.N % 2 == 0 &&
                              stats->rows_per_b == p.E0 * p.E1;
    if (stats && tma_epi && stats->ws && (size_t)grid * 8 * p.N <= stats->ws_doubles &&
        (stats->mask_mode == MOPOE_MASK_NONE || (nprob == 1 && stats->mask && p.N % 2 == 0) || bc_by_sample)) {
        fuse_stats = true;
        p.st.ws = stats->ws;
        p.st.mask = stats->mask_mode == MOPOE_MASK_NONE ? nullptr : stats->mask;
        p.st.mask_mode = stats->mask_mode;
        p.st.rows_per_b = stats->rows_per_b > 0 ? stats->rows_per_b : 1;
    }
    // CTA pairs (cta_group::2) when the TMA-store epilogue applies and there is at least a full wave of single-CTA tiles
    const int m_tiles = p.T0 * p.T1 * p.T2;
    // ... and the tile is wide: measured, N = 128 layers gain nothing from the pair (906 vs 925 TFLOP/s — whatever bounds
    // them, it is not the operand fill), N >= 192 layers gain 10-15 % (1441 -> 1605, 1213 -> 1403 TFLOP/s)
    static int pair_min_bn = -1;
    if (pair_min_bn < 0) {
        const char* e = getenv("MOPOE_GEMM_PAIR_MIN_BN");
        pair_min_bn = e ? atoi(e) : 192;
    }
    // ... and the reduction is long: a 1x1 conv (K = 256-512: 4-8 k-steps per tile) is bound by its epilogue and the
    // HBM, and pays for the pair's per-tile handshakes (x1 M=262144 N=256 K=256: 117 us paired, 93 us single)
    const int nkb = p.R * (p.KW >> 6);
    const bool pair = pair_enabled() && tma_epi && p.BN % 64 == 0 && p.BN >= pair_min_bn && nkb >= 12 &&
                      p.total_tiles >= g_num_sms && g_num_sms % 2 == 0;
    if (pair) {
        p.total_tiles = ((m_tiles + 1) / 2) * nprob * p.NT;          // pair tiles: two consecutive m-tiles x one n-tile
        const int clusters = p.total_tiles < g_num_sms / 2 ? p.total_tiles : g_num_sms / 2;
        grid = 2 * clusters;
        if (stats && stats->ws && (size_t)grid * 8 * p.N > stats->ws_doubles) { fuse_stats = false; p.st.ws = nullptr; }
    }
    // residual epilogue: needs the TMA-store epilogue, whole 64-column groups, 16-byte addressable rows of r, 8-byte aligned
    // mask rows, and (when statistics were asked for) the fused statistics
    bool fuse_res = false;
    if (res) {
        fuse_res = (bnb ? bnb_enabled() : res_enabled()) && tma_epi && p.N % 64 == 0 && res->R && res->mean && res->invstd &&
                   res->gamma && res->beta && (!stats || fuse_stats) && (bnb ? stats != nullptr : (!stats || stats->mask_mode == MOPOE_MASK_NONE));
        for (int i = 0; i < nprob && fuse_res; ++i)
            fuse_res = res->R[i].d_dtype == MOPOE_BF16 && res->R[i].N == p.N && res->R[i].d == res->R[0].d &&
                       res->R[i].s0 == res->R[0].s0 && res->R[i].s1 == res->R[0].s1 && res->R[i].s2 == res->R[0].s2 &&
                       res->R[i].s0 % 8 == 0 && res->R[i].s1 % 8 == 0 && res->R[i].s2 % 8 == 0 && res->R[i].d_off % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(res->R[i].d) & 15) == 0;
        if (fuse_res && !bnb && res->out)
            fuse_res = res->out->dtype == MOPOE_BF16 && res->out->C == p.N && res->out->sW % 8 == 0 && res->out->sH % 8 == 0 &&
                       res->out->sB % 8 == 0 && (reinterpret_cast<uintptr_t>(res->out->ptr) & 15) == 0;
        if (fuse_res && !bnb && res->mask_mode != MOPOE_MASK_NONE)
            fuse_res = res->mask && (reinterpret_cast<uintptr_t>(res->mask) & 7) == 0 &&
                       (res->mask_mode == MOPOE_MASK_BC || res->mask_mode == MOPOE_MASK_ELEM);
    }
    const int stage_bytes = 128 * 128 + p.BN * (pair ? 64 : 128);
    const int hdr_bytes = 16 * 8 + 48 + 64;
    const int extra = tma_epi ? (fuse_res ? 3 : 2) * (int)STG_BYTES + (fuse_stats ? 32 * p.nacc : 0) + (bias ? 4 * p.nacc : 0) +
                                    (fuse_res ? (bnb ? 16 : 8) * p.nacc : 0)
                              : 0;
    int stages = (TCP_SMEM_LIMIT - 1024 - hdr_bytes - extra) / stage_bytes;
    if (stages > 8) stages = 8;
    if (res) {
        if (stages < 3) fuse_res = false;
        if (res->dry_run) return fuse_res ? 0 : 2;
        if (!fuse_res) return 2;
    }
    if (stages < 2 && tma_epi) {                  // (never with the model's shapes) fall back to the register epilogue
        tma_epi = fuse_stats = false;
        p.st.ws = nullptr;
        stages = (TCP_SMEM_LIMIT - 1024 - hdr_bytes) / stage_bytes;
        if (stages > 8) stages = 8;
    }
    MOPOE_REQUIRE(stages >= 2, "conv_gemm_tc_batched: no room for 2 stages (BN=%d)", p.BN);
    p.stages = stages;
    TcMaps maps;
    for (int i = 0; i < nprob; ++i) {
        p.d_off[i] = D[i].d_off;
        const uint64_t dims[5] = {(uint64_t)A[i].KW, (uint64_t)A[i].E0, (uint64_t)A[i].E1, (uint64_t)A[i].R, (uint64_t)A[i].E2};
        const uint64_t str[5] = {1, (uint64_t)A[i].sA0, (uint64_t)A[i].sA1, (uint64_t)A[i].sAr, (uint64_t)A[i].sA2};
        const uint32_t box[5] = {64, (uint32_t)p.BX, (uint32_t)p.BY, 1, (uint32_t)p.NB};
        if (mopoe_tc_encode(&maps.a[i], reinterpret_cast<const bf16*>(A[i].a) + A[i].a_off, 5, dims, str, box, "conv_gemm_tc(A)"))
            return 1;
        const uint64_t K = (uint64_t)A[i].R * A[i].KW;
        const uint64_t dimsb[2] = {K, (uint64_t)p.N};
        const uint64_t strb[2] = {1, K};
        const uint32_t boxb[2] = {64, (uint32_t)(pair ? p.BN / 2 : p.BN)};
        if (mopoe_tc_encode(&maps.b[i], Wp[i], 2, dimsb, strb, boxb, "conv_gemm_tc(B)")) return 1;
        if (tma_epi) {
            const uint64_t dimsd[4] = {(uint64_t)p.N, (uint64_t)p.E0, (uint64_t)p.E1, (uint64_t)p.E2};
            const uint64_t strd[4] = {1, (uint64_t)p.s0, (uint64_t)p.s1, (uint64_t)p.s2};
            const uint32_t boxd[4] = {64, (uint32_t)p.BX, (uint32_t)p.BY, (uint32_t)p.NB};
            if (mopoe_tc_encode(&maps.d[i], reinterpret_cast<const bf16*>(p.d) + D[i].d_off, 4, dimsd, strd, boxd, "conv_gemm_tc(D)"))
                return 1;
        }
    }
    for (int i = nprob; i < TCP_MAXP; ++i) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; }
    if (!tma_epi) memset(maps.d, 0, sizeof(maps.d));
    else for (int i = nprob; i < TCP_MAXP; ++i) maps.d[i] = maps.d[0];
    memset(&p.rs, 0, sizeof(p.rs));
    if (fuse_res) {
        for (int i = 0; i < nprob; ++i) {
            const mopoe_rows_t& R = res->R[i];
            const uint64_t dimsr[4] = {(uint64_t)p.N, (uint64_t)p.E0, (uint64_t)p.E1, (uint64_t)p.E2};
            const uint64_t strr[4] = {1, (uint64_t)R.s0, (uint64_t)R.s1, (uint64_t)R.s2};
            const uint32_t boxr[4] = {64, (uint32_t)p.BX, (uint32_t)p.BY, (uint32_t)p.NB};
            if (mopoe_tc_encode(&maps.r[i], reinterpret_cast<const bf16*>(R.d) + R.d_off, 4, dimsr, strr, boxr, "conv_gemm_tc(R)"))
                return 1;
            // elementwise mask: one byte per element of r, same addressing; Dropout2d mask: [E2 = batch, N]
            p.rs.moff[i] = res->mask_mode == MOPOE_MASK_ELEM ? R.d_off : 0;
        }
        for (int i = nprob; i < TCP_MAXP; ++i) { maps.r[i] = maps.r[0]; p.rs.moff[i] = p.rs.moff[0]; }
        p.rs.mean = res->mean; p.rs.invstd = res->invstd; p.rs.gamma = res->gamma; p.rs.beta = res->beta;
        p.rs.a = res->a; p.rs.b = res->b;
        p.rs.mask = (bnb || res->mask_mode == MOPOE_MASK_NONE) ? nullptr : res->mask;
        p.rs.mask_mode = bnb ? MOPOE_MASK_NONE : res->mask_mode;
        if (!bnb && res->mask_mode == MOPOE_MASK_ELEM) { p.rs.ms0 = res->R[0].s0; p.rs.ms1 = res->R[0].s1; p.rs.ms2 = res->R[0].s2; }
        else { p.rs.ms0 = 0; p.rs.ms1 = 0; p.rs.ms2 = p.N; }
        if (!bnb && res->out && (res->out->ph > 0 || res->out->pw > 0)) {
            const mopoe_view_t* o = res->out;
            p.rs.zb = reinterpret_cast<bf16*>(o->ptr);
            p.rs.zB = o->B; p.rs.zH = o->H; p.rs.zW = o->W; p.rs.zph = o->ph; p.rs.zpw = o->pw;
            p.rs.zsB = o->sB; p.rs.zsH = o->sH; p.rs.zsW = o->sW;
        }
    } else {
        memset(maps.r, 0, sizeof(maps.r));
    }
    const int smem = 1024 + stages * stage_bytes + extra + hdr_bytes;
    if (fuse_res && !pair) {
        if (bnb) conv_gemm_tc_persist_kernel<true, false, false, true><<<grid, TCP_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
        else conv_gemm_tc_persist_kernel<true, false, true><<<grid, TCP_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    } else if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(TCP_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = !fuse_res ? cudaLaunchKernelEx(&cfg, conv_gemm_tc_persist_kernel<true, true>, maps, p)
                        : bnb     ? cudaLaunchKernelEx(&cfg, conv_gemm_tc_persist_kernel<true, true, false, true>, maps, p)
                                  : cudaLaunchKernelEx(&cfg, conv_gemm_tc_persist_kernel<true, true, true>, maps, p);
        if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_tc_pair: launch: %s", cudaGetErrorString(e));
    } else if (tma_epi)
        conv_gemm_tc_persist_kernel<true, false><<<grid, TCP_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    else
        conv_gemm_tc_persist_kernel<false, false><<<grid, TCP_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    MOPOE_CHECK_LAUNCH("conv_gemm_tc_persist");
    if (fuse_stats && stats->nchunk_out) *stats->nchunk_out = grid * 4;
    return 0;
}
int mopoe_conv_gemm_tc_batched(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                               const mopoe_rows_t* D, void* stream) {
    return mopoe_conv_gemm_tc_batched_ex(nprob, A, Wp, bias, D, nullptr, stream, nullptr);
}

// =====================================================================================================================
// Split-K for weight-bound GEMMs: few output tiles, long reduction (the 4x4 -> 1x1 convs and their gradients at the
// bottom of the image stacks: M = batch rows, N = 640, K = 8192-10240 — 10 CTAs streaming 10-13 MB of weights at the fill
// rate of 10 SMs: 40-58 us).  Work items = (tile, split): every SM streams 1/S of the weights into an fp32 partial tile;
// a finish kernel sums the S partials in a fixed order, adds the bias and writes the output rows.
// =====================================================================================================================
struct SplitKPlan {
    int S, kbps, tiles, nkb, BN;
    long long rows;
};
static bool splitk_plan(const mopoe_window_t* A, const mopoe_rows_t* D, SplitKPlan& q) {
    static int enabled = -1;
    if (enabled < 0) {
        // opt-in: measured on the bench step the split halves these launches (40-58 -> 20 us) and changes nothing in the
        // step time — on 10 SMs they were running UNDER the other modality branches' kernels anyway
        const char* e = getenv("MOPOE_GEMM_SPLITK");
        enabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (!enabled || !mopoe_tc_init_state()) return false;
    int BX, BY, NB;
    mopoe_tc_tile_split(A->E0, A->E1, 128, BX, BY, NB);
    const int T0 = (A->E0 + BX - 1) / BX, T1 = (A->E1 + BY - 1) / BY, T2 = (A->E2 + NB - 1) / NB;
    q.BN = mopoe_tc_pick_bn(D->N);
    const int NT = (D->N + q.BN - 1) / q.BN;
    q.tiles = T0 * T1 * T2 * NT;
    q.nkb = A->R * (A->KW >> 6);
    q.rows = (long long)A->E0 * A->E1 * A->E2;
    if (q.tiles > 37 || q.nkb < 32 || D->N % 4 != 0) return false;
    int S = q.nkb / 8;
    if (S > 148 / q.tiles) S = 148 / q.tiles;
    if (S > 32) S = 32;
    if (S < 2) return false;
    q.kbps = (q.nkb + S - 1) / S;
    q.S = (q.nkb + q.kbps - 1) / q.kbps;
    return q.S >= 2;
}
size_t mopoe_conv_gemm_tc_splitk_ws(const mopoe_window_t* A, const mopoe_rows_t* D) {
    SplitKPlan q;
    if (!splitk_plan(A, D, q)) return 0;
    return (size_t)q.S * q.rows * D->N * sizeof(float);
}

__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ ws, int S, long long rows, int N, int E0, int E1,
                                                            const float* __restrict__ bias, void* d, int d_is_bf16, long long d_off,
                                                            long long s0, long long s1, long long s2) {
    const int n4 = N >> 2;
    const long long total = rows * n4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long m = i / n4;
        const int n = (int)(i - m * n4) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < S; ++z) {                                   // fixed order: deterministic
            const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + ((long long)z * rows + m) * N + n));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (bias) { a.x += bias[n]; a.y += bias[n + 1]; a.z += bias[n + 2]; a.w += bias[n + 3]; }
        const long long m0 = m % E0, m1 = (m / E0) % E1, m2 = m / ((long long)E0 * E1);
        const long long o = d_off + m0 * s0 + m1 * s1 + m2 * s2 + n;
        if (d_is_bf16) {
            bf16* dp = reinterpret_cast<bf16*>(d) + o;
            dp[0] = __float2bfloat16_rn(a.x); dp[1] = __float2bfloat16_rn(a.y);
            dp[2] = __float2bfloat16_rn(a.z); dp[3] = __float2bfloat16_rn(a.w);
        } else {
            float* dp = reinterpret_cast<float*>(d) + o;
            dp[0] = a.x; dp[1] = a.y; dp[2] = a.z; dp[3] = a.w;
        }
    }
}

int mopoe_conv_gemm_tc_splitk(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* ws,
                              size_t ws_bytes, void* stream) {
    SplitKPlan q;
    MOPOE_REQUIRE(splitk_plan(A, D, q), "conv_gemm_splitk: problem not eligible (ask mopoe_conv_gemm_splitk_ws first)");
    MOPOE_REQUIRE(ws && ws_bytes >= (size_t)q.S * q.rows * D->N * sizeof(float), "conv_gemm_splitk: workspace too small");
    TcPersistParams p;
    p.E0 = A->E0; p.E1 = A->E1; p.E2 = A->E2; p.R = A->R; p.KW = A->KW; p.N = D->N;
    mopoe_tc_tile_split(p.E0, p.E1, 128, p.BX, p.BY, p.NB);
    p.T0 = (p.E0 + p.BX - 1) / p.BX; p.T1 = (p.E1 + p.BY - 1) / p.BY; p.T2 = (p.E2 + p.NB - 1) / p.NB;
    p.BN = q.BN;
    p.NT = (p.N + p.BN - 1) / p.BN;
    p.nacc = p.NT * p.BN;
    int pc = 32;
    while (pc < 2 * p.BN) pc <<= 1;
    MOPOE_REQUIRE(pc <= 512, "conv_gemm_splitk: BN=%d needs %d TMEM columns", p.BN, pc);
    p.tmem_cols = pc;
    p.nprob = 1;
    p.tiles_per_prob = q.tiles;
    p.total_tiles = q.tiles * q.S;
    p.ksplit = q.S; p.kb_per_split = q.kbps;
    p.split_stride = q.rows * p.N;
    p.d = ws; p.d_is_bf16 = 0; p.bias = nullptr;
    p.d_off[0] = 0; p.d_off[1] = p.d_off[2] = p.d_off[3] = 0;
    p.s0 = p.N; p.s1 = (long long)p.E0 * p.N; p.s2 = (long long)p.E0 * p.E1 * p.N;
    p.st.ws = nullptr; p.st.mask = nullptr; p.st.mask_mode = MOPOE_MASK_NONE; p.st.rows_per_b = 1;
    if (!g_persist_attr_set) {             // (first GEMM of the process: let the regular entry point set the attributes)
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_LIMIT);
        if (e != cudaSuccess) MOPOE_FAIL("conv_gemm_splitk: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int stage_bytes = 128 * 128 + p.BN * 128;
    const int hdr_bytes = 16 * 8 + 48 + 64;
    int stages = (TCP_SMEM_LIMIT - 1024 - hdr_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    MOPOE_REQUIRE(stages >= 2, "conv_gemm_splitk: no room for 2 stages (BN=%d)", p.BN);
    p.stages = stages;
    TcMaps maps;
    {
        const uint64_t dims[5] = {(uint64_t)A->KW, (uint64_t)A->E0, (uint64_t)A->E1, (uint64_t)A->R, (uint64_t)A->E2};
        const uint64_t str[5] = {1, (uint64_t)A->sA0, (uint64_t)A->sA1, (uint64_t)A->sAr, (uint64_t)A->sA2};
        const uint32_t box[5] = {64, (uint32_t)p.BX, (uint32_t)p.BY, 1, (uint32_t)p.NB};
        if (mopoe_tc_encode(&maps.a[0], reinterpret_cast<const bf16*>(A->a) + A->a_off, 5, dims, str, box, "conv_gemm_splitk(A)")) return 1;
        const uint64_t K = (uint64_t)A->R * A->KW;
        const uint64_t dimsb[2] = {K, (uint64_t)p.N};
        const uint64_t strb[2] = {1, K};
        const uint32_t boxb[2] = {64, (uint32_t)p.BN};
        if (mopoe_tc_encode(&maps.b[0], Wp, 2, dimsb, strb, boxb, "conv_gemm_splitk(B)")) return 1;
    }
    for (int i = 1; i < TCP_MAXP; ++i) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; }
    memset(maps.d, 0, sizeof(maps.d));
    memset(maps.r, 0, sizeof(maps.r));
    memset(&p.rs, 0, sizeof(p.rs));
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    const int smem = 1024 + stages * stage_bytes + hdr_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    conv_gemm_tc_persist_kernel<false, false><<<grid, TCP_THREADS, smem, st>>>(maps, p);
    MOPOE_CHECK_LAUNCH("conv_gemm_splitk");
    const long long total = q.rows * (p.N >> 2);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)ws, q.S, q.rows, p.N, p.E0, p.E1, bias, D->d,
                                                          D->d_dtype == MOPOE_BF16, D->d_off, D->s0, D->s1, D->s2);
    MOPOE_CHECK_LAUNCH("splitk_finish");
    return 0;
}
