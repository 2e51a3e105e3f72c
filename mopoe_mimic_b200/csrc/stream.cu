// stream.cu — the HBM-bound passes of a residual block (BatchNorm statistics / apply / backward sums / backward apply,
// residual combine) as PERSISTENT kernels that stage their operands through shared memory with bulk asynchronous copies.
//
// Why (ncu, profiles/r2_ncu_elementwise.txt): the register-staged kernels of elementwise.cu keep 16-64 bytes per
// thread in flight and spend 45-65 % of their cycles stalled on those loads: 4.3-5.3 TB/s where HBM3e delivers 6.5+.
// Here ONE producer thread per CTA issues `cp.async.bulk` copies of contiguous row chunks (up to 4 KB per operand) into a
// ring of shared-memory stages — 100+ KB in flight per SM at zero register cost — and 8 consumer warps turn the staged
// chunks into results.  Because a chunk is a whole number of pixels and the grid is persistent, a consumer thread owns
// ONE channel octet for its whole life: the per-channel coefficients are loaded once per CTA, there is no per-element
// index arithmetic at all (the producer decodes one (batch, row, chunk) triple per 4 KB), and the zero border of the
// outputs is written in a short loop at the end.
//
// Data layout: an operand is a channels-last view (mopoe_view_t, interior-origin pointer, element strides sB / sH,
// sW == C): the W*C interior elements of an image row are contiguous; rows are separated by the border.  A work item is
// (image row, chunk of PP pixels); items are dealt to the CTAs round-robin, so that at any moment the grid streams one
// compact window of every operand.
#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int VEC = 8;
constexpr int ST_CONSUMERS = 256;
constexpr int ST_THREADS = ST_CONSUMERS + 32;      // + one producer warp
constexpr int ST_MAX_OPS = 4;
constexpr int ST_HDR = 256;                        // barriers
constexpr int ST_MAX_STAGES = 12;
// sub-chunks (of <= 256 octets) per item, by number of staged operands: what sets the pace of the 1- and 2-operand
// passes is the instruction stream, not HBM (~100 instructions per 16-byte octet with one octet per thread and item:
// the whole chip issues 268 MB worth of them in ~90 us) — so an item carries several octets per thread and the per-item
// work (decode, mask fetch, barrier, release) is paid once
#ifndef ST_K1
#define ST_K1 4
#endif
#ifndef ST_K2
#define ST_K2 2
#endif
#ifndef ST_K3
#define ST_K3 2
#endif
#ifndef ST_K4
#define ST_K4 1
#endif

// fp32 storage (validation mode) holds twice the registers per staged octet: half the sub-chunks per item
template <typename T>
constexpr int ks_for(int k) { return sizeof(T) == 4 ? (k > 1 ? k / 2 : 1) : k; }

struct SOp {
    const void* p;
    int sB, sH;
};
struct SOut {
    void* p;
    int sB, sH, ph, pw;
};
struct StreamGeo {
    int B, H, W, C, CV;
    int PP, CH, IO, KC;            // pixels / octets per sub-chunk (CH = PP * CV <= 256), octets per item, items per row
    unsigned items;
    int stages;
    int reverse;                   // walk the items from the END of the tensors (see st_reverse())
    unsigned op_bytes;             // bytes of one operand slot of a stage
    FastDiv fKC, fH, fCV;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

// one poll of the phase parity (no spin)
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

template <typename T>
struct Oct;                       // 8 consecutive channels, packed as loaded
template <>
struct Oct<bf16> {
    uint4 v;
    __device__ __forceinline__ void lds(const uint8_t* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};
template <>
struct Oct<float> {
    float4 a, b;
    __device__ __forceinline__ void lds(const uint8_t* p) {
        a = reinterpret_cast<const float4*>(p)[0];
        b = reinterpret_cast<const float4*>(p)[1];
    }
    __device__ __forceinline__ void unpack(float (&o)[8]) const {
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
};
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&o)[8]) {
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
__device__ __forceinline__ void ld8f(const float* p, float (&o)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void mask8(const uint2& t, float (&o)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o[i] = ((t.x >> (8 * i)) & 0xffu) ? 2.f : 0.f;
        o[4 + i] = ((t.y >> (8 * i)) & 0xffu) ? 2.f : 0.f;
    }
}

// y = v * sc + sh == gamma * (v - mean) * invstd + beta, evaluated with ONE instruction sequence everywhere (forward apply,
// backward gate recompute): sc = invstd * gamma, sh = fma(-mean, sc, beta), y = fma(v, sc, sh) — explicit fma intrinsics, so
// that no contraction choice of the compiler can make two kernels disagree on a sign.
__device__ __forceinline__ void affine8(const float (&mu)[8], const float (&is)[8], const float (&ga)[8], const float (&be)[8],
                                        float (&sc)[8], float (&sh)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __fmul_rn(is[i], ga[i]);
        sh[i] = __fmaf_rn(-mu[i], sc[i], be[i]);
    }
}
// the stored activation relu(y) is > 0  <=>  y rounds to a non-zero value of the storage type
template <typename T>
__device__ __forceinline__ bool gate_open(float y) {
    return sizeof(T) == 2 ? y > 0x1p-134f : y > 0.f;       // bf16: half of the smallest subnormal rounds to zero (ties to even)
}

struct Item {
    int b, h, k;          // batch, image row, chunk
    int noct;             // octets in this item
    int eoff;             // element offset of the item inside its row (k * IO * 8)
};
__device__ __forceinline__ Item decode_item(const StreamGeo& g, unsigned item) {
    unsigned row, k, b, h;
    if (g.reverse) item = g.items - 1u - item;
    g.fKC.divmod(item, row, k);
    g.fH.divmod(row, b, h);
    Item it;
    it.b = (int)b; it.h = (int)h; it.k = (int)k;
    const int rem = g.W * g.CV - (int)k * g.IO;
    it.noct = rem < g.IO ? rem : g.IO;
    it.eoff = (int)k * g.IO * VEC;
    return it;
}

// zero the border of a bordered output (all consumer threads of all CTAs, grid-stride over border octets)
template <typename T>
__device__ __forceinline__ void zero_border(const SOut& o, const StreamGeo& g, int tid) {
    if (o.ph == 0 && o.pw == 0) return;
    const int Ws = g.W + 2 * o.pw;
    const int per = 2 * o.ph * Ws + 2 * o.pw * g.H;
    const long long total = (long long)g.B * per * g.CV;
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * ST_CONSUMERS + tid; i < total; i += (long long)gridDim.x * ST_CONSUMERS) {
        const int cv = (int)(i % g.CV);
        const long long q = i / g.CV;
        const int k = (int)(q % per);
        const int b = (int)(q / per);
        int hs, ws;
        if (k < o.ph * Ws) { hs = k / Ws; ws = k - hs * Ws; }
        else if (k < 2 * o.ph * Ws) { const int k2 = k - o.ph * Ws; hs = o.ph + g.H + k2 / Ws; ws = k2 % Ws; }
        else { const int k2 = k - 2 * o.ph * Ws; hs = o.ph + k2 / (2 * o.pw); const int j = k2 % (2 * o.pw); ws = j < o.pw ? j : g.W + j; }
        st8<T>(reinterpret_cast<T*>(o.p) + ((long long)b * o.sB + (long long)(hs - o.ph) * o.sH + (long long)(ws - o.pw) * g.C + cv * VEC), z);
    }
}

// ---- the pipeline skeleton --------------------------------------------------------------------------------------------
// An item is up to KS sub-chunks of CH octets (CH = PP pixels x C/8 octets <= 256) of one image row; consumer thread t
// owns octets t, t + CH, ... of the item — always the same 8 channels.  F provides:
//   KS (compile time), struct Pre (what prefetch hands to item),
//   void init(int c)                                   once; c = first channel of the thread's octets
//   Pre  prefetch(const Item&, int tid)                per item, BEFORE the wait on the stage (mask bytes from global)
//   void item(const Item&, int tid, const Oct<T> (&raw)[KS][NOPS], const Pre&)      per item
template <typename T, int NOPS, class F>
__device__ __forceinline__ void stream_pipeline(const StreamGeo& g, const SOp (&ops)[NOPS], uint8_t* smem, F& f) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t full0 = sbase, empty0 = sbase + 8u * ST_MAX_STAGES;
    uint8_t* const data = smem + ST_HDR;
    const uint32_t sdata = sbase + ST_HDR;
    const uint32_t stage_bytes = g.op_bytes * NOPS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, ST_CONSUMERS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == ST_CONSUMERS / 32) {
        // ===== producers: lane l < stages OWNS stage l and feeds it with the CTA's items l, l + stages, ...
        //  * a lane only ever waits for the release of the fill IT issued one ring revolution earlier, so the parity test of
        //    the barrier cannot alias (a lane polling two revolutions ahead of a fresh barrier would pass at once);
        //  * no lane SPINS: the compiler reconverges the warp behind a spin loop, and the lanes whose copies the spinning
        //    lane is (indirectly) waiting for would never get there.  Every lane polls once per trip of a warp-uniform loop.
        const unsigned stage = (unsigned)lane;
        unsigned round = 0;
        bool more = lane < g.stages && blockIdx.x + (unsigned long long)stage * gridDim.x < g.items;
        while (__any_sync(0xffffffffu, more)) {
            if (more && mbar_try(empty0 + 8 * stage, (round & 1u) ^ 1u)) {
                const unsigned n = round * (unsigned)g.stages + stage;
                const Item it = decode_item(g, blockIdx.x + n * gridDim.x);
                const uint32_t bytes = (uint32_t)it.noct * VEC * sizeof(T);
                const uint32_t bar = full0 + 8 * stage;
                mbar_expect_tx(bar, bytes * NOPS);
#pragma unroll
                for (int o = 0; o < NOPS; ++o)
                    bulk_load_1d(sdata + stage * stage_bytes + o * g.op_bytes,
                                 reinterpret_cast<const T*>(ops[o].p) + ((long long)it.b * ops[o].sB + (long long)it.h * ops[o].sH + it.eoff),
                                 bytes, bar);
                ++round;
                more = blockIdx.x + (unsigned long long)(n + g.stages) * gridDim.x < g.items;
            }
        }
    } else {
        constexpr int KS = F::KS;
        unsigned pq, cv;
        g.fCV.divmod((unsigned)tid, pq, cv);
        f.init((int)cv * VEC);
        int stage = 0;
        uint32_t phase = 0;
        for (unsigned item = blockIdx.x; item < g.items; item += gridDim.x) {
            const Item it = decode_item(g, item);
            const typename F::Pre pre = f.prefetch(it, tid);
            mbar_wait(full0 + 8 * stage, phase);
            Oct<T> raw[KS][NOPS];
            const uint8_t* sp = data + stage * stage_bytes + tid * (VEC * sizeof(T));
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                if (tid < g.CH && tid + j * g.CH < it.noct) {
#pragma unroll
                    for (int o = 0; o < NOPS; ++o) raw[j][o].lds(sp + o * g.op_bytes + j * g.CH * (VEC * sizeof(T)));
                }
            }
            f.item(it, tid, raw, pre);
            // release the stage only AFTER the staged values have been consumed: an arrive issued right behind the
            // ld.shared instructions is not ordered behind their completion (observed: stale-stage reads once the ring wraps)
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * stage);
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    }
}

// ---- bodies -----------------------------------------------------------------------------------------------------------
struct MaskRef {
    const uint8_t* m;
    int mode;
};
// keep-mask bytes: a Dropout2d mask [B, C] is ONE value per (sample, channel) — constant over an item; an elementwise mask
// [B, L, C] (1-D text) is fetched per octet.  All loads are issued before the wait on the stage.
template <int KS>
struct MaskPre {
    uint2 m[KS];
};
template <int KS>
__device__ __forceinline__ MaskPre<KS> mask_fetch(const MaskRef& mk, const StreamGeo& g, const Item& it, int tid, int c) {
    MaskPre<KS> p;
#pragma unroll
    for (int j = 0; j < KS; ++j) p.m[j] = make_uint2(0, 0);
    if (mk.mode == MOPOE_MASK_BC) {
        p.m[0] = *reinterpret_cast<const uint2*>(mk.m + (it.b * g.C + c));
    } else if (mk.mode == MOPOE_MASK_ELEM) {
        const long long el = ((long long)(it.b * g.H + it.h) * g.W) * g.C + it.eoff + tid * VEC;
#pragma unroll
        for (int j = 0; j < KS; ++j)
            if (tid < g.CH && tid + j * g.CH < it.noct) p.m[j] = *reinterpret_cast<const uint2*>(mk.m + el + (long long)j * g.CH * VEC);
    }
    return p;
}
// multipliers (2 keep / 0 drop / 1 no mask) of sub-chunk j
template <int KS>
__device__ __forceinline__ void mask_mul(const MaskRef& mk, const MaskPre<KS>& p, int j, float (&m)[8]) {
    if (mk.mode == MOPOE_MASK_NONE) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) m[i] = 1.f;
    } else {
        mask8(p.m[mk.mode == MOPOE_MASK_BC ? 0 : j], m);
    }
}

template <typename T>
struct ApplyBody {            // out = act(gamma * (x*2mask - mean) * invstd + beta)
    static constexpr int KS = ks_for<T>(ST_K1);
    typedef MaskPre<KS> Pre;
    StreamGeo g;
    MaskRef mk;
    const float *mean, *invstd, *gamma, *beta;
    int relu;
    SOut out;
    float sc[VEC], sh[VEC];
    int c;
    __device__ __forceinline__ void init(int c_) {
        c = c_;
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(beta + c, be);
        affine8(mu, is, ga, be, sc, sh);           // same sequence as bn_affine (elementwise.cu)
    }
    __device__ __forceinline__ Pre prefetch(const Item& it, int tid) const { return mask_fetch<KS>(mk, g, it, tid, c); }
    __device__ __forceinline__ void item(const Item& it, int tid, const Oct<T> (&raw)[KS][1], const Pre& pre) {
        T* const orow = reinterpret_cast<T*>(out.p) + ((long long)it.b * out.sB + (long long)it.h * out.sH + it.eoff + tid * VEC);
        // a Dropout2d mask is constant over the item: fold it into the scale.  (x * m) * sc == x * (m * sc) exactly for
        // m in {0, 2}, so the result is bit-identical to masking x first.
        float scm[VEC];
        const bool fold = mk.mode != MOPOE_MASK_ELEM;
        {
            float m[VEC];
            mask_mul<KS>(mk, pre, 0, m);
#pragma unroll
            for (int i = 0; i < VEC; ++i) scm[i] = fold ? sc[i] * m[i] : sc[i];
        }
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (!(tid < g.CH && tid + j * g.CH < it.noct)) continue;
            float xv[VEC], o[VEC];
            raw[j][0].unpack(xv);
            if (!fold) {
                float m[VEC];
                mask_mul<KS>(mk, pre, j, m);
#pragma unroll
                for (int i = 0; i < VEC; ++i) xv[i] *= m[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float y = __fmaf_rn(xv[i], scm[i], sh[i]);
                o[i] = (relu && y < 0.f) ? 0.f : y;
            }
            st8<T>(orow + j * g.CH * VEC, o);
        }
    }
};

template <typename T, bool STATS>
struct CombineBody {          // out = a * BN(r) + b * (c * 2mask)   [+ per-channel sum / sum of squares of the STORED out]
    static constexpr int KS = ks_for<T>(ST_K2);
    typedef MaskPre<KS> Pre;
    StreamGeo g;
    MaskRef mk;
    const float *mean, *invstd, *gamma, *beta;
    float a, bcoef;
    SOut out;
    float sc[VEC], sh[VEC];
    float f0[VEC], f1[VEC];
    int c;
    __device__ __forceinline__ void init(int c_) {
        c = c_;
        float mu[VEC], is[VEC], ga[VEC], be[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(beta + c, be);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            sc[i] = a * is[i] * ga[i];
            sh[i] = a * be[i] - mu[i] * sc[i];
            f0[i] = f1[i] = 0.f;
        }
    }
    __device__ __forceinline__ Pre prefetch(const Item& it, int tid) const { return mask_fetch<KS>(mk, g, it, tid, c); }
    __device__ __forceinline__ void item(const Item& it, int tid, const Oct<T> (&raw)[KS][2], const Pre& pre) {
        T* const orow = reinterpret_cast<T*>(out.p) + ((long long)it.b * out.sB + (long long)it.h * out.sH + it.eoff + tid * VEC);
        float m[VEC];
        mask_mul<KS>(mk, pre, 0, m);
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (!(tid < g.CH && tid + j * g.CH < it.noct)) continue;
            float rv[VEC], cv[VEC], o[VEC];
            raw[j][0].unpack(rv);
            raw[j][1].unpack(cv);
            if (mk.mode == MOPOE_MASK_ELEM) mask_mul<KS>(mk, pre, j, m);
            if (mk.mode == MOPOE_MASK_NONE) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = fmaf(rv[i], sc[i], sh[i]) + bcoef * cv[i];
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = fmaf(rv[i], sc[i], sh[i]) + bcoef * (cv[i] * m[i]);
            }
            st8<T>(orow + j * g.CH * VEC, o);
            if (STATS) {
                // the statistics the NEXT block's bn1 needs are those of the values it will read: rounded to the storage type
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float v = sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(o[i])) : o[i];
                    f0[i] += v;
                    f1[i] = fmaf(v, v, f1[i]);
                }
            }
        }
    }
};

// out  = gamma*invstd*(g - sums_g/cnt - xhat*sums_gx/cnt) * 2mask + addend,  g = gscale * dy * [gate > 0]
// out2 = scale2 * dy * 2mask2
// operand order: dy, x, [gate], [addend]
// RECOMP (with GATE): the gate is not a staged operand but recomputed from x (see ReduceRecompBody)
template <typename T, bool GATE, bool ADD, bool OUT2, bool RECOMP = false>
struct BwdBody {
    static constexpr int NOPS = 2 + (GATE && !RECOMP ? 1 : 0) + (ADD ? 1 : 0);
    static constexpr int KS = ks_for<T>(NOPS >= 4 ? ST_K4 : (NOPS == 3 ? ST_K3 : ST_K2));
    struct Pre {
        MaskPre<KS> a, b;
    };
    StreamGeo g;
    MaskRef mk, mk2;
    const float *mean, *invstd, *gamma, *sums, *gate_beta;
    float gscale, inv_cnt, scale2;
    SOut out, out2;
    float k1[VEC], ca[VEC], cb[VEC], gsh[VEC];       // (the gate's scale invstd * gamma IS k1)
    int c;
    __device__ __forceinline__ void init(int c_) {
        c = c_;
        float mu[VEC], is[VEC], ga[VEC], sg[VEC], sgx[VEC];
        ld8f(mean + c, mu); ld8f(invstd + c, is); ld8f(gamma + c, ga); ld8f(sums + c, sg); ld8f(sums + g.C + c, sgx);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {          // same sequence as the kernels of elementwise.cu
            k1[i] = __fmul_rn(is[i], ga[i]);                 // == affine8's sc
            const float mg = sg[i] * inv_cnt, mgx = sgx[i] * inv_cnt;
            ca[i] = -k1[i] * is[i] * mgx;
            cb[i] = k1[i] * (mu[i] * is[i] * mgx - mg);
        }
        if (GATE && RECOMP) {
            float be[VEC];
            ld8f(gate_beta + c, be);
#pragma unroll
            for (int i = 0; i < VEC; ++i) gsh[i] = __fmaf_rn(-mu[i], k1[i], be[i]);       // == affine8's sh
        }
    }
    __device__ __forceinline__ Pre prefetch(const Item& it, int tid) const {
        Pre p;
        p.a = mask_fetch<KS>(mk, g, it, tid, c);
        if (OUT2) p.b = mask_fetch<KS>(mk2, g, it, tid, c);
        return p;
    }
    __device__ __forceinline__ void item(const Item& it, int tid, const Oct<T> (&raw)[KS][NOPS], const Pre& pre) {
        const long long roff = (long long)it.eoff + tid * VEC;
        T* const orow = reinterpret_cast<T*>(out.p) + ((long long)it.b * out.sB + (long long)it.h * out.sH + roff);
        T* const o2row = reinterpret_cast<T*>(out2.p) + ((long long)it.b * out2.sB + (long long)it.h * out2.sH + roff);
        float m[VEC], m2[VEC];
        mask_mul<KS>(mk, pre.a, 0, m);
        if (OUT2) mask_mul<KS>(mk2, pre.b, 0, m2);
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (!(tid < g.CH && tid + j * g.CH < it.noct)) continue;
            float gv[VEC], v[VEC], o[VEC];
            raw[j][0].unpack(gv);
            raw[j][1].unpack(v);
            if (OUT2) {
                float o2[VEC];
                if (mk2.mode == MOPOE_MASK_ELEM) mask_mul<KS>(mk2, pre.b, j, m2);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o2[i] = scale2 * gv[i] * m2[i];
                st8<T>(o2row + j * g.CH * VEC, o2);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) gv[i] *= gscale;
            if (GATE && !RECOMP) {
                float gt[VEC];
                raw[j][2].unpack(gt);
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (!(gt[i] > 0.f)) gv[i] = 0.f;
            }
            if (mk.mode == MOPOE_MASK_ELEM) mask_mul<KS>(mk, pre.a, j, m);
            if (GATE && RECOMP) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float vm = mk.mode != MOPOE_MASK_NONE ? v[i] * m[i] : v[i];
                    if (!gate_open<T>(__fmaf_rn(vm, k1[i], gsh[i]))) gv[i] = 0.f;
                }
            }
            if (mk.mode != MOPOE_MASK_NONE) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], gv[i], fmaf(ca[i], v[i] * m[i], cb[i])) * m[i];
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = fmaf(k1[i], gv[i], fmaf(ca[i], v[i], cb[i]));
            }
            if (ADD) {
                float ad[VEC];
                raw[j][2 + (GATE && !RECOMP ? 1 : 0)].unpack(ad);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] += ad[i];
            }
            st8<T>(orow + j * g.CH * VEC, o);
        }
    }
};

// per-channel sums.  MODE 0: sum v, sum v^2 (v = x*2mask);  MODE 1: sum g, sum g*xhat (g = gscale*dy*[gate>0]);
// MODE 2: sum x.   operand order: x, [dy], [gate]
template <typename T, int MODE, bool GATE>
struct ReduceBody {
    static constexpr int NOPS = MODE == 1 ? (GATE ? 3 : 2) : 1;
    static constexpr int KS = ks_for<T>(NOPS == 1 ? ST_K1 : (NOPS == 2 ? ST_K2 : ST_K3));
    typedef MaskPre<KS> Pre;
    StreamGeo g;
    MaskRef mk;
    const float *mean, *invstd;
    float gscale;
    float f0[VEC], f1[VEC], mu[VEC], is[VEC];
    int c;
    __device__ __forceinline__ void init(int c_) {
        c = c_;
#pragma unroll
        for (int i = 0; i < VEC; ++i) f0[i] = f1[i] = 0.f;
        if (MODE == 1) {
            ld8f(mean + c, mu);
            ld8f(invstd + c, is);
        }
    }
    __device__ __forceinline__ Pre prefetch(const Item& it, int tid) const { return mask_fetch<KS>(mk, g, it, tid, c); }
    __device__ __forceinline__ void item(const Item& it, int tid, const Oct<T> (&raw)[KS][NOPS], const Pre& pre) {
        float m[VEC];
        mask_mul<KS>(mk, pre, 0, m);
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (!(tid < g.CH && tid + j * g.CH < it.noct)) continue;
            float xv[VEC];
            raw[j][0].unpack(xv);
            if (mk.mode == MOPOE_MASK_ELEM) mask_mul<KS>(mk, pre, j, m);
            if (MODE == 0) {
                if (mk.mode == MOPOE_MASK_NONE) {            // (the pass is instruction-bound: no multiply by 1)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        f0[i] += xv[i];
                        f1[i] = fmaf(xv[i], xv[i], f1[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float v = xv[i] * m[i];
                        f0[i] += v;
                        f1[i] = fmaf(v, v, f1[i]);
                    }
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) f0[i] += xv[i];
            } else {
                float gv[VEC], gt[VEC];
                raw[j][1].unpack(gv);
                if (GATE) raw[j][NOPS - 1].unpack(gt);
                if (mk.mode == MOPOE_MASK_NONE) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float gg = gscale * gv[i];
                        if (GATE && !(gt[i] > 0.f)) gg = 0.f;
                        const float xh = (xv[i] - mu[i]) * is[i];
                        f0[i] += gg;
                        f1[i] = fmaf(gg, xh, f1[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float gg = gscale * gv[i];
                        if (GATE && !(gt[i] > 0.f)) gg = 0.f;
                        const float xh = (xv[i] * m[i] - mu[i]) * is[i];
                        f0[i] += gg;
                        f1[i] = fmaf(gg, xh, f1[i]);
                    }
                }
            }
        }
    }
};

// BN-backward sums with the ReLU gate RECOMPUTED from the BatchNorm's input instead of read from the saved activation:
// gate = [relu(y) stored > 0], y = fma(x * 2mask, sc, sh) — the very instruction sequence of the forward apply pass
// (affine8 / gate_open), so the decision is bit-identical to the stored activation's sign.  Two staged operands (dy, x)
// instead of three (dy, x, a): the largest pass of the backward (72 launches) moves a third less, and stays exact.
// operand order: dy, x
template <typename T>
struct ReduceRecompBody {
    static constexpr int NOPS = 2;
    static constexpr int KS = ks_for<T>(ST_K2);
    typedef MaskPre<KS> Pre;
    StreamGeo g;
    MaskRef mk;
    const float *mean, *invstd, *gamma, *beta;
    float gscale;
    float f0[VEC], f1[VEC], sc[VEC], sh[VEC], mu[VEC], is[VEC];
    int c;
    __device__ __forceinline__ void init(int c_) {
        c = c_;
        float ga[VEC], be[VEC];
        ld8f(gamma + c, ga);
        ld8f(beta + c, be);
        ld8f(mean + c, mu);
        ld8f(invstd + c, is);
        affine8(mu, is, ga, be, sc, sh);
#pragma unroll
        for (int i = 0; i < VEC; ++i) f0[i] = f1[i] = 0.f;
    }
    __device__ __forceinline__ Pre prefetch(const Item& it, int tid) const { return mask_fetch<KS>(mk, g, it, tid, c); }
    __device__ __forceinline__ void item(const Item& it, int tid, const Oct<T> (&raw)[KS][2], const Pre& pre) {
        float m[VEC];
        mask_mul<KS>(mk, pre, 0, m);
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            if (!(tid < g.CH && tid + j * g.CH < it.noct)) continue;
            float gv[VEC], xv[VEC];
            raw[j][0].unpack(gv);
            raw[j][1].unpack(xv);
            if (mk.mode == MOPOE_MASK_ELEM) mask_mul<KS>(mk, pre, j, m);
            if (mk.mode != MOPOE_MASK_NONE) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) xv[i] *= m[i];          // exact (m in {0, 2}): the forward's masked input
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float gg = gscale * gv[i];
                if (!gate_open<T>(__fmaf_rn(xv[i], sc[i], sh[i]))) gg = 0.f;
                const float xh = (xv[i] - mu[i]) * is[i];
                f0[i] += gg;
                f1[i] = fmaf(gg, xh, f1[i]);
            }
        }
    }
};

// ---- kernels ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ST_THREADS, 2) staged_bn_apply_kernel(const SOp x, const ApplyBody<T> body_in) {
    extern __shared__ __align__(128) uint8_t smem[];
    ApplyBody<T> body = body_in;                     // (the body carries per-thread state: a private copy in registers)
    const SOp ops[1] = {x};
    stream_pipeline<T, 1>(body.g, ops, smem, body);
    if (threadIdx.x < ST_CONSUMERS) zero_border<T>(body.out, body.g, threadIdx.x);
}
// per-CTA partial sums of the consumers' fp32 strips: ws[(blockIdx.x * 2 + which) * C + channel], fp64; every CTA writes all
// 2*C entries.  Threads that share a channel octet (the PP pixel lanes of a sub-chunk) are summed in lane order: fixed order.
__device__ __forceinline__ void block_sums(const StreamGeo& g, uint8_t* smem, int c, const float (&f0)[VEC], const float (&f1)[VEC],
                                           double* ws) {
    __syncthreads();                                   // every stage has been consumed: reuse the data area
    float* red = reinterpret_cast<float*>(smem + ST_HDR);           // [2][PP][C]
    const int tid = threadIdx.x;
    if (tid < g.CH) {
        const int p = tid / g.CV;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            red[(0 * g.PP + p) * g.C + c + i] = f0[i];
            red[(1 * g.PP + p) * g.C + c + i] = f1[i];
        }
    }
    __syncthreads();
    for (int j = tid; j < 2 * g.C; j += ST_THREADS) {
        const int which = j / g.C, ch = j - which * g.C;
        double a = 0.0;
        for (int p = 0; p < g.PP; ++p) a += (double)red[(which * g.PP + p) * g.C + ch];
        ws[((long long)blockIdx.x * 2 + which) * g.C + ch] = a;
    }
}
template <typename T, bool STATS>
__global__ void __launch_bounds__(ST_THREADS, 2) staged_combine_kernel(const SOp r, const SOp c, const CombineBody<T, STATS> body_in,
                                                                       double* ws) {
    extern __shared__ __align__(128) uint8_t smem[];
    CombineBody<T, STATS> body = body_in;
    const SOp ops[2] = {r, c};
    stream_pipeline<T, 2>(body.g, ops, smem, body);
    if (threadIdx.x < ST_CONSUMERS) zero_border<T>(body.out, body.g, threadIdx.x);
    if (STATS) block_sums(body.g, smem, body.c, body.f0, body.f1, ws);
}
struct SOps4 {
    SOp o[4];
};
template <typename T, bool GATE, bool ADD, bool OUT2, bool RECOMP>
__global__ void __launch_bounds__(ST_THREADS, 2) staged_bn_bwd_apply_kernel(const SOps4 in, const BwdBody<T, GATE, ADD, OUT2, RECOMP> body_in) {
    extern __shared__ __align__(128) uint8_t smem[];
    BwdBody<T, GATE, ADD, OUT2, RECOMP> body = body_in;
    constexpr int NOPS = BwdBody<T, GATE, ADD, OUT2, RECOMP>::NOPS;
    SOp ops[NOPS];
#pragma unroll
    for (int i = 0; i < NOPS; ++i) ops[i] = in.o[i];
    stream_pipeline<T, NOPS>(body.g, ops, smem, body);
    if (threadIdx.x < ST_CONSUMERS) {
        zero_border<T>(body.out, body.g, threadIdx.x);
        if (OUT2) zero_border<T>(body.out2, body.g, threadIdx.x);
    }
}
// partial sums: ws[(blockIdx.x * 2 + which) * C + channel], fp64; every CTA writes all 2*C entries
template <typename T, int MODE, bool GATE>
__global__ void __launch_bounds__(ST_THREADS, 2) staged_reduce_kernel(const SOps4 in, const ReduceBody<T, MODE, GATE> body_in, double* ws) {
    extern __shared__ __align__(128) uint8_t smem[];
    ReduceBody<T, MODE, GATE> body = body_in;
    constexpr int NOPS = ReduceBody<T, MODE, GATE>::NOPS;
    SOp ops[NOPS];
#pragma unroll
    for (int i = 0; i < NOPS; ++i) ops[i] = in.o[i];
    stream_pipeline<T, NOPS>(body.g, ops, smem, body);
    block_sums(body.g, smem, body.c, body.f0, body.f1, ws);
}

template <typename T>
__global__ void __launch_bounds__(ST_THREADS, 2) staged_reduce_recomp_kernel(const SOp dy, const SOp x, const ReduceRecompBody<T> body_in,
                                                                             double* ws) {
    extern __shared__ __align__(128) uint8_t smem[];
    ReduceRecompBody<T> body = body_in;
    const SOp ops[2] = {dy, x};
    stream_pipeline<T, 2>(body.g, ops, smem, body);
    block_sums(body.g, smem, body.c, body.f0, body.f1, ws);
}

// ---- host side ------------------------------------------------------------------------------------------------------------
int g_num_sms = 0;
int num_sms() {
    if (g_num_sms <= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}
// The apply-type passes walk their tensors BACKWARDS: the pass that ran just before them (the statistics / sums reduction
// over the same operands, or the GEMM that produced the input) touched the END of those tensors last, so with a 126 MB L2
// the first ~100 MB an apply pass asks for are still on chip.  MOPOE_ST_REVERSE=0 restores the forward walk.
int gate_recompute() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_GATE_RECOMPUTE");
        v = e ? atoi(e) : 1;
    }
    return v;
}
int st_reverse() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOPOE_ST_REVERSE");
        v = e ? atoi(e) : 1;
    }
    return v;
}
bool view_ok(const mopoe_view_t* v) {
    return v && v->sW == v->C && v->C % VEC == 0 && v->C / VEC <= ST_CONSUMERS && (reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0 &&
           v->sB % VEC == 0 && v->sH % VEC == 0 && v->sB < (1ll << 31) && v->sH < (1ll << 31);
}
SOp sop(const mopoe_view_t* v) { return SOp{v->ptr, (int)v->sB, (int)v->sH}; }
SOut sout(const mopoe_view_t* v) { return SOut{v->ptr, (int)v->sB, (int)v->sH, v->ph, v->pw}; }

// geometry + launch shape; esize = bytes per element, nops = staged operands, ks = sub-chunks per item (the kernel's KS).
// smem_floor: bytes the kernel needs anyway
bool make_geo(const mopoe_view_t* v, int esize, int nops, int ks, StreamGeo& g, int& grid, size_t& smem, size_t smem_floor = 0) {
    g.B = v->B; g.H = v->H; g.W = v->W; g.C = v->C; g.CV = v->C / VEC;
    int pp = ST_CONSUMERS / g.CV;
    if (pp > g.W) pp = g.W;
    const int sc = (g.W + pp - 1) / pp;                   // sub-chunks per row
    g.PP = (g.W + sc - 1) / sc;
    g.CH = g.PP * g.CV;
    const int sc2 = (g.W + g.PP - 1) / g.PP;
    if (ks > sc2) ks = sc2;
    g.KC = (sc2 + ks - 1) / ks;
    const int ks_bal = (sc2 + g.KC - 1) / g.KC;           // balanced: 5 sub-chunks at ks = 4 -> items of 3 + 2
    g.IO = ks_bal * g.CH;
    const long long items = (long long)g.B * g.H * g.KC;
    if (items <= 0 || items >= (1ll << 31)) return false;
    g.items = (unsigned)items;
    g.op_bytes = (unsigned)((g.IO * VEC * esize + 127) / 128 * 128);
    const size_t budget = 100 * 1024;                     // two CTAs per SM
    int stages = (int)((budget - ST_HDR) / ((size_t)g.op_bytes * nops));
    if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
    if (stages < 2) return false;
    g.stages = stages;
    g.reverse = 0;
    g.fKC = FastDiv((unsigned)g.KC); g.fH = FastDiv((unsigned)g.H); g.fCV = FastDiv((unsigned)g.CV);
    smem = ST_HDR + (size_t)stages * g.op_bytes * nops;
    if (smem < smem_floor) smem = smem_floor;
    const int cap = 2 * num_sms();
    grid = (int)(items < cap ? items : cap);
    return true;
}
template <class K>
int set_smem(K kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) MOPOE_FAIL("staged kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return 0;
}
#define ST_ATTR(kernel)                        \
    do {                                       \
        static bool done_ = false;             \
        if (!done_) {                          \
            if (set_smem(kernel)) return 1;    \
            done_ = true;                      \
        }                                      \
    } while (0)

}  // namespace

// Each launcher returns 0 (launched), 1 (error, message set) or -1 (shape not eligible: caller uses its register-staged kernel).
int mopoe_staged_bn_apply(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, int relu, const mopoe_view_t* out, cudaStream_t st) {
    if (!view_ok(x) || !view_ok(out)) return -1;
    const int es = x->dtype == MOPOE_BF16 ? 2 : 4;
    StreamGeo g;
    int grid;
    size_t smem;
    const int ks = x->dtype == MOPOE_BF16 ? ApplyBody<bf16>::KS : ApplyBody<float>::KS;
    if (!make_geo(out, es, 1, ks, g, grid, smem)) return -1;
    g.reverse = st_reverse();
    MOPOE_DISPATCH_T(x->dtype, T, {
        ST_ATTR(staged_bn_apply_kernel<T>);
        ApplyBody<T> body;
        body.g = g; body.mk = MaskRef{mask, mask_mode};
        body.mean = mean; body.invstd = invstd; body.gamma = gamma; body.beta = beta; body.relu = relu; body.out = sout(out);
        staged_bn_apply_kernel<T><<<grid, ST_THREADS, smem, st>>>(sop(x), body);
    });
    MOPOE_CHECK_LAUNCH("staged_bn_apply");
    return 0;
}

static size_t sums_floor(const mopoe_view_t* v) {
    const int cv = v->C / VEC;
    const int pp_max = ST_CONSUMERS / cv > 0 ? ST_CONSUMERS / cv : 1;
    return ST_HDR + (size_t)2 * pp_max * v->C * sizeof(float);          // block_sums re-uses the data area: [2][PP][C] floats
}
// ws != NULL: also the per-channel statistics (sum, sum of squares) of the stored output, as *nchunk_used <= nchunk_cap
// partial slabs of [2][C] doubles
int mopoe_staged_combine(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         const mopoe_view_t* c, const uint8_t* mask, int mask_mode, float a, float b, const mopoe_view_t* out,
                         double* ws, int nchunk_cap, int* nchunk_used, cudaStream_t st) {
    if (!view_ok(r) || !view_ok(c) || !view_ok(out)) return -1;
    const int es = r->dtype == MOPOE_BF16 ? 2 : 4;
    StreamGeo g;
    int grid;
    size_t smem;
    const int ks = r->dtype == MOPOE_BF16 ? CombineBody<bf16, false>::KS : CombineBody<float, false>::KS;
    const size_t floor_bytes = ws ? sums_floor(r) : 0;
    if (floor_bytes > 100 * 1024) return -1;
    if (!make_geo(out, es, 2, ks, g, grid, smem, floor_bytes)) return -1;
    g.reverse = st_reverse();
    if (ws) {
        if (nchunk_cap < 1) return -1;
        if (grid > nchunk_cap) grid = nchunk_cap;
        *nchunk_used = grid;
    }
#define ST_CB(S)                                                                                                                 \
    do {                                                                                                                         \
        ST_ATTR((staged_combine_kernel<T, S>));                                                                                  \
        CombineBody<T, S> body;                                                                                                  \
        body.g = g; body.mk = MaskRef{mask, mask_mode};                                                                          \
        body.mean = mean; body.invstd = invstd; body.gamma = gamma; body.beta = beta; body.a = a; body.bcoef = b;                \
        body.out = sout(out);                                                                                                    \
        staged_combine_kernel<T, S><<<grid, ST_THREADS, smem, st>>>(sop(r), sop(c), body, ws);                                   \
    } while (0)
    MOPOE_DISPATCH_T(r->dtype, T, {
        if (ws) ST_CB(true);
        else ST_CB(false);
    });
#undef ST_CB
    MOPOE_CHECK_LAUNCH("staged_combine");
    return 0;
}

template <typename T, bool GATE, bool ADD, bool OUT2, bool RECOMP>
static int launch_bwd(const StreamGeo& g, int grid, size_t smem, cudaStream_t st, const mopoe_view_t* dy, const mopoe_view_t* gate,
                      float gscale, const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean,
                      const float* invstd, const float* gamma, const float* sums, float inv_cnt, const mopoe_view_t* addend,
                      const mopoe_view_t* out, const mopoe_view_t* out2, const uint8_t* mask2, int mask2_mode, float scale2,
                      const float* gate_beta) {
    ST_ATTR((staged_bn_bwd_apply_kernel<T, GATE, ADD, OUT2, RECOMP>));
    BwdBody<T, GATE, ADD, OUT2, RECOMP> body;
    body.gate_beta = gate_beta;
    body.g = g; body.mk = MaskRef{mask, mask_mode}; body.mk2 = MaskRef{mask2, mask2_mode};
    body.mean = mean; body.invstd = invstd; body.gamma = gamma; body.sums = sums;
    body.gscale = gscale; body.inv_cnt = inv_cnt; body.scale2 = scale2;
    body.out = sout(out);
    body.out2 = OUT2 ? sout(out2) : sout(out);
    SOps4 in;
    int n = 0;
    in.o[n++] = sop(dy);
    in.o[n++] = sop(x);
    if (GATE && !RECOMP) in.o[n++] = sop(gate);
    if (ADD) in.o[n++] = sop(addend);
    for (; n < 4; ++n) in.o[n] = sop(x);
    staged_bn_bwd_apply_kernel<T, GATE, ADD, OUT2, RECOMP><<<grid, ST_THREADS, smem, st>>>(in, body);
    return 0;
}

int mopoe_staged_bn_bwd_apply(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale, const mopoe_view_t* x,
                              const uint8_t* mask, int mask_mode, const float* mean, const float* invstd, const float* gamma,
                              const float* sums, const mopoe_view_t* addend, const mopoe_view_t* out, const mopoe_view_t* out2,
                              const uint8_t* mask2, int mask2_mode, float scale2, const float* gate_beta, cudaStream_t st) {
    if (!view_ok(dy) || !view_ok(x) || !view_ok(out) || (gate && !view_ok(gate)) || (addend && !view_ok(addend)) ||
        (out2 && !view_ok(out2)))
        return -1;
    const int es = x->dtype == MOPOE_BF16 ? 2 : 4;
    const bool recomp = gate && gate_beta && !out2 && gate_recompute();      // gate recomputed from x: not a staged operand
    const int nops = 2 + (gate && !recomp ? 1 : 0) + (addend ? 1 : 0);
    StreamGeo g;
    int grid;
    size_t smem;
    const int kb = nops >= 4 ? ST_K4 : (nops == 3 ? ST_K3 : ST_K2);
    const int ks = x->dtype == MOPOE_BF16 ? ks_for<bf16>(kb) : ks_for<float>(kb);
    if (!make_geo(out, es, nops, ks, g, grid, smem)) return -1;
    g.reverse = st_reverse();
    const float inv_cnt = 1.f / ((float)x->B * (float)x->H * (float)x->W);
#define ST_BW(G, A, O)                                                                                                          \
    if (launch_bwd<T, G, A, O, false>(g, grid, smem, st, dy, gate, gscale, x, mask, mask_mode, mean, invstd, gamma, sums, inv_cnt,  \
                                      addend, out, out2, mask2, mask2_mode, scale2, nullptr))                                      \
    return 1
#define ST_BWR(A)                                                                                                               \
    if (launch_bwd<T, true, A, false, true>(g, grid, smem, st, dy, gate, gscale, x, mask, mask_mode, mean, invstd, gamma, sums,    \
                                            inv_cnt, addend, out, out2, mask2, mask2_mode, scale2, gate_beta))                     \
    return 1
    MOPOE_DISPATCH_T(x->dtype, T, {
        if (recomp) {
            if (addend) { ST_BWR(true); } else { ST_BWR(false); }
        } else if (out2) {
            if (gate) { if (addend) { ST_BW(true, true, true); } else { ST_BW(true, false, true); } }
            else { if (addend) { ST_BW(false, true, true); } else { ST_BW(false, false, true); } }
        } else {
            if (gate) { if (addend) { ST_BW(true, true, false); } else { ST_BW(true, false, false); } }
            else { if (addend) { ST_BW(false, true, false); } else { ST_BW(false, false, false); } }
        }
    });
#undef ST_BW
#undef ST_BWR
    MOPOE_CHECK_LAUNCH("staged_bn_bwd_apply");
    return 0;
}

template <typename T, int MODE, bool GATE>
static int launch_reduce(const StreamGeo& g, int grid, size_t smem, cudaStream_t st, const mopoe_view_t* x, const mopoe_view_t* dy,
                         const mopoe_view_t* gate, float gscale, const uint8_t* mask, int mask_mode, const float* mean,
                         const float* invstd, double* ws) {
    ST_ATTR((staged_reduce_kernel<T, MODE, GATE>));
    ReduceBody<T, MODE, GATE> body;
    body.g = g; body.mk = MaskRef{mask, mask_mode}; body.mean = mean; body.invstd = invstd; body.gscale = gscale;
    SOps4 in;
    int n = 0;
    in.o[n++] = sop(x);
    if (MODE == 1) in.o[n++] = sop(dy);
    if (MODE == 1 && GATE) in.o[n++] = sop(gate);
    for (; n < 4; ++n) in.o[n] = sop(x);
    staged_reduce_kernel<T, MODE, GATE><<<grid, ST_THREADS, smem, st>>>(in, body, ws);
    return 0;
}

// mode: 0 statistics, 1 BN-backward sums, 2 column sums.  Writes *nchunk_used partial rows into ws (<= nchunk_cap).
// gate_gamma / gate_beta (mode 1, with a gate): the gate is relu(gamma * xhat + beta) of THIS BatchNorm -> the two-operand
// pass that recomputes it from x (ReduceRecompBody)
int mopoe_staged_reduce(int mode, const mopoe_view_t* x, const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                        const uint8_t* mask, int mask_mode, const float* mean, const float* invstd, double* ws, int nchunk_cap,
                        int* nchunk_used, const float* gate_gamma, const float* gate_beta, cudaStream_t st) {
    if (!view_ok(x) || (mode == 1 && !view_ok(dy)) || (gate && !view_ok(gate))) return -1;
    const int es = x->dtype == MOPOE_BF16 ? 2 : 4;
    const bool xg = mode == 1 && gate && gate_gamma && gate_beta && gate_recompute();
    const int nops = mode == 1 ? (gate && !xg ? 3 : 2) : 1;
    StreamGeo g;
    int grid;
    size_t smem;
    const size_t floor_bytes = sums_floor(x);
    if (floor_bytes > 100 * 1024) return -1;
    const int kb = nops == 1 ? ST_K1 : (nops == 2 ? ST_K2 : ST_K3);
    const int ks = x->dtype == MOPOE_BF16 ? ks_for<bf16>(kb) : ks_for<float>(kb);
    if (!make_geo(x, es, nops, ks, g, grid, smem, floor_bytes)) return -1;
    if (nchunk_cap < 1) return -1;
    if (grid > nchunk_cap) grid = nchunk_cap;
    *nchunk_used = grid;
    if (xg) {
        MOPOE_DISPATCH_T(x->dtype, T, {
            ST_ATTR(staged_reduce_recomp_kernel<T>);
            ReduceRecompBody<T> body;
            body.g = g; body.mk = MaskRef{mask, mask_mode};
            body.mean = mean; body.invstd = invstd; body.gamma = gate_gamma; body.beta = gate_beta; body.gscale = gscale;
            staged_reduce_recomp_kernel<T><<<grid, ST_THREADS, smem, st>>>(sop(dy), sop(x), body, ws);
        });
        MOPOE_CHECK_LAUNCH("staged_reduce_recomp");
        return 0;
    }
#define ST_RD(M, G)                                                                                                  \
    if (launch_reduce<T, M, G>(g, grid, smem, st, x, dy, gate, gscale, mask, mask_mode, mean, invstd, ws)) return 1
    MOPOE_DISPATCH_T(x->dtype, T, {
        if (mode == 0) { ST_RD(0, false); }
        else if (mode == 2) { ST_RD(2, false); }
        else if (gate) { ST_RD(1, true); }
        else { ST_RD(1, false); }
    });
#undef ST_RD
    MOPOE_CHECK_LAUNCH("staged_reduce");
    return 0;
}
