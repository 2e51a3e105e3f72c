// common.cuh — shared device/host helpers for libmopoe_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mopoe_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------------------------------
void mopoe_set_error(const char* fmt, ...);
#define MOPOE_FAIL(...)               \
    do {                              \
        mopoe_set_error(__VA_ARGS__); \
        return 1;                     \
    } while (0)
#define MOPOE_CHECK_LAUNCH(name)                                                      \
    do {                                                                              \
        cudaError_t e_ = cudaGetLastError();                                          \
        if (e_ != cudaSuccess) MOPOE_FAIL("%s: %s", name, cudaGetErrorString(e_));    \
    } while (0)
#define MOPOE_REQUIRE(cond, ...) \
    do {                         \
        if (!(cond)) MOPOE_FAIL(__VA_ARGS__); \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- division by a runtime constant as multiply-high + shift (valid for n < 2^31) ----------------------
struct FastDiv {
    unsigned d, mul, shr;
    __host__ __device__ FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(unsigned div) : d(div), mul(0), shr(0) {
        if (div > 1) {
            unsigned l = 0;
            while ((1u << l) < div) ++l;                      // ceil(log2 d)
            const unsigned p = 31 + l;
            mul = (unsigned)(((1ull << p) + div - 1) / div);
            shr = p - 32;
        }
    }
    __device__ __forceinline__ unsigned div(unsigned n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
    __device__ __forceinline__ void divmod(unsigned n, unsigned& q, unsigned& r) const {
        q = div(n);
        r = n - q * d;
    }
};

// ---- device view (mirror of mopoe_view_t with a typed pointer) --------------------------------------
template <typename T>
struct DView {
    T* p;
    int B, H, W, C, ph, pw;
    long long sB, sH, sW;
    FastDiv fW, fH, fWs, fHs, fCV8;      // W, H, W+2pw, H+2ph, C/8
};
template <typename T>
static inline DView<T> make_dview(const mopoe_view_t* v) {
    DView<T> d;
    d.p = (T*)v->ptr;
    d.B = v->B; d.H = v->H; d.W = v->W; d.C = v->C; d.ph = v->ph; d.pw = v->pw;
    d.sB = v->sB; d.sH = v->sH; d.sW = v->sW;
    d.fW = FastDiv((unsigned)v->W);
    d.fH = FastDiv((unsigned)v->H);
    d.fWs = FastDiv((unsigned)(v->W + 2 * v->pw));
    d.fHs = FastDiv((unsigned)(v->H + 2 * v->ph));
    d.fCV8 = FastDiv((unsigned)(v->C >= 8 ? v->C / 8 : 1));
    return d;
}

// ---- vector load/store: VEC consecutive channels as float --------------------------------------------
template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&o)[VEC]) {
    if constexpr (VEC == 4) {
        float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = p[i];
    }
}
template <int VEC>
__device__ __forceinline__ void ldv(const bf16* p, float (&o)[VEC]) {
    if constexpr (VEC == 4) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
        o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = __bfloat162float(p[i]);
    }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&o)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = o[i];
    }
}
template <int VEC>
__device__ __forceinline__ void stv(bf16* p, const float (&o)[VEC]) {
    if constexpr (VEC == 4) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&a);
        t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = __float2bfloat16_rn(o[i]);
    }
}
template <int VEC>
__device__ __forceinline__ void ldmask(const uint8_t* m, int mode, long long bc_idx, long long el_idx, float (&o)[VEC]) {
    if (mode == MOPOE_MASK_NONE) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = 1.f;
    } else {
        const uint8_t* q = m + (mode == MOPOE_MASK_BC ? bc_idx : el_idx);
        if constexpr (VEC == 4) {
            uchar4 t = *reinterpret_cast<const uchar4*>(q);
            o[0] = t.x ? 2.f : 0.f; o[1] = t.y ? 2.f : 0.f; o[2] = t.z ? 2.f : 0.f; o[3] = t.w ? 2.f : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = q[i] ? 2.f : 0.f;
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dtype dispatch for one typed operand
#define MOPOE_DISPATCH_T(dtype, T, ...)                               \
    do {                                                              \
        if ((dtype) == MOPOE_F32) { typedef float T; __VA_ARGS__; }   \
        else if ((dtype) == MOPOE_BF16) { typedef bf16 T; __VA_ARGS__; } \
        else MOPOE_FAIL("bad dtype %d", (int)(dtype));                \
    } while (0)
