// text_stem.cu — the first layer of the character-text encoder as a GATHER.
//
// The reference one-hot encodes every character on the host (dataio/MimicDataset.py:92-96, utils/text.py:13-34) and feeds
// nn.Conv1d(71, C, 4, 2, 1) (char_encoding/FeatureExtractorText.py:30-31, :71-72) with rows that hold a single 1.0: the
// convolution is then a sum of (at most) 4 weight columns,
//     y[b, l, :] = bias + sum_{t<4, 0 <= 2l-1+t < L} W[:, idx[b, 2l-1+t], t],
// When the step is fed the one-byte-per-token wire format the indices exist on the device: the forward is this gather (no
// fp32 -> bf16 layout copy of the 74 MB one-hot tensor, no K = 4 x 80 GEMM); the weight gradient stays a tensor-core GEMM,
// over one-hot rows that are built from the indices in the activation dtype when the backward pass needs them.  (A
// deterministic scatter into (tap, character) bins in shared memory was built and measured: one thread per channel walking
// ~900 positions held 145 KB of shared memory per SM for ~100 us and slowed the step by 0.35 ms.)
#include "common.cuh"

namespace {
constexpr int VEC = 8;

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&o)[8]) {
    if constexpr (sizeof(T) == 2) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else {
        const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
}
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

// table: full-form weights [(t * V + v), c] in the activation dtype (what the GEMM path multiplies by the one-hot rows)
template <typename T>
__global__ void __launch_bounds__(256) text_stem_gather_fwd_kernel(const uint8_t* __restrict__ idx, int B, int L, int V,
                                                                   const T* __restrict__ table, const float* __restrict__ bias, int C,
                                                                   T* __restrict__ out, long long sB, long long total) {
    const int CV = C / VEC, OL = L / 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cv = (int)(i % CV);
        const long long q = i / CV;
        const int l = (int)(q % OL), b = (int)(q / OL);
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const uint8_t* row = idx + (long long)b * L;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int pos = 2 * l - 1 + t;
            if (pos < 0 || pos >= L) continue;                       // zero padding of the convolution
            const int v = row[pos];
            if (v >= V) continue;                                    // (an out-of-range byte encodes "no character": all-zero row)
            float w[VEC];
            ld8<T>(table + ((long long)(t * V + v)) * C + cv * VEC, w);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += w[j];
        }
        if (bias) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += bias[cv * VEC + j];
        }
        st8<T>(out + (long long)b * sB + (long long)l * C + cv * VEC, acc);
    }
}

// the one-hot rows as the bordered, channel-padded activation the weight-gradient GEMM reads ([B, 1, L + 2pw, Vp], zero border
// and zero padding channels written here): built from the byte indices when the backward pass needs it — 1 byte read and
// 2*Vp bytes written per token instead of converting the fp32 one-hot tensor (4*V bytes read)
template <typename T>
__global__ void __launch_bounds__(256) text_onehot_act_kernel(const uint8_t* __restrict__ idx, int B, int L, int V, int Vp, int pw,
                                                              T* __restrict__ out, long long total) {
    const int CV = Vp / VEC, Ws = L + 2 * pw;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int cv = (int)(i % CV);
        const long long q = i / CV;
        const int ws = (int)(q % Ws), b = (int)(q / Ws);
        const int pos = ws - pw;
        int v = -1;
        if (pos >= 0 && pos < L) {
            v = idx[(long long)b * L + pos];
            if (v >= V) v = -1;
        }
        float o[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = (cv * VEC + j == v) ? 1.f : 0.f;
        st8<T>(out + ((long long)b * Ws + ws) * Vp + cv * VEC, o);
    }
}
}  // namespace

extern "C" int mopoe_text_stem_gather_fwd(const uint8_t* idx, int B, int L, int V, const void* table, int dtype, const float* bias,
                                          const mopoe_view_t* out, void* stream) {
    MOPOE_REQUIRE(idx && table && out, "text_stem_gather_fwd: null argument");
    MOPOE_REQUIRE(L % 2 == 0 && out->B == B && out->H == 1 && out->W == L / 2 && out->C % VEC == 0 && out->sW == out->C &&
                      out->dtype == dtype,
                  "text_stem_gather_fwd: output [%d,%d,%d,%d] does not match B=%d L=%d", out->B, out->H, out->W, out->C, B, L);
    const long long total = (long long)B * (L / 2) * (out->C / VEC);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    MOPOE_DISPATCH_T(dtype, T, {
        text_stem_gather_fwd_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, B, L, V, (const T*)table, bias, out->C,
                                                                                         (T*)out->ptr, out->sB, total);
    });
    MOPOE_CHECK_LAUNCH("text_stem_gather_fwd");
    return 0;
}

// idx [B, L] -> out: activation [B, 1, L, Vp] with border pw (storage [B, L + 2pw, Vp], contiguous), dtype f32 / bf16
extern "C" int mopoe_text_onehot_act(const uint8_t* idx, int B, int L, int V, const mopoe_view_t* out, void* stream) {
    MOPOE_REQUIRE(idx && out, "text_onehot_act: null argument");
    MOPOE_REQUIRE(out->B == B && out->H == 1 && out->W == L && out->C % VEC == 0 && out->C >= V && out->sW == out->C && out->ph == 0 &&
                      out->sB == (long long)(L + 2 * out->pw) * out->C,
                  "text_onehot_act: output [%d,%d,%d,%d] is not the contiguous bordered activation of B=%d L=%d", out->B, out->H,
                  out->W, out->C, B, L);
    const int Vp = out->C, pw = out->pw;
    const long long total = (long long)B * (L + 2 * pw) * (Vp / VEC);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    MOPOE_DISPATCH_T(out->dtype, T, {
        T* base = (T*)out->ptr - (long long)pw * Vp;                 // storage origin (the view addresses interior element 0)
        text_onehot_act_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, B, L, V, Vp, pw, base, total);
    });
    MOPOE_CHECK_LAUNCH("text_onehot_act");
    return 0;
}
