// conv_direct.cu — the two single-channel image layers.  K = 9 (first conv) and N = 1 (last deconv)
// are not tensor-core shapes, so they are CUDA-core kernels bound by the 128-channel
// side of their traffic.
//   first:  nn.Conv2d(1, C, 3, stride=2, padding=1, bias=False)          FeatureExtractorImg.py:29-34
//   last:   nn.ConvTranspose2d(C, 1, 3, stride=2, padding=1, output_padding=1)   DataGeneratorImg.py:84-90
#include "common.cuh"

constexpr int VEC = 4;

// ---- first conv forward: thread = (storage position of out, 4 channels) -------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv3x3s2_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               int H, int W, DView<T> out, long long total) {
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int CV = out.C / VEC;
    const int Ws = out.W + 2 * out.pw, Hs = out.H + 2 * out.ph;
    int c = (int)(idx % CV) * VEC;
    long long pos = idx / CV;
    int ws = (int)(pos % Ws);
    pos /= Ws;
    int hs = (int)(pos % Hs);
    int b = (int)(pos / Hs);
    int oy = hs - out.ph, ox = ws - out.pw;
    float o[VEC] = {0.f, 0.f, 0.f, 0.f};
    if (oy >= 0 && oy < out.H && ox >= 0 && ox < out.W) {
        const float* xb = x + (long long)b * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= W) continue;
                float xv = __ldg(xb + (long long)iy * W + ix);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] += xv * __ldg(w + (c + i) * 9 + ky * 3 + kx);
            }
        }
    }
    stv<VEC>(out.p + (long long)b * out.sB + (long long)oy * out.sH + (long long)ox * out.sW + c, o);
}
extern "C" int mopoe_conv3x3s2_c1_fwd(const float* x, const float* w, int B, int H, int W, const mopoe_view_t* out,
                                      void* stream) {
    MOPOE_REQUIRE(out->B == B && out->H == H / 2 && out->W == W / 2 && out->C % VEC == 0, "conv3x3s2_c1_fwd: bad out view");
    long long total = (long long)B * (out->H + 2 * out->ph) * (out->W + 2 * out->pw) * (out->C / VEC);
    MOPOE_DISPATCH_T(out->dtype, T, {
        conv3x3s2_c1_fwd_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
            x, w, H, W, make_dview<T>(out), total);
    });
    MOPOE_CHECK_LAUNCH("conv3x3s2_c1_fwd");
    return 0;
}

// ---- first conv weight gradient: dw[c, tap] = sum_{b,oy,ox} dy[b,oy,ox,c] * x[b, 2oy-1+ky, 2ox-1+kx] ------
template <typename T>
__global__ void __launch_bounds__(256) conv3x3s2_c1_wgrad_kernel(const float* __restrict__ x, DView<const T> dy, int H,
                                                                 int W, double* __restrict__ ws, int nchunk) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = (blockIdx.x * 32 + tx) * VEC;
    const bool cvalid = c < dy.C;
    const long long rows = (long long)dy.B * dy.H * dy.W;
    const long long rpc = (rows + nchunk - 1) / nchunk;
    const long long r0 = (long long)blockIdx.y * rpc, r1 = min(rows, r0 + rpc);
    float acc[9][VEC];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[t][i] = 0.f;
    double dacc[9][VEC];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) dacc[t][i] = 0.0;
    int it = 0;
    if (cvalid) {
        for (long long r = r0 + ty; r < r1; r += 8) {
            int ox = (int)(r % dy.W);
            long long t2 = r / dy.W;
            int oy = (int)(t2 % dy.H);
            int b = (int)(t2 / dy.H);
            float g[VEC];
            ldv<VEC>(dy.p + (long long)b * dy.sB + (long long)oy * dy.sH + (long long)ox * dy.sW + c, g);
            const float* xb = x + (long long)b * H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                int iy = 2 * oy - 1 + ky;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int ix = 2 * ox - 1 + kx;
                    float xv = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + (long long)iy * W + ix) : 0.f;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[ky * 3 + kx][i] += g[i] * xv;
                }
            }
            if (++it == 64) {   // flush the fp32 strip into fp64
                it = 0;
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { dacc[t][i] += (double)acc[t][i]; acc[t][i] = 0.f; }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) dacc[t][i] += (double)acc[t][i];
    __shared__ double sm[8][32][VEC];
    for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm[ty][tx][i] = dacc[t][i];
        __syncthreads();
        if (ty == 0 && cvalid) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) a += sm[j][tx][i];
                ws[((long long)blockIdx.y * 9 + t) * dy.C + c + i] = a;
            }
        }
        __syncthreads();
    }
}
// out[c*ntap + t] (+)= sum_k ws[(k*ntap + t)*C + c]
__global__ void tap_finalize_kernel(const double* ws, int nchunk, int ntap, int C, float* out, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * ntap) return;
    int c = i / ntap, t = i % ntap;
    double s = 0.0;
    for (int k = 0; k < nchunk; ++k) s += ws[((long long)k * ntap + t) * C + c];
    out[i] = (accumulate ? out[i] : 0.f) + (float)s;
}
extern "C" int mopoe_conv3x3s2_c1_wgrad(const float* x, const mopoe_view_t* dy, int B, int H, int W, float* dw,
                                        int accumulate, double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(dy->B == B && dy->H == H / 2 && dy->W == W / 2 && dy->C % VEC == 0, "conv3x3s2_c1_wgrad: bad dy view");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 block(32, 8), grid((dy->C + 127) / 128, nchunk);
    MOPOE_DISPATCH_T(dy->dtype, T, {
        conv3x3s2_c1_wgrad_kernel<T><<<grid, block, 0, st>>>(x, make_dview<const T>(dy), H, W, ws, nchunk);
    });
    MOPOE_CHECK_LAUNCH("conv3x3s2_c1_wgrad");
    tap_finalize_kernel<<<(dy->C * 9 + 127) / 128, 128, 0, st>>>(ws, nchunk, 9, dy->C, dw, accumulate);
    MOPOE_CHECK_LAUNCH("tap_finalize");
    return 0;
}

// ---- last deconv forward: one warp per 2x2 output quad ---------------------------------------------------
// out(2t,2s)     = x(t,s).w11
// out(2t,2s+1)   = x(t,s).w12 + x(t,s+1).w10
// out(2t+1,2s)   = x(t,s).w21 + x(t+1,s).w01
// out(2t+1,2s+1) = x(t,s).w22 + x(t,s+1).w20 + x(t+1,s).w02 + x(t+1,s+1).w00      (w[ky][kx], oy = 2iy-1+ky)
template <typename T>
__global__ void __launch_bounds__(256) deconv3x3s2_c1_fwd_kernel(DView<const T> x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                                 long long quads) {
    const int lane = threadIdx.x & 31;
    long long q = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= quads) return;
    int s = (int)(q % x.W);
    long long t2 = q / x.W;
    int t = (int)(t2 % x.H);
    int b = (int)(t2 / x.H);
    float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
    const bool rs = s + 1 < x.W, dn = t + 1 < x.H;
    for (int c = lane * VEC; c < x.C; c += 32 * VEC) {
        float a[VEC], ar[VEC], ad[VEC], adr[VEC];
        const T* p = x.p + (long long)b * x.sB + (long long)t * x.sH + (long long)s * x.sW + c;
        ldv<VEC>(p, a);
        if (rs) ldv<VEC>(p + x.sW, ar);
        if (dn) ldv<VEC>(p + x.sH, ad);
        if (rs && dn) ldv<VEC>(p + x.sH + x.sW, adr);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float* wc = w + (c + i) * 9;
            float w00 = __ldg(wc), w01 = __ldg(wc + 1), w02 = __ldg(wc + 2), w10 = __ldg(wc + 3), w11 = __ldg(wc + 4),
                  w12 = __ldg(wc + 5), w20 = __ldg(wc + 6), w21 = __ldg(wc + 7), w22 = __ldg(wc + 8);
            o00 += a[i] * w11;
            o01 += a[i] * w12;
            o10 += a[i] * w21;
            o11 += a[i] * w22;
            if (rs) { o01 += ar[i] * w10; o11 += ar[i] * w20; }
            if (dn) { o10 += ad[i] * w01; o11 += ad[i] * w02; }
            if (rs && dn) o11 += adr[i] * w00;
        }
    }
    o00 = warp_sum(o00); o01 = warp_sum(o01); o10 = warp_sum(o10); o11 = warp_sum(o11);
    if (lane == 0) {
        const float bb = bias[0];
        const int OW = 2 * x.W;
        float* ob = out + ((long long)b * 2 * x.H + 2 * t) * OW + 2 * s;
        *reinterpret_cast<float2*>(ob) = make_float2(o00 + bb, o01 + bb);
        *reinterpret_cast<float2*>(ob + OW) = make_float2(o10 + bb, o11 + bb);
    }
}
extern "C" int mopoe_deconv3x3s2_c1_fwd(const mopoe_view_t* x, const float* w, const float* bias, float* out,
                                        void* stream) {
    MOPOE_REQUIRE(x->C % VEC == 0, "deconv3x3s2_c1_fwd: C=%d", x->C);
    long long quads = (long long)x->B * x->H * x->W;
    MOPOE_DISPATCH_T(x->dtype, T, {
        deconv3x3s2_c1_fwd_kernel<T><<<(unsigned)ceil_div64(quads, 8), 256, 0, (cudaStream_t)stream>>>(
            make_dview<const T>(x), w, bias, out, quads);
    });
    MOPOE_CHECK_LAUNCH("deconv3x3s2_c1_fwd");
    return 0;
}

// ---- last deconv backward ------------------------------------------------------------------------------------
// dx[b,t,s,c] = sum_{ky,kx} dout[b, 2t-1+ky, 2s-1+kx] * w[c,ky,kx]
template <typename T>
__global__ void __launch_bounds__(256) deconv3x3s2_c1_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                                DView<T> dx, long long total) {
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int CV = dx.C / VEC;
    const int Ws = dx.W + 2 * dx.pw, Hs = dx.H + 2 * dx.ph;
    int c = (int)(idx % CV) * VEC;
    long long pos = idx / CV;
    int ws = (int)(pos % Ws);
    pos /= Ws;
    int hs = (int)(pos % Hs);
    int b = (int)(pos / Hs);
    int t = hs - dx.ph, s = ws - dx.pw;
    float o[VEC] = {0.f, 0.f, 0.f, 0.f};
    if (t >= 0 && t < dx.H && s >= 0 && s < dx.W) {
        const int OH = 2 * dx.H, OW = 2 * dx.W;
        const float* db = dout + (long long)b * OH * OW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            int oy = 2 * t - 1 + ky;
            if (oy < 0 || oy >= OH) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                int ox = 2 * s - 1 + kx;
                if (ox < 0 || ox >= OW) continue;
                float g = __ldg(db + (long long)oy * OW + ox);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] += g * __ldg(w + (c + i) * 9 + ky * 3 + kx);
            }
        }
    }
    stv<VEC>(dx.p + (long long)b * dx.sB + (long long)t * dx.sH + (long long)s * dx.sW + c, o);
}
// dw[c,ky,kx] = sum_{b,t,s} x[b,t,s,c] * dout[b, 2t-1+ky, 2s-1+kx];  ws layout [chunk][9][C]
template <typename T>
__global__ void __launch_bounds__(256) deconv3x3s2_c1_dw_kernel(DView<const T> x, const float* __restrict__ dout,
                                                                double* __restrict__ ws, int nchunk) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = (blockIdx.x * 32 + tx) * VEC;
    const bool cvalid = c < x.C;
    const long long rows = (long long)x.B * x.H * x.W;
    const long long rpc = (rows + nchunk - 1) / nchunk;
    const long long r0 = (long long)blockIdx.y * rpc, r1 = min(rows, r0 + rpc);
    const int OH = 2 * x.H, OW = 2 * x.W;
    float acc[9][VEC];
    double dacc[9][VEC];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[t][i] = 0.f; dacc[t][i] = 0.0; }
    int it = 0;
    if (cvalid) {
        for (long long r = r0 + ty; r < r1; r += 8) {
            int s = (int)(r % x.W);
            long long t2 = r / x.W;
            int t = (int)(t2 % x.H);
            int b = (int)(t2 / x.H);
            float xv[VEC];
            ldv<VEC>(x.p + (long long)b * x.sB + (long long)t * x.sH + (long long)s * x.sW + c, xv);
            const float* db = dout + (long long)b * OH * OW;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                int oy = 2 * t - 1 + ky;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    int ox = 2 * s - 1 + kx;
                    float g = (oy >= 0 && oy < OH && ox >= 0 && ox < OW) ? __ldg(db + (long long)oy * OW + ox) : 0.f;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[ky * 3 + kx][i] += g * xv[i];
                }
            }
            if (++it == 64) {
                it = 0;
#pragma unroll
                for (int tt = 0; tt < 9; ++tt)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { dacc[tt][i] += (double)acc[tt][i]; acc[tt][i] = 0.f; }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) dacc[t][i] += (double)acc[t][i];
    __shared__ double sm[8][32][VEC];
    for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm[ty][tx][i] = dacc[t][i];
        __syncthreads();
        if (ty == 0 && cvalid) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) a += sm[j][tx][i];
                ws[((long long)blockIdx.y * 9 + t) * x.C + c + i] = a;
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ v, long long n, double* __restrict__ part) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc += (double)v[i];
    __shared__ double sm[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += sm[i];
        part[blockIdx.x] = s;
    }
}
__global__ void sum_final_kernel(const double* part, int n, float* out, int accumulate) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += part[i];
        out[0] = (accumulate ? out[0] : 0.f) + (float)s;
    }
}
extern "C" int mopoe_deconv3x3s2_c1_bwd(const mopoe_view_t* x, const float* w, const float* dout, const mopoe_view_t* dx,
                                        float* dw, float* dbias, int accumulate, double* ws, int nchunk, void* stream) {
    MOPOE_REQUIRE(x->C % VEC == 0 && dx->C == x->C && dx->B == x->B && dx->H == x->H && dx->W == x->W &&
                      dx->dtype == x->dtype, "deconv3x3s2_c1_bwd: bad views");
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)dx->B * (dx->H + 2 * dx->ph) * (dx->W + 2 * dx->pw) * (dx->C / VEC);
    dim3 block(32, 8), grid((x->C + 127) / 128, nchunk);
    MOPOE_DISPATCH_T(x->dtype, T, {
        deconv3x3s2_c1_dx_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(dout, w, make_dview<T>(dx), total);
        deconv3x3s2_c1_dw_kernel<T><<<grid, block, 0, st>>>(make_dview<const T>(x), dout, ws, nchunk);
    });
    MOPOE_CHECK_LAUNCH("deconv3x3s2_c1_bwd");
    tap_finalize_kernel<<<(x->C * 9 + 127) / 128, 128, 0, st>>>(ws, nchunk, 9, x->C, dw, accumulate);
    MOPOE_CHECK_LAUNCH("tap_finalize");
    double* part = ws + (long long)nchunk * 9 * x->C;
    long long n = (long long)x->B * 4 * x->H * x->W;
    sum_partial_kernel<<<nchunk, 256, 0, st>>>(dout, n, part);
    sum_final_kernel<<<1, 32, 0, st>>>(part, nchunk, dbias, accumulate);
    MOPOE_CHECK_LAUNCH("dbias");
    return 0;
}
