// gemm_api.cu — public implicit-GEMM entry points: pick the tcgen05/TMA kernel when the problem is
// eligible (bf16 operands, tile-aligned shapes, sm_100 device), else the CUDA-core kernel.
#include "common.cuh"

int mopoe_conv_gemm_simt(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* stream);
size_t mopoe_conv_wgrad_ws_simt(const mopoe_window_t* A, const mopoe_rows_t* dY);
int mopoe_conv_wgrad_simt(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate, void* ws,
                          size_t ws_bytes, void* stream, const int* fin);
size_t mopoe_conv_wgrad_ws_simt_fin(const mopoe_window_t* A, const mopoe_rows_t* dY);
size_t mopoe_conv_wgrad_ws_tc_fin(const mopoe_window_t* A, const mopoe_rows_t* dY);
// gemm_tc.cu
int mopoe_tc_fwd_eligible(const mopoe_window_t* A, const mopoe_rows_t* D);
int mopoe_conv_gemm_tc(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* stream);
int mopoe_tc_wgrad_eligible(const mopoe_window_t* A, const mopoe_rows_t* dY);
size_t mopoe_conv_wgrad_ws_tc(const mopoe_window_t* A, const mopoe_rows_t* dY);
int mopoe_conv_wgrad_tc(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate, void* ws,
                        size_t ws_bytes, void* stream, const int* fin);

extern "C" int mopoe_conv_gemm(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D,
                               int impl, void* stream) {
    if (impl == 2 && !mopoe_tc_fwd_eligible(A, D)) MOPOE_FAIL("conv_gemm: tcgen05 path forced but problem not eligible");
    if (impl != 1 && mopoe_tc_fwd_eligible(A, D)) return mopoe_conv_gemm_tc(A, Wp, bias, D, stream);
    return mopoe_conv_gemm_simt(A, Wp, bias, D, stream);
}
extern "C" size_t mopoe_conv_wgrad_ws(const mopoe_window_t* A, const mopoe_rows_t* dY, int impl) {
    if (impl != 1 && mopoe_tc_wgrad_eligible(A, dY)) return mopoe_conv_wgrad_ws_tc(A, dY);
    return mopoe_conv_wgrad_ws_simt(A, dY);
}
extern "C" int mopoe_conv_wgrad(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate, void* ws,
                                size_t ws_bytes, int impl, void* stream) {
    if (impl == 2 && !mopoe_tc_wgrad_eligible(A, dY)) MOPOE_FAIL("conv_wgrad: tcgen05 path forced but problem not eligible");
    if (impl != 1 && mopoe_tc_wgrad_eligible(A, dY)) return mopoe_conv_wgrad_tc(A, dY, dWp, accumulate, ws, ws_bytes, stream, nullptr);
    return mopoe_conv_wgrad_simt(A, dY, dWp, accumulate, ws, ws_bytes, stream, nullptr);
}

// ---- batched fprop/dgrad: up to 4 problems of identical shape in one launch (sub-pixel phases) -------------------------
int mopoe_conv_gemm_tc_batched(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                               const mopoe_rows_t* D, void* stream);

extern "C" int mopoe_conv_gemm_batched(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                       const mopoe_rows_t* D, int impl, void* stream) {
    MOPOE_REQUIRE(nprob >= 1 && nprob <= 4, "conv_gemm_batched: nprob=%d (1..4)", nprob);
    bool tc = impl != 1;
    for (int i = 0; i < nprob && tc; ++i) tc = mopoe_tc_fwd_eligible(&A[i], &D[i]) != 0;
    if (impl == 2 && !tc) MOPOE_FAIL("conv_gemm_batched: tcgen05 path forced but problem not eligible");
    if (tc) return mopoe_conv_gemm_tc_batched(nprob, A, Wp, bias, D, stream);
    for (int i = 0; i < nprob; ++i)
        if (mopoe_conv_gemm_simt(&A[i], Wp[i], bias, &D[i], stream)) return 1;
    return 0;
}

// ---- weight gradient straight into the parameter's layout --------------------------------------------------------------
// grad[a][b][t] (+)= conv-form( sum_m dY[m, n=a] * A[m, (t, b')] ), b' < bpad: split reduction, re-layout and
// accumulation into the (flat) gradient buffer are ONE kernel after the GEMM.
extern "C" size_t mopoe_conv_wgrad_param_ws(const mopoe_window_t* A, const mopoe_rows_t* dY, int impl) {
    if (impl != 1 && mopoe_tc_wgrad_eligible(A, dY)) return mopoe_conv_wgrad_ws_tc_fin(A, dY);
    return mopoe_conv_wgrad_ws_simt_fin(A, dY);
}
extern "C" int mopoe_conv_wgrad_param(const mopoe_window_t* A, const mopoe_rows_t* dY, float* grad, int pa, int pb, int taps,
                                      int bpad, int accumulate, void* ws, size_t ws_bytes, int impl, void* stream) {
    const int fin[4] = {pa, pb, taps, bpad};
    if (impl == 2 && !mopoe_tc_wgrad_eligible(A, dY)) MOPOE_FAIL("conv_wgrad_param: tcgen05 path forced but problem not eligible");
    if (impl != 1 && mopoe_tc_wgrad_eligible(A, dY)) return mopoe_conv_wgrad_tc(A, dY, grad, accumulate, ws, ws_bytes, stream, fin);
    return mopoe_conv_wgrad_simt(A, dY, grad, accumulate, ws, ws_bytes, stream, fin);
}

// ---- fprop with the output's BatchNorm statistics fused into the epilogue ------------------------------------------------
struct TcStatsReq {
    double* ws;
    size_t ws_doubles;
    const uint8_t* mask;
    int mask_mode;
    int rows_per_b;
    int* nchunk_out;
};
struct TcResReq {
    const mopoe_rows_t* R;
    const float *mean, *invstd, *gamma, *beta;
    float a, b;
    const uint8_t* mask;
    int mask_mode;
    int dry_run;
    const mopoe_view_t* out;
    int kind;
};
int mopoe_conv_gemm_tc_batched_ex(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                  const mopoe_rows_t* D, const TcStatsReq* stats, void* stream, const TcResReq* res);
int mopoe_bn_finalize_launch(const double* ws, int nchunk, int C, double count, float eps, float momentum, float* mean,
                             float* invstd, float* rmean, float* rvar, void* stream);

extern "C" int mopoe_conv_gemm_bn(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                  const mopoe_rows_t* D, int impl, const mopoe_bn_req_t* bn, void* stream) {
    MOPOE_REQUIRE(nprob >= 1 && nprob <= 4, "conv_gemm_bn: nprob=%d (1..4)", nprob);
    MOPOE_REQUIRE(bn && bn->ws && bn->mean && bn->invstd, "conv_gemm_bn: null statistics request");
    MOPOE_REQUIRE(bn->out.C == D[0].N, "conv_gemm_bn: the output view has %d channels, the GEMM %d columns", bn->out.C, D[0].N);
    bool tc = impl != 1;
    for (int i = 0; i < nprob && tc; ++i) tc = mopoe_tc_fwd_eligible(&A[i], &D[i]) != 0;
    if (impl == 2 && !tc) MOPOE_FAIL("conv_gemm_bn: tcgen05 path forced but problem not eligible");
    int fused_chunks = 0;
    if (tc) {
        TcStatsReq req;
        req.ws = bn->ws;
        req.ws_doubles = (size_t)bn->ws_doubles;
        req.mask = bn->mask;
        req.mask_mode = bn->mask_mode;
        req.rows_per_b = (A[0].E2 > 1 || A[0].E1 > 1) ? A[0].E0 * A[0].E1 : bn->out.H * bn->out.W;
        req.nchunk_out = &fused_chunks;
        if (mopoe_conv_gemm_tc_batched_ex(nprob, A, Wp, bias, D, &req, stream, nullptr)) return 1;
    } else {
        for (int i = 0; i < nprob; ++i)
            if (mopoe_conv_gemm_simt(&A[i], Wp[i], bias, &D[i], stream)) return 1;
    }
    if (fused_chunks > 0)
        return mopoe_bn_finalize_launch(bn->ws, fused_chunks, bn->out.C, (double)bn->out.B * bn->out.H * bn->out.W, bn->eps,
                                        bn->momentum, bn->mean, bn->invstd, bn->running_mean, bn->running_var, stream);
    MOPOE_REQUIRE((long long)2 * bn->nchunk * bn->out.C <= bn->ws_doubles, "conv_gemm_bn: workspace too small for the fallback reduction");
    return mopoe_bn_stats(&bn->out, bn->mask, bn->mask_mode, bn->ws, bn->nchunk, bn->eps, bn->momentum, bn->mean, bn->invstd,
                          bn->running_mean, bn->running_var, nullptr, stream);
}

// ---- fprop with the block's residual combine (and the next BatchNorm's statistics) fused into the epilogue -----------------
static int res_call(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias, const mopoe_rows_t* D, int impl,
                    const mopoe_res_req_t* res, const mopoe_bn_req_t* bn, int dry, int* fused_chunks, void* stream) {
    if (impl == 1 || !res || !res->r) return 2;
    for (int i = 0; i < nprob; ++i)
        if (!mopoe_tc_fwd_eligible(&A[i], &D[i])) return 2;
    TcResReq rq;
    rq.R = res->r;
    rq.mean = res->mean; rq.invstd = res->invstd; rq.gamma = res->gamma; rq.beta = res->beta;
    rq.a = res->a; rq.b = res->b;
    rq.mask = res->mask; rq.mask_mode = res->mask_mode;
    rq.dry_run = dry;
    rq.out = res->out;
    rq.kind = 0;
    TcStatsReq sq;
    if (bn) {
        sq.ws = bn->ws;
        sq.ws_doubles = (size_t)bn->ws_doubles;
        sq.mask = nullptr;
        sq.mask_mode = bn->mask_mode;            // must be MOPOE_MASK_NONE: the next BatchNorm reads the block output as stored
        sq.rows_per_b = 1;
        sq.nchunk_out = fused_chunks;
    }
    return mopoe_conv_gemm_tc_batched_ex(nprob, A, Wp, bias, D, bn ? &sq : nullptr, stream, &rq);
}
extern "C" int mopoe_conv_gemm_res_eligible(int nprob, const mopoe_window_t* A, const float* bias, const mopoe_rows_t* D, int impl,
                                            const mopoe_res_req_t* res, const mopoe_bn_req_t* bn) {
    if (nprob < 1 || nprob > 4) return 0;
    int chunks = 0;
    return res_call(nprob, A, nullptr, bias, D, impl, res, bn, 1, &chunks, nullptr) == 0;
}
extern "C" int mopoe_conv_gemm_res(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                                   const mopoe_rows_t* D, int impl, const mopoe_res_req_t* res, const mopoe_bn_req_t* bn,
                                   void* stream) {
    MOPOE_REQUIRE(nprob >= 1 && nprob <= 4, "conv_gemm_res: nprob=%d (1..4)", nprob);
    MOPOE_REQUIRE(res && res->r && res->mean && res->invstd && res->gamma && res->beta, "conv_gemm_res: null residual request");
    if (bn) {
        MOPOE_REQUIRE(bn->ws && bn->mean && bn->invstd, "conv_gemm_res: null statistics request");
        MOPOE_REQUIRE(bn->out.C == D[0].N, "conv_gemm_res: the output view has %d channels, the GEMM %d columns", bn->out.C, D[0].N);
    }
    int chunks = 0;
    const int rc = res_call(nprob, A, Wp, bias, D, impl, res, bn, 0, &chunks, stream);
    if (rc == 2) MOPOE_FAIL("conv_gemm_res: the fused residual epilogue does not apply (ask mopoe_conv_gemm_res_eligible first)");
    if (rc) return 1;
    if (bn) {
        MOPOE_REQUIRE(chunks > 0, "conv_gemm_res: statistics were not produced");
        return mopoe_bn_finalize_launch(bn->ws, chunks, bn->out.C, (double)bn->out.B * bn->out.H * bn->out.W, bn->eps,
                                        bn->momentum, bn->mean, bn->invstd, bn->running_mean, bn->running_var, stream);
    }
    return 0;
}

// ---- input gradient with the following BatchNorm backward's per-channel sums fused into the epilogue ----------------------
int mopoe_sums_finalize_launch(const double* ws, int nchunk, int C, float* dbeta, float* dgamma, int accumulate, float* sums,
                               void* stream);
static int bnbwd_call(int nprob, const mopoe_window_t* A, const void* const* Wp, const mopoe_rows_t* D, int impl,
                      const mopoe_bnbwd_req_t* rq_in, int dry, int* fused_chunks, void* stream) {
    if (impl == 1 || !rq_in || !rq_in->x || !rq_in->ws) return 2;
    for (int i = 0; i < nprob; ++i)
        if (!mopoe_tc_fwd_eligible(&A[i], &D[i])) return 2;
    TcResReq rq;
    rq.R = rq_in->x;
    rq.mean = rq_in->mean; rq.invstd = rq_in->invstd; rq.gamma = rq_in->gamma; rq.beta = rq_in->beta;
    rq.a = rq.b = 0.f;
    rq.mask = nullptr; rq.mask_mode = MOPOE_MASK_NONE;
    rq.dry_run = dry;
    rq.out = nullptr;
    rq.kind = 1;
    TcStatsReq sq;
    sq.ws = rq_in->ws;
    sq.ws_doubles = (size_t)rq_in->ws_doubles;
    sq.mask = rq_in->mask;
    sq.mask_mode = rq_in->mask_mode;
    sq.rows_per_b = A[0].E0 * A[0].E1;           // Dropout2d mask: sample = E2 index
    sq.nchunk_out = fused_chunks;
    return mopoe_conv_gemm_tc_batched_ex(nprob, A, Wp, nullptr, D, &sq, stream, &rq);
}
extern "C" int mopoe_conv_gemm_bnbwd_eligible(int nprob, const mopoe_window_t* A, const mopoe_rows_t* D, int impl,
                                              const mopoe_bnbwd_req_t* req) {
    if (nprob < 1 || nprob > 4) return 0;
    int chunks = 0;
    return bnbwd_call(nprob, A, nullptr, D, impl, req, 1, &chunks, nullptr) == 0;
}
extern "C" int mopoe_conv_gemm_bnbwd(int nprob, const mopoe_window_t* A, const void* const* Wp, const mopoe_rows_t* D, int impl,
                                     const mopoe_bnbwd_req_t* req, void* stream) {
    MOPOE_REQUIRE(nprob >= 1 && nprob <= 4, "conv_gemm_bnbwd: nprob=%d (1..4)", nprob);
    MOPOE_REQUIRE(req && req->x && req->mean && req->invstd && req->gamma && req->beta && req->ws && req->sums,
                  "conv_gemm_bnbwd: null request fields");
    int chunks = 0;
    const int rc = bnbwd_call(nprob, A, Wp, D, impl, req, 0, &chunks, stream);
    if (rc == 2) MOPOE_FAIL("conv_gemm_bnbwd: the fused epilogue does not apply (ask mopoe_conv_gemm_bnbwd_eligible first)");
    if (rc) return 1;
    MOPOE_REQUIRE(chunks > 0, "conv_gemm_bnbwd: sums were not produced");
    return mopoe_sums_finalize_launch(req->ws, chunks, D[0].N, req->dbeta, req->dgamma, req->accumulate, req->sums, stream);
}

// ---- split-K for weight-bound problems (few output tiles, long reduction) ------------------------------------------------
size_t mopoe_conv_gemm_tc_splitk_ws(const mopoe_window_t* A, const mopoe_rows_t* D);
int mopoe_conv_gemm_tc_splitk(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* ws,
                              size_t ws_bytes, void* stream);
extern "C" size_t mopoe_conv_gemm_splitk_ws(const mopoe_window_t* A, const mopoe_rows_t* D, int impl) {
    if (impl == 1 || !mopoe_tc_fwd_eligible(A, D)) return 0;
    if (D->d_dtype != MOPOE_BF16 && D->d_dtype != MOPOE_F32) return 0;
    return mopoe_conv_gemm_tc_splitk_ws(A, D);
}
extern "C" int mopoe_conv_gemm_splitk(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* ws,
                                      size_t ws_bytes, int impl, void* stream) {
    MOPOE_REQUIRE(impl != 1 && mopoe_tc_fwd_eligible(A, D), "conv_gemm_splitk: tcgen05 path not eligible");
    return mopoe_conv_gemm_tc_splitk(A, Wp, bias, D, ws, ws_bytes, stream);
}
