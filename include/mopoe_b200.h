/* mopoe_b200.h — C ABI of libmopoe_b200.so, the sm_100a compute library behind the MoPoE-MIMIC
 * training step.
 *
 * The reference (Jimmy2027/MoPoE-MIMIC) is pure Python on torch.nn: it has NO FFI/plugin interface
 * for this path.  Each entry point below therefore replaces a torch library call
 * site of the reference, cited as file:line relative to /root/reference/mimic/.  The Python host
 * (mopoe_mimic_b200/*.py) mirrors the reference's module API and reaches this library with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer (inputs, outputs, workspaces) is owned by the
 *    caller (device memory from the PyTorch caching allocator); the library never allocates device
 *    memory, so the reference's 'CUDA out of memory.' contract (run_epochs.py:37-49) is untouched;
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t) and returns; it is re-entrant and
 *    may be called from the autograd worker thread;
 *  - return value 0 = ok, otherwise a non-zero code; mopoe_last_error() gives the text
 *    (thread-local);
 *  - activations are channels-last: [B, H, W, C] (1-D text: H = 1, W = L) described by a
 *    mopoe_view_t whose `ptr` addresses interior element (0,0,0,0); padded tensors carry a zero
 *    border of ph rows / pw columns that producing kernels write themselves.
 */
#ifndef MOPOE_B200_H
#define MOPOE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MOPOE_F32 = 0, MOPOE_BF16 = 1 };
/* dropout-mask addressing: none, one byte per (b,c) [Dropout2d], one byte per element in the
 * plain [B,H,W,C] order [Dropout].  Byte value 1 = keep (scaled by 2), 0 = drop. */
enum { MOPOE_MASK_NONE = 0, MOPOE_MASK_BC = 1, MOPOE_MASK_ELEM = 2 };

typedef struct {
    void*   ptr;        /* interior element (b=0,h=0,w=0,c=0) */
    int32_t dtype;      /* MOPOE_F32 | MOPOE_BF16 */
    int32_t B, H, W, C;
    int32_t ph, pw;     /* zero border (rows, cols) present around the interior in storage */
    int32_t _pad;
    int64_t sB, sH, sW; /* element strides; channel stride is 1 */
} mopoe_view_t;

/* implicit-GEMM operand description shared by the fprop/dgrad and wgrad entry points:
 *   A[m, r, k] = a[a_off + m0*sA0 + m1*sA1 + m2*sA2 + r*sAr + k],  m = (m2, m1, m0) over E2 x E1 x E0,
 *   r < R window rows, k < KW contiguous window elements (taps x channels).  K = R*KW. */
typedef struct {
    const void* a;
    int32_t a_dtype;
    int32_t E0, E1, E2;
    int32_t R, KW;
    int32_t _pad;
    int64_t a_off, sA0, sA1, sA2, sAr;
} mopoe_window_t;

/* row addressing of a GEMM output / wgrad "dY" operand: row m at d_off + m0*s0 + m1*s1 + m2*s2, n contiguous */
typedef struct {
    void*   d;
    int32_t d_dtype;
    int32_t N;
    int64_t d_off, s0, s1, s2;
} mopoe_rows_t;

const char* mopoe_last_error(void);
int mopoe_version(void);
/* 1 when the tcgen05/TMA kernels are usable on the current device (cc 10.x + driver entry point found) */
int mopoe_tc_available(void);
/* 1 when the tcgen05 weight-gradient kernel is compiled in */
int mopoe_tc_wgrad_built(void);

/* ---- implicit-GEMM convolution family -------------------------------------------------------------
 * D[m, n] = sum_{r,k} A[m,r,k] * Wp[n, r*KW + k] + bias[n]
 * Replaces nn.Conv{1,2}d / nn.ConvTranspose{1,2}d / nn.Linear forward and their input-gradient
 * (networks/ResidualBlocks.py:8-16,71-80,103-111; FeatureExtractorImg.py:29-34; DataGeneratorImg.py:10-16;
 * FeatureCompressor.py:13-19; char_encoding/FeatureExtractorText.py:30-31, DataGeneratorText.py:44-49).
 * Wp has the dtype of A.  impl: 0 = auto (tcgen05 when eligible), 1 = force SIMT fp32-accumulate path. */
int mopoe_conv_gemm(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D,
                    int impl, void* stream);

/* Same contraction for up to 4 problems of IDENTICAL shape (E0,E1,E2,R,KW,N, output tensor and row strides) in one
 * launch — the 2^nd sub-pixel phases of a stride-2 nn.ConvTranspose{1,2}d (ResidualBlocks.py:44-46,108-110) or of a
 * strided conv's input gradient: they differ only in window origin, weight slice and output origin.  On the tcgen05
 * path this is a persistent kernel (one CTA per SM, double-buffered TMEM accumulators). */
int mopoe_conv_gemm_batched(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias,
                            const mopoe_rows_t* D, int impl, void* stream);

/* Split-K variant of mopoe_conv_gemm for weight-bound problems — few output tiles and a long reduction (the 4x4 -> 1x1
 * convolutions at the bottom of the image stacks and their gradients, FeatureExtractorImg.py / DataGeneratorImg.py: M = batch
 * rows, K = 8192-10240): every SM reduces a slice of K into an fp32 partial tile in `ws`, a finish kernel sums the slices in
 * a fixed order (deterministic), adds the bias and writes D.  mopoe_conv_gemm_splitk_ws returns the workspace bytes, or 0
 * when the problem is not of that kind (then use mopoe_conv_gemm). */
size_t mopoe_conv_gemm_splitk_ws(const mopoe_window_t* A, const mopoe_rows_t* D, int impl);
int mopoe_conv_gemm_splitk(const mopoe_window_t* A, const void* Wp, const float* bias, const mopoe_rows_t* D, void* ws,
                           size_t ws_bytes, int impl, void* stream);

/* The same launch with the training-mode BatchNorm STATISTICS of its output fused into the epilogue: the GEMM that
 * produces a tensor also produces the per-channel mean / 1/sqrt(var + eps) the following nn.BatchNorm needs
 * (ResidualBlocks.py:84-97: conv1 -> dropout1 -> bn2, and shortcut conv -> BatchNorm), and updates the running statistics
 * (momentum, unbiased variance) — instead of a separate reduction pass over the activation.  `out` describes the whole
 * tensor the launch writes (all problems together cover every pixel once); mask / mask_mode: the dropout keep-mask that
 * sits between the GEMM and the BatchNorm (statistics of x * 2 * mask).  The statistics are those of the STORED (bf16-
 * rounded) values.  ws: >= max(2 * nchunk, 8 * #SMs) * out.C doubles.  When the fused epilogue does not apply (fp32 /
 * SIMT path, unaligned rows, masked multi-problem launches) the library runs mopoe_bn_stats itself: same results. */
typedef struct {
    mopoe_view_t out;
    const uint8_t* mask;
    int32_t mask_mode;
    int32_t nchunk;
    double* ws;
    int64_t ws_doubles;
    float eps, momentum;
    float* mean;
    float* invstd;
    float* running_mean;
    float* running_var;
} mopoe_bn_req_t;
int mopoe_conv_gemm_bn(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias, const mopoe_rows_t* D,
                       int impl, const mopoe_bn_req_t* bn, void* stream);

/* The same launch with the block's RESIDUAL COMBINE fused into the epilogue: the GEMM is the block's conv2
 * (ResidualBlocks.py:92-93 `out = self.conv2(out); out = self.dropout2(out)`), and instead of its result it stores the
 * block output `self.a * residual + self.b * out` (ResidualBlocks.py:94-96) with the shortcut's BatchNorm folded in:
 *     D[m, n] = a * BN(r[m, n]) + b * ((acc[m, n] + bias[n]) * 2mask)      (mask_mode NONE: no factor 2, no mask)
 * acc is the fp32 accumulator: the conv2 result never goes through HBM (mopoe_conv_gemm + mopoe_combine write it, read it
 * back, and launch twice).  r: one row addressing per problem over the shortcut branch (bf16, N columns), same geometry as
 * D.  mask: MOPOE_MASK_BC = [E2, N] keep-bytes (Dropout2d, E2 the batch), MOPOE_MASK_ELEM = one byte per element of r,
 * laid out like r.  bn (may be NULL): also the training-mode statistics of the STORED output — the next block's bn1 —
 * with bn->mask_mode == MOPOE_MASK_NONE.  out (may be NULL): the activation D addresses — the launch then also writes its
 * zero border (idle warps of the same kernel); with out == NULL the caller runs mopoe_zero_border.
 * mopoe_conv_gemm_res_eligible answers 1 when this epilogue applies (tcgen05 path, bf16, N % 64 == 0, 16-byte aligned rows,
 * MOPOE_GEMM_RES != 0); otherwise the caller runs mopoe_conv_gemm(_batched) + mopoe_combine(_bn); mopoe_conv_gemm_res fails
 * on a problem that is not eligible. */
typedef struct {
    const mopoe_rows_t* r;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
    float a, b;
    const uint8_t* mask;
    int32_t mask_mode;
    const mopoe_view_t* out;
} mopoe_res_req_t;
int mopoe_conv_gemm_res_eligible(int nprob, const mopoe_window_t* A, const float* bias, const mopoe_rows_t* D, int impl,
                                 const mopoe_res_req_t* res, const mopoe_bn_req_t* bn);
int mopoe_conv_gemm_res(int nprob, const mopoe_window_t* A, const void* const* Wp, const float* bias, const mopoe_rows_t* D,
                        int impl, const mopoe_res_req_t* res, const mopoe_bn_req_t* bn, void* stream);

/* The same launch as an INPUT GRADIENT whose result dy feeds a BatchNorm(+ReLU, +dropout) backward (autograd's
 * conv2 -> relu -> bn2 -> dropout1 chain of ResidualBlocks.py:88-92, run_epochs.py:130 total_loss.backward()): the
 * statistics warps of the epilogue read the matching tile of the BatchNorm's input x next to every stored tile of dy and
 * produce the two per-channel sums of the BatchNorm backward — sums[0][n] = sum_m g, sums[1][n] = sum_m g * xhat with
 * g = dy * [relu(BN(x * 2mask)) > 0] (gate recomputed bit-exactly, as in mopoe_bn_bwd_reduce with gate_gamma / gate_beta) and
 * xhat = (x * 2mask - mean) * invstd — plus dgamma (+)= sums[1], dbeta (+)= sums[0].  This replaces the mopoe_bn_bwd_reduce
 * pass (two reads of the activation); mopoe_bn_bwd_apply then runs as usual on the stored dy.  x: one row addressing per
 * problem over the BatchNorm's input (bf16, N columns), same geometry as D.  mask: MOPOE_MASK_BC [E2, N] for any nprob,
 * MOPOE_MASK_ELEM ([rows, N] in D's flat row order) for nprob == 1.  ws: >= 8 * #SMs * N doubles.  The sums are those of
 * the STORED (bf16-rounded) dy.  mopoe_conv_gemm_bnbwd_eligible answers 1 when this epilogue applies (MOPOE_GEMM_BNB != 0). */
typedef struct {
    const mopoe_rows_t* x;
    const uint8_t* mask;
    int32_t mask_mode;
    int32_t accumulate;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
    double* ws;
    int64_t ws_doubles;
    float* dgamma;
    float* dbeta;
    float* sums;
} mopoe_bnbwd_req_t;
int mopoe_conv_gemm_bnbwd_eligible(int nprob, const mopoe_window_t* A, const mopoe_rows_t* D, int impl,
                                   const mopoe_bnbwd_req_t* req);
int mopoe_conv_gemm_bnbwd(int nprob, const mopoe_window_t* A, const void* const* Wp, const mopoe_rows_t* D, int impl,
                          const mopoe_bnbwd_req_t* req, void* stream);

/* dWp[n, r*KW + k] (+)= sum_m dY[m, n] * A[m,r,k]   (fp32 output).  `ws` holds split partials
 * (ws_bytes from mopoe_conv_wgrad_ws); replaces the weight-gradient half of autograd's conv backward
 * (run_epochs.py:130 total_loss.backward()). */
size_t mopoe_conv_wgrad_ws(const mopoe_window_t* A, const mopoe_rows_t* dY, int impl);
int mopoe_conv_wgrad(const mopoe_window_t* A, const mopoe_rows_t* dY, float* dWp, int accumulate,
                     void* ws, size_t ws_bytes, int impl, void* stream);

/* Weight gradient delivered in the PARAMETER's layout: grad[a][b][t] (+)= sum_m dY[m, a] * A[m, (t, b')]  (b' < bpad,
 * the window's channel padding).  The split-K reduction, the conv-form -> [a][b][taps] re-layout and the accumulation
 * into the caller's (flat) gradient buffer are one kernel after the GEMM.  ws_bytes from mopoe_conv_wgrad_param_ws. */
size_t mopoe_conv_wgrad_param_ws(const mopoe_window_t* A, const mopoe_rows_t* dY, int impl);
int mopoe_conv_wgrad_param(const mopoe_window_t* A, const mopoe_rows_t* dY, float* grad, int pa, int pb, int taps,
                           int bpad, int accumulate, void* ws, size_t ws_bytes, int impl, void* stream);

/* out[c] (+)= sum over rows of v[.., c]  (bias gradients).  ws: 2*nchunk*C doubles. */
int mopoe_colsum(const mopoe_view_t* v, float* out, int accumulate, double* ws, int nchunk, int* counters,
                 void* stream);

/* ---- BatchNorm (training statistics) + fused pre-activation ----------------------------------------
 * nn.BatchNorm{1,2}d in train mode (ResidualBlocks.py:8,12,73-75,...): per-channel mean / biased var
 * of v = x * (2*mask); running stats updated with momentum and the unbiased var.
 * ws: 2*nchunk*C doubles.  running_* may be NULL (no update).
 * counters (may be NULL): >= ceil(C/128) zero-initialised ints; when given, the last-arriving block of each channel
 * group finalises in the same launch (fixed summation order -> deterministic) and resets its counter. */
int mopoe_bn_stats(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, double* ws, int nchunk,
                   float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, int* counters, void* stream);
/* out = act(gamma * (x*2mask - mean) * invstd + beta), act = relu when relu != 0; writes the zero border
 * of `out`.  bn1->relu / dropout1->bn2->relu of ResidualBlocks.py:20-33 in one pass. */
int mopoe_bn_apply(const mopoe_view_t* x, const uint8_t* mask, int mask_mode, const float* mean,
                   const float* invstd, const float* gamma, const float* beta, int relu,
                   const mopoe_view_t* out, void* stream);
/* out = a * BN(r) + b * (c * 2mask)   — `out = self.a * residual + self.b * out` with the shortcut's
 * BatchNorm and dropout2 folded in (ResidualBlocks.py:29-33). */
int mopoe_combine(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, const mopoe_view_t* c, const uint8_t* mask, int mask_mode,
                  float a, float b, const mopoe_view_t* out, void* stream);
/* mopoe_combine + the training-mode BatchNorm statistics of `out` (mean, 1/sqrt(var+eps), running statistics): the bn1 of the
 * NEXT residual block reads exactly this tensor (ResidualBlocks.py:84-86), so its statistics pass folds into this one.
 * Statistics are those of the STORED (storage-dtype-rounded) values.  ws: 2*nchunk*C doubles. */
int mopoe_combine_bn(const mopoe_view_t* r, const float* mean, const float* invstd, const float* gamma,
                     const float* beta, const mopoe_view_t* c, const uint8_t* mask, int mask_mode, float a, float b,
                     const mopoe_view_t* out, double* ws, int nchunk, float eps, float momentum, float* out_mean,
                     float* out_invstd, float* running_mean, float* running_var, void* stream);
/* BN backward, reduction half: g = gscale * dy * [gate > 0];  xhat = (x*2mask - mean)*invstd;
 * dbeta (+)= sum g, dgamma (+)= sum g*xhat, sums[0:C] = sum g, sums[C:2C] = sum g*xhat.
 * The ReLU gate is read from `gate` (the saved activation) or absent (NULL).  gate_gamma / gate_beta (optional): the
 * affine parameters of THIS BatchNorm when `gate` = relu(gamma*xhat + beta) is its own output — the library may then
 * RECOMPUTE the gate from x with the forward pass's own instruction sequence (bit-identical decisions) instead of reading
 * the activation: a third less traffic, same result.  NULL: always read `gate`. */
int mopoe_bn_bwd_reduce(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                        const mopoe_view_t* x, const uint8_t* mask, int mask_mode,
                        const float* mean, const float* invstd, double* ws, int nchunk,
                        float* dgamma, float* dbeta, int accumulate, float* sums,
                        const float* gate_gamma, const float* gate_beta, int* counters, void* stream);
/* BN backward, apply half: out = gamma*invstd*(g - sums_g/cnt - xhat*sums_gx/cnt) * 2mask + addend.
 * gate_beta (optional, with `gate`): the BatchNorm's bias -> the gate may be recomputed from x (see above). */
int mopoe_bn_bwd_apply(const mopoe_view_t* dy, const mopoe_view_t* gate, float gscale,
                       const mopoe_view_t* x, const uint8_t* mask, int mask_mode,
                       const float* mean, const float* invstd, const float* gamma, const float* sums,
                       const mopoe_view_t* addend, const mopoe_view_t* out, const float* gate_beta, void* stream);
/* backward of `y = a*BN(r) + b*(c*2mask2)` in ONE pass over dy: dr = BN-backward of g = a*dy (sums from
 * mopoe_bn_bwd_reduce), dc = b * dy * 2mask2; dr and dc must share their border widths. */
int mopoe_combine_bwd_apply(const mopoe_view_t* dy, float a, const mopoe_view_t* r, const float* mean,
                            const float* invstd, const float* gamma, const float* sums,
                            const uint8_t* mask2, int mask2_mode, float b, const mopoe_view_t* dr,
                            const mopoe_view_t* dc, void* stream);
/* out = scale * dy * 2mask (+ zero border): dropout backward / the b*dropout2 branch. */
int mopoe_scale_mask(const mopoe_view_t* dy, const uint8_t* mask, int mask_mode, float scale,
                     const mopoe_view_t* out, void* stream);
/* generic layout/dtype conversion between views of equal B,H,W; channels beyond src->C are zero filled.
 * src_nchw != 0: src->ptr is an NCHW fp32 user tensor [B,C,H,W] (strides ignored). */
int mopoe_convert(const mopoe_view_t* src, int src_nchw, const mopoe_view_t* dst, void* stream);
/* Bernoulli(0.5) keep-mask bytes from a counter-based generator (Philox-4x32-10). */
/* step_ptr (may be NULL): device counter of the training step, mixed into the Philox counter so a captured
 * CUDA graph draws fresh masks on every replay. */
int mopoe_dropout_mask(uint8_t* mask, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_ptr,
                       void* stream);
/* fp32 master weight W[A][B][KH][KW] -> packed GEMM operand (dst_dtype): form 0 conv-form [A, KH*KW*bpad],
 * 1 phase-form (py,px) [B, taps*A], 2 full-form [KH*KW*B, A], 3 [A,B], 4 [B,A]  (layouts in DESIGN.md). */
int mopoe_pack_weight(const float* W, int A, int B, int KH, int KW, int form, int py, int px, int bpad,
                      void* dst, int dst_dtype, void* stream);
/* same re-layouts through a shared-memory tile (sector-efficient reads AND writes); form 1 fills dsts[0..3]
 * (2-D) / dsts[0..1] (1-D) with all sub-pixel phases in one launch; form 3 = [A,B], form 4 = [B,A] for 1x1 kernels. */
int mopoe_pack_weight_tiled(const float* W, int A, int B, int KH, int KW, int form, int bpad, void* const* dsts,
                            int dst_dtype, void* stream);
/* advance the device-side step state (dropout step counter, Adam step + bias-correction coefficients) */
int mopoe_step_advance(uint64_t* rng_step, int32_t* adam_step, float* adam_coef, float lr, float beta1,
                       float beta2, void* stream);

/* ---- single-channel image layers (CUDA-core direct kernels) ---------------------------------------
 * first conv  nn.Conv2d(1, C, 3, stride 2, pad 1, bias=False)  (FeatureExtractorImg.py:29-34)
 * x: fp32 [B, 1, H, W]; w: fp32 [C, 1, 3, 3]; out view [B, H/2, W/2, C]. */
int mopoe_conv3x3s2_c1_fwd(const float* x, const float* w, int B, int H, int W, const mopoe_view_t* out,
                           void* stream);
int mopoe_conv3x3s2_c1_wgrad(const float* x, const mopoe_view_t* dy, int B, int H, int W, float* dw,
                             int accumulate, double* ws, int nchunk, void* stream);
/* last deconv nn.ConvTranspose2d(C, 1, 3, stride 2, pad 1, output_padding 1) (DataGeneratorImg.py:84-90)
 * x view [B, H, W, C]; w fp32 [C,1,3,3]; out fp32 [B,1,2H,2W].  ws (mopoe_deconv3x3s2_c1_fwd_ws bytes, 16-B aligned):
 * per-pixel tap products of the two-phase form; ws == NULL selects the single-kernel form. */
size_t mopoe_deconv3x3s2_c1_fwd_ws(const mopoe_view_t* x);
int mopoe_deconv3x3s2_c1_fwd(const mopoe_view_t* x, const float* w, const float* bias, float* out, void* ws,
                             size_t ws_bytes, void* stream);
/* given dout fp32 [B,1,2H,2W]: dx view, dw fp32 [C,1,3,3], dbias[1]. ws: nchunk*(9*C+1) doubles. */
int mopoe_deconv3x3s2_c1_bwd(const mopoe_view_t* x, const float* w, const float* dout, const mopoe_view_t* dx,
                             float* dw, float* dbias, int accumulate, double* ws, int nchunk, void* stream);
/* 3x3 / stride-2 / pad-1 patches of a single-channel fp32 image [B, SH, SW] as a bf16 matrix [B*(SH/2)*(SW/2), 16]
 * (column t = ky*3 + kx, columns 9..15 zero).  The weight gradients of the two single-channel layers are then
 * mopoe_conv_wgrad launches (activation = window operand, patches = 16-wide row operand); mopoe_deconv3x3s2_c1_bwd accepts
 * dw = NULL and dx = NULL for these cases. */
int mopoe_im2col3x3s2(const float* src, int B, int SH, int SW, int cols, void* out_bf16, void* stream);
/* (cols: 16, or 64 = one k-block of the tcgen05 GEMMs — the patches are then also the A operand of the layer itself:
 * first conv forward and last deconv input gradient are out[m, c] = sum_t patches[m, t] * w[c, t], mopoe_conv_gemm launches.)
 * The last deconv's forward as a GEMM: taps[m, t] = sum_c x[m, c] * w[c, t] (16 zero-padded filter rows, fp32 [M, stride])
 * followed by the assembly of the 2x2 output quads (+ bias): */
int mopoe_deconv3x3s2_c1_assemble(const float* taps, int stride, const float* bias, float* out, int B, int H, int W,
                                  void* stream);
/* zero the border of a bordered channels-last activation whose producer wrote only the interior (a GEMM) */
int mopoe_zero_border(const mopoe_view_t* v, void* stream);

/* ---- fused MoPoE kernel (north_star item 2) --------------------------------------------------------
 * BaseMMVae.inference (utils/BaseMMVae.py:139-196) + poe (mm_div.py:10-17) + mixture_component_selection
 * (utils/utils.py:55-77) + reparameterize (:45-48) + calc_kl_divergence for every subset (kl_div.py:8-16).
 * All tensors fp32.  mu/logvar: M pointers to [B, D].  Subset s has member bitmask members[s] (bit i =
 * modality i, in the caller's modality order).  fuse_mode 0 = product of experts, 1 = mixture (batch-range
 * selection among members, uniform weights).  prior_expert != 0 appends N(0,I) to every product (poe method).
 * stacked[j] (j < S) are the subsets forming the joint mixture (-1 = the N(0,I) prior component of jsd mode:
 * joint mu = logvar = 0, z = eps on its rows); sel_end[j] their exclusive batch-row ends
 * (host-computed with the reference's fp32 floor rule).
 * Outputs: sub_mu/sub_lv [nsub, B, D]; joint_mu/joint_lv/z [B, D]; kl[nsub] = KL(subset || N(0,I)) / norm;
 * nan_flag[0] != 0 when an encoder mean/logvar is NaN (utils.check_latents, utils/utils.py:201-208).
 * ws: nsub * kl_chunks doubles. */
typedef struct {
    int32_t M, B, D, nsub, S;
    int32_t fuse_mode, prior_expert, kl_chunks;
    int32_t members[16];
    int32_t stacked[16];
    int32_t sel_end[16];
    int32_t mem_cnt[16];    /* number of members of subset s */
    int32_t mem_idx[16][4]; /* members of subset s in the reference's stacking order (sorted by name) */
    int32_t mem_end[16][4]; /* mixture mode: exclusive batch-row end of the j-th stacked member */
    float   norm;
    float   _pad;
} mopoe_fusion_cfg_t;
int mopoe_fusion_fwd(const mopoe_fusion_cfg_t* cfg, const float* const* mu, const float* const* logvar,
                     const float* eps, float* sub_mu, float* sub_lv, float* joint_mu, float* joint_lv,
                     float* z, float* kl, int32_t* nan_flag, double* ws, void* stream);
/* Backward: upstream d_z [B,D], d_joint_mu/d_joint_lv [B,D], d_sub_mu/d_sub_lv [nsub,B,D], d_kl [nsub]
 * (any may be NULL = zero) -> d_mu/d_lv: M pointers to [B, D] (overwritten). */
int mopoe_fusion_bwd(const mopoe_fusion_cfg_t* cfg, const float* const* mu, const float* const* logvar,
                     const float* eps, const float* sub_mu, const float* sub_lv,
                     const float* d_z, const float* d_joint_mu, const float* d_joint_lv,
                     const float* d_sub_mu, const float* d_sub_lv, const float* d_kl,
                     float* const* d_mu, float* const* d_lv, void* stream);

/* ---- fused reconstruction log-likelihoods (north_star item 3) --------------------------------------
 * Laplace(loc, scale).log_prob(x).sum()  (modalities/Modality.py:25-30, scale 0.75 from
 * networks/ConvNetworksImgMimic.py:54): out[0] = sum_i -(log(2*scale) + |x_i - loc_i| / scale).
 * ws: nchunk doubles. */
int mopoe_laplace_logprob_sum(const float* loc, const float* x, int64_t n, float scale, float* out,
                              double* ws, int nchunk, void* stream);
/* dloc_i = (*gout) * sign(x_i - loc_i) / scale   (gradient of the log-prob sum; gout on device) */
int mopoe_laplace_logprob_bwd(const float* loc, const float* x, int64_t n, float scale, const float* gout,
                              float* dloc, void* stream);
/* out_i = -(log(2*scale) + |x_i - loc_i| / scale): the elementwise density the evaluation callers reduce per sample
 * (utils/likelihood.py:120-121) */
int mopoe_laplace_logprob_elem(const float* loc, const float* x, int64_t n, float scale, float* out, void* stream);
/* LogSoftmax(dim=vocab) of the text decoder (char_encoding/DataGeneratorText.py:51,75) fused with
 * OneHotCategorical(logits).log_prob(target).sum() (modalities/utils.py:7-8, MimicText.py:37-40).
 * y: pre-softmax [rows, V] fp32 (rows = B*L); target: ONE-HOT fp32 [rows, V] (the argmax is taken, first maximum wins:
 * rows that are not strictly one-hot — all-zero or soft targets — deviate from OneHotCategorical.log_prob's
 * sum(target * logits); the reference's data is always one-hot) or NULL with idx int32 [rows].  logits_out [rows, V] =
 * log_softmax(y) (may be NULL); idx_out (may be NULL) receives the argmax; lse_out (may be NULL) the row logsumexp
 * for the flat backward — filled only when mopoe_categorical_has_lse(y, target, V) is 1 (V <= 256, 16-byte aligned
 * buffers: the staged-row kernel), NaN otherwise; out[0] = sum_rows logits[row, idx]. */
int mopoe_categorical_logprob_sum(const float* y, const float* target, const int32_t* idx, int64_t rows,
                                  int V, float* logits_out, int32_t* idx_out, float* lse_out, float* out, double* ws,
                                  int nchunk, void* stream);
int mopoe_categorical_has_lse(const float* y, const float* target, int V);
/* dy[row, v] = (*gout) * (onehot(idx)[v] - softmax(y)[row, v]).  lse: the row logsumexp saved by the forward (one flat
 * elementwise pass) or NULL (the row softmax is recomputed). */
int mopoe_categorical_logprob_bwd(const float* y, const int32_t* idx, const float* lse, int64_t rows, int V,
                                  const float* gout, float* dy, void* stream);

/* ---- optimizer: torch.optim.Adam defaults (experiment.py:171-178) over the FLAT
 * parameter / gradient / moment buffers (all 420 tensors are views into one allocation): one launch.
 * `step` is the 1-based step count; g is multiplied by grad_scale first (1/world_size under DP). */
int mopoe_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, int step, float grad_scale, void* stream);

int mopoe_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* coef,
                        float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- alpha-JSD divergence with a dynamic prior (jsd mode) -----------------------------------------
 * BaseMMVae.divergence_dynamic_prior (utils/BaseMMVae.py:87-99) -> calc_alphaJSD_modalities (mm_div.py:67-87):
 * dynamic prior = alpha_poe of the K stacked experts (mm_div.py:20-32; the N(0,I) prior is passed as one of them),
 * kl[k] = KL(expert k || dynamic prior) / norm (kl_div.py:11-13).  mu / logvar: K device pointers to [B, D] fp32; alpha: K
 * HOST floats; ws: K*B doubles.  bwd: d_kl[K] (device) -> d_mu[k], d_lv[k] ([B, D] each, overwritten). */
int mopoe_jsd_divergence_fwd(int K, int B, int D, const float* const* mu, const float* const* logvar, const float* alpha,
                             float norm, float* dyn_mu, float* dyn_lv, float* kl, double* ws, void* stream);
int mopoe_jsd_divergence_bwd(int K, int B, int D, const float* const* mu, const float* const* logvar, const float* alpha,
                             float norm, const float* d_kl, float* const* d_mu, float* const* d_lv, void* stream);

/* Token indices (fp32 values, as the reference ships them; clamped to [0, V)) -> one-hot rows [rows, Vp] (Vp >= V,
 * multiple of 8) in MOPOE_F32 / MOPOE_BF16, plus the int32 indices: nn.Embedding (word_encoding/mmvae_text_enc.py:27-28,69)
 * becomes a GEMM of these rows with the embedding matrix, its gradient the matching weight-gradient GEMM. */
int mopoe_onehot(const float* idx, int64_t rows, int V, int Vp, void* out, int out_dtype, int32_t* idx_out, void* stream);
/* uint8 character indices [rows] -> fp32 one-hot rows [rows, V] (V <= 256): the device side of the 1-byte-per-token wire
 * format; the reference builds these rows on the host (dataio/MimicDataset.py:92-96, utils/text.py:13-34). */
int mopoe_onehot_u8(const uint8_t* idx, int64_t rows, int V, float* out, void* stream);
/* The first layer of the character-text encoder, nn.Conv1d(V, C, 4, 2, 1) on one-hot rows
 * (char_encoding/FeatureExtractorText.py:30-31, :71-72), as a GATHER over the byte indices [B, L] of the wire format:
 *   out[b, l, :] = bias + sum_{t<4, 0 <= 2l-1+t < L} W[:, idx[b, 2l-1+t], t]
 * table: the layer's weights in full form [(t*V + v), c] in the activation dtype (mopoe_pack_weight_tiled form 2);
 * out: [B, 1, L/2, C] (bordered allowed; only the interior is written).  An index >= V contributes nothing.
 * mopoe_text_onehot_act builds the one-hot rows as the bordered, channel-padded activation [B, 1, L, Vp] (border pw, zero
 * border and zero padding channels included) that the layer's weight-gradient GEMM reads, from the same indices. */
int mopoe_text_stem_gather_fwd(const uint8_t* idx, int B, int L, int V, const void* table, int dtype, const float* bias,
                               const mopoe_view_t* out, void* stream);
int mopoe_text_onehot_act(const uint8_t* idx, int B, int L, int V, const mopoe_view_t* out, void* stream);
/* 8-bit images on the wire: dst[i] = float(src[i]) / 255 — torchvision ToTensor(), which the reference's loader
 * applies on the host (dataio/MimicDataset.py), evaluated on the device: 1 byte per pixel crosses PCIe instead of 4. */
int mopoe_u8_to_unit(const uint8_t* src, int64_t n, float* dst, void* stream);

/* All weight re-layouts of a step in ONE launch.  jobs_dev: DEVICE array of njobs descriptors (same meaning as the
 * arguments of mopoe_pack_weight_tiled; form 1 fills dst[0..3] / dst[0..1], the others dst[0]); tile0 = index of the
 * job's first tile (a run of 256 work items = one thread block) in the launch grid; mopoe_pack_job_tiles returns the
 * job's tile count (nx: unused, set to 0).  Jobs must be sorted by tile0.  total_tiles = sum of the tile counts.  Replaces nothing in the reference (torch.nn consumes its
 * fp32 weights in place); it exists because the tcgen05 kernels want K-major bf16 operands. */
typedef struct {
    const float* W;
    void*   dst[4];
    int32_t A, B, KH, KW, form, bpad, tile0, nx;
} mopoe_pack_job_t;
int mopoe_pack_job_tiles(int A, int B, int KH, int KW, int form, int bpad, int* nx);
int mopoe_pack_weights_batched(const mopoe_pack_job_t* jobs_dev, int njobs, int total_tiles, int dst_dtype, void* stream);

/* ---- data-parallel exchange (SURVEY §8 a25 / e): gradient reduce-scatter + Adam + parameter all-gather as ONE kernel
 * over NVLink peer memory.  Replaces DistributedDataParallel's all-reduce followed by optimizer.step()
 * (main_mimic.py:44-48, utils/utils.py:179-185, run_epochs.py:130-131).
 * peers: for every rank r, the address (mapped into THIS process: CUDA IPC / symmetric memory) of its flat gradient
 * buffer, its flat parameter buffer and its flag array (>= 2*world uint32, zero-initialised before the first call).
 * Rank `rank` sums slice [rank*ceil(n/4/world)*4, ...) of all gradient buffers in rank order, multiplies by grad_scale,
 * applies Adam (moments m, v: local, only the own slice is touched) and stores the new parameters into every rank's
 * parameter buffer.  state: 4 local uint32 {epoch (initialise to 1), 0, error, 0}.  coef as mopoe_adam_flat_dev.  Every
 * rank must enqueue the same sequence of calls; the kernel completes only when all peers are done with this rank's
 * buffers.  CUDA-graph capturable (nothing step-dependent is a launch argument).  The flag waits are bounded by
 * MOPOE_DP_TIMEOUT_S seconds (environment, default 600, 0 = unbounded); an expired wait never traps: it stores
 * 1 + 16*barrier + missing_rank into state[2] and returns, and the caller must treat the step as failed.
 * mc_grad / mc_param (both or neither): NVSwitch multicast addresses of the gradient / parameter buffers.  When given,
 * the sum is formed inside the switch (multimem.ld_reduce) and the parameters are broadcast by it (multimem.st); the
 * summation order is then the switch's, not rank order. */
typedef struct {
    const float* grad[16];
    float*       param[16];
    uint32_t*    flags[16];
} mopoe_dp_peers_t;
int mopoe_dp_adam_exchange(const mopoe_dp_peers_t* peers, const float* mc_grad, float* mc_param, float* m, float* v,
                           int64_t n, int rank, int world, uint32_t* state, const float* coef, float beta1,
                           float beta2, float eps, float grad_scale, void* stream);
/* The same on at most max_blocks thread blocks (0 = no limit).  A training step exchanges its gradients in BUCKETS — the
 * decoders' slice of the flat buffers as soon as the decoders' backward pass is done, under the encoders' backward (DDP's
 * bucketed overlap, run_epochs.py:245-247); every bucket is one call on its own sub-range (pointers offset by the bucket's
 * start), with its own flags / state, and the overlapped one runs on a small grid. */
int mopoe_dp_adam_exchange_ex(const mopoe_dp_peers_t* peers, const float* mc_grad, float* mc_param, float* m, float* v,
                              int64_t n, int rank, int world, uint32_t* state, const float* coef, float beta1,
                              float beta2, float eps, float grad_scale, int max_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif
