/* b200at_model -- C ABI of the memory-bound ConvNeXt-CvSt layer kernels (libb200at.so).
 *
 * The reference runs these layers as eager torch ops (timm ConvNeXtBlock == the vendored
 * /root/reference/models/convnext.py:37-50; stem/downsample LayerNorm = utils_architecture.py:57-81);
 * each entry point cites what it replaces.  Conventions as in b200at.h: raw device pointers, caller's
 * stream, no allocation / synchronisation, returns the cudaError_t of the launch.
 * Activations: NHWC, bf16 (`void*`), M = B*H*W pixel rows of C channels.  Parameters and statistics: fp32.
 */
#ifndef B200AT_MODEL_H
#define B200AT_MODEL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* F.layer_norm over C per pixel (models/convnext.py:29,41 `self.norm`; utils_architecture.py:76-81 on NHWC
 * memory), optionally followed by GELU (the stems' LN -> GELU, utils_architecture.py:205-211).
 * Saves mean/rstd [M] for the backward.  C % 4 == 0, C <= 1536. */
int b200at_ln_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t M,
                  int64_t C, float eps, int fuse_gelu, void* stream);

/* input gradient of the above (autograd of F.layer_norm [+ GELU]); if dw/db are non-null also ACCUMULATES
 * the gamma/beta gradients into them (fp32 [C], caller zeroes). */
int b200at_ln_bwd(const void* dy, const void* x, const float* w, const float* b, const float* mean, const float* rstd,
                  void* dx, float* dw, float* db, int64_t M, int64_t C, int fuse_gelu, void* stream);

/* The same two with a per-channel bias added to x first: y = LN(x + pre_bias) [-> GELU].  This is how the CvSt stem's
 * `Conv2d(..., bias=True)` -> LayerNorm (utils_architecture.py:205-211) runs when the convolution is a library call
 * without its bias: the bias add (a full extra pass over the stem's activation, the largest of the network) rides in the
 * LayerNorm kernel; the convolution's bias gradient is the column sum of dx (b200at_colsum_bf16).  pre_bias: fp32 [C] or
 * null. */
int b200at_ln_fwd_bias(const void* x, const float* pre_bias, const float* w, const float* b, void* y, float* mean,
                       float* rstd, int64_t M, int64_t C, float eps, int fuse_gelu, void* stream);
int b200at_ln_bwd_bias(const void* dy, const void* x, const float* pre_bias, const float* w, const float* b,
                       const float* mean, const float* rstd, void* dx, float* dw, float* db, int64_t M, int64_t C,
                       int fuse_gelu, void* stream);

/* The LayerNorm in front of a downsample layer (models/convnext.py:79-82: LayerNorm(channels_first) -> Conv2d(k=2, s=2)),
 * writing its result in the "2x2 patch" layout [B][H/2][W/2][2][2][C] (pixel (b,h,w) -> row 4*((b*H/2+h/2)*W/2+w/2) +
 * 2*(h&1) + (w&1)), so that the stride-2 2x2 convolution becomes b200at_gemm_bf16 over [B*H/2*W/2][4C] rows with the
 * weight reordered to [Cout][kh][kw][Cin].  x: NHWC bf16 [B][H][W][C]; mean / rstd are indexed by the INPUT pixel.
 * H, W even. */
int b200at_ln_fwd_patch2(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int64_t B,
                         int64_t H, int64_t W, int64_t C, float eps, void* stream);
/* input gradient of the above: dy is in the patch layout (what the GEMM's input gradient produces), dx NHWC */
int b200at_ln_bwd_patch2(const void* dy, const void* x, const float* w, const float* b, const float* mean,
                         const float* rstd, void* dx, float* dw, float* db, int64_t B, int64_t H, int64_t W, int64_t C,
                         void* stream);

/* models/convnext.py:30-31: h = GELU(z + bias) on the 4C hidden (z = x @ W1^T from the GEMM), N % 8 == 0 */
int b200at_bias_gelu_fwd(const void* z, const float* bias, void* h, int64_t M, int64_t N, void* stream);
/* dz = dh * GELU'(z + bias); if dbias is non-null also ACCUMULATES the pwconv1 bias gradient
 * dbias[n] += sum_m dz[m][n] (fp32 [N], caller zeroes) in the same pass.  N <= 8192. */
int b200at_bias_gelu_bwd(const void* dh, const void* z, const float* bias, void* dz, float* dbias, int64_t M, int64_t N,
                         void* stream);
/* out[n] += sum_m a[m][n] (a bf16 [M][N], out fp32 [N], caller zeroes): bias gradients of `pwconv2` /
 * column sums of an upstream gradient (models/convnext.py:32,45).  N % 8 == 0, N <= 8192. */
int b200at_colsum_bf16(const void* a, float* out, int64_t M, int64_t N, void* stream);

/* models/convnext.py:45-49: out = res + gamma * (z + bias)  (layer scale + residual), N % 8 == 0 */
int b200at_scale_residual_fwd(const void* z, const float* bias, const float* gamma, const void* res, void* out,
                              int64_t M, int64_t N, void* stream);
/* dz = dout * gamma */
int b200at_scale_bwd(const void* dout, const float* gamma, void* dz, int64_t M, int64_t N, void* stream);
/* c = a + b (bf16): join of the residual and branch gradients */
int b200at_add_bf16(const void* a, const void* b, void* c, int64_t total, void* stream);

/* models/convnext.py:28,39 `self.dwconv` (7x7 depthwise, pad 3) on NHWC bf16.  wt is tap-major fp32 [49][C]
 * (wt[i*7+j][c] = weight[c,0,i,j]); bias may be null.  The input gradient is the same call with the taps
 * flipped (wt'[i*7+j] = wt[(6-i)*7+(6-j)]) and bias = null; `add` (nullable, same shape as y) is summed into
 * the result -- the residual-gradient join of the block (models/convnext.py:49).  C % 32 == 0. */
int b200at_dwconv7_fwd(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B,
                       int64_t H, int64_t W, int64_t C, void* stream);
/* weight / bias gradients, ACCUMULATED into dw [49][C] and db [C] (fp32, caller zeroes) */
int b200at_dwconv7_wgrad(const void* x, const void* dy, float* dw, float* db, int64_t B, int64_t H, int64_t W,
                         int64_t C, void* stream);

/* First stage of the CvSt stems as one kernel (utils_architecture.py:198-217 ConvBlock1, :174-195 ConvBlock3:
 * `nn.Conv2d(3, C0, 3, stride=2, padding=1)` -> channels-first LayerNorm (:57-81) -> GELU), with the
 * ImageNormalizer (:86-98) folded in:  y = GELU(LN(conv((x - mean) / std) + bias)).
 * x: fp32 NCHW [B][3][H][W] (the attack's iterate, read as is); y: NHWC bf16 [B][Ho][Wo][C0], Ho = (H-1)/2+1.
 * wk: fp32 [27][C0], wk[c*9+kh*3+kw][co] = weight[co][c][kh][kw].  mean3 / std3: HOST pointers to 3 floats, or
 * null for no normalisation.  C0 in {48, 64, 96}.  b200at_stem0_fwd serves the attack's evaluations (no weight
 * gradients), b200at_stem0_fwd_save the training forward. */
int b200at_stem0_fwd(const float* x, const float* mean3, const float* std3, const float* wk, const float* bias,
                     const float* ln_w, const float* ln_b, void* y, int64_t B, int64_t H, int64_t W, int64_t C0,
                     float eps, void* stream);
/* The same stage for the TRAINING forward: additionally y_pre [B][Ho][Wo][C0] bf16 = the convolution output WITHOUT its bias
 * and mean / rstd [B*Ho*Wo] fp32 = the LayerNorm statistics, i.e. exactly what b200at_ln_bwd_bias(dy, y_pre, pre_bias = bias,
 * ..., fuse_gelu = 1) needs; the convolution's weight gradient is then a library call on (normalised x, d y_pre). */
int b200at_stem0_fwd_save(const float* x, const float* mean3, const float* std3, const float* wk, const float* bias,
                          const float* ln_w, const float* ln_b, void* y, void* y_pre, float* mean, float* rstd, int64_t B,
                          int64_t H, int64_t W, int64_t C0, float eps, void* stream);
/* dL/dx (fp32 NCHW, every element written) of the above given dy = dL/dy (NHWC bf16): the attack's
 * `torch.autograd.grad(loss, [x_adv])` (autopgd_train_clean.py:185,283) through the first layer.  Recomputes the
 * pre-LayerNorm activation from x; nothing is saved by the forward. */
int b200at_stem0_bwd_input(const void* dy, const float* x, const float* mean3, const float* std3, const float* wk,
                           const float* bias, const float* ln_w, const float* ln_b, float* dx, int64_t B, int64_t H,
                           int64_t W, int64_t C0, float eps, void* stream);

/* models/convnext.py:30,32 `pwconv1` / `pwconv2` (nn.Linear on the NHWC rows) and their input gradients:
 *     C[M,N] = epilogue(A[M,K] . B[N,K]^T),  A, B, C bf16 row-major, fp32 accumulation in TMEM.
 * Hand-written tcgen05.mma (cta_group::1, UMMA 128 x BLOCK_N x 16) fed by TMA (SWIZZLE_128B), persistent and
 * warp specialised; the block's elementwise tail is fused into the epilogue:
 *   NONE       C = acc
 *   BIAS       C = acc + bias[n]
 *   BIAS_GELU  C = GELU(acc + bias[n]);  if c2 != null also c2 = acc + bias[n]   (pre-activation for the backward)
 *   RESIDUAL   C = aux + acc + bias[n]                                           (layer scale folded into B, bias)
 *   GELU_GRAD  C = acc * GELU'(aux)                                              (aux = saved pre-activation)
 * N % 16 == 0, K % 8 == 0, pointers 16-byte aligned.  The weight-gradient GEMMs (contraction over M) stay on
 * cuBLAS. */
#define B200AT_EPI_NONE 0
#define B200AT_EPI_BIAS 1
#define B200AT_EPI_BIAS_GELU 2
#define B200AT_EPI_RESIDUAL 3
#define B200AT_EPI_GELU_GRAD 4
int b200at_gemm_bf16(const void* a, const void* b, void* c, void* c2, const void* aux, const float* bias,
                     int64_t M, int64_t N, int64_t K, int epilogue, void* stream);

/* b200at_gemm_bf16(..., B200AT_EPI_GELU_GRAD) that also accumulates colsum[n] += sum_m C[m][n] (fp32 [N], the values as
 * rounded to bf16): the pwconv1 bias gradient of models/convnext.py:42-44 in the backward GEMM's epilogue instead of a
 * b200at_colsum_bf16 pass over dz. */
int b200at_gemm_gelu_grad_colsum(const void* a, const void* b, void* c, const void* aux, float* colsum, int64_t M, int64_t N,
                                 int64_t K, void* stream);

/* Conv2d(kernel 3, stride 2, padding 1) WITHOUT bias on NHWC bf16 input -- the second convolution of the CvSt stems
 * (utils_architecture.py:205-211 ConvBlock1: conv -> LayerNorm -> GELU; the bias rides in b200at_ln_fwd_bias) -- as an
 * implicit GEMM on the tcgen05 kernel: M = output pixels in tiles of whole output rows, N = Cout, K = 9 taps x 64
 * channels; the A tile of a tap is ONE TMA box of the input with element stride 2 along W and H (zero fill = padding
 * ring and the channels >= Cin).  x [B][H][W][Cin], y [B][H/2][W/2][Cout] bf16; wk [Cout][9*64] bf16 with
 * wk[co][(kh*3+kw)*64 + ci] = w[co][ci][kh][kw], zero for ci >= Cin.  H, W even, Cin % 8 == 0, Cin <= 64, Cout % 16 == 0,
 * W/2 <= 128.  Returns -1 (and launches nothing) for a shape it does not take. */
int b200at_conv3x3s2_fwd(const void* x, const void* wk, void* y, int64_t B, int64_t H, int64_t W, int64_t Cin,
                         int64_t Cout, void* stream);

/* Kernel-side copies of one ConvNeXt block's MLP weights, rebuilt once per optimiser step (models/convnext.py:42-49 with
 * the layer scale `gamma` folded into pwconv2): w1b = bf16(W1) [4C][C], w1t = w1b^T [C][4C], w2g = bf16(gamma[:,None] W2)
 * [C][4C], w2gt = w2g^T [4C][C] (the four operands b200at_mlp_fused / b200at_gemm_bf16 take in the two directions) and
 * b2g = gamma * b2.  fp32 parameters in, one launch.  C % 32 == 0. */
int b200at_prepare_mlp_weights(const float* w1, const float* w2, const float* b2, const float* gamma, void* w1b, void* w1t,
                               void* w2g, void* w2gt, float* b2g, int64_t C, void* stream);
/* Tail of the block's parameter gradients: given dW2g (gradient w.r.t. the folded gamma[:,None] W2, fp32 [C][4C]) and
 * col = column sum of the upstream gradient:  dW2 = gamma[:,None] dW2g,  db2 = col gamma,  dgamma = rowsum(dW2g W2) + col b2
 * (the chain rule of models/convnext.py:45-47 `x = self.gamma * x`). */
int b200at_finish_mlp_grads(const float* dw2g, const float* w2, const float* col, const float* b2, const float* gamma,
                            float* dw2, float* db2, float* dgamma, int64_t C, void* stream);

/* The ImageNormalizer (utils_architecture.py:86-98) in front of the TRAINING forward's first (library) convolution, with
 * the cast and the layout change: y[B][H][W][3] bf16 = (x[B][3][H][W] - mean) / std, fp32 arithmetic rounded once.
 * mean3 / std3: 3 host floats each, null = identity. */
int b200at_normalize_nhwc_bf16(const float* x, const float* mean3, const float* std3, void* y, int64_t B, int64_t H,
                               int64_t W, void* stream);

/* The whole MLP of a ConvNeXt block as one tcgen05 kernel per direction (models/convnext.py:42-49: pwconv1 -> GELU ->
 * pwconv2 -> layer scale -> + residual), hidden activation kept on chip:
 *   backward == 0:  out = residual + GELU(a wa^T + bias1) wb^T + bias2      a = LN output t2 [M][C], wa = W1 [4C][C],
 *                   wb = gamma*W2 [C][4C], bias2 = gamma*b2; z [M][4C] RECEIVES the pre-activation a wa^T (no bias);
 *                   p_out (nullable) receives GELU(z + bias1) (kept only when weight gradients will be needed)
 *   backward != 0:  out = ((a wa^T) * GELU'(z + bias1)) wb^T                a = dL/dout [M][C], wa = (gamma*W2)^T [4C][C],
 *                   wb = W1^T [C][4C]; z is READ; p_out (nullable) receives dz = dL/dz; bias2 / residual ignored
 * All matrices bf16 row-major (K contiguous), fp32 accumulation in TMEM, biases fp32.  C in {96, 192}
 * (b200at_mlp_fused_supported); other widths: b200at_gemm_bf16 + b200at_bias_gelu_*. */
int b200at_mlp_fused_supported(int64_t C);
int b200at_mlp_fused(const void* a, const void* wa, const void* wb, const float* bias1, const float* bias2,
                     const void* residual, void* z, void* p_out, void* out, int64_t M, int64_t C, int backward,
                     void* stream);

/* Multi-head self-attention of the ViT-S-CvSt blocks (timm 0.8 vision_transformer.Attention.forward, un-vendored;
 * call sites utils_architecture.py:271-301):  q,k,v = qkv.reshape(B,N,3,H,64).permute(2,0,3,1,4);
 * o = softmax(q k^T * scale) v, written as [B][N][H*64].  qkv: bf16 [B][N][3][H][64]; lse: fp32 [B][H][N], the
 * per-row log2-sum-exp of the scaled scores (saved for the backward).  Head dimension 64, N <= 208 (the whole
 * sequence of one (image, head) lives in one CTA's shared memory). */
int b200at_attn_fwd(const void* qkv, void* o, float* lse, int64_t B, int64_t N, int64_t H, float scale, void* stream);
/* dqkv (same layout as qkv, every element written) given d_o = dL/do; recomputes the probabilities from qkv and
 * lse.  Deterministic (no atomics). */
int b200at_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int64_t B, int64_t N,
                    int64_t H, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200AT_MODEL_H */
