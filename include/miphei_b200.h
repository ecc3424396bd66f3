/*
 * miphei_b200 — C ABI of the B200-native (sm_100a) kernels behind the MIPHEI-ViT generator hot path.
 *
 * The reference (Sanofi-Public/MIPHEI-ViT) has no native layer: every op below replaces a PyTorch/timm library
 * dispatch on the path  get_generator("myvitmatte") -> ViTMatte.forward  (src/generators/mipheivit.py:106-110)
 * and its backward / optimiser step (src/models.py:87-139).  Each entry point cites the reference line whose
 * arithmetic it implements.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MV_ERR_* code or a positive cudaError_t otherwise;
 *     mv_last_error() returns a human-readable message for the calling thread.  Nothing throws.
 *   - all pointers are DEVICE pointers owned by the caller (torch tensors in practice); kernels never allocate
 *     or free; `stream` is a cudaStream_t passed as void*.
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS; "tokens" are [M, C] rows,
 *     decoder feature maps are NHWC.
 *   - no CPU fallback exists: without a visible sm_100 device mv_init() fails.
 */
#ifndef MIPHEI_B200_H_
#define MIPHEI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MV_OK 0
#define MV_ERR_ARG (-1)      /* bad shape / alignment / null pointer */
#define MV_ERR_DEVICE (-2)   /* no sm_100 device, driver entry point missing */
#define MV_ERR_LAUNCH (-3)   /* kernel launch failed */

int mv_init(int device);
const char* mv_last_error(void);
int mv_version(void);
int mv_num_sms(void);
/* number of kernels this library has launched since load / since the last reset (bench.py "gpu_launches") */
int64_t mv_launch_count(void);
void mv_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM  D = epilogue(A[M,K] . B[N,K]^T), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma),
 * operands staged by TMA.  Replaces every nn.Linear on the path (timm Attention.qkv / proj, GluMlp.fc1 / fc2 —
 * created at src/generators/foundation_models.py:53-57; LoRA-extended qkv src/generators/lora.py:29-33), the
 * patch-embedding conv (timm PatchEmbed.proj) and, through im2col, the decoder 3x3 convs
 * (src/generators/mipheivit.py:20-41).
 *
 *   MV_GEMM_LINEAR     v = acc*scale[n] + shift[n]; v = act(v); v += resid[r(m), n]; out[o(m), n] = v
 *                      (scale/shift/resid optional; out bf16 or fp32; optional bf16 copy to aux)
 *                      o(m) = m                                   if rows_per_group == 0
 *                           = (m / rpg) * group_stride + m % rpg + row_offset   otherwise (patch tokens -> token rows)
 *                      r(m) = m % rpg if resid_row_mod else o(m)
 *   MV_GEMM_SWIGLU     B holds [gate rows | value rows] (N = 2*H): out[m, j] = silu(g) * v with
 *                      g = acc[m, j] + shift[j], v = acc[m, H + j] + shift[H + j]   (timm GluMlp, gate first);
 *                      aux (optional, bf16 [M, N]) receives the pre-activations [g | v] for the backward pass.
 *   MV_GEMM_SWIGLU_BWD acc = dU[M, H]; in2 = saved pre-activations [g | v] (bf16 [M, 2H]);
 *                      out[m, j] = acc * v * silu'(g), out[m, H + j] = acc * silu(g)     (bf16 [M, 2H])
 *   MV_GEMM_HEAD_GATE  the attention gate of all SegmentationHeads at once (AttentionBlock.psi, src/generators/
 *                      unet.py:407-422): B = the heads' 1x1 convs stacked [16*heads, C]; per head h
 *                      out[m, h] = sigmoid(b2[h] + sum_j w2[16h+j] * relu(acc[m,16h+j]*scale[16h+j] + shift[16h+j]))
 *                      with w2 = in2 (fp32 [16*heads]), b2 = resid (fp32 [heads]); out bf16 [M, heads].
 *   MV_GEMM_HEAD_CONV  the gated 3x3 output conv + tanh of all SegmentationHeads at once (unet.py:425-438 looped at
 *                      mipheivit.py:213-218): conv = 1 over the C<=64 feature map, B = [heads<=16, 9*64] (per tap the
 *                      head's 3x3 weights over channels); out[b, h, y, x] = tanh(shift[h] + sum_tap g[b, y', x', h] *
 *                      sum_c B[h, tap, c] f[b, y', x', c]) with g = in2 (bf16 [M, ldin2], the HEAD_GATE output).
 *                      out is NCHW: out_f32 = 1 fp32, 0 bf16, 2 uint8 through the inference sink mapping
 *                      ((p+0.9)/1.8).clamp(0,1)*255 truncated (src/callbacks.py:345-346).
 *   MV_GEMM_NN_ATOMIC  out[M, N] (fp32, caller-zeroed) += A[M, K] . B[K, N] with B ROW-major [K, N] (the reduction index is
 *                      the slow one: weight-gradient form, K = tokens / pixels), split-K across CTAs, fp32 atomics.
 *                      Used for the LoRA gradients dA = x^T dT, dB = (xA)^T dQ (src/generators/lora.py:16-18) and, with
 *                      conv = 1, for the 3x3 conv weight gradients: A = dz^T [Cout, pixels], B = the conv's NHWC input
 *                      map(s) b (/ a2) read as tap-shifted TMA tiles; out[co, tap*Cpad + ci] in the packed weight layout.
 *   MV_ACT_GATE_MASK   (LINEAR, bf16 out) out[m, n] = in2[m, n / 16] if acc*scale+shift > 0 else 0 — the ReLU-masked
 *                      gradient of the heads' gate hidden units (AttentionBlock.psi backward).
 *
 * Implicit-GEMM 3x3 convolution (pad 1, stride 1|2; Basic_Conv3x3, src/generators/mipheivit.py:20-41): set conv = 1.
 * A is then read from one or two NHWC bf16 feature maps (channel concat: a = [B,Hin,Win,c0], a2 = [B,Hin,Win,c1],
 * Hin = conv_h*stride) as TMA 4-D tiles whose out-of-bounds zero fill is the padding; rows of D are output pixels
 * (NHWC), M = batch*conv_h*conv_w, and B is [N, 9 * 64 * (ceil(c0/64)+ceil(c1/64))]: per tap (ky,kx) the source-0
 * channels zero-padded to a multiple of 64, then the source-1 channels likewise.
 * ---------------------------------------------------------------------------------------------------------- */
enum { MV_GEMM_LINEAR = 0, MV_GEMM_SWIGLU = 1, MV_GEMM_SWIGLU_BWD = 2, MV_GEMM_HEAD_GATE = 3, MV_GEMM_HEAD_CONV = 4, MV_GEMM_NN_ATOMIC = 5 };
enum { MV_ACT_NONE = 0, MV_ACT_RELU = 1, MV_ACT_GATE_MASK = 2 };

typedef struct mv_gemm_args {
  const void* a;      /* bf16 [M, K] */
  int64_t lda;
  const void* b;      /* bf16 [N, K] */
  int64_t ldb;
  int32_t m, n, k;
  int32_t mode;       /* MV_GEMM_* */
  int32_t act;        /* MV_ACT_* (LINEAR only) */
  int32_t out_f32;    /* 1: out is fp32, 0: bf16 */
  void* out;
  int64_t ldo;
  void* aux;          /* optional bf16 second output */
  int64_t ldaux;
  const float* scale; /* [N] or NULL */
  const float* shift; /* [N] or NULL */
  const float* resid; /* fp32 or NULL */
  int64_t ldr;
  const void* in2;    /* bf16, SWIGLU_BWD only */
  int64_t ldin2;
  int32_t rows_per_group, group_stride, row_offset, resid_row_mod;
  int32_t block_n;    /* 0 = choose */
  int32_t conv;       /* 1: implicit 3x3 convolution, see above */
  const void* a2;     /* second NHWC source (channel concat) or NULL */
  int32_t conv_batch, conv_h, conv_w, conv_stride; /* OUTPUT height/width */
  int32_t conv_c0, conv_c1;
  int32_t reserved_splits; /* NN_ATOMIC: split-K factor, 0 = choose */
  int32_t reserved2;
  float* colstats;    /* LINEAR + bf16 out: fp32 [2, N] += per-column (sum, sum of squares) of the stored values —
                         BatchNorm batch statistics of a train-mode conv; out may be NULL (statistics only) */
  int32_t kskip_begin, kskip_end; /* LINEAR: K range [begin, end) (multiples of 64) whose B columns are all zero and is
                                     not loaded at all (dT = [dQ | dK | dV] . [aB_q ; 0 ; aB_v]^T skips the dK third) */
  int32_t ab_f16;     /* 0: A, a2 and B hold bf16; 3: all hold fp16 (kind::f16 takes either format, but both operands must
                         share it: a mixed descriptor is an illegal instruction on sm_100a — measured).
                         The training-mode decoder keeps feature maps and conv weights in fp16: the
                         reference trains under fp16 autocast (configs/config.yaml:23) and the LoRA gradients need the
                         extra mantissa bits. Outputs are unaffected. */
  int32_t reserved3;  /* schedule experiments for the last, partial wave of tiles (both measured slower than leaving it partly
                         idle): 0 = stream-K when the library judges it worthwhile (needs `workspace`),
                         1 = stream-K whenever legal, 2 = never, 3 = re-tile the last wave as 128 x 128 tiles in a 2nd launch */
  void* workspace;    /* optional, 256-byte aligned, ZERO-INITIALISED ONCE by the caller and then left to the library (it
                         restores the zeros): lets LINEAR / SWIGLU / SWIGLU_BWD GEMMs spread the k blocks of their last,
                         partial wave of tiles over all SMs (stream-K; partial sums meet here). One workspace per stream:
                         two GEMMs that may run concurrently must not share it. >= 20 MB covers every shape; NULL disables. */
  int64_t workspace_bytes;
} mv_gemm_args;

int mv_gemm_bf16(const mv_gemm_args* args, void* stream);
/* Diagnostics: when buf != NULL every following mv_gemm_bf16 launch writes 16 int64 per CTA into buf[16 * blockIdx.x ..]:
 * cycles [0] producer waiting for smem, [1] MMA waiting for operands, [2] MMA waiting for an accumulator, [3] epilogue
 * waiting for an accumulator, [4] epilogue busy, [5] producer lifetime, [6] tiles; %globaltimer ns at [8] kernel entry,
 * [9] after the prologue / grid-dependency wait, [10] CTA end.  buf must hold 16 * 2 * num_sms int64. NULL disables. */
void mv_gemm_set_profile_buffer(void* buf);
/* Same idea for the attention forward (diagnostic build only): 16 int64 per CTA, layout in csrc/attention.cu. */
void mv_attn_set_profile_buffer(void* buf);

/* ------------------------------------------------------------------------------------------------------------
 * LayerNorm over channels of token rows (timm Block.norm1/norm2, final norm; eps 1e-6).
 *   fwd: x fp32 [M, D] -> y bf16 [M, ldy]; mean/rstd (fp32 [M]) optional, both or neither.
 *   bwd: dx[M, D] = dres + dLN(x, w)(dy); dy bf16 or fp32; dres optional; optional bf16 copy of dx.
 *        The affine parameters are frozen on this path (apply_lora, src/generators/lora.py:66-68): no dw / db.
 * ---------------------------------------------------------------------------------------------------------- */
/* y_f16: store y as fp16 instead of bf16 (the final norm of the TRAINING forward feeds the fp16 decoder, see ab_f16) */
int mv_layernorm_fwd(const float* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy, int y_f16,
                     float* mean, float* rstd, int m, int d, float eps, void* stream);
int mv_layernorm_bwd(const float* x, int64_t ldx, const float* w, const void* dy, int64_t lddy, int dy_f32,
                     const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb, int m,
                     int d, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Self-attention forward, head_dim 64, n_tok >= 16 (flash-style key blocks; timm Attention.forward -> F.scaled_dot_product_attention).
 *   qkv bf16 [batch*n_tok, 3*heads*64] (q | k | v per token, as nn.Linear(dim, 3*dim) lays them out),
 *   out bf16 [batch*n_tok, heads*64] token-major, lse fp32 [batch, heads, n_tok] (natural log; optional).
 * ---------------------------------------------------------------------------------------------------------- */
int mv_attn_fwd(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, float* lse, int batch, int n_tok, int heads,
                float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Memory-bound glue of the forward pass.
 *   mv_prep_input     x fp32 NCHW [B,3,S,S] -> img_nhwc8 bf16 [B,S,S,8] (channels 3..7 zero; D0 of ConvStream,
 *                     src/generators/mipheivit.py:66-73) and/or patch_matrix bf16 [B*(S/14)^2, 592]: timm
 *                     PatchEmbed's conv(k14, s14) as a GEMM operand, K = (c, ky, kx) = 588 zero-padded to 592.
 *   mv_fill_prefix    residual-stream rows of cls + register tokens (timm _pos_embed, no_embed_class=True).
 *   mv_tokens_to_map  Encoder.forward tail (mipheivit.py:158-162): prefix tokens dropped, g x g token map
 *                     bicubic-resized (A=-0.75, align_corners False, scale factor target/grid) -> NHWC bf16.
 *   mv_upsample2x     Fusion_Block's bilinear x2 (mipheivit.py:89), NHWC bf16.
 * ---------------------------------------------------------------------------------------------------------- */
/* img_f16 / f16 flags below: the 16-bit maps are fp16 instead of bf16 (training-mode decoder, see mv_gemm_args.ab_f16);
 * the patch matrix always stays bf16 (it meets the frozen bf16 ViT weights). */
int mv_prep_input(const float* x, void* img_nhwc8, int img_f16, void* patch_matrix, int batch, int size, int ldk,
                  void* stream);
/* Input staging for whole-slide inference: raw uint8 H&E tiles NHWC [batch, size, size, 3] normalised on the
 * device, v = u8 * scale[c] + bias[c] with the H-Optimus statistics of src/dataset.py:600-601 (scale = 1/(255 std_c),
 * bias = -mean_c/std_c; HOST pointers to 3 floats each) — the H2D copy is 4x smaller than the fp32 NCHW tensor. */
int mv_prep_input_u8(const void* tiles_u8, const float* scale3, const float* bias3, void* img_nhwc8, int img_f16,
                     void* patch_matrix, int batch, int size, int ldk, void* stream);
int mv_fill_prefix(float* x, int64_t ldx, const float* prefix, int batch, int n_tok, int n_prefix, int dim, void* stream);
int mv_tokens_to_map(const void* tokens, int64_t ldt, void* out, int batch, int n_tok, int prefix, int grid, int target,
                     int dim, int f16, void* stream);
int mv_upsample2x(const void* in, void* out, int batch, int h, int w, int c, int f16, void* stream);

/* adjoint of mv_tokens_to_map: d_map NHWC bf16 [B,t,t,D] -> d_tokens bf16 [B*n_tok, D] (prefix rows zero) */
int mv_tokens_to_map_bwd(const void* dmap, void* dtokens, int64_t ldt, int batch, int n_tok, int prefix, int grid,
                         int target, int dim, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Self-attention backward (autograd of timm Attention's scaled_dot_product_attention), head_dim 64, any n_tok.
 *   qkv / out / lse as saved by mv_attn_fwd, dout bf16 [B*n_tok, D]; dsum fp32 [B, heads, n_tok] workspace;
 *   dqkv bf16 [B*n_tok, >= 3D] receives dq | dk | dv.
 * ---------------------------------------------------------------------------------------------------------- */
int mv_attn_bwd(const void* qkv, int64_t ldqkv, const void* out, int64_t ldo, const void* dout, int64_t lddo,
                const float* lse, float* dsum, void* dqkv, int64_t lddqkv, int batch, int n_tok, int heads, float scale,
                void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * LoRA gradients of one block (QkvWithLoRA, src/generators/lora.py:16-18,29-33): see csrc/lora.cu.
 *   xn_ext bf16 [M, >= D+16]: LayerNorm output with T = xn [A_q | A_v] in columns D..D+15;
 *   dqkv_ext bf16 [M, >= 3D+16]: dq | dk | dv with dT in columns 3D..3D+15.
 *   dA_* fp32 [D, 8], dB_* fp32 [8, D] are overwritten.  workspace: 256-byte aligned.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t mv_lora_grads_workspace_bytes(int m, int d);
int mv_lora_grads(const void* xn_ext, int64_t ldx, const void* dqkv_ext, int64_t ldq, int m, int d, float alpha,
                  float* dA_q, float* dA_v, float* dB_q, float* dB_v, void* workspace, int64_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss and optimiser.
 *   mv_loss_fwd_bwd     mode 0: WeightedMSELoss (src/loss.py:54-57; weights NULL = plain MSE, get_mse_loss 41-44),
 *                       mode 1: get_mae_loss (35-38), mode 2: L1_L2_Loss (113-123). pred/target/grad NCHW fp32
 *                       [B, C, HW]; grad (optional) = grad_scale * dloss/dpred; loss: one fp32.
 *   mv_grad_norm        global L2 norm of a flat fp32 gradient buffer + clip coefficient min(1, max_norm/(norm+1e-6))
 *                       (clip_gradients(..., 1.0, "norm"), src/models.py:136); norm_out: 2 floats; workspace 1024 floats.
 *   mv_adam_clip_step   torch.optim.Adam(betas, eps, weight_decay 0) update (src/models.py:361-362) on flat buffers with
 *                       the gradient scaled by norm_coef[1] * grad_mul on the fly; `step` counts from 1.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t mv_loss_workspace_floats(int batch, int chans, int hw);
int mv_loss_fwd_bwd(const float* pred, const float* target, float* grad, const float* weights, int batch, int chans, int hw,
                    int mode, float lambda, float grad_scale, float* loss, float* workspace, int64_t workspace_floats,
                    void* stream);
int mv_grad_norm(const float* grads, int64_t n, float max_norm, float* norm_out, float* workspace, void* stream);
int mv_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const float* norm_coef, float grad_mul, float lr, float beta1, float beta2, float eps, int step,
                      void* stream);
/* The same update with the step count and learning rate living on the DEVICE, so that a captured CUDA graph of the whole
 * training step replays correctly:
 *   mv_adam_schedule      *step (int64, optimiser steps taken so far) -> hyper[4] = {lr / (1 - beta1^t), sqrt(1 - beta2^t),
 *                         lr, t} with t = *step + 1 and lr = base_lr * LambdaLR factor of pix2pix_lr_scheduler
 *                         (src/utils.py:217-230 as built at src/models.py:363-369: linear warm-up, flat to total/2, linear to
 *                         zero) evaluated at *step; then *step += 1.
 *   mv_adam_clip_step_dev Adam update reading hyper[0..1] written by mv_adam_schedule. */
int mv_adam_schedule(int64_t* step, float base_lr, int64_t total_steps, int64_t warmup_steps, float beta1, float beta2,
                     float* hyper, void* stream);
int mv_adam_clip_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const float* norm_coef, float grad_mul, const float* hyper, float beta1, float beta2, float eps,
                          void* stream);
/* Table-driven re-layout, dst[i] = convert(src[idx[i]]) (idx -1: zero, -2: leave dst[i]); mode 0 fp32, 1 bf16, 2 fp16.
 * Packs the trainable decoder weights (reference layouts, one flat fp32 buffer) into the kernels' operand layouts and
 * scatters the packed weight-gradient accumulators back into parameter layout — one launch per arena, no framework ops
 * inside a training step.  mv_add_i64: BatchNorm num_batches_tracked counters (+= inc).  mv_memset_async: cudaMemsetAsync. */
/* LoRA operands of every block from the flat parameter buffer, one launch: lora_flat fp32 [depth, 4, 8*d] = per block
 * (A_q [d,8], B_q [8,d], A_v [d,8], B_v [8,d]) as QkvWithLoRA registers them (src/generators/lora.py:21-27); ptrs int64
 * [depth, 4] = device pointers to the block's bf16 acat [16, d], K-extended QKV weight [3d, ldw] (columns d..d+15
 * written), K-extended dX weight [d, ldb] (columns 3d..3d+15; may be 0) and bcat [16, 3d] (may be 0). */
int mv_lora_refresh(const float* lora_flat, const int64_t* ptrs, int depth, int d, float alpha, int64_t ldw, int64_t ldb,
                    void* stream);
int mv_gather_cast(const float* src, const int32_t* idx, void* dst, int64_t n, int mode, void* stream);
int mv_add_i64(int64_t* p, int n, int64_t inc, void* stream);
int mv_memset_async(void* p, int value, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Train-mode BatchNorm2d + ReLU around the convs (Basic_Conv3x3, src/generators/mipheivit.py:20-41; AttentionBlock.psi[1],
 * src/generators/unet.py:411-415); activations are NHWC bf16 seen as [M, C] rows.  See csrc/bn.cu.
 *   mv_bn_finalize    (sum, sumsq) from the conv epilogue -> batch mean / rstd, folded (scale, shift), running-stat update
 *                     (momentum, unbiased variance). pre_bias: bias added before the BN (psi[0].bias) or NULL.
 *   mv_bn_relu_apply  y = relu(z*scale + shift); z is the raw conv output, bf16 or fp32 (z_f32) — the training path keeps
 *                     it in fp32 so that bf16 rounding does not flip ReLU masks against the fp32 reference
 *   mv_bn_relu_bwd    dz = BN'(dy * [y > 0]); sums (fp32 [2, C], overwritten) = (dbeta, dgamma); y may be bf16 or fp16
 *                     (only its sign is read)
 * ---------------------------------------------------------------------------------------------------------- */
int mv_bn_finalize(const float* colstats, double count, const float* gamma, const float* beta, const float* pre_bias,
                   float* running_mean, float* running_var, float momentum, float eps, int c, float* scale, float* shift,
                   float* mean, float* rstd, void* stream);
/* Closed-form train-mode BatchNorm statistics of the SegmentationHead gates (AttentionBlock.psi[0..1],
 * src/generators/unet.py:407-422): the BN input W1 f + b1 is linear in the 32-channel map f, so its batch mean / variance
 * follow from E[f] and E[f f^T] (csrc/heads_stats.cu).
 *   mv_gram32             gram (fp32 [40, 32], zeroed by the caller): rows 0..31 += f^T f, row 32 += 1^T f; f bf16 [m, 32]
 *   mv_heads_bn_from_gram (gram, count = m) -> folded (scale, shift) for relu(scale * (W1 f) + shift), saved batch
 *                         (mean, rstd) of W1 f + b1, running-stat update; w1 fp32 [c, 32] = the weights the gate GEMM uses */
int mv_gram32(const void* f, int64_t ldf, int64_t m, int f16, float* gram, void* stream);
/* w1_fmt: W1 is rounded to the gate GEMM's operand format before use (0 none, 1 bf16, 2 fp16) */
int mv_heads_bn_from_gram(const float* gram, double count, const float* w1, const float* b1, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float momentum, float eps, int c,
                          float* scale, float* shift, float* mean, float* rstd, int w1_fmt, void* stream);
/* Closed-form backward of the heads' gate MLP through its train-mode BatchNorm (csrc/heads_stats.cu): from
 * E fp32 [40, 256] (rows 0..31 f^T e, row 32 1^T e), FF = the mv_gram32 output, fin = the [4, 256] (scale, shift, mean,
 * rstd) of mv_heads_bn_from_gram and count = pixels:  dw1 [256, 32], dgamma / dbeta / dw2 [256] (parameter gradients of
 * psi[0].weight, psi[1].weight / bias, psi[3].weight) and the operands of the two GEMMs that assemble d f:
 * ca_t bf16 [32, 256], mx_n [32, 64] (fp16 when mx_f16 else bf16; stored times a power of two that keeps these
 * gradient-sized values out of the fp16 underflow range, its reciprocal in mx_scale fp32 [32] = the `scale` of the GEMM that
 * consumes mx_n), kshift fp32 [32]. */
int mv_heads_bwd_algebra(const float* E, const float* FF, const float* w1, const float* b1, const float* gamma,
                         const float* w2, const float* fin, double count, int n_units, float* dw1, float* dgamma,
                         float* dbeta, float* dw2, void* ca_t_bf16, void* mx_n, int mx_f16, float* mx_scale, float* kshift,
                         void* stream);
int mv_bn_relu_apply(const void* z, int z_f32, const float* scale, const float* shift, void* y, int y_f16, int64_t m,
                     int c, void* stream);
int mv_bn_relu_bwd(const void* dy, int64_t lddy, const void* y, const void* z, int z_f32, const float* mean, const float* rstd,
                   const float* gamma, float* sums, void* dz, int64_t m, int c, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Memory-bound kernels of the decoder backward pass (csrc/decoder_bwd_ew.cu).
 * ---------------------------------------------------------------------------------------------------------- */
/* [m, c] -> bf16 [c (+ a row of ones), ldo]; in_f16: the input holds fp16 and is converted on the way (the weight-gradient
 * GEMMs pair it with bf16 gradients, and both tensor-core operands must share one format) */
int mv_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t m, int c, int ones_row, int in_f16,
                      void* stream);
/* dst bf16 [n] = src fp16 [n] (n % 8 == 0): bf16 twins of the fp16 activation maps for the weight-gradient GEMMs */
int mv_f16_to_bf16(const void* src, void* dst, int64_t n, void* stream);
int mv_upsample2x_bwd(const void* dup, int64_t ldu, void* dx, int batch, int h, int w, int c, void* stream);
int mv_zero_insert2x(const void* dz, void* u, int batch, int h, int w, int c, void* stream);
int mv_add_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t m, int c, void* stream);
int mv_heads_ds(const float* dpred, const float* pred, void* ds, float* dbias, int batch, int heads, int hw, void* stream);
int mv_heads_bwd_stencil(const void* t, const void* ds, const void* gate, void* dt, void* du, float* db2, int batch, int h,
                         int w, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-nucleus mean intensities: MeanCellExtrator.extract_mean, src/utils.py:49-121 (and the same
 * reduction in CellMetrics.update, src/metrics.py:32-74).  pred / target fp32 NCHW [batch, chans, hw]; nuclei labels
 * [batch, hw] int32 or int64 (label_bytes 4 / 8), 0 = background.  One CTA per image (csrc/cell_means.cu).
 *   mv_cell_means       per image: ids (ascending, like torch.unique), means_pred / means_target [batch, cap, chans],
 *                       counts [batch, cap], n_unique [batch]; rows >= n_unique[b] are untouched. cap = power of two
 *                       >= the number of nuclei of any image; *overflow is set to 1 if an image has more (the caller
 *                       must pre-zero it and re-run with a larger cap).  target / means_target may both be NULL.
 *                       workspace (256-byte aligned, mv_cell_means_workspace_bytes() bytes, contents irrelevant): while the
 *                       tables fit in shared memory (cap * (2 chans + 10) * 4 bytes <= 220 KB) it holds the partial sums that
 *                       let SEVERAL CTAs share one image (optional: NULL = one CTA per image); beyond that (thousands of
 *                       nuclei in one tile) it holds the tables themselves and is required.
 *   mv_cell_means_bwd   the extractor is differentiable in the reference (training_step's cell loss, src/models.py:
 *                       120-131): d map[b, c, p] = d means[row(b, label p), c] / count[row], 0 on background; ids / counts /
 *                       n_unique are the packed forward outputs.
 *   mv_cell_means_pack  concatenates the per-image rows in batch order (the reference's torch.cat): out_* hold
 *                       sum(n_unique) rows.
 * ---------------------------------------------------------------------------------------------------------- */
int mv_cell_means(const float* pred, const float* target, const void* nuclei, int label_bytes, int batch, int chans, int hw,
                  int cap, float* means_pred, float* means_target, int64_t* ids, float* counts, int32_t* n_unique,
                  int32_t* overflow, void* workspace, int64_t workspace_bytes, void* stream);
int64_t mv_cell_means_workspace_bytes(int batch, int chans, int cap);
int mv_cell_means_bwd(const float* dmeans, const int64_t* ids, const float* counts, const int32_t* n_unique,
                      const void* nuclei, int label_bytes, int batch, int chans, int hw, float* dmap, void* stream);
int mv_cell_means_pack(const float* means_pred, const float* means_target, const int64_t* ids, const float* counts,
                       const int32_t* n_unique, int batch, int chans, int cap, float* out_pred, float* out_target,
                       int64_t* out_ids, float* out_counts, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Whole-slide plumbing (csrc/wsi.cu).
 *   mv_thumb_std_hist  get_locs_otsu (slidevips-python/slidevips/tiling.py:25-31): thumb uint8 [n_pix, chans] interleaved ->
 *                      std_u8 [n_pix] = np.uint8(thumbnail.std(axis=-1)) bit-exactly (chans == 1: copy) and its histogram
 *                      (uint32 [256], overwritten).
 *   mv_otsu_threshold  the threshold cv2.threshold(src, 0, 255, THRESH_BINARY + THRESH_OTSU) selects (tiling.py:30), from the
 *                      histogram; *thresh (device int32).
 *   mv_tile_tissue     tiling.py:52-60: counts[i] = #{mask > threshold} in box i = (x0, y0, x1, y1) (int32 [n, 4], clipped to
 *                      the map by the caller); threshold from thresh_dev (device, may be NULL) else fixed_thresh.
 *   mv_stitch_tiles    preprocessings/cycle_gan/cycle_gan_wsi_inference.py:98-104: tiles uint8 [batch, chans, size, size];
 *                      the window [crop, crop + keep)^2 of tile b goes to canvas[:, y.., x..] with (x, y) = xy[b] (int32
 *                      [batch, 2]), clipped to the canvas uint8 [chans, canvas_h, canvas_w] (pyvips insert). The canvas may be
 *                      device memory or mapped pinned host memory. sequential != 0 inserts tile by tile (later wins) for
 *                      windows that overlap each other.
 *   mv_host_alloc_mapped / mv_host_free      zeroed pinned host memory addressable by kernels (*device_ptr) — slide canvases.
 *   mv_host_register / mv_host_unregister    page-lock an existing host range (shared-memory ring filled by loader
 *                      worker processes, src/dataset.py:545-575 / slidevips torch_datasets.py:88-127).
 * ---------------------------------------------------------------------------------------------------------- */
int mv_thumb_std_hist(const void* thumb, int64_t n_pix, int chans, void* std_u8, uint32_t* hist256, void* stream);
int mv_otsu_threshold(const uint32_t* hist256, int64_t n_pix, int32_t* thresh, void* stream);
int mv_tile_tissue(const void* mask_u8, int width, const int32_t* thresh_dev, int fixed_thresh, const int32_t* boxes,
                   int n_boxes, int32_t* counts, void* stream);
int mv_stitch_tiles(const void* tiles_u8, const int32_t* xy, int batch, int chans, int size, int crop, int keep,
                    void* canvas_u8, int64_t canvas_h, int64_t canvas_w, int sequential, void* stream);
int mv_host_alloc_mapped(int64_t bytes, void** host_ptr, void** device_ptr);
int mv_host_free(void* host_ptr);
int mv_host_register(void* ptr, int64_t bytes);
int mv_host_unregister(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* MIPHEI_B200_H_ */
