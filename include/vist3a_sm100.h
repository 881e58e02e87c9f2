/*
 * libvist3a_sm100 — C ABI of the Blackwell (sm_100a) kernels behind the VIST3A hot path.
 *
 * The reference (gohyojun15/VIST3A @ 32253c3) has no FFI of its own: its hot path is PyTorch
 * modules calling cuBLAS / SDPA / cuDNN.  Every entry point below therefore names the reference
 * operation (file:line, or the diffusers-0.33.1 op for the un-vendored DiT) it replaces.  The
 * Python host code in vist3a_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer (torch `tensor.data_ptr()`), 16-byte aligned;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued, never synchronised;
 *   - no hidden device allocation; TMA descriptors are built on the host per call (no global state
 *     except a per-device attribute cache);
 *   - return value: 0 = ok, <0 = error (see VIST3A_ERR_*); vist3a_last_error() gives the message
 *     of the last failing call on the calling thread;
 *   - bf16 tensors are row-major with explicit leading dimensions given in ELEMENTS.
 */
#ifndef VIST3A_SM100_H_
#define VIST3A_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIST3A_OK 0
#define VIST3A_ERR_INVALID (-1)      /* bad shape / stride / alignment / null pointer */
#define VIST3A_ERR_ARCH (-2)         /* current device is not sm_100 */
#define VIST3A_ERR_CUDA (-3)         /* a CUDA runtime / driver call failed */
#define VIST3A_ERR_UNSUPPORTED (-4)  /* valid request, no kernel variant for it */

#define VIST3A_DTYPE_BF16 0
#define VIST3A_DTYPE_F32 1

const char* vist3a_last_error(void);
int vist3a_abi_version(void);
/* number of kernels this library has launched from the calling process (all threads) */
int64_t vist3a_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = epilogue( A[M,K] · W[N,K]^T )      tcgen05.mma + TMA + TMEM accumulators
 * replaces: every nn.Linear on the path — diffusers WanTransformerBlock q/k/v/out/ffn
 *   (call sites inference_t23d.py:94-103), AS/model/encoder/vggt/layers/attention.py:55,75
 *   (qkv, proj), AS/.../layers/mlp.py:34-40 (fc1, fc2), the 1x1 convs of
 *   AS/.../heads/dpt_head.py:69-79, and (after im2col) models/stitching_layer_builder.py:32-42.
 * Epilogue, in order:  v = acc + bias[n];  v = act(v);  if round_linear: v = bf16(v);
 *   if gate: v *= gate[(row / rows_per_batch) * gate_bstride + n];  if round_gate: v = bf16(v);
 *   if residual: v += residual[row, n];  store as out_dtype.
 * ------------------------------------------------------------------------------------------ */
#define VIST3A_ACT_NONE 0
#define VIST3A_ACT_GELU_TANH 1 /* FeedForward(activation_fn="gelu-approximate"), Wan text embedder */
#define VIST3A_ACT_GELU_ERF 2  /* AS/.../layers/mlp.py:25 (nn.GELU) */
#define VIST3A_ACT_SILU 3      /* Wan time embedder */
#define VIST3A_ACT_RELU 4      /* DPT heads */

#define VIST3A_GEMM_FLAG_2CTA 1u       /* use cta_group::2 pairs (256-row tiles) */
#define VIST3A_GEMM_FLAG_1CTA 2u       /* force single-CTA tiles */

typedef struct vist3a_gemm_args {
  const void* A;          /* [M, K] in_dtype, row stride lda */
  const void* W;          /* [N, K] in_dtype, row stride ldw (nn.Linear weight layout) */
  void* C;                /* [M, N] out_dtype, row stride ldc */
  const float* bias;      /* [N] fp32 or NULL */
  const float* gate;      /* fp32, index (row / rows_per_batch) * gate_bstride + n, or NULL */
  const void* residual;   /* [M, N] out_dtype, row stride ldr, or NULL (may alias C) */
  int64_t M, N, K;
  int64_t lda, ldw, ldc, ldr;
  int64_t rows_per_batch; /* >= 1 */
  int64_t gate_bstride;   /* 0 => one gate vector shared by all rows (LayerScale) */
  int32_t in_dtype;       /* VIST3A_DTYPE_BF16 (kind::f16) or VIST3A_DTYPE_F32 (kind::tf32) */
  int32_t out_dtype;
  int32_t act;
  int32_t round_linear;   /* round (acc+bias, act) to bf16 before gate/residual (autocast Linear) */
  int32_t round_gate;     /* round the gated product to bf16 (bf16 LayerScale) */
  uint32_t flags;
} vist3a_gemm_args;

int vist3a_gemm(const vist3a_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head attention forward, non-causal, no mask:  O = softmax(Q K^T * scale) V
 * replaces: F.scaled_dot_product_attention in diffusers WanAttnProcessor2_0 (self: L=4096, d=128;
 *   cross: kv=512) and AS/model/encoder/vggt/layers/attention.py:64-69 (d=64; L=1029 / 13377).
 * Q element (b, i, h, c) at q + b*q_bs + i*q_rs + h*q_hs + c (strides in elements; c contiguous),
 * likewise K, V, O.  head_dim in {64, 128}.
 * ------------------------------------------------------------------------------------------ */
typedef struct vist3a_fmha_args {
  const void* Q;
  const void* K;
  const void* V;
  void* O;
  int64_t batch, heads, len_q, len_kv, head_dim;
  int64_t q_bs, q_rs, q_hs;
  int64_t k_bs, k_rs, k_hs;
  int64_t v_bs, v_rs, v_hs;
  int64_t o_bs, o_rs, o_hs;
  float scale; /* multiplies QK^T; 1/sqrt(head_dim) in both references */
  uint32_t flags;
} vist3a_fmha_args;

int vist3a_fmha_fwd(const vist3a_fmha_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm with optional modulation:  out[r,:] = LN(x[r,:]) * mul[b,:] + add[b,:],  b = r / rows_per_batch
 * LN statistics in fp32 (biased variance), eps inside the sqrt.
 * replaces: diffusers FP32LayerNorm + AdaLN-zero "(norm(x) * (1 + scale) + shift)" (mul = 1+scale,
 *   add = shift, batch stride = mul_bstride), FP32LayerNorm(elementwise_affine=True) of norm2
 *   (mul = weight, add = bias, bstride 0), and nn.LayerNorm in AS/.../layers/block.py:62,73.
 * ------------------------------------------------------------------------------------------ */
int vist3a_layernorm(const void* x, int32_t x_dtype, int64_t ldx, void* out, int32_t out_dtype, int64_t ldo,
                     int64_t rows, int64_t dim, int64_t rows_per_batch, const float* mul, int64_t mul_bstride,
                     const float* add, int64_t add_bstride, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * RMSNorm across all heads (+ optional interleaved-pair RoPE), in place on a strided bf16 matrix.
 *   x[r, 0:dim] <- x * rsqrt(mean(x^2) + eps) * weight ; then per head h and pair j:
 *   (x[2j], x[2j+1]) <- (x[2j] c - x[2j+1] s, x[2j] s + x[2j+1] c), (c,s) = cos/sin[(r % rope_len), j]
 * replaces: diffusers WanAttnProcessor2_0: attn.norm_q / norm_k (RMSNorm "rms_norm_across_heads",
 *   eps 1e-6) and apply_rotary_emb with WanRotaryPosEmbed frequencies.
 * cos/sin: [rope_len, head_dim/2] fp32, or NULL for no RoPE (cross-attention).
 * ------------------------------------------------------------------------------------------ */
int vist3a_rmsnorm_rope(void* x, int64_t ldx, int64_t rows, int64_t dim, int64_t head_dim, const float* weight,
                        float eps, const float* rope_cos, const float* rope_sin, int64_t rope_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * AdaLN modulation vectors:  out[b, j, :] = table[j, :] + mod[b, j, :] (+ 1 where bit j of
 * one_plus_mask is set).  mod may be [B, J*D] (DiT blocks: timestep_proj) or broadcast [B, D]
 * (output head: temb) when mod_is_broadcast != 0.
 * replaces: "(self.scale_shift_table + temb.float()).chunk(6, dim=1)" in WanTransformerBlock and
 *   "(self.scale_shift_table + temb.unsqueeze(1)).chunk(2, dim=1)" in WanTransformer3DModel.
 * ------------------------------------------------------------------------------------------ */
int vist3a_modulation(const float* table, const void* mod, int32_t mod_dtype, int32_t mod_is_broadcast, float* out,
                      int64_t batch, int64_t nvec, int64_t dim, uint32_t one_plus_mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Skinny linear for M <= 16 rows (HBM-bound weight streaming):  y = act(x W^T + b)
 * replaces: Wan TimestepEmbedding / time_proj (M = batch) and every Linear of
 *   AS/model/encoder/vggt/heads/camera_head.py:87-170 (M = views).
 * ------------------------------------------------------------------------------------------ */
int vist3a_skinny_linear(const void* x, int32_t x_dtype, int64_t ldx, const void* W, int32_t w_dtype, int64_t ldw,
                         const float* bias, void* y, int32_t y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                         int32_t pre_act, int32_t act, void* stream);

/* sinusoidal timestep features: out[b, 0:half] = cos(t_b f_i), out[b, half:] = sin(t_b f_i),
 * f_i = exp(-ln(10000) i / half)  (diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0)) */
int vist3a_timestep_features(const float* t, void* out, int32_t out_dtype, int64_t batch, int64_t dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * Wan patch embedding gather (Conv3d k = s = (1,2,2) as a K=64 GEMM) and its inverse.
 *   patchify:   x[B, C, T, H, W] (fp32 or bf16) -> A[B*T*(H/2)*(W/2), C*4] bf16, k = c*4 + dy*2 + dx
 *   unpatchify: P[B*T*(H/2)*(W/2), 4*C] -> out[B, C, T, H, W], column = (dy*2 + dx)*C + c
 * replaces: WanTransformer3DModel.patch_embedding + flatten/transpose, and the final
 *   reshape/permute(0,7,1,4,2,5,3,6).
 * ------------------------------------------------------------------------------------------ */
int vist3a_patchify(const void* x, int32_t x_dtype, void* A, int64_t B, int64_t C, int64_t T, int64_t H, int64_t W,
                    void* stream);
int vist3a_unpatchify(const void* P, int32_t p_dtype, int64_t ldp, void* out, int32_t out_dtype, int64_t B, int64_t C,
                      int64_t T, int64_t H, int64_t W, void* stream);

/* ------------------------------------------------------------------------------------------
 * Classifier-free guidance + linear multistep update on the latent (fp32):
 *   eps = uncond + g * (cond - uncond);
 *   out = c_x * x + c_e * eps + sum_i c_h[i] * hist[i]       (n_hist <= 3)
 * and eps_out <- the converted model output the scheduler keeps (x0 prediction = x - sigma * eps).
 * replaces: "noise_pred = noise_uncond + guidance_scale * (noise_pred - noise_uncond)" and the
 *   tensor arithmetic of UniPCMultistepScheduler.step (inference_t23d.py:65-70,94-103).
 * ------------------------------------------------------------------------------------------ */
int vist3a_cfg_combine(const void* cond, const void* uncond, int32_t in_dtype, float guidance, float* out, int64_t n,
                       void* stream);
int vist3a_axpby_n(float* out, int32_t n_terms, const float* const* terms, const float* coeffs, int64_t n,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIST3A_SM100_H_ */
