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
int vist3a_abi_version(void); /* 8 */
/* number of kernels this library has launched from the calling process (all threads) */
int64_t vist3a_launch_count(void);
/* Programmatic dependent launch (PDL) of the hot kernels (GEMM, attention, LayerNorm, RMSNorm+RoPE, row_rinv): each is launched
 * with the programmatic-stream-serialization attribute and executes griddepcontrol.wait before its first global-memory access,
 * so its prologue (mbarrier init, TMEM allocation, tensor-map prefetch, CTA scheduling) overlaps the tail of the kernel before
 * it in the stream -- also inside a captured CUDA graph.  Results are identical either way.  Default: on (env VIST3A_PDL=0 turns
 * it off at load).  Returns the previous setting.  No counterpart in the reference (PyTorch launches are fully serialised). */
int vist3a_set_pdl(int32_t enable);

/* ------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = epilogue( A[M,K] · W[N,K]^T )      tcgen05.mma + TMA + TMEM accumulators
 * replaces: every nn.Linear on the path — diffusers WanTransformerBlock q/k/v/out/ffn
 *   (call sites inference_t23d.py:94-103), AS/model/encoder/vggt/layers/attention.py:55,75
 *   (qkv, proj), AS/.../layers/mlp.py:34-40 (fc1, fc2), the 1x1 convs of
 *   AS/.../heads/dpt_head.py:69-79, and (after im2col) models/stitching_layer_builder.py:32-42.
 * Epilogue, in order:  v = acc + bias[n];  v = act(v);  if round_linear: v = bf16(v);
 *   if gate: v *= gate[(row / rows_per_batch) * gate_bstride + n];  if round_gate: v = bf16(v);
 *   if residual: v += residual[rmap(row), n];  if residual2: v += residual2[cmap(row), n];  v = post_act(v);
 *   store C[cmap(row), n] as out_dtype.
 * ------------------------------------------------------------------------------------------ */
#define VIST3A_ACT_NONE 0
#define VIST3A_ACT_GELU_TANH 1 /* FeedForward(activation_fn="gelu-approximate"), Wan text embedder */
#define VIST3A_ACT_GELU_ERF 2  /* AS/.../layers/mlp.py:25 (nn.GELU) */
#define VIST3A_ACT_SILU 3      /* Wan time embedder */
#define VIST3A_ACT_RELU 4      /* DPT heads */

#define VIST3A_GEMM_FLAG_2CTA 1u       /* use cta_group::2 pairs (256-row tiles) */
#define VIST3A_GEMM_FLAG_1CTA 2u       /* force single-CTA tiles */
#define VIST3A_GEMM_FLAG_MULTICAST 8u  /* pairs only, even number of 256-wide column tiles: clusters of two pairs share the A rows by TMA multicast */
#define VIST3A_GEMM_FLAG_STAGED 16u    /* A/B: output / residual through the shared-memory staging buffer instead of 256-bit per-thread accesses */
#define VIST3A_GEMM_FLAG_BN176 4u      /* A/B: 176-wide column tiles for CTA pairs where they fill the waves better (measured slower) */

/* row -> memory-row mapping of C / residual:  mem_row = (row / rpg) * gstride + goff + row % rpg   (rpg == 0: identity).
 * Lets a GEMM write into (or add from) a token buffer whose groups of rows are separated by other rows, e.g. the
 * stitching conv writing patch tokens behind the 5 special tokens of every view (models/anysplat_stitched.py:181-202),
 * or add one [rpg, N] table to every group (gstride = 0: DINO / DPT positional embeddings). */
typedef struct vist3a_rowmap {
  int64_t rpg, gstride, goff;
} vist3a_rowmap;

/* implicit-GEMM convolution over an NHWC activation (A operand fetched by 4-D TMA, zero padding by out-of-bounds fill):
 *   A = x[n_img, h_in, w_in, c_in];  W = [N, kh*kw*c_in] with k = (dy*kw + dx)*c_in + c;  stride 1;
 *   output row = pixel (n, y, x) of the h_out x w_out map, M = n_img*h_out*w_out.
 * pix_stride / row_stride / img_stride (elements, multiples of 4 floats / 8 bf16; 0 = dense NHWC) describe where pixel
 * (n, y, x) starts.  A pix_stride smaller than c_in gives OVERLAPPING windows: with a 4-channel (RGB0) image whose rows are
 * physically zero-padded, c_in = 32, pix_stride = 4, kw = 1 one TMA box row holds the 8 horizontal taps x 4 channels of a
 * pixel, i.e. a 7x7 RGB convolution becomes kh = 7 k-blocks without any im2col buffer.
 * replaces: the 3x3 nn.Conv2d layers of AS/model/encoder/vggt/heads/dpt_head.py:346-359,375-376,437-439 and the 7x7
 *   input_merger of AS/model/encoder/heads/vggt_dpt_gs_head.py:73-76 (cuDNN). */
typedef struct vist3a_conv {
  int32_t enabled;
  int32_t kh, kw, pad_y, pad_x;
  int32_t n_img, h, w, c_in; /* stride-1 conv: h_out = h + 2*pad_y - kh + 1, w_out = w + 2*pad_x - kw + 1 */
  int32_t kt;                /* 0 / 1: 2-D conv.  > 1: causal temporal taps over the image (= frame) index: output frame t reads frames
                                t + dt - (kt - 1), dt = 0..kt-1, frames before the first are zero (WanCausalConv3d, utils/wan_utils.py:96-147);
                                K = kt*kh*kw*c_in, taps ordered (dt, dy, dx).  One clip per call (the frame shift must not cross clips). */
  int64_t pix_stride, row_stride, img_stride;
} vist3a_conv;

typedef struct vist3a_gemm_args {
  const void* A;          /* [M, K] in_dtype, row stride lda  (conv mode: NHWC tensor) */
  const void* W;          /* [N, K] in_dtype, row stride ldw (nn.Linear weight layout) */
  void* C;                /* [M, N] out_dtype, row stride ldc */
  const float* bias;      /* [N] fp32 or NULL */
  const float* gate;      /* fp32, index (row / rows_per_batch) * gate_bstride + n, or NULL */
  const void* residual;   /* out_dtype, row stride ldr, rows mapped by rmap, or NULL (may alias C) */
  const void* residual2;  /* out_dtype, row stride ldc, rows mapped like C, or NULL */
  int64_t M, N, K;
  int64_t lda, ldw, ldc, ldr;
  int64_t rows_per_batch; /* >= 1 */
  int64_t gate_bstride;   /* 0 => one gate vector shared by all rows (LayerScale) */
  vist3a_rowmap cmap;     /* row mapping of C and residual2 */
  vist3a_rowmap rmap;     /* row mapping of residual */
  vist3a_conv conv;
  int32_t in_dtype;       /* VIST3A_DTYPE_BF16 (kind::f16) or VIST3A_DTYPE_F32 (kind::tf32) */
  int32_t out_dtype;
  int32_t act;            /* applied to acc + bias */
  int32_t post_act;       /* applied after the residual adds (VIST3A_ACT_NONE / VIST3A_ACT_RELU) */
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
  uint32_t flags; /* 0 = tuned default; kernel-variant selectors for A/B measurements only (results identical up to rounding):
                     bit0 one thread per query row, bit1 128-key steps (d=128), bit2 single MMA-issuing warp, bits 3.. = 1 + FMA-pipe exp2 share,
                     bit8 | variant << 9 the CTA-pair kernel, bit13 / 14 / 16 former default / two threads per row / one tile per CTA;
                     work decomposition, no kernel selection: bit17 no key split of the last wave, bit20 one cluster per query block
                     instead of persistent clusters, bit21 CUDA-core CTAs for 1..8 tail rows (head_dim 64);
                     timing only (WRONG results): bit12 softmax skipped, bit22 merge kernel skipped */
  const float* q_row_scale; /* optional [batch * len_q] fp32: extra positive factor on the logits of query row (b, i), all heads */
  void* workspace;          /* optional scratch (128-byte aligned), see vist3a_fmha_workspace_bytes; NULL / too small: the call still succeeds */
  int64_t workspace_bytes;
} vist3a_fmha_args;

int vist3a_fmha_fwd(const vist3a_fmha_args* args, void* stream);
/* Scratch the call would use for these arguments (>= 0; < 0: a VIST3A_ERR_* code).  head_dim 128 on CTA pairs: when the grid's last wave would
 * leave most SMs idle, the query blocks of that wave are cut along the KEYS into chunks that fill the machine (flash-decoding style partial
 * results: bf16 O + fp32 (max, sum) per row, 264 B per query row and chunk) and a merge kernel follows; without workspace the kernel runs
 * unsplit (same result up to rounding).  No reference counterpart: torch SDPA owns its scratch. */
int64_t vist3a_fmha_workspace_bytes(const vist3a_fmha_args* args);

/* ------------------------------------------------------------------------------------------
 * LayerNorm with optional modulation:  out[r,:] = LN(x[r,:]) * mul[b,:] + add[b,:],  b = r / rows_per_batch
 * LN statistics in fp32 (biased variance), eps inside the sqrt.  mul_plus_one: use (1 + mul).  in_map / out_map (may be
 * NULL) map row r to a memory row of x / out (e.g. the patch tokens behind the 5 special tokens of every view).
 * replaces: diffusers FP32LayerNorm + AdaLN-zero "(norm(x) * (1 + scale) + shift)" (mul = 1+scale,
 *   add = shift, batch stride = mul_bstride), FP32LayerNorm(elementwise_affine=True) of norm2
 *   (mul = weight, add = bias, bstride 0), and nn.LayerNorm in AS/.../layers/block.py:62,73.
 * ------------------------------------------------------------------------------------------ */
int vist3a_layernorm(const void* x, int32_t x_dtype, int64_t ldx, void* out, int32_t out_dtype, int64_t ldo,
                     int64_t rows, int64_t dim, int64_t rows_per_batch, const float* mul, int64_t mul_bstride,
                     const float* add, int64_t add_bstride, float eps, int32_t mul_plus_one,
                     const vist3a_rowmap* in_map, const vist3a_rowmap* out_map, void* stream);

/* ------------------------------------------------------------------------------------------
 * RMSNorm across all heads (+ optional interleaved-pair RoPE), in place on a strided bf16 matrix.
 *   x[r, 0:dim] <- x * rsqrt(mean(x^2) + eps) * weight ; then per head h and pair j:
 *   (x[2j], x[2j+1]) <- (x[2j] c - x[2j+1] s, x[2j] s + x[2j+1] c), (c,s) = cos/sin[(r % rope_len), j]
 * replaces: diffusers WanAttnProcessor2_0: attn.norm_q / norm_k (RMSNorm "rms_norm_across_heads",
 *   eps 1e-6) and apply_rotary_emb with WanRotaryPosEmbed frequencies.
 * cos/sin: [rope_len, head_dim/2] fp32, or NULL for no RoPE (cross-attention).
 * nseg > 1 processes nseg column segments of the same rows in one launch (q and k of a fused qkv buffer): segment s starts
 * at column s * seg_stride and uses weight[s * dim : (s + 1) * dim].
 * ------------------------------------------------------------------------------------------ */
int vist3a_rmsnorm_rope(void* x, int64_t ldx, int64_t rows, int64_t dim, int64_t head_dim, const float* weight,
                        float eps, const float* rope_cos, const float* rope_sin, int64_t rope_len, int64_t nseg,
                        int64_t seg_stride, void* stream);

/* out[r] = rsqrt(mean(x[r, 0:dim]^2) + eps) for a bf16 matrix (read-only pass).  With vist3a_fmha_args.q_row_scale this is the
 * RMSNorm of the cross-attention queries: (q * rinv_r * w_q) . k == rinv_r * (q . (w_q * k)), so the per-row factor scales
 * the logits inside the attention kernel and w_q is folded into the (per-prompt cached) text keys.
 * replaces: attn2.norm_q of diffusers WanAttnProcessor2_0. */
int vist3a_row_rinv(const void* x, int64_t ldx, int64_t rows, int64_t dim, float eps, float* out, void* stream);

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
 * Skinny linear for M <= 16 rows (HBM-bound weight streaming):  y = residual + gate[n] * act(pre_act(x) W^T + b)
 * (gate [N] fp32 = LayerScale, residual [M, N] fp32; both optional)
 * replaces: Wan TimestepEmbedding / time_proj (M = batch) and every Linear of
 *   AS/model/encoder/vggt/heads/camera_head.py:87-170 (M = views).
 * ------------------------------------------------------------------------------------------ */
int vist3a_skinny_linear(const void* x, int32_t x_dtype, int64_t ldx, const void* W, int32_t w_dtype, int64_t ldw,
                         const float* bias, void* y, int32_t y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                         int32_t pre_act, int32_t act, const float* gate, const float* residual, int64_t ldres,
                         void* stream);

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


/* ------------------------------------------------------------------------------------------
 * Stitched latent -> 3D-Gaussian decoder kernels
 * ------------------------------------------------------------------------------------------ */

/* im2col of the stitching Conv3d with the trilinear T-upsample fused in:
 *   latent [B, C, T, h, w] (fp32 or bf16) -> A [B*V*(h/2)*(w/2), C*45] bf16, V = 4(T-1)+1,
 *   A[(b,v,oy,ox), c*45 + kt*9 + ky*3 + kx] = up(latent)[b, c, clamp(v+kt-2), clamp(2oy+ky-1), clamp(2ox+kx-1)]
 *   with up() the align_corners=True linear interpolation along T.
 * replaces: upsampling_layer (models/stitched_model.py:92-107) + the input side of
 *   nn.Conv3d(k=(5,3,3), s=(1,2,2), p=(2,1,1), padding_mode="replicate") (models/stitching_layer_builder.py:32-42). */
int vist3a_im2col_stitch(const void* latent, int32_t dtype, void* A, int64_t B, int64_t C, int64_t T, int64_t h,
                         int64_t w, void* stream);

/* generic NHWC im2col (fp32 -> fp32):  A[(n,oy,ox), (dy*kw+dx)*C + c] = x[n, oy*stride+dy-pad, ox*stride+dx-pad, c] (0 outside),
 * row pitch ldA >= kh*kw*C (extra columns zeroed).  Used for the 7x7 RGB `input_merger`
 * (AS/model/encoder/heads/vggt_dpt_gs_head.py:73-76) and the stride-2 3x3 resize conv (dpt_head.py:85-90). */
int vist3a_im2col_nhwc(const float* x, float* A, int64_t ldA, int64_t n_img, int64_t h, int64_t w, int64_t C,
                       int32_t kh, int32_t kw, int32_t stride, int32_t pad, void* stream);

/* RGB views -> zero-padded 4-channel NHWC image for the overlapping-window 7x7 convolution (see vist3a_conv):
 *   image [B, 3, V, H, W] in [-1, 1] (fp32 or bf16)  ->  out [B*V, H, W + 8, 4] fp32,
 *   out[(b,v), y, 3 + x, c] = (image[b, c, v, y, x] + 1) / 2 for c < 3, zero elsewhere.
 * replaces: "context_image = (rearrange(context_image, 'b c v h w -> b v c h w') + 1) / 2" (models/anysplat_stitched.py:172-174)
 *   and the zero padding of the 7x7 input_merger convolution. */
int vist3a_rgb_to_nhwc4pad(const void* image, int32_t dtype, float* out, int64_t B, int64_t V, int64_t H, int64_t W, void* stream);
/* the same for the un-stitched (image -> 3DGS) path: image [B, V, 3, H, W] already in [0, 1], copied unscaled
 * (EncoderAnySplat.forward passes `image` to the Gaussian head as is, AS/model/encoder/anysplat.py:449-455) */
int vist3a_rgb01_views_to_nhwc4pad(const void* image, int32_t dtype, float* out, int64_t B, int64_t V, int64_t H, int64_t W, void* stream);

/* DINOv2 patch embedding (stride-p, p x p convolution, no overlap) as a GEMM operand:
 *   image [n_img, 3, H, W] in [0, 1] (fp32 or bf16)  ->  A [n_img * (H/p) * (W/p), ldA] bf16,
 *   A[(n, gy, gx), c*p*p + py*p + px] = (bf16(image[n, c, gy*p+py, gx*p+px]) - mean3[c]) / std3[c], columns >= 3 p^2 zero;
 *   follow with vist3a_gemm against the flattened conv weight [embed_dim, 3 p^2] (zero-padded to ldA columns).
 * mean3 / std3: HOST pointers to 3 floats (ImageNet statistics).
 * replaces: Aggregator.forward normalisation (AS/model/encoder/vggt/models/aggregator.py:228) + PatchEmbed.proj
 *   (AS/model/encoder/vggt/layers/patch_embed.py:65,68-81) of the un-stitched AnySplat encoder. */
int vist3a_patch_embed_im2col(const void* image, int32_t dtype, void* A, int64_t ldA, int64_t n_img, int64_t H, int64_t W, int32_t patch,
                              const float* mean3, const float* std3, void* stream);

/* per-head LayerNorm(head_dim = 64) of q and k + 2-D rotary embedding, in place on a fused bf16 [rows, 3*heads*64] qkv buffer.
 *   token p = row % tokens_per_view; p < n_special: position (0,0); else (1 + (p-n_special)/grid_w, 1 + (p-n_special)%grid_w)
 *   rope: each 32-wide half (y then x) rotates pairs (j, j+16) by pos * base^(-j/16); cos/sin tables [max_pos, 16] fp32.
 * replaces: q_norm/k_norm (AS/.../layers/attention.py:42-43,57) + RotaryPositionEmbedding2D (layers/rope.py:133-188). */
int vist3a_qknorm_rope2d(void* qkv, int64_t ld, int64_t rows, int64_t heads, const float* qw, const float* qb,
                         const float* kw, const float* kb, float eps, const float* cos_tab, const float* sin_tab,
                         int64_t max_pos, int64_t tokens_per_view, int64_t n_special, int64_t grid_w, void* stream);

/* bilinear resize, align_corners=True, NHWC fp32, with optional fused adds:
 *   out[n,y,x,c] = lerp(in)[n,y,x,c] + (add ? add[n,y,x,c] : 0) + (pos_x ? (c < C/2 ? pos_x[x, c] : pos_y[y, c - C/2]) : 0)
 * replaces: custom_interpolate (dpt_head.py:477-502), "out + direct_img_feat" (vggt_dpt_gs_head.py:168-169) and
 *   _apply_pos_embed on the full-resolution map (dpt_head.py:267-277). */
int vist3a_bilinear_nhwc(const float* in, float* out, int64_t n_img, int64_t h_in, int64_t w_in, int64_t h_out,
                         int64_t w_out, int64_t C, const float* add, const float* pos_x, const float* pos_y,
                         void* stream);

/* depth-to-space after a k=s ConvTranspose2d expressed as a GEMM:  in [n*h*w, k*k*C] (col = (dy*k+dx)*C + c)
 * -> out NHWC [n, h*k, w*k, C].   replaces: nn.ConvTranspose2d(k=4,s=4 / k=2,s=2) output scatter (dpt_head.py:85-90). */
int vist3a_depth_to_space(const float* in, float* out, int64_t n_img, int64_t h, int64_t w, int64_t C, int32_t k,
                          void* stream);

/* fp32 attention for short sequences (L <= 32):  qkv [B, L, 3, H, D] -> out [B, L, H*D].
 * replaces: F.scaled_dot_product_attention inside the camera-head trunk (AS/.../heads/camera_head.py:58-69; 13 tokens). */
int vist3a_attention_small(const float* qkv, float* out, int64_t B, int64_t L, int64_t H, int64_t D, float scale,
                           void* stream);

/* out[r, :] = a[r, :] * b[r, :] + c[r, :]  (fp32, row strides in elements)
 * replaces: "gate_msa * modulate(...) + pose_tokens" (camera_head.py:140-144). */
int vist3a_fma_rows(float* out, int64_t ldo, const float* a, int64_t lda, const float* b, int64_t ldb, const float* c,
                    int64_t ldc, int64_t rows, int64_t dim, void* stream);

/* transposed epilogue of a swapped-operand skinny linear.  For M <= 16 tokens the weight matrix is the streamed (A) operand
 * of vist3a_gemm:  ct[n, m] = sum_k W[n, k] x[m, k]  (ct is [N, ldct >= 16] fp32); this finishes
 *   y[m, n] = residual[m, n] + gate[n] * act(ct[n, m] + bias[n])        (bias / gate / residual optional)
 * splits = S > 1: split-K by reshaping, so that a 2048-row weight matrix fills the GPU instead of 16 CTAs -- the caller ran the GEMM on
 *   W viewed as [N*S, K/S] (row n*S + s = the s-th K-slice of row n: the same memory) against x viewed as [16*S, K/S]; the partial sums
 *   of output (n, m) are the diagonal blocks ct[n*S + s, m*S + s], summed here (the off-diagonal products are wasted tensor work on a
 *   weight-bandwidth-bound operation).
 * replaces: bias, activation, LayerScale and residual add of the nn.Linear layers inside the camera-head trunk
 *   (AS/.../heads/camera_head.py:87-170; Block.forward AS/.../layers/block.py:81-107 at 13 tokens). */
int vist3a_bias_act_t(const float* ct, int64_t ldct, const float* bias, int32_t act, const float* gate, const float* residual,
                      int64_t ldr, float* y, int64_t ldy, int64_t M, int64_t N, int32_t splits, void* stream);

/* camera head output: activate pose encodings (relu on the 2 FoV entries) and build cameras.
 *   pose_raw [S, 9] -> pose_act [S, 9]; extr [S, 3, 4] world->cam; intr [S, 3, 3] in pixels;
 *   c2w [S, 4, 4] = inverse([R|t; 0 0 0 1]); intr_norm [S, 3, 3] (rows 0/1 divided by W/H).  Any output may be NULL.
 * replaces: activate_pose (head_act.py:12-35), pose_encoding_to_extri_intri (utils/pose_enc.py:65-130),
 *   pose packing (models/anysplat_stitched.py:475-494). */
int vist3a_pose_to_cameras(const float* pose_raw, float* pose_act, float* extr, float* intr, float* c2w,
                           float* intr_norm, int64_t S, int64_t H, int64_t W, void* stream);

/* fused per-pixel epilogue of the decoder (HBM bound):
 *   depth[p] = exp(dot(depth_feat[p, 0:cd], depth_w) + depth_b)                       (dpt_head.py output_conv2[2] + "exp")
 *   means[p] = R^T ((u-cu) d / fx, (v-cv) d / fy, d) - R^T t                          (utils/geometry.py:10-58)
 *   opacity = sigmoid(raw[0]); scales = min(0.001 softplus(raw[1:4]), 0.3); rot = q / (|q| + 1e-8);
 *   harmonics = raw[8:8+3*d_sh] * sh_mask; cov = R S S^T R^T                          (gaussian_adapter.py:114-147)
 *   scene_sum += |means[p]| (one atomicAdd per block)
 * gs_raw [P, ld_raw] fp32 (P = S*H*W pixels, view-major), outputs contiguous fp32.
 * replaces: models/anysplat_stitched.py:358-376,410-474. */
int vist3a_gaussian_epilogue(const float* depth_feat, int64_t ld_df, int64_t cd, const float* depth_w, float depth_b,
                             const float* gs_raw, int64_t ld_raw, const float* extr, const float* intr,
                             const float* sh_mask, int64_t d_sh, int64_t S, int64_t H, int64_t W, float* depth,
                             float* means, float* scales, float* rotations, float* opacities, float* harmonics,
                             float* covariances, float* scene_sum, void* stream);

/* Gaussian adapter on given positions (the voxelize=True branch): rows of feats hold (density, scales 3, quaternion xyzw 4,
 * SH 3*d_sh) = raw_gs_dim values; means = pts, the other outputs as vist3a_gaussian_epilogue.
 * replaces: map_pdf_to_opacity + UnifiedGaussianAdapter.forward on the fused voxels (models/anysplat_stitched.py:457-474,
 *   AS/model/encoder/common/gaussian_adapter.py:114-147). */
int vist3a_gaussian_adapter(const float* pts, const float* feats, int64_t ld_feats, const float* sh_mask, int64_t d_sh, int64_t P,
                            float* means, float* scales, float* rotations, float* opacities, float* harmonics, float* covariances,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * Voxelised Gaussian fusion (integer / index work, HBM bound; no tensor cores)
 * replaces: EncoderAnySplat.voxelizaton_with_fusion (AS/model/encoder/anysplat.py:298-335), called per batch element from
 *   models/anysplat_stitched.py:419-440 when cfg.voxelize is set (released AnySplat configs: voxelize true, voxel_size 0.002):
 *     voxel   = round_half_even(pts / voxel_size) -> int32 per axis  (IEEE fp32 division, as the reference's CPU path computes it;
 *               torch's CUDA kernel for tensor / python-scalar multiplies by the rounded reciprocal, which differs by one ulp for a
 *               few points per million at voxel_size 0.002 -- cell membership of those border points follows the CPU result)
 *     unique voxels in lexicographic (x, y, z) order (torch.unique(dim=0)), inverse index, counts
 *     w_i     = exp(conf_i - max_voxel conf) / (sum_voxel exp(conf - max) + 1e-6)
 *     voxel_pts = sum_i w_i pts_i ; voxel_feats = sum_i w_i feats_i      (summed in point order: the sort is stable)
 * pts [N, 3] fp32; feats rows of feat_dim (<= 125) fp32 at stride ld_feats; conf_i = conf[i * conf_stride] (the confidence may be
 * a column of the same rows).  voxel_pts [N, 3] and voxel_feats [N, feat_dim] have capacity for N voxels; the first *n_voxels
 * rows are written.  inverse [N] / counts [N] (int32) are optional (NULL).  *n_voxels (device or pinned-host int64) is written by
 * the last kernels of the call: read it after synchronising `stream`; -1 = the coordinate ranges need more than 64 key bits.
 * workspace: vist3a_voxel_fusion_workspace_bytes(N) bytes of device memory, 256-byte aligned, caller-owned.
 * Every stage is asynchronous on `stream` (no host round trip): radix passes beyond ceil(key bits / 8) exit immediately. */
int64_t vist3a_voxel_fusion_workspace_bytes(int64_t n_points);
int vist3a_voxel_fusion(const float* pts, const float* feats, int64_t ld_feats, int64_t feat_dim, const float* conf, int64_t conf_stride,
                        int64_t n_points, float voxel_size, float* voxel_pts, float* voxel_feats, int32_t* inverse, int32_t* counts,
                        int64_t* n_voxels, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3D-Gaussian rasteriser, forward (the consumer of the Gaussians: novel-view rendering)
 * replaces: gsplat 1.4.0 `rasterization(means, quats, scales, opacities, sh, viewmat, K, W, H, sh_degree=4, render_mode="RGB+D",
 *   packed=False, near_plane=1e-10, backgrounds, radius_clip=0.1, covars, rasterize_mode="classic")`, called once per view from
 *   DecoderSplattingCUDA.rendering_fn (AS/model/decoder/decoder_splatting_cuda.py:43-125, call :92-112).  gsplat is absent from the
 *   reference tree (requirements.txt:17); its published algorithm is restated (see csrc/gs_render.cu, oracle/gsplat_ref.py).
 * Phase 1, vist3a_gs_project: per Gaussian camera-space projection (covariances [N,3,3] given), eps2d blur, conic, 3-sigma radius, culling,
 *   degree <= 4 spherical-harmonic colour from harmonics [N, 3, d_sh] (the decoder's layout), 16x16-tile box, exclusive scan of the tile
 *   counts.  viewmat (row-major 4x4 world->camera) and K (row-major 3x3, pixels) are HOST pointers.  *n_isect (device int64) receives the
 *   number of (tile, Gaussian) intersections: read it after synchronising, then size phase 2's workspace.
 * Phase 2, vist3a_gs_rasterize: intersection keys (tile id | depth bits), stable radix sort, per-tile ranges, front-to-back compositing.
 *   background: HOST pointer to 3 floats.  Outputs rgb [H, W, 3] (unclamped, as gsplat), depth [H, W] (accumulated alpha-weighted depth,
 *   "RGB+D"), alpha [H, W].
 * Workspaces are caller-owned device memory, 256-byte aligned; the project workspace is an input of phase 2. */
int64_t vist3a_gs_project_workspace_bytes(int64_t n_gaussians);
int vist3a_gs_project(const float* means, const float* covariances, const float* opacities, const float* harmonics, int64_t d_sh, int32_t sh_degree,
                      int64_t n_gaussians, const float* viewmat, const float* K, int64_t W, int64_t H, float near_plane, float far_plane,
                      float radius_clip, float eps2d, void* workspace, int64_t workspace_bytes, int64_t* n_isect, void* stream);
int64_t vist3a_gs_rasterize_workspace_bytes(int64_t n_isect, int64_t W, int64_t H);
int vist3a_gs_rasterize(const void* project_workspace, int64_t n_gaussians, int64_t n_isect, int64_t W, int64_t H, const float* background,
                        void* workspace, int64_t workspace_bytes, float* rgb, float* depth, float* alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * Wan-2.1 VAE decode (pipe.vae.decode between the denoiser and the stitched decoder, inference_t23d.py:104-114; arithmetic vendored at
 * utils/wan_utils.py:96-1180).  Activations are NDHWC bf16, one clip [T, H, W, ld]; every convolution is vist3a_gemm in conv mode
 * (`kt` temporal taps = WanCausalConv3d :96-147).  These are the HBM-bound passes between the convolutions.
 * ------------------------------------------------------------------------------------------ */
/* y[r, c] = silu?(x[r, c] / max(||x[r, :C]||_2, 1e-12) * sqrt(C) * gamma[c]), c < C; y[r, C:ldy] = 0 (zero padding channels of the next
 * convolution's operand).  replaces: WanRMS_norm (:150-184, channel_first images) + nn.SiLU of WanResidualBlock (:366-372, :395-399) and
 * the un-activated norm of WanAttentionBlock (:449). */
int vist3a_vae_rmsnorm(const void* x, int64_t ldx, const float* gamma, void* y, int64_t ldy, int64_t rows, int64_t C, int32_t silu, void* stream);
/* p[r, :] = softmax(scale * s[r, :]) in bf16 from fp32 logits.  replaces: the softmax inside F.scaled_dot_product_attention of
 * WanAttentionBlock (:463-467: one head of width C over the H*W positions of a frame). */
int vist3a_softmax_rows(const float* s, void* p, int64_t rows, int64_t L, int64_t valid, int64_t ldp, float scale, void* stream);
/* out[2t + half, p, :] = y[t, p, half*C : (half+1)*C]  (y [T, P, 2C] -> out [2T, P, C], bf16).  replaces: the channel-halves-to-time
 * interleave of WanResample "upsample3d" (:304-306). */
int vist3a_time_interleave(const void* y, int64_t ldy, void* out, int64_t ldo, int64_t T, int64_t P, int64_t C, void* stream);
/* out[c, r] = in[r, c]  (bf16 [R, C] with row stride ld_in -> [C, R]): V^T operand of the mid-block attention's P V GEMM */
int vist3a_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t R, int64_t C, void* stream);
/* depth-to-space (k = 2) of the parity-decomposed up-sampling convolution: in [n*h*w, 4*C] (col = (py*2 + px)*C + c) -> NHWC bf16
 * [n, 2h, 2w, ldo] (channels [0, C)).  replaces: WanUpsample (nearest-exact 2x) + Conv2d 3x3 output layout (:226-238). */
int vist3a_depth_to_space2_bf16(const void* in, void* out, int64_t n_img, int64_t h, int64_t w, int64_t C, int64_t ldo, void* stream);
/* latent [C, T*h*w] (fp32 or bf16) -> NDHWC bf16 [T*h*w, ld], channels [C, ld) zero: operand of post_quant_conv / decoder.conv_in */
int vist3a_latent_to_ndhwc(const void* z, int32_t z_dtype, void* out, int64_t C, int64_t THW, int64_t ld, void* stream);
/* conv_out result [T*H*W, ld] fp32 (channels 0..2) -> frames [3, T*H*W] fp32 clamped to [-1, 1] (AutoencoderKLWan._decode :1115) */
int vist3a_vae_frames_out(const float* y, int64_t ld, float* out, int64_t THW, void* stream);

/* planar bilinear resize, half-pixel centres: in [planes, h_in, w_in] fp32 -> out [planes, h_out, w_out].  replaces: F.interpolate(samples,
 * (T, 448, 448), mode="trilinear", align_corners=False) of the decoded frames (inference_t23d.py:116-123; the frame count is kept, so the
 * temporal weights are the identity). */
int vist3a_resize_planes(const float* in, float* out, int64_t planes, int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Confidence-quantile branches of the stitched decoder (EncoderAnySplatCfg.render_conf / opacity_conf; models/anysplat_stitched.py:381-387,
 * 442-455, 463-467).  HBM-bound index work on caller-owned workspaces (256-byte aligned device memory).
 * ------------------------------------------------------------------------------------------ */
/* conf[p] = 1 + exp(dot(feat[p, 0:C], w) + bias): the depth head's confidence channel ("expp1", AS/.../heads/head_act.py:102-103) from the
 * 32-wide feature rows in front of its last 1x1 convolution */
int vist3a_depth_conf(const float* feat, int64_t ld, int64_t C, const float* w, float bias, float* conf, int64_t n_pixels, void* stream);
/* *out (device float) = torch.quantile(x, q) over n floats, interpolation "linear" (rank q (n - 1) in fp32; torch.lerp between the two order
 * statistics): stable radix sort of order-preserving keys.  replaces: torch.quantile(depth_conf.flatten(0, 1), conf_threshold) (:382-384, :464) */
int64_t vist3a_quantile_workspace_bytes(int64_t n);
int vist3a_quantile_f32(const float* x, int64_t n, float q, float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* ordered compaction: rows i with conf[i] > *threshold (all rows if use_threshold == 0), in index order -> out_feats [count, C] (from feats rows
 * of stride ld_feats), out_pts [count, 3], out_damp [count] = sigmoid(conf[i] - *threshold) (may be NULL); *count (device int64) = kept rows.
 * replaces: `anchor_feats[b_i].permute(0, 2, 3, 1)[conf_valid_mask[b_i]]`, `pts_all[b_i][conf_valid_mask[b_i]]` (:442-446) and
 * `torch.sigmoid(depth_conf - shift)[conf_valid_mask]` (:465-467). */
int64_t vist3a_compact_rows_workspace_bytes(int64_t n);
int vist3a_compact_rows(const float* conf, const float* threshold, int32_t use_threshold, int64_t n, const float* feats, int64_t ld_feats, int64_t C,
                        const float* pts, float* out_feats, float* out_pts, float* out_damp, int64_t* count, void* workspace, int64_t workspace_bytes,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIST3A_SM100_H_ */
