// extern "C" surface of libvist3a_sm100 (declared in include/vist3a_sm100.h) plus the host
// utilities shared by the kernels' launchers.
#include <stdlib.h>

#include <atomic>

#include "host_util.cuh"

namespace v3a {

// entry points implemented next to their kernels
int gemm_entry(const vist3a_gemm_args*, cudaStream_t);
int fmha_entry(const vist3a_fmha_args*, cudaStream_t, long long* ws_query);
int layernorm_entry(const void*, int, long long, void*, int, long long, long long, long long, long long, const float*,
                    long long, const float*, long long, float, int, const vist3a_rowmap*, const vist3a_rowmap*, cudaStream_t);
int rmsnorm_rope_entry(void*, long long, long long, long long, long long, const float*, float, const float*,
                       const float*, long long, long long, long long, cudaStream_t);
int row_rinv_entry(const void*, long long, long long, long long, float, float*, cudaStream_t);
int modulation_entry(const float*, const void*, int, int, float*, long long, long long, long long, unsigned,
                     cudaStream_t);
int skinny_linear_entry(const void*, int, long long, const void*, int, long long, const float*, void*, int, long long,
                        long long, long long, long long, int, int, const float*, const float*, long long, cudaStream_t);
int im2col_stitch_entry(const void*, int, void*, long long, long long, long long, long long, long long, cudaStream_t);
int im2col_nhwc_entry(const float*, float*, long long, long long, long long, long long, long long, int, int, int, int, cudaStream_t);
int qknorm_rope2d_entry(void*, long long, long long, long long, const float*, const float*, const float*, const float*, float,
                        const float*, const float*, long long, long long, long long, long long, cudaStream_t);
int bilinear_nhwc_entry(const float*, float*, long long, long long, long long, long long, long long, long long, const float*,
                        const float*, const float*, cudaStream_t);
int depth_to_space_entry(const float*, float*, long long, long long, long long, long long, int, cudaStream_t);
int attention_small_entry(const float*, float*, long long, long long, long long, long long, float, cudaStream_t);
int bias_act_t_entry(const float*, long long, const float*, int, const float*, const float*, long long, float*, long long, long long,
                     long long, int, cudaStream_t);
int rgb_to_nhwc4pad_entry(const void*, int, float*, long long, long long, long long, long long, int, float, float, cudaStream_t);
int patch_embed_im2col_entry(const void*, int, void*, long long, long long, long long, long long, int, const float*, const float*, cudaStream_t);
int fma_rows_entry(float*, long long, const float*, long long, const float*, long long, const float*, long long, long long,
                   long long, cudaStream_t);
int pose_to_cameras_entry(const float*, float*, float*, float*, float*, float*, long long, long long, long long, cudaStream_t);
int gaussian_epilogue_entry(const float*, long long, long long, const float*, float, const float*, long long, const float*,
                            const float*, const float*, long long, long long, long long, long long, float*, float*, float*, float*,
                            float*, float*, float*, float*, cudaStream_t);
int gaussian_adapter_entry(const float*, const float*, long long, const float*, long long, long long, float*, float*, float*, float*, float*,
                           float*, cudaStream_t);
long long voxel_fusion_workspace_bytes(long long);
int voxel_fusion_entry(const float*, const float*, long long, long long, const float*, long long, long long, float, float*, float*, int*, int*,
                       long long*, void*, long long, cudaStream_t);
long long gs_project_workspace_bytes(long long);
long long gs_rasterize_workspace_bytes(long long, long long, long long);
int gs_project_entry(const float*, const float*, const float*, const float*, long long, int, long long, const float*, const float*, long long, long long,
                     float, float, float, float, void*, long long, long long*, cudaStream_t);
int gs_rasterize_entry(const void*, long long, long long, long long, long long, const float*, void*, long long, float*, float*, float*, cudaStream_t);
int timestep_features_entry(const float*, void*, int, long long, long long, cudaStream_t);
int patchify_entry(const void*, int, void*, long long, long long, long long, long long, long long, cudaStream_t);
int unpatchify_entry(const void*, int, long long, void*, int, long long, long long, long long, long long, long long,
                     cudaStream_t);
int cfg_combine_entry(const void*, const void*, int, float, float*, long long, cudaStream_t);
int axpby_entry(float*, int, const float* const*, const float*, long long, cudaStream_t);
int vae_rmsnorm_entry(const void*, long long, const float*, void*, long long, long long, long long, int, cudaStream_t);
int softmax_rows_entry(const float*, void*, long long, long long, long long, long long, float, cudaStream_t);
int time_interleave_entry(const void*, long long, void*, long long, long long, long long, long long, cudaStream_t);
int transpose_bf16_entry(const void*, long long, void*, long long, long long, long long, cudaStream_t);
int depth_to_space2_bf16_entry(const void*, void*, long long, long long, long long, long long, long long, cudaStream_t);
int latent_to_ndhwc_entry(const void*, int, void*, long long, long long, long long, cudaStream_t);
int vae_frames_out_entry(const float*, long long, float*, long long, cudaStream_t);
int resize_planes_entry(const float*, float*, long long, long long, long long, long long, long long, cudaStream_t);
int depth_conf_entry(const float*, long long, long long, const float*, float, float*, long long, cudaStream_t);
long long quantile_workspace_bytes(long long);
int quantile_entry(const float*, long long, float, float*, void*, long long, cudaStream_t);
long long compact_rows_workspace_bytes(long long);
int compact_rows_entry(const float*, const float*, int, long long, const float*, long long, long long, const float*, float*, float*, float*, long long*, void*,
                       long long, cudaStream_t);

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
std::atomic<int>& pdl_flag() {
  static std::atomic<int> f{[] {
    const char* e = getenv("VIST3A_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }()};
  return f;
}
}  // namespace
bool pdl_enabled() { return pdl_flag().load(std::memory_order_relaxed) != 0; }
int set_pdl(int enable) { return pdl_flag().exchange(enable ? 1 : 0); }

std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}

namespace {
struct DevInfo {
  int sms = 0;
  int major = 0, minor = 0;
  bool valid = false;
};
DevInfo g_dev[64];
std::mutex g_dev_mu;

const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[dev];
  if (!d.valid) {
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev);
    d.valid = true;
  }
  return &d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;
}  // namespace

int num_sms() {
  const DevInfo* d = dev_info();
  return d ? d->sms : 148;
}

int check_arch() {
  const DevInfo* d = dev_info();
  if (!d) return set_error(VIST3A_ERR_CUDA, "no CUDA device available (the kernels of libvist3a_sm100 need an sm_100 GPU)");
  if (d->major != 10) return set_error(VIST3A_ERR_ARCH, "device is sm_%d%d; libvist3a_sm100 is built for sm_100a only", d->major, d->minor);
  return VIST3A_OK;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, bool is_float32, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) return set_error(VIST3A_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapDataType dt = is_float32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                            : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VIST3A_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] stride0 %llu box [%u,%u]",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
  return VIST3A_OK;
}

}  // namespace v3a

using namespace v3a;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* vist3a_last_error(void) { return last_error_buf(); }
int vist3a_abi_version(void) { return 8; }
int64_t vist3a_launch_count(void) { return (int64_t)launch_counter().load(); }
int vist3a_set_pdl(int32_t enable) { return set_pdl(enable); }

int vist3a_gemm(const vist3a_gemm_args* args, void* stream) { return gemm_entry(args, ST(stream)); }
int vist3a_fmha_fwd(const vist3a_fmha_args* args, void* stream) { return fmha_entry(args, ST(stream), nullptr); }
int64_t vist3a_fmha_workspace_bytes(const vist3a_fmha_args* args) {
  long long bytes = 0;
  const int rc = fmha_entry(args, nullptr, &bytes);
  return rc ? (int64_t)rc : (int64_t)bytes;
}
int vist3a_vae_rmsnorm(const void* x, int64_t ldx, const float* gamma, void* y, int64_t ldy, int64_t rows, int64_t C, int32_t silu, void* stream) {
  return vae_rmsnorm_entry(x, ldx, gamma, y, ldy, rows, C, silu, ST(stream));
}
int vist3a_softmax_rows(const float* s, void* p, int64_t rows, int64_t L, int64_t valid, int64_t ldp, float scale, void* stream) {
  return softmax_rows_entry(s, p, rows, L, valid, ldp, scale, ST(stream));
}
int vist3a_time_interleave(const void* y, int64_t ldy, void* out, int64_t ldo, int64_t T, int64_t P, int64_t C, void* stream) {
  return time_interleave_entry(y, ldy, out, ldo, T, P, C, ST(stream));
}
int vist3a_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t R, int64_t C, void* stream) {
  return transpose_bf16_entry(in, ld_in, out, ld_out, R, C, ST(stream));
}
int vist3a_depth_to_space2_bf16(const void* in, void* out, int64_t n_img, int64_t h, int64_t w, int64_t C, int64_t ldo, void* stream) {
  return depth_to_space2_bf16_entry(in, out, n_img, h, w, C, ldo, ST(stream));
}
int vist3a_latent_to_ndhwc(const void* z, int32_t z_dtype, void* out, int64_t C, int64_t THW, int64_t ld, void* stream) {
  return latent_to_ndhwc_entry(z, z_dtype, out, C, THW, ld, ST(stream));
}
int vist3a_vae_frames_out(const float* y, int64_t ld, float* out, int64_t THW, void* stream) { return vae_frames_out_entry(y, ld, out, THW, ST(stream)); }
int vist3a_depth_conf(const float* feat, int64_t ld, int64_t C, const float* w, float bias, float* conf, int64_t n_pixels, void* stream) {
  return depth_conf_entry(feat, ld, C, w, bias, conf, n_pixels, ST(stream));
}
int64_t vist3a_quantile_workspace_bytes(int64_t n) { return quantile_workspace_bytes(n); }
int vist3a_quantile_f32(const float* x, int64_t n, float q, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  return quantile_entry(x, n, q, out, workspace, workspace_bytes, ST(stream));
}
int64_t vist3a_compact_rows_workspace_bytes(int64_t n) { return compact_rows_workspace_bytes(n); }
int vist3a_compact_rows(const float* conf, const float* threshold, int32_t use_threshold, int64_t n, const float* feats, int64_t ld_feats, int64_t C,
                        const float* pts, float* out_feats, float* out_pts, float* out_damp, int64_t* count, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  return compact_rows_entry(conf, threshold, use_threshold, n, feats, ld_feats, C, pts, out_feats, out_pts, out_damp, (long long*)count, workspace,
                            workspace_bytes, ST(stream));
}
int vist3a_resize_planes(const float* in, float* out, int64_t planes, int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out, void* stream) {
  return resize_planes_entry(in, out, planes, h_in, w_in, h_out, w_out, ST(stream));
}

int vist3a_layernorm(const void* x, int32_t x_dtype, int64_t ldx, void* out, int32_t out_dtype, int64_t ldo,
                     int64_t rows, int64_t dim, int64_t rows_per_batch, const float* mul, int64_t mul_bstride,
                     const float* add, int64_t add_bstride, float eps, int32_t mul_plus_one,
                     const vist3a_rowmap* in_map, const vist3a_rowmap* out_map, void* stream) {
  return layernorm_entry(x, x_dtype, ldx, out, out_dtype, ldo, rows, dim, rows_per_batch, mul, mul_bstride, add,
                         add_bstride, eps, mul_plus_one, in_map, out_map, ST(stream));
}

int vist3a_rmsnorm_rope(void* x, int64_t ldx, int64_t rows, int64_t dim, int64_t head_dim, const float* weight,
                        float eps, const float* rope_cos, const float* rope_sin, int64_t rope_len, int64_t nseg,
                        int64_t seg_stride, void* stream) {
  return rmsnorm_rope_entry(x, ldx, rows, dim, head_dim, weight, eps, rope_cos, rope_sin, rope_len, nseg, seg_stride, ST(stream));
}

int vist3a_row_rinv(const void* x, int64_t ldx, int64_t rows, int64_t dim, float eps, float* out, void* stream) {
  return row_rinv_entry(x, ldx, rows, dim, eps, out, ST(stream));
}

int vist3a_modulation(const float* table, const void* mod, int32_t mod_dtype, int32_t mod_is_broadcast, float* out,
                      int64_t batch, int64_t nvec, int64_t dim, uint32_t one_plus_mask, void* stream) {
  return modulation_entry(table, mod, mod_dtype, mod_is_broadcast, out, batch, nvec, dim, one_plus_mask, ST(stream));
}

int vist3a_skinny_linear(const void* x, int32_t x_dtype, int64_t ldx, const void* W, int32_t w_dtype, int64_t ldw,
                         const float* bias, void* y, int32_t y_dtype, int64_t ldy, int64_t M, int64_t N, int64_t K,
                         int32_t pre_act, int32_t act, const float* gate, const float* residual, int64_t ldres,
                         void* stream) {
  return skinny_linear_entry(x, x_dtype, ldx, W, w_dtype, ldw, bias, y, y_dtype, ldy, M, N, K, pre_act, act, gate, residual,
                             ldres, ST(stream));
}

int vist3a_timestep_features(const float* t, void* out, int32_t out_dtype, int64_t batch, int64_t dim, void* stream) {
  return timestep_features_entry(t, out, out_dtype, batch, dim, ST(stream));
}

int vist3a_patchify(const void* x, int32_t x_dtype, void* A, int64_t B, int64_t C, int64_t T, int64_t H, int64_t W,
                    void* stream) {
  return patchify_entry(x, x_dtype, A, B, C, T, H, W, ST(stream));
}

int vist3a_unpatchify(const void* P, int32_t p_dtype, int64_t ldp, void* out, int32_t out_dtype, int64_t B, int64_t C,
                      int64_t T, int64_t H, int64_t W, void* stream) {
  return unpatchify_entry(P, p_dtype, ldp, out, out_dtype, B, C, T, H, W, ST(stream));
}

int vist3a_cfg_combine(const void* cond, const void* uncond, int32_t in_dtype, float guidance, float* out, int64_t n,
                       void* stream) {
  return cfg_combine_entry(cond, uncond, in_dtype, guidance, out, n, ST(stream));
}

int vist3a_axpby_n(float* out, int32_t n_terms, const float* const* terms, const float* coeffs, int64_t n,
                   void* stream) {
  return axpby_entry(out, n_terms, terms, coeffs, n, ST(stream));
}

int vist3a_im2col_stitch(const void* latent, int32_t dtype, void* A, int64_t B, int64_t C, int64_t T, int64_t h,
                         int64_t w, void* stream) {
  return im2col_stitch_entry(latent, dtype, A, B, C, T, h, w, ST(stream));
}
int vist3a_im2col_nhwc(const float* x, float* A, int64_t ldA, int64_t n_img, int64_t h, int64_t w, int64_t C,
                       int32_t kh, int32_t kw, int32_t stride, int32_t pad, void* stream) {
  return im2col_nhwc_entry(x, A, ldA, n_img, h, w, C, kh, kw, stride, pad, ST(stream));
}
int vist3a_qknorm_rope2d(void* qkv, int64_t ld, int64_t rows, int64_t heads, const float* qw, const float* qb,
                         const float* kw, const float* kb, float eps, const float* cos_tab, const float* sin_tab,
                         int64_t max_pos, int64_t tokens_per_view, int64_t n_special, int64_t grid_w, void* stream) {
  return qknorm_rope2d_entry(qkv, ld, rows, heads, qw, qb, kw, kb, eps, cos_tab, sin_tab, max_pos, tokens_per_view,
                             n_special, grid_w, ST(stream));
}
int vist3a_bilinear_nhwc(const float* in, float* out, int64_t n_img, int64_t h_in, int64_t w_in, int64_t h_out,
                         int64_t w_out, int64_t C, const float* add, const float* pos_x, const float* pos_y,
                         void* stream) {
  return bilinear_nhwc_entry(in, out, n_img, h_in, w_in, h_out, w_out, C, add, pos_x, pos_y, ST(stream));
}
int vist3a_depth_to_space(const float* in, float* out, int64_t n_img, int64_t h, int64_t w, int64_t C, int32_t k,
                          void* stream) {
  return depth_to_space_entry(in, out, n_img, h, w, C, k, ST(stream));
}
int vist3a_attention_small(const float* qkv, float* out, int64_t B, int64_t L, int64_t H, int64_t D, float scale,
                           void* stream) {
  return attention_small_entry(qkv, out, B, L, H, D, scale, ST(stream));
}
int vist3a_fma_rows(float* out, int64_t ldo, const float* a, int64_t lda, const float* b, int64_t ldb, const float* c,
                    int64_t ldc, int64_t rows, int64_t dim, void* stream) {
  return fma_rows_entry(out, ldo, a, lda, b, ldb, c, ldc, rows, dim, ST(stream));
}
int vist3a_bias_act_t(const float* ct, int64_t ldct, const float* bias, int32_t act, const float* gate, const float* residual,
                      int64_t ldr, float* y, int64_t ldy, int64_t M, int64_t N, int32_t splits, void* stream) {
  return bias_act_t_entry(ct, ldct, bias, act, gate, residual, ldr, y, ldy, M, N, splits, ST(stream));
}
int vist3a_rgb_to_nhwc4pad(const void* image, int32_t dtype, float* out, int64_t B, int64_t V, int64_t H, int64_t W, void* stream) {
  return rgb_to_nhwc4pad_entry(image, dtype, out, B, V, H, W, 0, 0.5f, 0.5f, ST(stream));
}
int vist3a_rgb01_views_to_nhwc4pad(const void* image, int32_t dtype, float* out, int64_t B, int64_t V, int64_t H, int64_t W, void* stream) {
  return rgb_to_nhwc4pad_entry(image, dtype, out, B, V, H, W, 1, 1.0f, 0.0f, ST(stream));
}
int vist3a_patch_embed_im2col(const void* image, int32_t dtype, void* A, int64_t ldA, int64_t n_img, int64_t H, int64_t W, int32_t patch,
                              const float* mean3, const float* std3, void* stream) {
  return patch_embed_im2col_entry(image, dtype, A, ldA, n_img, H, W, patch, mean3, std3, ST(stream));
}
int vist3a_pose_to_cameras(const float* pose_raw, float* pose_act, float* extr, float* intr, float* c2w,
                           float* intr_norm, int64_t S, int64_t H, int64_t W, void* stream) {
  return pose_to_cameras_entry(pose_raw, pose_act, extr, intr, c2w, intr_norm, S, H, W, ST(stream));
}
int vist3a_gaussian_epilogue(const float* depth_feat, int64_t ld_df, int64_t cd, const float* depth_w, float depth_b,
                             const float* gs_raw, int64_t ld_raw, const float* extr, const float* intr,
                             const float* sh_mask, int64_t d_sh, int64_t S, int64_t H, int64_t W, float* depth,
                             float* means, float* scales, float* rotations, float* opacities, float* harmonics,
                             float* covariances, float* scene_sum, void* stream) {
  return gaussian_epilogue_entry(depth_feat, ld_df, cd, depth_w, depth_b, gs_raw, ld_raw, extr, intr, sh_mask, d_sh, S, H, W,
                                 depth, means, scales, rotations, opacities, harmonics, covariances, scene_sum, ST(stream));
}

int vist3a_gaussian_adapter(const float* pts, const float* feats, int64_t ld_feats, const float* sh_mask, int64_t d_sh, int64_t P,
                            float* means, float* scales, float* rotations, float* opacities, float* harmonics, float* covariances,
                            void* stream) {
  return gaussian_adapter_entry(pts, feats, ld_feats, sh_mask, d_sh, P, means, scales, rotations, opacities, harmonics, covariances, ST(stream));
}
int64_t vist3a_voxel_fusion_workspace_bytes(int64_t n_points) { return voxel_fusion_workspace_bytes(n_points); }
int vist3a_voxel_fusion(const float* pts, const float* feats, int64_t ld_feats, int64_t feat_dim, const float* conf, int64_t conf_stride,
                        int64_t n_points, float voxel_size, float* voxel_pts, float* voxel_feats, int32_t* inverse, int32_t* counts,
                        int64_t* n_voxels, void* workspace, int64_t workspace_bytes, void* stream) {
  return voxel_fusion_entry(pts, feats, ld_feats, feat_dim, conf, conf_stride, n_points, voxel_size, voxel_pts, voxel_feats, inverse, counts,
                            reinterpret_cast<long long*>(n_voxels), workspace, workspace_bytes, ST(stream));
}

int64_t vist3a_gs_project_workspace_bytes(int64_t n_gaussians) { return gs_project_workspace_bytes(n_gaussians); }
int vist3a_gs_project(const float* means, const float* covariances, const float* opacities, const float* harmonics, int64_t d_sh, int32_t sh_degree,
                      int64_t n_gaussians, const float* viewmat, const float* K, int64_t W, int64_t H, float near_plane, float far_plane,
                      float radius_clip, float eps2d, void* workspace, int64_t workspace_bytes, int64_t* n_isect, void* stream) {
  return gs_project_entry(means, covariances, opacities, harmonics, d_sh, sh_degree, n_gaussians, viewmat, K, W, H, near_plane, far_plane, radius_clip,
                          eps2d, workspace, workspace_bytes, reinterpret_cast<long long*>(n_isect), ST(stream));
}
int64_t vist3a_gs_rasterize_workspace_bytes(int64_t n_isect, int64_t W, int64_t H) { return gs_rasterize_workspace_bytes(n_isect, W, H); }
int vist3a_gs_rasterize(const void* project_workspace, int64_t n_gaussians, int64_t n_isect, int64_t W, int64_t H, const float* background,
                        void* workspace, int64_t workspace_bytes, float* rgb, float* depth, float* alpha, void* stream) {
  return gs_rasterize_entry(project_workspace, n_gaussians, n_isect, W, H, background, workspace, workspace_bytes, rgb, depth, alpha, ST(stream));
}

}  // extern "C"
