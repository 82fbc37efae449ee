// common.cuh -- shared declarations for the nerfb200 CUDA library (sm_100a).
//
// The "plan" is the host-built description of one FlexibleNeRFModel (nerf/models.py:185-256 of the
// reference) as a chain of GEMM layers (outputs of 64*NJ columns, evaluated tile-by-tile) and up
// to two narrow "heads" (fc_alpha: 1 output, fc_rgb: 3, fc_out: 4) that are evaluated as dot
// products.  It is passed by value to every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nerfb200.h"

namespace nerfb200 {

constexpr int kMaxGemm = 20;
constexpr int kTileRows = 128;   // points per CTA tile
constexpr int kThreads = 256;    // threads per CTA in the MLP kernels

struct GemmLayer {
  int k_h;       // input columns taken from the previous activation (0 for layer1, else hidden)
  int k_enc;     // padded (multiple of 8) encoded-input columns appended after k_h (0 if none)
  int enc_real;  // real encoded columns (63 / 27 / ...)
  int enc_sel;   // 0: xyz encoding, 1: direction encoding
  int n;         // output features (hidden or hidden/2), multiple of 64
  int relu;      // ReLU after the bias (layer1 has none: models.py:238)
  int wt_off;    // blob offset of Wt[k_h + k_enc][n]   (forward operand)
  int wh_off;    // blob offset of Wh[n][k_h]           (dgrad operand; h-part only)
  int b_off;     // blob offset of bias[n]
  int cum_n;     // sum of n over previous gemm layers: stash slice = base + n_points * cum_n
  int flat_w;    // offset of weight[n][k_h + enc_real] in the flat (torch-layout) vector
  int flat_b;    // offset of bias[n] in the flat vector
  int src;       // gemm index whose output is this layer's h-input (-1 for layer1)
  // tcgen05 operand copy (mlp_tc.cu): k_tc = K the tensor-core path contracts over (the direction
  // encoding of layers_dir[0] is hoisted out per ray, so k_tc = k_h there); per k-step of 16 the blob
  // holds [hi slab0 | hi slab1 | lo slab0 | lo slab1], slab = [n][8 fp16] (UMMA canonical K-major,
  // no swizzle: 8-row core matrices 128 B apart, the two K halves one slab apart); fp16 x 2 split with a
  // 2^11-scaled residual (tc_common.cuh split_f16x2).
  int k_tc;
  int tc_off;
  // dgrad operand copy: Wt[k < k_h][n] as K-major slabs over the reduction index n, hi|lo per 16 n
  int tcd_off;
  int mask_cum;  // ReLU bit-mask words (uint32, one bit per output) per point before this layer's
};

struct HeadLayer {
  int k;        // input features
  int n_out;    // 1, 3 or 4
  int out_col;  // first column of raw[...,4] it produces
  int w_off;    // blob offset of W[n_out][k] (torch layout), 4-float aligned
  int b_off;    // blob offset of bias[4]
  int src;      // gemm index whose output it reads
  int flat_w, flat_b;
};

struct Plan {
  int n_gemm;
  int n_trunk;  // gemm layers 0..n_trunk-1 are layer1 + layers_xyz
  int n_head;
  int hidden;
  int sum_n;  // floats stashed per point: every gemm layer's output + the two padded encodings + the mask words
  int enc_cum[2];  // stash slice of the xyz / direction encoding: base + n_points * enc_cum[sel], width enc_tile_w / dim_dir_pad
  int mask_base;   // stash slice of the ReLU bit masks: base + n_points * (mask_base + g.mask_cum), n/32 words per point
  int dim_xyz, dim_xyz_pad, dim_dir, dim_dir_pad;
  int enc_tile_w;  // dim_xyz padded to 16: width of the xyz-encoding operand tile of the tcgen05 path (tc_common.cuh)
  int n_freq_xyz, n_freq_dir, inc_xyz, inc_dir;
  int blob_floats, flat_floats;
  int use_viewdirs;
  GemmLayer g[kMaxGemm];
  HeadLayer h[2];  // viewdirs: h[0] = fc_alpha (after trunk), h[1] = fc_rgb (after dir layer); else h[0] = fc_out
  float freq_xyz[NERFB200_MAX_FREQS];
  float freq_dir[NERFB200_MAX_FREQS];
};

// host side (api.cu)
int build_plan(const nerfb200_arch_t* arch, Plan* plan);  // 0 or NERFB200_ERR_*
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch();  // one call per kernel launched (nerfb200_launch_count)

// launchers (one per .cu), all return NERFB200_OK or an error code
int launch_pack(const Plan& p, const float* flat, float* blob, cudaStream_t s);
int launch_sample_coarse(const float* rays, int ray_stride, int64_t n_rays, const float* t_vals,
                         const float* t_rand, int n_coarse, int perturb, int lindisp, float* z, cudaStream_t s);
int launch_gen_rays(const float* c2w12_host, int height, int width, float focal, const int64_t* pix, int64_t n, int ndc,
                    float near, float far, int use_viewdirs, int out_stride, float* out, cudaStream_t s);
int launch_pack_rays(const float* ro, const float* rd, int64_t n, int height, int width, float focal, int ndc, float near,
                     float far, int use_viewdirs, int out_stride, float* out, cudaStream_t s);
int launch_encode(const Plan& p, int which, const float* x, int64_t n, float* out, cudaStream_t s);
int launch_mlp_fwd_simt(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z,
                        int64_t n_rays, int n_samples, float* raw, float* stash, cudaStream_t s);
int launch_mlp_bwd(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z,
                   int64_t n_rays, int n_samples, const float* d_raw, const float* stash, float* gstash,
                   float* flat_grad, int impl, cudaStream_t s);
// the two halves of launch_mlp_bwd (stage-level entry points: tests and per-kernel timing)
int launch_mlp_dgrad(const Plan& p, const float* blob, const float* d_raw, const float* stash, float* gstash, int64_t P,
                     int impl, cudaStream_t s);
int launch_mlp_wgrad(const Plan& p, const float* rays, int ray_stride, const float* z, int64_t n_rays, int n_samples,
                     const float* d_raw, const float* stash, const float* gstash, float* flat_grad, int impl,
                     cudaStream_t s);
// tcgen05 backward (mlp_tc_bwd.cu): data-gradient chain + every weight gradient in one kernel; `scratch` holds the
// L2-resident gradient blob (bwd_tc_scratch_floats) the weight-gradient tiles are reduced into
int launch_mlp_bwd_tc(const Plan& p, const float* blob, const float* rays, int ray_stride, int64_t n_rays, int n_samples,
                      const float* d_raw, const float* stash, float* scratch, float* flat_grad, cudaStream_t s);
int64_t bwd_tc_scratch_floats(const Plan& p);
int tc_supported(const Plan& p, int n_samples, const char* what, bool training);  // forward (training: + stash)
int bwd_tc_supported(const Plan& p, int n_samples, const char* what);  // forward + backward
int launch_composite_fwd(const float* raw, const float* z, const float* rays, int ray_stride, const float* noise,
                         int64_t n_rays, int n_samples, float noise_std, int white_bkgd, float* out,
                         float* weights, cudaStream_t s);
int launch_composite_bwd(const float* raw, const float* z, const float* rays, int ray_stride, const float* noise,
                         const float* g_out, int64_t n_rays, int n_samples, float noise_std, int white_bkgd,
                         float* d_raw, cudaStream_t s);
int launch_sample_pdf_merge(const float* z_coarse, const float* weights_coarse, const float* u, int u_stride,
                            const float* cdf_in, int64_t n_rays, int n_coarse, int n_fine, float* z_fine,
                            float* z_samples, int32_t* inds, float* cdf_out, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, int step, float lr, float b1, float b2,
                float eps, float grad_scale, cudaStream_t s);
// tcgen05 path (mlp_tc.cu, mlp_tc_bwd.cu)
int launch_mlp_fwd_tc(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z,
                      int64_t n_rays, int n_samples, float* raw, float* stash, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Stash row swizzle.  The activation / gradient stashes ([points][n] fp32, workspace private to this library) keep
// every 16-byte chunk of a row at chunk index (q ^ (point & 7)) inside its 128-byte segment.  Global access
// patterns are unchanged (a permutation inside each 128 B line), and a 128-row tile of the stash is byte-identical
// to a bank-conflict-free shared-memory tile, so the tensor-core kernels write it with ONE cp.async.bulk per layer.
__device__ __forceinline__ int swz_col(int col, int64_t point) { return (((col >> 2) ^ (int)(point & 7)) << 2) | (col & 3); }

// Positional encoding of one scalar coordinate (nerf_helpers.py:113-157): writes x (if include) and
// sin/cos(x * f_i) at the reference's channel positions for coordinate c of 3:
//   [x y z | sin(f0 xyz) | cos(f0 xyz) | sin(f1 xyz) | ...]
// `row` points at the start of this point's encoding.  x * f is rounded to fp32 first, as the
// reference evaluates `tensor * freq` before sin/cos (nerf_helpers.py:149-151).
__device__ __forceinline__ void encode_coord(float x, int c, int include, int f_begin, int f_end,
                                             const float* __restrict__ freqs, float* __restrict__ row) {
  const int base = include ? 3 : 0;
  if (include && f_begin == 0) row[c] = x;
  for (int f = f_begin; f < f_end; ++f) {
    float s, co;
    sincosf(__fmul_rn(x, freqs[f]), &s, &co);
    row[base + 6 * f + c] = s;
    row[base + 6 * f + 3 + c] = co;
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

#endif  // __CUDACC__

}  // namespace nerfb200
