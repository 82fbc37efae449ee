// api.cu -- the extern "C" surface declared in include/nerfb200.h: argument checking, plan
// building, workspace carving and kernel orchestration for the whole per-ray path
// (predict_and_render_radiance, nerf/train_utils.py:28-127 of the reference).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>

#include "common.cuh"
#include "wgrad_items.cuh"

namespace nerfb200 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NERFB200_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return NERFB200_ERR_CUDA;
}

static inline int pad8(int x) { return (x + 7) & ~7; }
static inline int pad4(int x) { return (x + 3) & ~3; }

int build_plan(const nerfb200_arch_t* a, Plan* p) {
  if (!a || !p) {
    set_error("build_plan: null argument");
    return NERFB200_ERR_INVALID;
  }
  memset(p, 0, sizeof(*p));
  if (a->hidden != 128 && a->hidden != 256) {
    set_error("unsupported hidden_size %d (kernels implement 128 and 256)", a->hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (a->num_layers < 1 || a->num_layers > 16) {
    set_error("unsupported num_layers %d (1..16)", a->num_layers);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (a->skip_every <= 0) {
    set_error("skip_connect_every must be > 0 (got %d)", a->skip_every);
    return NERFB200_ERR_INVALID;
  }
  if (a->n_freq_xyz < 0 || a->n_freq_xyz > NERFB200_MAX_FREQS || a->n_freq_dir < 0 ||
      a->n_freq_dir > NERFB200_MAX_FREQS) {
    set_error("number of encoding functions out of range (max %d)", NERFB200_MAX_FREQS);
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int H = a->hidden;
  p->hidden = H;
  p->use_viewdirs = a->use_viewdirs ? 1 : 0;
  p->n_freq_xyz = a->n_freq_xyz;
  p->n_freq_dir = a->n_freq_dir;
  p->inc_xyz = a->include_input_xyz ? 1 : 0;
  p->inc_dir = a->include_input_dir ? 1 : 0;
  p->dim_xyz = 3 * p->inc_xyz + 6 * a->n_freq_xyz;
  p->dim_dir = p->use_viewdirs ? 3 * p->inc_dir + 6 * a->n_freq_dir : 0;
  if (p->dim_xyz <= 0 || (p->use_viewdirs && p->dim_dir <= 0)) {
    set_error("empty encoding (dim_xyz %d, dim_dir %d)", p->dim_xyz, p->dim_dir);
    return NERFB200_ERR_UNSUPPORTED;
  }
  p->dim_xyz_pad = pad8(p->dim_xyz);
  p->dim_dir_pad = pad8(p->dim_dir);
  p->enc_tile_w = (p->dim_xyz + 15) & ~15;
  for (int i = 0; i < NERFB200_MAX_FREQS; ++i) {
    p->freq_xyz[i] = a->freq_xyz[i];
    p->freq_dir[i] = a->freq_dir[i];
  }

  int blob = 0, flat = 0, cum = 0, ng = 0;
  auto add_gemm = [&](int k_h, int enc_sel, int enc_real, int n, int relu, int src) {
    GemmLayer& g = p->g[ng];
    g.k_h = k_h;
    g.enc_sel = enc_sel;
    g.enc_real = enc_real;
    g.k_enc = pad8(enc_real);
    g.n = n;
    g.relu = relu;
    g.src = src;
    g.wt_off = blob;
    blob += (g.k_h + g.k_enc) * n;
    g.wh_off = blob;
    blob += n * g.k_h;
    g.b_off = blob;
    blob += n;
    g.k_tc = enc_sel == 1 ? k_h : k_h + g.k_enc;
    g.tc_off = blob;
    blob += 3 * ((g.k_tc + 15) & ~15) * n / 2;  // three fp16 copies (hs | h | l) of [K padded to 16][n]
    g.tcd_off = blob;
    blob += 3 * g.k_h * n / 2;                  // three fp16 copies of [n][k_h] for the dgrad chain
    g.cum_n = cum;
    cum += n;
    g.flat_w = flat;
    flat += n * (k_h + enc_real);
    g.flat_b = flat;
    flat += n;
    return ng++;
  };
  auto set_head = [&](int idx, int k, int n_out, int out_col, int src) {
    HeadLayer& h = p->h[idx];
    h.k = k;
    h.n_out = n_out;
    h.out_col = out_col;
    h.src = src;
    h.w_off = blob;
    blob += pad4(n_out * k);
    h.b_off = blob;
    blob += 4;
    h.flat_w = flat;
    flat += n_out * k;
    h.flat_b = flat;
    flat += n_out;
  };

  // trunk: layer1 (no activation, models.py:238) + layers_xyz (ReLU; wide iff allocated wide, :210-215)
  add_gemm(0, 0, p->dim_xyz, H, 0, -1);
  for (int i = 0; i < a->num_layers - 1; ++i) {
    const bool wide = (i % a->skip_every == 0) && i > 0 && i != a->num_layers - 1;
    add_gemm(H, 0, wide ? p->dim_xyz : 0, H, 1, ng - 1);
  }
  p->n_trunk = ng;
  if (p->use_viewdirs) {
    // flat order follows the canonical slot order: fc_feat, fc_alpha, layers_dir[0], fc_rgb
    const int feat = add_gemm(H, 0, 0, H, 1, p->n_trunk - 1);      // models.py:248
    set_head(0, H, 1, 3, p->n_trunk - 1);                          // fc_alpha, models.py:249
    const int dir = add_gemm(H, 1, p->dim_dir, H / 2, 1, feat);    // models.py:250-252
    set_head(1, H / 2, 3, 0, dir);                                 // fc_rgb, models.py:253
    p->n_head = 2;
  } else {
    set_head(0, H, 4, 0, p->n_trunk - 1);  // fc_out, models.py:256
    p->n_head = 1;
  }
  p->n_gemm = ng;
  p->enc_cum[0] = cum;
  cum += p->enc_tile_w;
  p->enc_cum[1] = cum;
  cum += p->dim_dir_pad;
  p->mask_base = cum;
  int mw = 0;
  for (int i = 0; i < ng; ++i) {
    p->g[i].mask_cum = mw;
    mw += p->g[i].n / 32;
  }
  cum += mw;
  p->sum_n = cum;
  p->blob_floats = blob;
  p->flat_floats = flat;
  return NERFB200_OK;
}

// slot -> (is_head, index)
static bool slot_lookup(const Plan& p, int slot, bool* is_head, int* idx) {
  const int nt = p.n_trunk;
  if (slot < 0) return false;
  if (slot < nt) {
    *is_head = false;
    *idx = slot;
    return true;
  }
  if (p.use_viewdirs) {
    switch (slot - nt) {
      case 0: *is_head = false; *idx = nt; return true;
      case 1: *is_head = true; *idx = 0; return true;
      case 2: *is_head = false; *idx = nt + 1; return true;
      case 3: *is_head = true; *idx = 1; return true;
      default: return false;
    }
  }
  if (slot == nt) {
    *is_head = true;
    *idx = 0;
    return true;
  }
  return false;
}

static inline int64_t align256(int64_t bytes) { return (bytes + 255) & ~int64_t(255); }

struct Workspace {
  int64_t z_c, raw_c, w_c, z_f, raw_f, stash_c, stash_f, gstash, d_raw, total;
};

static void carve(const Plan& pc, const Plan* pf, const nerfb200_render_opts_t& o, int64_t n, int training,
                  Workspace* w) {
  const int64_t nc = o.n_coarse, ns = o.n_coarse + o.n_fine;
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    int64_t r = off;
    off += align256(floats * 4);
    return r;
  };
  w->z_c = take(n * nc);
  w->raw_c = take(n * nc * 4);
  w->w_c = take(n * nc);
  const bool fine = o.n_fine > 0 && pf != nullptr;
  w->z_f = take(fine ? n * ns : 0);
  w->raw_f = take(fine ? n * ns * 4 : 0);
  if (training) {
    auto pad128 = [](int64_t v) { return (v + 127) & ~int64_t(127); };
    const int64_t sc = pad128(n * nc) * pc.sum_n;
    const int64_t sf = fine ? pad128(n * ns) * pf->sum_n : 0;
    w->stash_c = take(sc);
    w->stash_f = take(sf);
    // backward scratch: the gradient stash of the CUDA-core path / the gradient blob of the tcgen05 path
    int64_t gs = sc > sf ? sc : sf;
    const int64_t bc = pc.hidden == 128 ? bwd_tc_scratch_floats(pc) : 0;
    const int64_t bf = (fine && pf->hidden == 128) ? bwd_tc_scratch_floats(*pf) : 0;
    if (bc > gs) gs = bc;
    if (bf > gs) gs = bf;
    w->gstash = take(gs);
    w->d_raw = take(n * (fine ? ns : nc) * 4);
  } else {
    w->stash_c = w->stash_f = w->gstash = w->d_raw = off;
  }
  w->total = off;
}

static int check_opts(const nerfb200_render_opts_t* o) {
  if (!o) {
    set_error("null render options");
    return NERFB200_ERR_INVALID;
  }
  if (o->n_coarse < 2 || o->n_coarse > 1024 || o->n_fine < 0 || o->n_coarse + o->n_fine > 1024) {
    set_error("unsupported sample counts (n_coarse %d, n_fine %d; need 2 <= nc, nc + nf <= 1024)", o->n_coarse,
              o->n_fine);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (o->n_fine > 0 && o->n_coarse < 4) {
    set_error("hierarchical sampling needs n_coarse >= 4");
    return NERFB200_ERR_UNSUPPORTED;
  }
  return NERFB200_OK;
}

}  // namespace nerfb200

using namespace nerfb200;

#define NB_TRY(expr)                    \
  do {                                  \
    int _rc = (expr);                   \
    if (_rc != NERFB200_OK) return _rc; \
  } while (0)

extern "C" {

int32_t nerfb200_version(void) { return NERFB200_VERSION; }
const char* nerfb200_last_error(void) { return g_err; }
int64_t nerfb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int64_t nerfb200_bwd_bytes_per_point(const nerfb200_arch_t* arch) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK) return -1;
  int64_t bytes = 16;  // d_raw
  for (int t = 0; t < p.n_gemm; ++t) {
    const GemmLayer& g = p.g[t];
    if (g.k_h > 0) bytes += 4 * g.k_h;                               // the layer input tile (hi + lo)
    if (g.k_enc > 0 && g.enc_sel == 0) bytes += 4 * p.enc_tile_w;    // the encoding tile
    if (g.relu) bytes += g.n / 8;                                    // ReLU bit mask
  }
  for (int h = 0; h < p.n_head; ++h) bytes += 4 * p.h[h].k;          // the heads' input tiles
  return bytes;
}

int32_t nerfb200_impl_supported(const nerfb200_arch_t* arch, int32_t n_samples, int32_t impl) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  if (impl == 0) return NERFB200_OK;
  if (impl == 1) return bwd_tc_supported(p, n_samples, "impl_supported");
  if (impl == 2) return tc_supported(p, n_samples, "impl_supported", /*training=*/false);  // tcgen05 forward only (inference)
  set_error("unknown impl %d", impl);
  return NERFB200_ERR_INVALID;
}

int64_t nerfb200_bwd_scratch_floats(const nerfb200_arch_t* arch, int64_t n_points, int32_t impl) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK || n_points < 0) return -1;
  if (impl == 1) return bwd_tc_scratch_floats(p);
  return ((n_points + 127) & ~int64_t(127)) * p.sum_n;
}

int64_t nerfb200_num_linear(const nerfb200_arch_t* arch) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK) return -1;
  return p.n_gemm + p.n_head;
}
int64_t nerfb200_flat_param_count(const nerfb200_arch_t* arch) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK) return -1;
  return p.flat_floats;
}
int64_t nerfb200_blob_floats(const nerfb200_arch_t* arch) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK) return -1;
  return p.blob_floats;
}
int64_t nerfb200_stash_floats(const nerfb200_arch_t* arch, int64_t n_points) {
  Plan p;
  if (build_plan(arch, &p) != NERFB200_OK || n_points < 0) return -1;
  return ((n_points + 127) & ~int64_t(127)) * p.sum_n;  // whole 128-point tiles (the tcgen05 path stores tiles)
}

int32_t nerfb200_flat_layout(const nerfb200_arch_t* arch, int32_t slot, int64_t* w_off, int64_t* b_off,
                             int32_t* in_features, int32_t* out_features) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  bool is_head;
  int idx;
  if (!slot_lookup(p, slot, &is_head, &idx)) {
    set_error("flat_layout: slot %d out of range", slot);
    return NERFB200_ERR_INVALID;
  }
  if (is_head) {
    if (w_off) *w_off = p.h[idx].flat_w;
    if (b_off) *b_off = p.h[idx].flat_b;
    if (in_features) *in_features = p.h[idx].k;
    if (out_features) *out_features = p.h[idx].n_out;
  } else {
    if (w_off) *w_off = p.g[idx].flat_w;
    if (b_off) *b_off = p.g[idx].flat_b;
    if (in_features) *in_features = p.g[idx].k_h + p.g[idx].enc_real;
    if (out_features) *out_features = p.g[idx].n;
  }
  return NERFB200_OK;
}

int32_t nerfb200_pack_weights(const nerfb200_arch_t* arch, const float* flat, float* blob, void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  if (!flat || !blob) {
    set_error("pack_weights: null pointer");
    return NERFB200_ERR_INVALID;
  }
  return launch_pack(p, flat, blob, static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_sample_coarse(const float* rays, int32_t ray_stride, int64_t n_rays, const float* t_vals,
                               const float* t_rand, int32_t n_coarse, int32_t perturb, int32_t lindisp,
                               float* z, void* stream) {
  if (!rays || !t_vals || !z || n_rays <= 0 || n_coarse < 2 || ray_stride < 8 || (perturb && !t_rand)) {
    set_error("sample_coarse: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_sample_coarse(rays, ray_stride, n_rays, t_vals, perturb ? t_rand : nullptr, n_coarse, perturb,
                              lindisp, z, static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_gen_rays(const float* c2w12_host, int32_t height, int32_t width, float focal, const int64_t* pixel_ids,
                          int64_t n, int32_t ndc, float near, float far, int32_t use_viewdirs, int32_t out_stride,
                          float* out, void* stream) {
  if (!c2w12_host || !out || n <= 0) {
    set_error("gen_rays: null pointer or empty batch");
    return NERFB200_ERR_INVALID;
  }
  return launch_gen_rays(c2w12_host, height, width, focal, pixel_ids, n, ndc, near, far, use_viewdirs, out_stride, out,
                         static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_pack_rays(const float* ro, const float* rd, int64_t n, int32_t height, int32_t width, float focal,
                           int32_t ndc, float near, float far, int32_t use_viewdirs, int32_t out_stride, float* out,
                           void* stream) {
  if (!ro || !rd || !out || n <= 0) {
    set_error("pack_rays: null pointer or empty batch");
    return NERFB200_ERR_INVALID;
  }
  return launch_pack_rays(ro, rd, n, height, width, focal, ndc, near, far, use_viewdirs, out_stride, out,
                          static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_encode(const nerfb200_arch_t* arch, int32_t which, const float* x, int64_t n, float* out,
                        void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  if (!x || !out || n <= 0 || which < 0 || which > 1 || (which == 1 && !p.use_viewdirs)) {
    set_error("encode: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_encode(p, which, x, n, out, static_cast<cudaStream_t>(stream));
}

static int check_mlp_args(const Plan& p, const void* blob, const float* rays, int ray_stride, const float* z,
                          int64_t n_rays, int n_samples) {
  if (!blob || !rays || !z || n_rays <= 0 || n_samples <= 0) {
    set_error("mlp: null pointer or empty input");
    return NERFB200_ERR_INVALID;
  }
  if (ray_stride < (p.use_viewdirs ? 11 : 8)) {
    set_error("mlp: ray_stride %d too small (need %d columns)", ray_stride, p.use_viewdirs ? 11 : 8);
    return NERFB200_ERR_INVALID;
  }
  if (n_rays * (int64_t)n_samples >= (int64_t(1) << 31)) {
    set_error("mlp: more than 2^31 points in one call; chunk the rays");
    return NERFB200_ERR_UNSUPPORTED;
  }
  return NERFB200_OK;
}

int32_t nerfb200_mlp_fwd(const nerfb200_arch_t* arch, const float* blob, const float* rays, int32_t ray_stride,
                         const float* z, int64_t n_rays, int32_t n_samples, float* raw, float* stash,
                         int32_t impl, void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  NB_TRY(check_mlp_args(p, blob, rays, ray_stride, z, n_rays, n_samples));
  if (!raw) {
    set_error("mlp_fwd: null output");
    return NERFB200_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0) return launch_mlp_fwd_simt(p, blob, rays, ray_stride, z, n_rays, n_samples, raw, stash, s);
  if (impl == 1) return launch_mlp_fwd_tc(p, blob, rays, ray_stride, z, n_rays, n_samples, raw, stash, s);
  set_error("mlp_fwd: unknown impl %d", impl);
  return NERFB200_ERR_INVALID;
}

int32_t nerfb200_mlp_bwd(const nerfb200_arch_t* arch, const float* blob, const float* rays, int32_t ray_stride,
                         const float* z, int64_t n_rays, int32_t n_samples, const float* d_raw,
                         const float* stash, float* gstash, float* flat_grad, int32_t impl, void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  NB_TRY(check_mlp_args(p, blob, rays, ray_stride, z, n_rays, n_samples));
  if (!d_raw || !stash || !gstash || !flat_grad) {
    set_error("mlp_bwd: null pointer");
    return NERFB200_ERR_INVALID;
  }
  return launch_mlp_bwd(p, blob, rays, ray_stride, z, n_rays, n_samples, d_raw, stash, gstash, flat_grad, impl,
                        static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_mlp_dgrad(const nerfb200_arch_t* arch, const float* blob, const float* d_raw, const float* stash,
                           float* gstash, int64_t n_points, int32_t impl, void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  if (!blob || !d_raw || !stash || !gstash || n_points <= 0) {
    set_error("mlp_dgrad: null pointer or empty input");
    return NERFB200_ERR_INVALID;
  }
  return launch_mlp_dgrad(p, blob, d_raw, stash, gstash, n_points, impl, static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_mlp_wgrad(const nerfb200_arch_t* arch, const float* rays, int32_t ray_stride, const float* z,
                           int64_t n_rays, int32_t n_samples, const float* d_raw, const float* stash,
                           const float* gstash, float* flat_grad, int32_t impl, void* stream) {
  Plan p;
  NB_TRY(build_plan(arch, &p));
  if (!rays || !z || !d_raw || !stash || !gstash || !flat_grad || n_rays <= 0 || n_samples <= 0) {
    set_error("mlp_wgrad: null pointer or empty input");
    return NERFB200_ERR_INVALID;
  }
  return launch_mlp_wgrad(p, rays, ray_stride, z, n_rays, n_samples, d_raw, stash, gstash, flat_grad, impl,
                          static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_composite_fwd(const float* raw, const float* z, const float* rays, int32_t ray_stride,
                               const float* noise, int64_t n_rays, int32_t n_samples, float noise_std,
                               int32_t white_bkgd, float* out, float* weights, void* stream) {
  if (!raw || !z || !rays || !out || n_rays <= 0 || n_samples <= 0 || n_samples > 1024 || ray_stride < 8) {
    set_error("composite_fwd: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_composite_fwd(raw, z, rays, ray_stride, noise, n_rays, n_samples, noise_std, white_bkgd, out,
                              weights, static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_composite_bwd(const float* raw, const float* z, const float* rays, int32_t ray_stride,
                               const float* noise, const float* g_out, int64_t n_rays, int32_t n_samples,
                               float noise_std, int32_t white_bkgd, float* d_raw, void* stream) {
  if (!raw || !z || !rays || !g_out || !d_raw || n_rays <= 0 || n_samples <= 0 || n_samples > 1024 ||
      ray_stride < 8) {
    set_error("composite_bwd: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_composite_bwd(raw, z, rays, ray_stride, noise, g_out, n_rays, n_samples, noise_std, white_bkgd,
                              d_raw, static_cast<cudaStream_t>(stream));
}

int32_t nerfb200_sample_pdf_merge(const float* z_coarse, const float* weights_coarse, const float* u,
                                  int32_t u_stride, const float* cdf_in, int64_t n_rays, int32_t n_coarse,
                                  int32_t n_fine, float* z_fine, float* z_samples, int32_t* inds,
                                  float* cdf_out, void* stream) {
  if (!z_coarse || (!weights_coarse && !cdf_in) || !u || n_rays <= 0 || n_coarse < 4 || n_fine <= 0 ||
      n_coarse + n_fine > 1024 || (u_stride != 0 && u_stride != n_fine)) {
    set_error("sample_pdf_merge: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_sample_pdf_merge(z_coarse, weights_coarse, u, u_stride, cdf_in, n_rays, n_coarse, n_fine, z_fine,
                                 z_samples, inds, cdf_out, static_cast<cudaStream_t>(stream));
}

int64_t nerfb200_render_workspace_bytes(const nerfb200_arch_t* coarse, const nerfb200_arch_t* fine,
                                        const nerfb200_render_opts_t* opts, int64_t n_rays, int32_t training) {
  Plan pc, pf;
  if (build_plan(coarse, &pc) != NERFB200_OK) return -1;
  if (fine && build_plan(fine, &pf) != NERFB200_OK) return -1;
  if (check_opts(opts) != NERFB200_OK || n_rays <= 0) return -1;
  Workspace w;
  carve(pc, fine ? &pf : nullptr, *opts, n_rays, training, &w);
  return w.total;
}

/* Byte offsets of the named sections of the render workspace (test hook): order
 * z_coarse, raw_coarse, weights_coarse, z_fine, raw_fine, stash_coarse, stash_fine, gstash, d_raw. */
int32_t nerfb200_render_workspace_layout(const nerfb200_arch_t* coarse, const nerfb200_arch_t* fine,
                                         const nerfb200_render_opts_t* opts, int64_t n_rays, int32_t training,
                                         int64_t* offsets9) {
  Plan pc, pf;
  NB_TRY(build_plan(coarse, &pc));
  if (fine) NB_TRY(build_plan(fine, &pf));
  NB_TRY(check_opts(opts));
  if (n_rays <= 0 || !offsets9) {
    set_error("workspace_layout: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  Workspace w;
  carve(pc, fine ? &pf : nullptr, *opts, n_rays, training, &w);
  const int64_t v[9] = {w.z_c, w.raw_c, w.w_c, w.z_f, w.raw_f, w.stash_c, w.stash_f, w.gstash, w.d_raw};
  for (int i = 0; i < 9; ++i) offsets9[i] = v[i];
  return NERFB200_OK;
}

int32_t nerfb200_render_fwd(const nerfb200_arch_t* arch_c, const nerfb200_arch_t* arch_f,
                            const nerfb200_render_opts_t* opts, const float* blob_c, const float* blob_f,
                            const float* rays, int32_t ray_stride, int64_t n_rays, const float* t_vals,
                            const float* t_rand, const float* noise_c, const float* u, int32_t u_stride,
                            const float* noise_f, float* out_coarse, float* out_fine, void* workspace,
                            int32_t training, int32_t impl, void* stream) {
  Plan pc, pf;
  NB_TRY(build_plan(arch_c, &pc));
  NB_TRY(check_opts(opts));
  const bool fine = opts->n_fine > 0;
  if (fine) {
    if (!arch_f || !blob_f || !out_fine || !u) {
      set_error("render_fwd: fine pass requested but fine model / output / u missing");
      return NERFB200_ERR_INVALID;
    }
    NB_TRY(build_plan(arch_f, &pf));
  }
  if (!rays || !t_vals || !out_coarse || !workspace || !blob_c || n_rays <= 0) {
    set_error("render_fwd: null pointer or empty batch");
    return NERFB200_ERR_INVALID;
  }
  if (opts->perturb && !t_rand) {
    set_error("render_fwd: perturb set but t_rand missing");
    return NERFB200_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace w;
  carve(pc, fine ? &pf : nullptr, *opts, n_rays, training, &w);
  char* base = static_cast<char*>(workspace);
  auto F = [&](int64_t off) { return reinterpret_cast<float*>(base + off); };
  const int nc = opts->n_coarse, ns = opts->n_coarse + opts->n_fine;

  // coarse pass: stratified depths -> MLP -> compositing (train_utils.py:45-94)
  NB_TRY(launch_sample_coarse(rays, ray_stride, n_rays, t_vals, opts->perturb ? t_rand : nullptr, nc,
                              opts->perturb, opts->lindisp, F(w.z_c), s));
  float* stash_c = training ? F(w.stash_c) : nullptr;
  NB_TRY(check_mlp_args(pc, blob_c, rays, ray_stride, F(w.z_c), n_rays, nc));
  if (impl == 1)
    NB_TRY(launch_mlp_fwd_tc(pc, blob_c, rays, ray_stride, F(w.z_c), n_rays, nc, F(w.raw_c), stash_c, s));
  else
    NB_TRY(launch_mlp_fwd_simt(pc, blob_c, rays, ray_stride, F(w.z_c), n_rays, nc, F(w.raw_c), stash_c, s));
  NB_TRY(launch_composite_fwd(F(w.raw_c), F(w.z_c), rays, ray_stride, opts->noise_std > 0.f ? noise_c : nullptr,
                              n_rays, nc, opts->noise_std, opts->white_bkgd, out_coarse, F(w.w_c), s));
  if (!fine) return NERFB200_OK;

  // hierarchical resampling + merge (train_utils.py:96-105), fine pass (:107-125)
  NB_TRY(launch_sample_pdf_merge(F(w.z_c), F(w.w_c), u, u_stride, nullptr, n_rays, nc, opts->n_fine, F(w.z_f),
                                 nullptr, nullptr, nullptr, s));
  float* stash_f = training ? F(w.stash_f) : nullptr;
  NB_TRY(check_mlp_args(pf, blob_f, rays, ray_stride, F(w.z_f), n_rays, ns));
  if (impl == 1)
    NB_TRY(launch_mlp_fwd_tc(pf, blob_f, rays, ray_stride, F(w.z_f), n_rays, ns, F(w.raw_f), stash_f, s));
  else
    NB_TRY(launch_mlp_fwd_simt(pf, blob_f, rays, ray_stride, F(w.z_f), n_rays, ns, F(w.raw_f), stash_f, s));
  NB_TRY(launch_composite_fwd(F(w.raw_f), F(w.z_f), rays, ray_stride, opts->noise_std > 0.f ? noise_f : nullptr,
                              n_rays, ns, opts->noise_std, opts->white_bkgd, out_fine, nullptr, s));
  return NERFB200_OK;
}

int32_t nerfb200_render_bwd(const nerfb200_arch_t* arch_c, const nerfb200_arch_t* arch_f,
                            const nerfb200_render_opts_t* opts, const float* blob_c, const float* blob_f,
                            const float* rays, int32_t ray_stride, int64_t n_rays, const float* noise_c,
                            const float* noise_f, const float* g_coarse, const float* g_fine, void* workspace,
                            float* flat_grad_c, float* flat_grad_f, int32_t impl, int32_t parts, void* stream) {
  Plan pc, pf;
  NB_TRY(build_plan(arch_c, &pc));
  NB_TRY(check_opts(opts));
  const bool fine = opts->n_fine > 0;
  if (fine) {
    if (!arch_f || !blob_f || !g_fine || !flat_grad_f) {
      set_error("render_bwd: fine pass present but fine model / grads missing");
      return NERFB200_ERR_INVALID;
    }
    NB_TRY(build_plan(arch_f, &pf));
  }
  if (!rays || !g_coarse || !workspace || !blob_c || !flat_grad_c || n_rays <= 0) {
    set_error("render_bwd: null pointer or empty batch");
    return NERFB200_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace w;
  carve(pc, fine ? &pf : nullptr, *opts, n_rays, /*training=*/1, &w);
  char* base = static_cast<char*>(workspace);
  auto F = [&](int64_t off) { return reinterpret_cast<float*>(base + off); };
  const int nc = opts->n_coarse, ns = opts->n_coarse + opts->n_fine;

  if ((parts & 3) == 0) {
    set_error("render_bwd: parts must select the fine (1) and / or the coarse (2) backward");
    return NERFB200_ERR_INVALID;
  }
  if (fine && (parts & 1)) {
    NB_TRY(launch_composite_bwd(F(w.raw_f), F(w.z_f), rays, ray_stride, opts->noise_std > 0.f ? noise_f : nullptr,
                                g_fine, n_rays, ns, opts->noise_std, opts->white_bkgd, F(w.d_raw), s));
    NB_TRY(launch_mlp_bwd(pf, blob_f, rays, ray_stride, F(w.z_f), n_rays, ns, F(w.d_raw), F(w.stash_f), F(w.gstash),
                          flat_grad_f, impl, s));
  }
  if (!(parts & 2)) return NERFB200_OK;
  // the coarse weights feed the resampler only through a detach (train_utils.py:103), so the
  // coarse net's gradient comes from rgb/disp/acc_coarse alone.
  NB_TRY(launch_composite_bwd(F(w.raw_c), F(w.z_c), rays, ray_stride, opts->noise_std > 0.f ? noise_c : nullptr,
                              g_coarse, n_rays, nc, opts->noise_std, opts->white_bkgd, F(w.d_raw), s));
  NB_TRY(launch_mlp_bwd(pc, blob_c, rays, ray_stride, F(w.z_c), n_rays, nc, F(w.d_raw), F(w.stash_c), F(w.gstash),
                        flat_grad_c, impl, s));
  return NERFB200_OK;
}

int32_t nerfb200_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                           float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) {
    set_error("adam_step: invalid argument");
    return NERFB200_ERR_INVALID;
  }
  return launch_adam(p, g, m, v, n, step, lr, beta1, beta2, eps, grad_scale, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
