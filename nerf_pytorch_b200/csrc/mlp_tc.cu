// mlp_tc.cu -- tcgen05 (5th-generation tensor core) implementation of the fused per-point forward
//     points -> positional encoding -> FlexibleNeRFModel (nerf/train_utils.py:67, :8-25; nerf/models.py:233-256)
// for hidden_size 128, fp32-faithful through a 3xTF32 split:
//     x * w  ~=  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo        (x_hi = tf32(x), x_lo = x - x_hi, same for w)
// accumulated in fp32 in tensor memory (single-pass TF32/BF16 misses the
// 1e-4 bar on the shipped checkpoints; the 3-term split meets it).
//
// Persistent kernel, one CTA per SM, 192 threads:
//   warps 0-3  prologue/epilogue: thread r owns row r of the 128-point tile (= TMEM lane r).  Prologue:
//              point generation + sin/cos encoding, split into hi/lo, written to shared memory in the UMMA
//              canonical K-major layout.  Epilogue of every layer: tcgen05.ld the fp32 accumulator, + bias,
//              ReLU, (stash), narrow heads (fc_alpha / fc_rgb / fc_out) as register dot products, split
//              into hi/lo and tcgen05.st back into tensor memory as the NEXT layer's A operand.
//   warp 4     MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=128|64, K=8 per
//              instruction; three instructions per k-step), A from tensor memory (hidden activations) or
//              shared memory (encodings), B = pre-split weights from the shared-memory ring.
//   warp 5     weight producer: cp.async.bulk of one k-step of (hi, lo) weights per ring stage from the
//              L2-resident blob, mbarrier complete_tx.
// Tensor memory (512 columns): [0,128) accumulator, [128,256) A_hi, [256,384) A_lo.
// The direction encoding enters layers_dir[0] through a per-ray bias computed on the CUDA cores in fp32
// (it is constant along a ray: SURVEY.md section 7.3 item 5), so that layer contracts over K = 128 only.
#include "common.cuh"
#include "tc_common.cuh"

namespace nerfb200 {

static long long* g_tc_prof = nullptr;  // debug hook: per-CTA cycle counters (nerfb200_debug_tc_profile)
void set_tc_profile(void* p) { g_tc_prof = static_cast<long long*>(p); }

namespace tc {

constexpr int kEpiThreads = 128;
constexpr int kThreadsTc = 192;
constexpr int kStages = 12;          // weight ring depth
constexpr int kStageBytes = 8192;    // one k-step of hi+lo weights for N = 128
constexpr int kSlabBytes = 2048;     // 128 rows x 16 B
constexpr int kMaxRaysPerTile = 10;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColAcc = 0, kColAhi = 128, kColAlo = 256;

struct Smem {
  // byte offsets from the 1024-aligned base
  static constexpr int e_hi = 0;                         // 16 slabs (K <= 64) x 2 KB
  static constexpr int e_lo = e_hi + 16 * kSlabBytes;
  static constexpr int ring = e_lo + 16 * kSlabBytes;    // kStages x 8 KB
  static constexpr int bias = ring + kStages * kStageBytes;   // kMaxGemm x 128 floats
  static constexpr int headw = bias + kMaxGemm * 128 * 4;     // 4*128 + 3*64 floats (+pad) and 8 bias floats
  static constexpr int viewb = headw + (4 * 128 + 3 * 64 + 16) * 4;  // kMaxRaysPerTile x 64
  static constexpr int encd = viewb + kMaxRaysPerTile * 64 * 4;      // kMaxRaysPerTile x 32
  static constexpr int bars = encd + kMaxRaysPerTile * 32 * 4;       // mbarriers
  static constexpr int total = bars + 256;
};

}  // namespace tc

using namespace tc;

__global__ void __launch_bounds__(kThreadsTc, 1)
mlp_fwd_tc_kernel(const __grid_constant__ Plan p, const float* __restrict__ blob, const float* __restrict__ rays,
                  int ray_stride, const float* __restrict__ z, int64_t P, int S, int64_t n_tiles,
                  float* __restrict__ raw, float* __restrict__ stash, long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1 KB aligned base
  float* s_bias = reinterpret_cast<float*>(sm + Smem::bias);
  float* s_headw = reinterpret_cast<float*>(sm + Smem::headw);
  float* s_viewb = reinterpret_cast<float*>(sm + Smem::viewb);
  float* s_encd = reinterpret_cast<float*>(sm + Smem::encd);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Smem::bars);
  uint64_t* bar_full = bars;                 // [kStages]  weights landed
  uint64_t* bar_empty = bars + kStages;      // [kStages]  stage consumed by the MMAs
  uint64_t* bar_a = bars + 2 * kStages;      // A operand of the next layer is ready (128 arrivals)
  uint64_t* bar_acc = bars + 2 * kStages + 1;  // accumulator of the current layer is complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_a, kEpiThreads);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // biases + head weights: once per CTA
  for (int gi = 0; gi < p.n_gemm; ++gi)
    for (int i = tid; i < p.g[gi].n; i += kThreadsTc) s_bias[gi * 128 + i] = blob[p.g[gi].b_off + i];
  const int hw1 = p.h[0].n_out * p.h[0].k;
  const int hw2 = p.n_head > 1 ? p.h[1].n_out * p.h[1].k : 0;
  for (int i = tid; i < hw1; i += kThreadsTc) s_headw[i] = blob[p.h[0].w_off + i];
  for (int i = tid; i < hw2; i += kThreadsTc) s_headw[hw1 + i] = blob[p.h[1].w_off + i];
  float* s_headb = s_headw + ((hw1 + hw2 + 3) & ~3);
  if (tid < 4) s_headb[tid] = blob[p.h[0].b_off + tid];
  if (tid >= 4 && tid < 8) s_headb[tid] = p.n_head > 1 ? blob[p.h[1].b_off + tid - 4] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 5) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int gi = 0; gi < p.n_gemm; ++gi) {
          const GemmLayer& g = p.g[gi];
          const uint32_t bytes = 64u * g.n;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(blob + g.tc_off);
          const int ksteps = g.k_tc >> 3;
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bar_full[stage], bytes);
            bulk_g2s(sm + Smem::ring + stage * kStageBytes, src + (size_t)ks * bytes, bytes, &bar_full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t e_hi = smem_u32(sm + Smem::e_hi), e_lo = smem_u32(sm + Smem::e_lo);
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int gi = 0; gi < p.n_gemm; ++gi) {
          const GemmLayer& g = p.g[gi];
          const uint32_t idesc = make_idesc(g.n);
          const uint32_t slab_b = 16u * g.n;  // bytes of one weight slab
          const int ksteps = g.k_tc >> 3, ksteps_h = g.k_h >> 3;
          mbar_wait(bar_a, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            const uint32_t wb = smem_u32(sm + Smem::ring + stage * kStageBytes);
            const uint64_t b_hi = make_desc(wb, slab_b, 128);
            const uint64_t b_lo = make_desc(wb + 2 * slab_b, slab_b, 128);
            const uint32_t acc0 = ks > 0 ? 1u : 0u;
            if (ks < ksteps_h) {
              const uint32_t a_hi = tmem + kColAhi + 8 * ks, a_lo = tmem + kColAlo + 8 * ks;
              mma_ts(tmem + kColAcc, a_hi, b_hi, idesc, acc0);
              mma_ts(tmem + kColAcc, a_lo, b_hi, idesc, 1u);
              mma_ts(tmem + kColAcc, a_hi, b_lo, idesc, 1u);
            } else {
              const uint32_t off = (uint32_t)(ks - ksteps_h) * 2 * kSlabBytes;
              const uint64_t a_hi = make_desc(e_hi + off, kSlabBytes, 128);
              const uint64_t a_lo = make_desc(e_lo + off, kSlabBytes, 128);
              mma_ss(tmem + kColAcc, a_hi, b_hi, idesc, acc0);
              mma_ss(tmem + kColAcc, a_lo, b_hi, idesc, 1u);
              mma_ss(tmem + kColAcc, a_hi, b_lo, idesc, 1u);
            }
            mma_commit(&bar_empty[stage]);  // frees the ring stage once these MMAs have read it
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          mma_commit(bar_acc);  // accumulator of layer gi complete
        }
      }
    }
  } else {
    // ===================== prologue / epilogue warps (thread = row) =====================
    const int row = tid;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    uint32_t acc_phase = 0;
    uint8_t* e_hi = sm + Smem::e_hi;
    uint8_t* e_lo = sm + Smem::e_lo;
    long long t_pro = 0, t_wait = 0, t_epi = 0, t_all = clock64();
    for (int64_t it = 0; it < my_tiles; ++it) {
      long long t0 = clock64();
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t p0 = tile * kTileRows;
      int64_t pt = p0 + row;
      const bool valid = pt < P;
      if (!valid) pt = P - 1;
      const int64_t ray = pt / S;
      const int64_t first_ray = p0 / S;
      const int64_t last_pt = (p0 + kTileRows - 1 < P) ? p0 + kTileRows - 1 : P - 1;
      const int n_rays_tile = (int)(last_pt / S - first_ray) + 1;
      const int ray_slot = (int)(ray - first_ray);

      // ---- prologue: encodings of this row -> E_hi / E_lo (canonical K-major slabs) ----
      {
        const float* rr = rays + ray * ray_stride;
        const float zz = z[pt];
        float* sx = (stash && valid) ? stash + (size_t)P * p.enc_cum[0] + (size_t)pt * p.dim_xyz_pad : nullptr;
        for (int c = 0; c < 3; ++c) {
          const float x = __fadd_rn(rr[c], __fmul_rn(rr[3 + c], zz));  // pts = ro + rd * z (train_utils.py:67)
          auto put = [&](int k, float v) {
            if (sx) sx[k] = v;
            const uint32_t hi = tf32_hi(v);
            const float lo = v - __uint_as_float(hi);
            const int off = (k >> 2) * kSlabBytes + row * 16 + (k & 3) * 4;
            *reinterpret_cast<uint32_t*>(e_hi + off) = hi;
            *reinterpret_cast<float*>(e_lo + off) = lo;
          };
          const int base = p.inc_xyz ? 3 : 0;
          if (p.inc_xyz) put(c, x);
          for (int f = 0; f < p.n_freq_xyz; ++f) {
            float sn, cs;
            sincosf(__fmul_rn(x, p.freq_xyz[f]), &sn, &cs);
            put(base + 6 * f + c, sn);
            put(base + 6 * f + 3 + c, cs);
          }
        }
        for (int k = p.dim_xyz; k < p.dim_xyz_pad; ++k) {
          if (sx) sx[k] = 0.f;
          const int off = (k >> 2) * kSlabBytes + row * 16 + (k & 3) * 4;
          *reinterpret_cast<uint32_t*>(e_hi + off) = 0u;
          *reinterpret_cast<uint32_t*>(e_lo + off) = 0u;
        }
      }
      // ---- per-ray direction term of layers_dir[0]: vb[ray][n] = sum_k enc_dir(ray)[k] * W[n][H + k] ----
      if (p.use_viewdirs) {
        if (row < n_rays_tile * 3) {
          const int j = row / 3, c = row - 3 * j;
          const float v = rays[(first_ray + j) * ray_stride + 8 + c];
          encode_coord(v, c, p.inc_dir, 0, p.n_freq_dir, p.freq_dir, s_encd + j * 32);
        }
        epi_bar();
        const GemmLayer& gd = p.g[p.n_gemm - 1];
        const float* wv = blob + gd.wt_off + (size_t)gd.k_h * gd.n;  // rows k_h.. of Wt[k][n]
        for (int i = row; i < n_rays_tile * gd.n; i += kEpiThreads) {
          const int j = i / gd.n, n = i - j * gd.n;
          float a = 0.f;
          for (int k = 0; k < p.dim_dir; ++k) a = fmaf(s_encd[j * 32 + k], wv[k * gd.n + n], a);
          s_viewb[j * 64 + n] = a;
        }
      }
      fence_proxy_async();  // E_hi / E_lo were written through the generic proxy; the MMAs read them via the async proxy
      epi_bar();            // also publishes s_viewb
      if (stash && valid && p.use_viewdirs) {
        float* sd = stash + (size_t)P * p.enc_cum[1] + (size_t)pt * p.dim_dir_pad;
        for (int k = 0; k < p.dim_dir_pad; ++k) sd[k] = k < p.dim_dir ? s_encd[ray_slot * 32 + k] : 0.f;
      }
      mbar_arrive(bar_a);
      t_pro += clock64() - t0;

      // ---- layers ----
      float hacc[4];
      for (int gi = 0; gi < p.n_gemm; ++gi) {
        const GemmLayer& g = p.g[gi];
        const bool has_next = gi + 1 < p.n_gemm;
        int hsel = -1;
        if (p.h[0].src == gi) hsel = 0;
        if (p.n_head > 1 && p.h[1].src == gi) hsel = 1;
        const float* hw = hsel == 1 ? s_headw + hw1 : s_headw;
        const int hk = hsel >= 0 ? p.h[hsel].k : 0, hn = hsel >= 0 ? p.h[hsel].n_out : 0;
        if (hsel >= 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) hacc[c] = s_headb[hsel * 4 + c];
        }
        const bool is_dir = p.use_viewdirs && gi == p.n_gemm - 1;
        float* st = (stash && valid) ? stash + (size_t)P * g.cum_n + (size_t)pt * g.n : nullptr;

        t0 = clock64();
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const long long t1 = clock64();
        t_wait += t1 - t0;
        for (int c0 = 0; c0 < g.n; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + kColAcc + c0, v);
          tmem_wait_ld();
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + gi * 128 + c0 + j);
            float x[4] = {__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                          __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w};
            if (is_dir) {
              const float4 vb = *reinterpret_cast<const float4*>(s_viewb + ray_slot * 64 + c0 + j);
              x[0] += vb.x; x[1] += vb.y; x[2] += vb.z; x[3] += vb.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float y = g.relu ? fmaxf(x[q], 0.f) : x[q];
              if (hsel >= 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (c < hn) hacc[c] = fmaf(y, hw[c * hk + c0 + j + q], hacc[c]);
              }
              x[q] = y;
              hi[j + q] = tf32_hi(y);
              lo[j + q] = __float_as_uint(y - __uint_as_float(hi[j + q]));
            }
            if (st) *reinterpret_cast<float4*>(st + c0 + j) = make_float4(x[0], x[1], x[2], x[3]);
          }
          if (has_next) {
            tmem_st32(tmem + lane_base + kColAhi + c0, hi);
            tmem_st32(tmem + lane_base + kColAlo + c0, lo);
          }
        }
        if (hsel >= 0 && valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < hn) raw[pt * 4 + p.h[hsel].out_col + c] = hacc[c];
        }
        if (has_next) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(bar_a);
        } else {
          // the accumulator has been drained (wait::ld above); the next tile's first MMA is additionally
          // ordered behind this thread by the bar_a arrival after the next prologue.
          tc_fence_before();
        }
        t_epi += clock64() - t1;
      }
    }
    if (prof && tid == 0) {
      prof[blockIdx.x * 4 + 0] = t_pro;
      prof[blockIdx.x * 4 + 1] = t_wait;
      prof[blockIdx.x * 4 + 2] = t_epi;
      prof[blockIdx.x * 4 + 3] = clock64() - t_all;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

int launch_mlp_fwd_tc(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z,
                      int64_t n_rays, int n_samples, float* raw, float* stash, cudaStream_t s) {
  if (p.hidden != 128) {
    set_error("mlp_fwd impl=1 (tcgen05): hidden_size %d not supported (128 only); use impl=0", p.hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (p.dim_xyz_pad > 64 || p.dim_dir > 32) {
    set_error("mlp_fwd impl=1 (tcgen05): encodings wider than 64 (xyz) / 32 (dir) not supported; use impl=0");
    return NERFB200_ERR_UNSUPPORTED;
  }
  if ((kTileRows + n_samples - 1) / n_samples + 1 > kMaxRaysPerTile) {
    set_error("mlp_fwd impl=1 (tcgen05): fewer than 16 samples per ray not supported; use impl=0");
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int64_t P = n_rays * n_samples;
  const int64_t tiles = (P + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < sms ? tiles : sms);
  const size_t bytes = Smem::total + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(mlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                      "mlp_fwd_tc smem attribute");
  if (rc) return rc;
  mlp_fwd_tc_kernel<<<grid, kThreadsTc, bytes, s>>>(p, blob, rays, ray_stride, z, P, n_samples, tiles, raw, stash, g_tc_prof);
  count_launch();
  return check_cuda(cudaGetLastError(), "mlp_fwd_tc launch");
}



// =============================================================================================
// backward, tensor-core dgrad chain: for one 128-point tile walk the layers in reverse,
//     G_t = ( G_s * W_s[:, :hidden]  +  d_raw * W_head ) (.) [stash_t > 0]          (s = consumer of t)
// with G_s resident in tensor memory as the A operand (hi/lo), the transposed weights streamed through
// the same bulk-copy ring as the forward, and every G_t written to `gstash` for the wgrad kernel.
// =============================================================================================
namespace tcd {
struct SmemD {
  static constexpr int ring = 0;                                     // kStages x 8 KB
  static constexpr int headw = ring + tc::kStages * tc::kStageBytes;  // head weights (4*128 + 3*64 floats)
  static constexpr int bars = headw + (4 * 128 + 3 * 64 + 16) * 4;
  static constexpr int total = bars + 256;
};
}  // namespace tcd

__global__ void __launch_bounds__(kThreadsTc, 1)
mlp_dgrad_tc_kernel(const __grid_constant__ Plan p, const float* __restrict__ blob, const float* __restrict__ d_raw,
                    const float* __restrict__ stash, float* __restrict__ gstash, int64_t P, int64_t n_tiles) {
  using tcd::SmemD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_headw = reinterpret_cast<float*>(sm + SmemD::headw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SmemD::bars);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kStages;
  uint64_t* bar_a = bars + 2 * kStages;
  uint64_t* bar_acc = bars + 2 * kStages + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_a, kEpiThreads);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const int hw1 = p.h[0].n_out * p.h[0].k;
  const int hw2 = p.n_head > 1 ? p.h[1].n_out * p.h[1].k : 0;
  for (int i = tid; i < hw1; i += kThreadsTc) s_headw[i] = blob[p.h[0].w_off + i];
  for (int i = tid; i < hw2; i += kThreadsTc) s_headw[hw1 + i] = blob[p.h[1].w_off + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // consumer gemm of layer t (the layer whose h-input is t's output), or -1
  auto consumer = [&](int t) {
    int s = -1;
    for (int c = t + 1; c < p.n_gemm; ++c)
      if (p.g[c].src == t) s = c;
    return s;
  };

  if (warp == 5) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int t = p.n_gemm - 1; t >= 0; --t) {
          const int s = consumer(t);
          if (s < 0) continue;
          const GemmLayer& g = p.g[s];
          const uint32_t bytes = 64u * g.k_h;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(blob + g.tcd_off);
          const int ksteps = g.n >> 3;
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bar_full[stage], bytes);
            bulk_g2s(sm + SmemD::ring + stage * kStageBytes, src + (size_t)ks * bytes, bytes, &bar_full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int t = p.n_gemm - 1; t >= 0; --t) {
          const int s = consumer(t);
          if (s < 0) continue;
          const GemmLayer& g = p.g[s];
          const uint32_t idesc = make_idesc(g.k_h);
          const uint32_t slab_b = 16u * g.k_h;
          const int ksteps = g.n >> 3;
          mbar_wait(bar_a, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            const uint32_t wb = smem_u32(sm + SmemD::ring + stage * kStageBytes);
            const uint64_t b_hi = make_desc(wb, slab_b, 128);
            const uint64_t b_lo = make_desc(wb + 2 * slab_b, slab_b, 128);
            const uint32_t a_hi = tmem + kColAhi + 8 * ks, a_lo = tmem + kColAlo + 8 * ks;
            mma_ts(tmem + kColAcc, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
            mma_ts(tmem + kColAcc, a_lo, b_hi, idesc, 1u);
            mma_ts(tmem + kColAcc, a_hi, b_lo, idesc, 1u);
            mma_commit(&bar_empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          mma_commit(bar_acc);
        }
      }
    }
  } else {
    const int row = tid;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    uint32_t acc_phase = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t pt = tile * kTileRows + row;
      const bool valid = pt < P;
      const float4 dr4 = valid ? reinterpret_cast<const float4*>(d_raw)[pt] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float dr[4] = {dr4.x, dr4.y, dr4.z, dr4.w};
      for (int t = p.n_gemm - 1; t >= 0; --t) {
        const GemmLayer& gt = p.g[t];
        const int s = consumer(t);
        int hsel = -1;
        for (int c = 0; c < p.n_head; ++c)
          if (p.h[c].src == t) hsel = c;
        const float* hw = hsel == 1 ? s_headw + hw1 : s_headw;
        const int hk = hsel >= 0 ? p.h[hsel].k : 0, hn = hsel >= 0 ? p.h[hsel].n_out : 0;
        const int hcol = hsel >= 0 ? p.h[hsel].out_col : 0;
        const float* st = stash + (size_t)P * gt.cum_n + (size_t)(valid ? pt : 0) * gt.n;
        float* gs = gstash + (size_t)P * gt.cum_n + (size_t)(valid ? pt : 0) * gt.n;
        if (s >= 0) {
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
        }
        for (int c0 = 0; c0 < gt.n; c0 += 32) {
          uint32_t v[32];
          if (s >= 0) {
            tmem_ld32(tmem + lane_base + kColAcc + c0, v);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float x[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3])};
            if (hsel >= 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < hn) {
                  const float4 w = *reinterpret_cast<const float4*>(hw + c * hk + c0 + j);
                  const float d = dr[(hcol + c) & 3];
                  x[0] = fmaf(d, w.x, x[0]); x[1] = fmaf(d, w.y, x[1]); x[2] = fmaf(d, w.z, x[2]); x[3] = fmaf(d, w.w, x[3]);
                }
            }
            if (gt.relu) {
              const float4 a = valid ? __ldg(reinterpret_cast<const float4*>(st + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              x[0] = a.x > 0.f ? x[0] : 0.f; x[1] = a.y > 0.f ? x[1] : 0.f;
              x[2] = a.z > 0.f ? x[2] : 0.f; x[3] = a.w > 0.f ? x[3] : 0.f;
            }
            if (valid) *reinterpret_cast<float4*>(gs + c0 + j) = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              hi[j + q] = tf32_hi(x[q]);
              lo[j + q] = __float_as_uint(x[q] - __uint_as_float(hi[j + q]));
            }
          }
          if (t > 0) {
            tmem_st32(tmem + lane_base + kColAhi + c0, hi);
            tmem_st32(tmem + lane_base + kColAlo + c0, lo);
          }
        }
        if (t > 0) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(bar_a);
        } else {
          tc_fence_before();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

int launch_dgrad_tc(const Plan& p, const float* blob, const float* d_raw, const float* stash, float* gstash, int64_t P,
                    cudaStream_t s) {
  if (p.hidden != 128) {
    set_error("dgrad impl=1 (tcgen05): hidden_size %d not supported (128 only)", p.hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int64_t tiles = (P + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < sms ? tiles : sms);
  const size_t bytes = tcd::SmemD::total + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(mlp_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                      "dgrad_tc smem attribute");
  if (rc) return rc;
  mlp_dgrad_tc_kernel<<<grid, kThreadsTc, bytes, s>>>(p, blob, d_raw, stash, gstash, P, tiles);
  count_launch();
  return check_cuda(cudaGetLastError(), "dgrad_tc launch");
}

}  // namespace nerfb200
