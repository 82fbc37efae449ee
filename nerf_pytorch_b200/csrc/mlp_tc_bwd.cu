// mlp_tc_bwd.cu -- tcgen05 weight-gradient kernel:  dW[n][k] = sum_p dY[p][n] * X[p][k]   (3xTF32)
//
// One CTA = one (layer, row-block, column-block) item of wgrad_items.cuh over a contiguous range of
// points.  The reduction runs over POINTS while the stash rows dY[p][:] / X[p][:] are feature-contiguous,
// so the eight producer warps transpose on the way in: each thread loads the same 4 features of 4
// consecutive points (coalesced 512 B per warp and point), transposes the 4x4 block in registers, splits
// every value into tf32 hi + lo and stores 64 contiguous bytes per operand into the UMMA canonical
// K-major no-swizzle layout (slab = [feature][4 points]); the positional-encoding columns come from the
// encoding the forward pass stashed.  32 points per stage; one elected thread issues three tcgen05.mma.kind::tf32 per
// 8 points (hi*hi, lo*hi, hi*lo) into a 128 x N fp32 accumulator in tensor memory that lives for the
// whole range; at the end four warps drain it with tcgen05.ld and atomically add into the flat gradient.
// HBM-bound by design: per point and layer it reads (n + k) * 4 bytes once.
#include "common.cuh"
#include "tc_common.cuh"
#include "wgrad_items.cuh"

namespace nerfb200 {

namespace tcw {
constexpr int kProducers = 256;              // warps 0-7
constexpr int kThreadsW = 320;               // + warp 8 (MMA issue, TMEM alloc) + warp 9 (idle)
constexpr int kStagePts = 32;                // points per pipeline stage = 4 MMA k-groups
constexpr int kStagesW = 3;
constexpr int kSboW = 144;                     // 8-feature core matrices 128 B + 16 B pad apart: conflict-free 64 B/lane stores
constexpr int kSlabW = 16 * kSboW;             // 128 features x 4 points (one K-major slab), padded
constexpr int kOpBytes = (kStagePts / 4) * kSlabW;  // one operand tile (hi or lo): 8 slabs
constexpr int kStageBytesW = 4 * kOpBytes;     // A_hi | A_lo | B_hi | B_lo
}  // namespace tcw

using namespace tc;
using namespace tcw;

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, int off, float4 v) {
  uint4 h;
  h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
  float4 l = make_float4(v.x - __uint_as_float(h.x), v.y - __uint_as_float(h.y), v.z - __uint_as_float(h.z),
                         v.w - __uint_as_float(h.w));
  *reinterpret_cast<uint4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

__global__ void __launch_bounds__(kThreadsW, 1)
mlp_wgrad_tc_kernel(const __grid_constant__ Plan p, const float* __restrict__ rays, int ray_stride,
                    const float* __restrict__ z, int S, const float* __restrict__ stash,
                    const float* __restrict__ gstash, const float* __restrict__ d_raw, int64_t P,
                    float* __restrict__ flat_grad, int n_items, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kStagesW * kStageBytesW);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + kStagesW;
  uint64_t* bar_done = bars + 2 * kStagesW;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStagesW + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((int)blockIdx.y >= n_items) return;
  const WgItem it = wg_decode(p, blockIdx.y);
  // operand sources.  gemm items: A rows = dY_t (gstash), B rows = the producing layer's output or the stashed
  // encoding.  Head items (fc_alpha / fc_rgb / fc_out): A rows = d_raw[p][0..3] masked to the head's columns
  // (4 live rows of the 128-row tile), B rows = the output of the layer the head reads.
  const bool head = it.kind == 2;
  const GemmLayer& g = p.g[head ? p.h[it.t].src : it.t];
  const int hcol0 = head ? p.h[it.t].out_col : 0, hcols = head ? p.h[it.t].n_out : 0;

  // contiguous point range of this CTA, in units of one stage
  const int64_t stages_total = (P + kStagePts - 1) / kStagePts;
  const int64_t per = (stages_total + gridDim.x - 1) / gridDim.x;
  int64_t pt_begin = (int64_t)blockIdx.x * per * kStagePts;
  int64_t pt_end = pt_begin + per * kStagePts;
  if (pt_begin > P) pt_begin = P;
  if (pt_end > P) pt_end = P;
  if (pt_begin >= pt_end) return;
  const int64_t n_stage = (pt_end - pt_begin + kStagePts - 1) / kStagePts;

  if (tid == 0) {
    for (int i = 0; i < kStagesW; ++i) {
      mbar_init(&bar_full[i], kProducers);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // zero every operand tile once: rows / columns outside the item stay zero for the whole kernel
  for (int i = tid; i < kStagesW * kStageBytesW / 16; i += kThreadsW)
    reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int n_mma = it.kblk;  // accumulator columns (multiple of 16)

  if (warp < 8) {
    // ===================== producers =====================
    // warps 0-3 build the A tiles (dY), warps 4-7 the B tiles (X); lane = 4-feature chunk, warp%4 = point quad
    const int c = lane;               // features 4c .. 4c+3
    const bool is_a = warp < 4;
    const int quad0 = warp & 3;       // this thread transposes quads quad0 and quad0 + 4 of every stage
    // X rows: the producing layer's stashed output (kind 0) or the stashed, zero-padded encoding (kind 1)
    const int xw = head ? g.n : (it.kind == 0 ? p.g[g.src].n : (g.enc_sel ? p.dim_dir_pad : p.dim_xyz_pad));
    const bool on = is_a ? (head ? c == 0 : 4 * c < it.nblk) : (4 * c < it.kblk && 4 * c < xw);
    // stash / gstash rows are chunk-swizzled (swz_col); d_raw and the stashed encodings are plain
    const bool swz = is_a ? !head : (head || it.kind == 0);
    const float* src =
        is_a ? (head ? d_raw : gstash + (size_t)P * g.cum_n + it.n0)
             : (head ? stash + (size_t)P * g.cum_n
                     : (it.kind == 0 ? stash + (size_t)P * p.g[g.src].cum_n + it.k0
                                     : stash + (size_t)P * p.enc_cum[g.enc_sel]));
    const int ld = is_a ? (head ? 4 : g.n) : xw;
    // head items: keep only this head's columns of d_raw
    const float hm0 = (!head || (hcol0 <= 0 && 0 < hcol0 + hcols)) ? 1.f : 0.f;
    const float hm1 = (!head || (hcol0 <= 1 && 1 < hcol0 + hcols)) ? 1.f : 0.f;
    const float hm2 = (!head || (hcol0 <= 2 && 2 < hcol0 + hcols)) ? 1.f : 0.f;
    const float hm3 = (!head || (hcol0 <= 3 && 3 < hcol0 + hcols)) ? 1.f : 0.f;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t stage = 0, phase = 0;

    float4 cur[8];
    auto issue_loads = [&](int64_t q0, float4 (&v)[8]) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t pt = q0 + 4 * (quad0 + 4 * h) + i;
          const int cc = swz ? (c ^ (int)(pt & 7)) : c;
          // (no arithmetic on the loaded value here: it must stay in flight until the stage is built)
          v[4 * h + i] = (on && pt < pt_end) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)pt * ld + 4 * cc)) : zero4;
        }
    };
    issue_loads(pt_begin, cur);
    for (int64_t s = 0; s < n_stage; ++s) {
      const int64_t q0 = pt_begin + s * kStagePts;
      float4 nxt[8];
      issue_loads(q0 + kStagePts, nxt);  // keep the next stage's loads in flight (out-of-range points load zeros)
      mbar_wait(&bar_empty[stage], phase ^ 1);
      uint8_t* st = sm + stage * kStageBytesW;
      if (on) {
        uint8_t* hi_b = st + (is_a ? 0 : 2 * kOpBytes);
        uint8_t* lo_b = hi_b + kOpBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 r0 = cur[4 * h + 0], r1 = cur[4 * h + 1], r2 = cur[4 * h + 2], r3 = cur[4 * h + 3];
          if (head && is_a) {  // keep only this head's columns of d_raw
            r0.x *= hm0; r1.x *= hm0; r2.x *= hm0; r3.x *= hm0;
            r0.y *= hm1; r1.y *= hm1; r2.y *= hm1; r3.y *= hm1;
            r0.z *= hm2; r1.z *= hm2; r2.z *= hm2; r3.z *= hm2;
            r0.w *= hm3; r1.w *= hm3; r2.w *= hm3; r3.w *= hm3;
          }
          // slab (quad) -> [feature/8][feature%8][4 points]; this thread owns 4 consecutive features = 64 B
          const int off = (quad0 + 4 * h) * kSlabW + (c >> 1) * kSboW + (c & 1) * 64;
          split_store(hi_b, lo_b, off + 0, make_float4(r0.x, r1.x, r2.x, r3.x));
          split_store(hi_b, lo_b, off + 16, make_float4(r0.y, r1.y, r2.y, r3.y));
          split_store(hi_b, lo_b, off + 32, make_float4(r0.z, r1.z, r2.z, r3.z));
          split_store(hi_b, lo_b, off + 48, make_float4(r0.w, r1.w, r2.w, r3.w));
          if (is_a) {
            bsum.x += (r0.x + r1.x) + (r2.x + r3.x); bsum.y += (r0.y + r1.y) + (r2.y + r3.y);
            bsum.z += (r0.z + r1.z) + (r2.z + r3.z); bsum.w += (r0.w + r1.w) + (r2.w + r3.w);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[stage]);
      if (++stage == kStagesW) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
    }
    if (it.bias && is_a && on) {
      if (head) {
        float* gb = flat_grad + p.h[it.t].flat_b - hcol0;
        const float bs[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i >= hcol0 && i < hcol0 + hcols) atomicAdd(gb + i, bs[i]);
      } else {
        float* gb = flat_grad + g.flat_b + it.n0 + 4 * c;
        atomicAdd(gb + 0, bsum.x); atomicAdd(gb + 1, bsum.y); atomicAdd(gb + 2, bsum.z); atomicAdd(gb + 3, bsum.w);
      }
    }
    // ===================== drain the accumulator (warps 0-3: TMEM lane = output row n) =====================
    if (warp < 4) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int row = tid;
      const int in_real = head ? p.h[it.t].k : g.k_h + g.enc_real;
      const int coff = head ? 0 : (it.kind == 0 ? it.k0 : g.k_h);
      const int kreal = head ? p.h[it.t].k : (it.kind == 0 ? it.kblk : g.enc_real);
      const bool row_live = head ? (row >= hcol0 && row < hcol0 + hcols) : (row < it.nblk);
      float* dst = head ? flat_grad + p.h[it.t].flat_w + (size_t)(row - hcol0) * in_real
                        : flat_grad + g.flat_w + (size_t)(it.n0 + row) * in_real + coff;
      for (int c0 = 0; c0 < n_mma; c0 += 32) {
        uint32_t v[32];
        if (n_mma - c0 >= 32) {
          tmem_ld32(tmem + (((uint32_t)(warp * 32)) << 16) + c0, v);
        } else {  // 16 remaining columns (N = 48): read them with two x8 loads
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(v[8 * h8 + 0]), "=r"(v[8 * h8 + 1]), "=r"(v[8 * h8 + 2]), "=r"(v[8 * h8 + 3]),
                           "=r"(v[8 * h8 + 4]), "=r"(v[8 * h8 + 5]), "=r"(v[8 * h8 + 6]), "=r"(v[8 * h8 + 7])
                         : "r"(tmem + (((uint32_t)(warp * 32)) << 16) + c0 + 8 * h8)
                         : "memory");
          }
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0u;
        }
        tmem_wait_ld();
        if (row_live) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < kreal) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(n_mma);
      uint32_t stage = 0, phase = 0;
      for (int64_t s = 0; s < n_stage; ++s) {
        mbar_wait(&bar_full[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(sm + stage * kStageBytesW);
#pragma unroll
        for (int j = 0; j < kStagePts / 8; ++j) {  // 8 points (two slabs) per instruction
          const uint64_t a_hi = make_desc(sb + j * 2 * kSlabW, kSlabW, kSboW);
          const uint64_t a_lo = make_desc(sb + kOpBytes + j * 2 * kSlabW, kSlabW, kSboW);
          const uint64_t b_hi = make_desc(sb + 2 * kOpBytes + j * 2 * kSlabW, kSlabW, kSboW);
          const uint64_t b_lo = make_desc(sb + 3 * kOpBytes + j * 2 * kSlabW, kSlabW, kSboW);
          mma_ss(tmem, a_hi, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);
          mma_ss(tmem, a_lo, b_hi, idesc, 1u);
          mma_ss(tmem, a_hi, b_lo, idesc, 1u);
        }
        mma_commit(&bar_empty[stage]);
        if (++stage == kStagesW) { stage = 0; phase ^= 1; }
      }
      mma_commit(bar_done);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

int launch_wgrad_tc(const Plan& p, const float* rays, int ray_stride, const float* z, int64_t n_rays, int n_samples,
                    const float* stash, const float* gstash, const float* d_raw, float* flat_grad, cudaStream_t s) {
  if (p.hidden != 128) {
    set_error("wgrad impl=1 (tcgen05): hidden_size %d not supported (128 only)", p.hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int64_t P = n_rays * n_samples;
  const int items = wg_item_count(p);
  const size_t bytes = (size_t)kStagesW * kStageBytesW + 256 + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                      "wgrad_tc smem attribute");
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t stages = (P + kStagePts - 1) / kStagePts;
  int split = (2 * sms) / items;  // two full waves of one-CTA-per-SM, no ragged third wave
  if (split > stages) split = (int)stages;
  if (split < 1) split = 1;
  dim3 grid(split, items);
  mlp_wgrad_tc_kernel<<<grid, kThreadsW, bytes, s>>>(p, rays, ray_stride, z, n_samples, stash, gstash, d_raw, P,
                                                     flat_grad, items, get_tc_flags());
  count_launch();
  return check_cuda(cudaGetLastError(), "wgrad_tc launch");
}

}  // namespace nerfb200
