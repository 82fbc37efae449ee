// mlp_tc_bwd.cu -- tcgen05 weight-gradient kernel:  dW[n][k] = sum_p dY[p][n] * X[p][k]   (3xTF32)
//
// One CTA = one (layer, row-block, column-block) item of wgrad_items.cuh -- the 1..4-row heads included -- over a
// contiguous range of points.  The reduction runs over POINTS while the stash rows dY[p][:] / X[p][:] are
// feature-contiguous, so both operands are transposed on chip, 32 points per pipeline stage:
//   copy warp    both stashes are dense row-major arrays, so the 32 rows of a stage are ONE contiguous block per
//                operand: one cp.async.bulk each into a 4-deep ring of raw tiles (mbarrier complete_tx).  128 KB
//                per SM stay in flight, which is what keeps HBM busy (register prefetch could not).
//   A = dY^T     (M = 128 output features) lives in TENSOR MEMORY (lane = feature, column = point): every thread
//                reads ITS feature of the 32 raw rows (conflict-free 4-byte loads: the stash chunk swizzle is a
//                permutation inside each 128-byte segment), splits hi/lo and writes both with two tcgen05.st.
//   B = X^T      (N = 128|64|48.. input features) goes to shared memory: float4 reads of 4 features x 4 points,
//                4x4 register transpose, hi/lo split, 64 contiguous bytes per operand into the UMMA canonical
//                K-major no-swizzle layout (slab = [feature][4 points], core matrices padded to 144 B apart so the
//                stores are bank-conflict free).
//   Two transposer groups (8 warps each) alternate stages: the chain wait -> read -> split -> store -> fence ->
//   arrive is latency-bound for one warp, so two stages are always being transposed.
//   MMA warp     three tcgen05.mma.kind::tf32 per 8 points (hi*hi, lo*hi, hi*lo; A from tensor memory) into a
//                128 x N fp32 accumulator in tensor memory that lives for the whole point range; at the end the
//                A warps drain it with tcgen05.ld and atomically add into the flat gradient.
// Per point and item the kernel reads (n + k) * 4 bytes from HBM exactly once; it runs at ~80 % of the measured
// HBM bandwidth.
#include "common.cuh"
#include "tc_common.cuh"
#include "wgrad_items.cuh"

namespace nerfb200 {

namespace tcw {
constexpr int kThreadsW = 576;               // warps 0-3/8-11: A, 4-7/12-15: B (even/odd stages), 16: MMA + TMEM alloc, 17: copies
constexpr int kWarpMma = 16;
constexpr int kStagePts = 32;                // points per pipeline stage = 4 MMA k-groups
// Both rings have an EVEN number of slots, so a slot always belongs to the same transposer group: a parity wait
// is only safe if the waiting thread itself consumed the slot's previous phase.
constexpr int kRawStages = 4;                // ring of raw row blocks filled by the copy engine
constexpr int kOpStages = 2;                 // ring of transposed operands (B in shared memory, A in tensor memory)
constexpr int kSboW = 144;                   // 8-feature core matrices 128 B + 16 B pad apart
constexpr int kSlabW = 16 * kSboW;           // 128 features x 4 points (one K-major slab), padded
constexpr int kOpBytes = (kStagePts / 4) * kSlabW;  // one B tile (hi or lo): 8 slabs = 18 KB
constexpr int kStageBytesW = 2 * kOpBytes;   // B_hi | B_lo
constexpr int kRawHalf = kStagePts * 512;    // raw copy of 32 rows of <= 128 features
constexpr int kRawBytes = 2 * kRawHalf;      // dY rows | X rows
constexpr uint32_t kTmemColsW = 512;
constexpr uint32_t kColAccW = 0;             // accumulator: 128 columns
constexpr uint32_t kColA = 128;              // per operand stage 64 columns: A_hi (32 points) | A_lo (32 points)
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
mlp_wgrad_tc_kernel(const __grid_constant__ Plan p, const float* __restrict__ stash, const float* __restrict__ gstash,
                    const float* __restrict__ d_raw, int64_t P, float* __restrict__ flat_grad,
                    const __grid_constant__ WgGrid grid) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw = sm + kOpStages * kStageBytesW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + kRawStages * kRawBytes);
  uint64_t* raw_full = bars;                          // copy engine -> transposers (transaction bytes)
  uint64_t* raw_empty = raw_full + kRawStages;        // 8 transposer warps -> copy warp
  uint64_t* op_full = raw_empty + kRawStages;         // 256 transposer threads -> MMA warp
  uint64_t* op_empty = op_full + kOpStages;           // tcgen05.commit -> transposers
  uint64_t* bar_done = op_empty + kOpStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int item = 0;
  while (item + 1 < grid.n_items && (int)blockIdx.x >= grid.start[item + 1]) ++item;
  const int part = (int)blockIdx.x - grid.start[item], parts = grid.start[item + 1] - grid.start[item];
  const WgItem it = wg_decode(p, item);
  // operand sources.  gemm items: A rows = dY_t (gstash), B rows = the producing layer's output or the stashed
  // encoding.  Head items (fc_alpha / fc_rgb / fc_out): A rows = d_raw[p][0..3] restricted to the head's columns
  // (<= 4 live rows of the 128-row tile), B rows = the output of the layer the head reads.
  const bool head = it.kind == 2;
  const GemmLayer& g = p.g[head ? p.h[it.t].src : it.t];
  const int hcol0 = head ? p.h[it.t].out_col : 0, hcols = head ? p.h[it.t].n_out : 0;
  // both sources are dense row-major [P][width] arrays, so 32 consecutive points are ONE contiguous block
  int wa, wb;  // floats per dY row / per X row
  wg_row_widths(p, it, &wa, &wb);
  const float* src_a = head ? d_raw : gstash + (size_t)P * g.cum_n;
  const float* src_b = head ? stash + (size_t)P * g.cum_n
                            : (it.kind == 0 ? stash + (size_t)P * p.g[g.src].cum_n
                                            : stash + (size_t)P * p.enc_cum[g.enc_sel]);

  // contiguous point range of this CTA, in units of one stage
  const int64_t stages_total = (P + kStagePts - 1) / kStagePts;
  const int64_t per = (stages_total + parts - 1) / parts;
  int64_t pt_begin = (int64_t)part * per * kStagePts;
  int64_t pt_end = pt_begin + per * kStagePts;
  if (pt_begin > P) pt_begin = P;
  if (pt_end > P) pt_end = P;
  if (pt_begin >= pt_end) return;
  const int64_t n_stage = (pt_end - pt_begin + kStagePts - 1) / kStagePts;

  if (tid == 0) {
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], 8);
    }
    for (int i = 0; i < kOpStages; ++i) {
      mbar_init(&op_full[i], 256);
      mbar_init(&op_empty[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(kTmemColsW));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // zero the B tiles once: feature rows outside the item stay zero for the whole kernel
  for (int i = tid; i < kOpStages * kStageBytesW / 16; i += kThreadsW)
    reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int n_mma = it.kblk;  // accumulator columns (multiple of 16)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // two transposer groups alternate stages (group 0: warps 0-7, group 1: warps 8-15): the per-stage chain
  // wait -> shared-memory reads -> split -> tcgen05.st / st.shared -> fence -> arrive is latency-bound for one
  // warp, so two stages are kept in flight
  const int grp = warp >> 3, wg = warp & 7;
  if (warp < kWarpMma && wg < 4) {
    // ===================== A: raw dY rows -> tensor memory (lane = feature, column = point) =====================
    const int n = tid & 127;
    const bool on_n = head ? (n >= hcol0 && n < hcol0 + hcols) : (n < it.nblk && n < wa);
    const uint32_t lane_base = ((uint32_t)(wg * 32)) << 16;
    const int row_b = wa * 4;
    float bsum = 0.f;
    for (int64_t s = grp; s < n_stage; s += 2) {
      const uint32_t rs = (uint32_t)(s % kRawStages), rph = (uint32_t)(s / kRawStages) & 1u;
      const uint32_t os = (uint32_t)(s % kOpStages), oph = (uint32_t)(s / kOpStages) & 1u;
      const int rows = (int)min((int64_t)kStagePts, pt_end - (pt_begin + s * kStagePts));
      uint32_t hi[32], lo[32];
      mbar_wait(&raw_full[rs], rph);
      if (on_n) {
        const uint8_t* ar = raw + rs * kRawBytes + (n & 3) * 4;
        // head rows are plain (feature n = column n of d_raw); stash rows are chunk-swizzled by (point & 7), and
        // every stage starts at a multiple of 32 points
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int chunk = head ? 0 : ((n >> 2) ^ (j & 7));
          float a = *reinterpret_cast<const float*>(ar + j * row_b + chunk * 16);
          a = j < rows ? a : 0.f;
          hi[j] = tf32_hi(a);
          lo[j] = __float_as_uint(a - __uint_as_float(hi[j]));
          bsum += a;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) hi[j] = lo[j] = 0u;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[rs]);
      mbar_wait(&op_empty[os], oph ^ 1);
      __syncwarp();  // (the spin loops above may leave the warp diverged; tcgen05.st is .sync.aligned)
      tc_fence_after();
      tmem_st32(tmem + lane_base + kColA + 64 * os, hi);
      tmem_st32(tmem + lane_base + kColA + 64 * os + 32, lo);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&op_full[os]);
    }
    if (it.bias && on_n) {
      float* gb = head ? flat_grad + p.h[it.t].flat_b + (n - hcol0) : flat_grad + g.flat_b + it.n0 + n;
      atomicAdd(gb, bsum);
    }
    // ===================== drain the accumulator (TMEM lane = output row n) =====================
    mbar_wait(bar_done, 0);
    __syncwarp();
    tc_fence_after();
    const int in_real = head ? p.h[it.t].k : g.k_h + g.enc_real;
    const int coff = head ? 0 : (it.kind == 0 ? it.k0 : g.k_h);
    const int kreal = head ? p.h[it.t].k : (it.kind == 0 ? it.kblk : g.enc_real);
    float* dst = head ? flat_grad + p.h[it.t].flat_w + (size_t)(n - hcol0) * in_real
                      : flat_grad + g.flat_w + (size_t)(it.n0 + n) * in_real + coff;
    for (int c0 = 32 * grp; c0 < n_mma; c0 += 64) {  // the two groups share the drain
      uint32_t v[32];
      if (n_mma - c0 >= 32) {
        tmem_ld32(tmem + lane_base + kColAccW + c0, v);
      } else {  // 16 remaining columns (N = 48): read them with two x8 loads
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                       : "=r"(v[8 * h8 + 0]), "=r"(v[8 * h8 + 1]), "=r"(v[8 * h8 + 2]), "=r"(v[8 * h8 + 3]),
                         "=r"(v[8 * h8 + 4]), "=r"(v[8 * h8 + 5]), "=r"(v[8 * h8 + 6]), "=r"(v[8 * h8 + 7])
                       : "r"(tmem + lane_base + kColAccW + c0 + 8 * h8)
                       : "memory");
        }
#pragma unroll
        for (int j = 16; j < 32; ++j) v[j] = 0u;
      }
      tmem_wait_ld();
      if (on_n) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < kreal) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  } else if (warp < kWarpMma) {
    // ===================== B: raw X rows -> K-major hi/lo slabs.  lane = 4-feature chunk, warp = point quad =====
    const int c = lane;               // features 4c .. 4c+3
    const int quad0 = warp & 3;       // this thread transposes quads quad0 and quad0 + 4 of every stage
    const bool on = 4 * c < it.kblk && 4 * c < wb;
    const bool swz = head || it.kind == 0;  // layer outputs are chunk-swizzled, the stashed encodings are plain
    const int row_b = wb * 4;
    for (int64_t s = grp; s < n_stage; s += 2) {
      const uint32_t rs = (uint32_t)(s % kRawStages), rph = (uint32_t)(s / kRawStages) & 1u;
      const uint32_t os = (uint32_t)(s % kOpStages), oph = (uint32_t)(s / kOpStages) & 1u;
      const int rows = (int)min((int64_t)kStagePts, pt_end - (pt_begin + s * kStagePts));
      float4 v[8];
      mbar_wait(&raw_full[rs], rph);
      if (on) {
        const uint8_t* br = raw + rs * kRawBytes + kRawHalf;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = 4 * (quad0 + 4 * h) + i;
            const int cc = swz ? (c ^ (j & 7)) : c;
            const float4 x = *reinterpret_cast<const float4*>(br + j * row_b + cc * 16);
            v[4 * h + i] = j < rows ? x : zero4;
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[rs]);
      mbar_wait(&op_empty[os], oph ^ 1);
      if (on) {
        uint8_t* hi_b = sm + os * kStageBytesW;
        uint8_t* lo_b = hi_b + kOpBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 r0 = v[4 * h + 0], r1 = v[4 * h + 1], r2 = v[4 * h + 2], r3 = v[4 * h + 3];
          // slab (quad) -> [feature/8][feature%8][4 points]; this thread owns 4 consecutive features = 64 B
          const int off = (quad0 + 4 * h) * kSlabW + (c >> 1) * kSboW + (c & 1) * 64;
          split_store(hi_b, lo_b, off + 0, make_float4(r0.x, r1.x, r2.x, r3.x));
          split_store(hi_b, lo_b, off + 16, make_float4(r0.y, r1.y, r2.y, r3.y));
          split_store(hi_b, lo_b, off + 32, make_float4(r0.z, r1.z, r2.z, r3.z));
          split_store(hi_b, lo_b, off + 48, make_float4(r0.w, r1.w, r2.w, r3.w));
        }
      }
      fence_proxy_async();
      mbar_arrive(&op_full[os]);
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = make_idesc(n_mma);
    uint32_t os = 0, oph = 0;
    for (int64_t s = 0; s < n_stage; ++s) {
      mbar_wait(&op_full[os], oph);
      tc_fence_after();
      const uint32_t sb = smem_u32(sm + os * kStageBytesW);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < kStagePts / 8; ++j) {  // 8 points (two slabs) per instruction
          const uint32_t a_hi = tmem + kColA + 64 * os + 8 * j;
          const uint32_t a_lo = a_hi + 32;
          const uint64_t b_hi = make_desc(sb + j * 2 * kSlabW, kSlabW, kSboW);
          const uint64_t b_lo = make_desc(sb + kOpBytes + j * 2 * kSlabW, kSlabW, kSboW);
          mma_ts(tmem + kColAccW, a_hi, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);
          mma_ts(tmem + kColAccW, a_lo, b_hi, idesc, 1u);
          mma_ts(tmem + kColAccW, a_hi, b_lo, idesc, 1u);
        }
        mma_commit(&op_empty[os]);
      }
      __syncwarp();
      if (++os == kOpStages) { os = 0; oph ^= 1; }
    }
    if (elect_one()) mma_commit(bar_done);
    __syncwarp();
  } else {
    // ===================== copy warp: one bulk copy per operand and stage, kRawStages deep =====================
    uint32_t rs = 0, rph = 0;
    for (int64_t s = 0; s < n_stage; ++s) {
      const int64_t q0 = pt_begin + s * kStagePts;
      const uint32_t rows = (uint32_t)min((int64_t)kStagePts, pt_end - q0);
      mbar_wait(&raw_empty[rs], rph ^ 1);
      if (elect_one()) {
        uint8_t* dst = raw + rs * kRawBytes;
        const uint32_t ba = rows * wa * 4, bb = rows * wb * 4;
        mbar_arrive_expect_tx(&raw_full[rs], ba + bb);
        bulk_g2s(dst, src_a + (size_t)q0 * wa, ba, &raw_full[rs]);
        bulk_g2s(dst + kRawHalf, src_b + (size_t)q0 * wb, bb, &raw_full[rs]);
      }
      __syncwarp();
      if (++rs == kRawStages) { rs = 0; rph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemColsW));
}

int launch_wgrad_tc(const Plan& p, const float* rays, int ray_stride, const float* z, int64_t n_rays, int n_samples,
                    const float* stash, const float* gstash, const float* d_raw, float* flat_grad, cudaStream_t s) {
  (void)rays; (void)ray_stride; (void)z;  // the encodings come from the stash
  if (p.hidden != 128) {
    set_error("wgrad impl=1 (tcgen05): hidden_size %d not supported (128 only)", p.hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int64_t P = n_rays * n_samples;
  const int items = wg_item_count(p);
  const size_t bytes = (size_t)kOpStages * kStageBytesW + (size_t)kRawStages * kRawBytes + 256 + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                      "wgrad_tc smem attribute");
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // two full waves of one-CTA-per-SM, every item the same number of CTAs: a stage costs about the same for every
  // item (measured: shares proportional to bytes per point are 10-30 % slower), and equal shares line the items'
  // point ranges up, so rows two items share (dY of a layer with a skip input, X of the layer a head reads) are
  // served from L2 the second time.
  if (items > kWgMaxItems) {
    set_error("wgrad impl=1 (tcgen05): %d work items exceed the grid table (%d)", items, kWgMaxItems);
    return NERFB200_ERR_UNSUPPORTED;
  }
  const int64_t stages = (P + kStagePts - 1) / kStagePts;
  int share = (2 * sms) / items;
  if (share > stages) share = (int)stages;
  if (share < 1) share = 1;
  WgGrid grid;
  grid.n_items = items;
  for (int i = 0; i <= items; ++i) grid.start[i] = (short)(i * share);
  mlp_wgrad_tc_kernel<<<grid.start[items], kThreadsW, bytes, s>>>(p, stash, gstash, d_raw, P, flat_grad, grid);
  count_launch();
  return check_cuda(cudaGetLastError(), "wgrad_tc launch");
}

}  // namespace nerfb200
