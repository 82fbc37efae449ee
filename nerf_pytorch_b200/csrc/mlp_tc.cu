// mlp_tc.cu -- tcgen05 (5th-generation tensor core) implementation of the two layer-chained kernels
//   forward : points -> positional encoding -> FlexibleNeRFModel   (nerf/train_utils.py:67, :8-25; nerf/models.py:233-256)
//   dgrad   : the same chain walked backwards, G_t = (G_s W_s[:, :hidden] + d_raw W_head) (.) relu'(layer t)
// for hidden_size 128, fp32-faithful through a 3-term split
//     x * w  ~=  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// accumulated in fp32 in tensor memory (single-pass TF32/BF16 misses the 1e-4
// bar on the shipped checkpoints; the 3-term split meets it).
// with an fp16 x 2 split of both operands (x_hi = fp16(x), x_lo = fp16((x - x_hi) 2^11): 22 significant bits, products
// exact in the fp32 accumulators; hi*hi and the 2^11-scaled cross terms accumulate separately) and tcgen05.mma.kind::f16, K = 16 per instruction -- half the tensor time of a 3xTF32 split.
//   forward: activations, encodings and weights of a NeRF sit far inside the split's range (|v| < 65504);
//            conversions saturate instead of producing infinities.
//   dgrad  : gradients span many decades ACROSS points (a sample's compositing weight scales its whole row), so
//            every row runs in its own power-of-two scale: d_raw[row] is scaled to max-abs in [1, 2) on load, the
//            chain is linear in it, and the gradient stash receives the exactly unscaled fp32 values.
//
// Persistent kernels, one CTA per SM, 320 threads:
//   warps 0-7  prologue/epilogue: thread (row = tid % 128, half = tid / 128) owns half of the columns of row
//              `row` of the 128-point tile (= TMEM lane row).  Per layer: tcgen05.ld the fp32 accumulator,
//              + bias / ReLU (forward) or + head term / ReLU mask (dgrad), narrow heads as register dot
//              products, split into fp16 hi + lo and tcgen05.st back into tensor memory as the NEXT layer's A
//              operand; training outputs (activation stash, ReLU bit mask, gradient stash) leave through a
//              per-warp swizzled shared-memory transpose so that every global store is a full 128-byte row.
//   warp 8     MMA issuer: one elected lane, tcgen05.mma.kind::f16 M=128, N=128|64, K=16, three per k-step;
//              A from tensor memory (hidden activations / gradients) or shared memory (encodings), B = pre-split
//              weights from the shared-memory ring (four k-steps per stage).
//   warp 9     weight producer: one cp.async.bulk per stage from the L2-resident blob, mbarrier complete_tx.
// Tensor memory (512 columns): [0,128) accumulator of hi*hi, [384,512) accumulator of the cross terms (scaled by
// 2^11, see tc_common.cuh split_f16x2), [128,192) A_hi, [256,320) A_lo (two fp16 per column).
// The direction encoding enters layers_dir[0] through a per-ray fp32 bias computed on the CUDA cores (it is
// constant along a ray: SURVEY.md section 7.3 item 5), so that layer contracts over K = 128 only.
#include "common.cuh"
#include "tc_common.cuh"

namespace nerfb200 {

static long long* g_tc_prof = nullptr;  // debug hook: per-CTA cycle counters (nerfb200_debug_tc_profile)
void set_tc_profile(void* p) { g_tc_prof = static_cast<long long*>(p); }
static int g_tc_flags = 0;  // debug: 1 = skip the weight copies, 2 = skip the MMAs (timing experiments only)
void set_tc_flags(int f) { g_tc_flags = f; }
int get_tc_flags() { return g_tc_flags; }

namespace tc {
constexpr int kEpiThreads = 256;
constexpr int kThreadsTc = 320;
constexpr int kMaxStages = 3;         // weight ring depth (2 while a forward also stages its stash tile)
constexpr int kStepsPerStage = 4;     // k-steps (8 KB each for N = 128) per ring stage
constexpr int kStageBytes = kStepsPerStage * 8192;
constexpr int kSlabBytes = 2048;      // 128 rows x 16 B
constexpr int kMaxRaysPerTile = 10;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColAcc = 0, kColAhi = 128, kColAlo = 256, kColAcc2 = 384;  // Acc2: cross terms, scaled 2^11

// Shared memory map (bytes from the 1 KB-aligned base).  The hi parts of the encodings are an SS-mode A operand in
// shared memory, their lo parts live in tensor memory (columns [384, 448)).  `stage` is the 64 KB stash staging tile:
// a byte-exact image of one 128-row tile of a stash layer, written by the epilogue and shipped with one
// cp.async.bulk; its first 32 KB double as the per-warp transpose tiles of the encoding stash.
struct Smem {
  static constexpr int e_hi = 0;                               // encodings, fp16: 8 slabs (K <= 64) x 2 KB hi, then 8 slabs lo
  static constexpr int ring = e_hi + 16 * kSlabBytes;          // kMaxStages x 32 KB
  static constexpr int stage = ring + kMaxStages * kStageBytes;  // 64 KB staging tile (training)
  static constexpr int tbuf = stage;                           // 8 warps x 4 KB transpose tiles (encoding stash)
  static constexpr int bias = stage + 65536;                   // kMaxGemm x 128 floats
  static constexpr int headw = bias + kMaxGemm * 128 * 4;      // 4*128 + 3*64 floats (+pad) and 8 bias floats
  static constexpr int viewb = headw + (4 * 128 + 3 * 64 + 16) * 4;   // kMaxRaysPerTile x 64
  static constexpr int encd = viewb + kMaxRaysPerTile * 64 * 4;       // kMaxRaysPerTile x 32
  static constexpr int hpart = encd + kMaxRaysPerTile * 32 * 4;       // 2 heads x 128 rows x 4 partial sums
  static constexpr int bars = hpart + 2 * 128 * 4 * 4;
  static constexpr int total = bars + 256;
};

__device__ __forceinline__ void epi_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 rows (lane = row) x 32 fp32 columns held one row per lane -> global rows of `ld` floats, coalesced:
// swizzled 16-byte chunks through a 4 KB per-warp tile, then 4 rows x 128 B per store instruction.
__device__ __forceinline__ void store_tile_coalesced(float* tbuf, const float (&x)[32], float* gdst, int ld, int lane,
                                                     int rows_valid) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(tbuf + lane * 32 + ((q ^ (lane & 7)) << 2)) =
        make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  __syncwarp();
  const int q = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    const float4 v = *reinterpret_cast<const float4*>(tbuf + r * 32 + ((q ^ (r & 7)) << 2));
    if (r < rows_valid) __stcs(reinterpret_cast<float4*>(gdst + (size_t)r * ld + 4 * q), v);
  }
  __syncwarp();
}

struct Pipe {  // role-local ring state
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n_stages) {
    if (++stage == n_stages) { stage = 0; phase ^= 1; }
  }
};

}  // namespace tc

using namespace tc;

// consumer gemm of layer t (the layer whose h-input is t's output), or -1
__device__ __forceinline__ int consumer_of(const Plan& p, int t) {
  int s = -1;
  for (int c = t + 1; c < p.n_gemm; ++c)
    if (p.g[c].src == t) s = c;
  return s;
}


// One 32-column chunk of one row in the epilogue (the hot loop of the chain kernels), specialised at compile time
// on the two things that change its instruction mix: a narrow head reading this layer (kHead) and training
// outputs (kTrain: ReLU bit mask + coalesced stash store).
//   forward: y = max(acc + bias, lb)                      (lb = 0 with ReLU, -inf without; bias already holds the
//                                                           per-ray direction term for layers_dir[0])
//   dgrad  : y = (acc + sum_c d_raw[c] * W_head[c]) masked by the forward ReLU bit
// then y -> fp16 hi / lo -> tensor memory (next layer's A operand).
struct ChunkArgs {
  const float* bias;      // fwd: 128 floats for this layer (per-thread pointer)
  float lb;               // fwd: ReLU lower bound
  const float* hw;        // head weights [hn][hk] in smem
  int hk, hn, hcol;
  float dr[4];            // dgrad: d_raw of this row
  uint32_t mword_in;      // dgrad: ReLU mask word of this chunk
  uint32_t* mword_out;    // fwd train: where to store the mask word (or nullptr when the row is out of range)
  uint8_t* stg_row;       // train: this row inside the staging tile (row * n * 4 bytes in), or nullptr
  int row7;               // row & 7: the stash chunk swizzle of this row (common.cuh swz_col)
  float unscale;          // dgrad: 2^-s of this row's power-of-two scale (applied to what goes to the gradient stash)
  uint32_t tmem_hi, tmem_lo;  // destination addresses (already offset to column c0)
  bool has_next;
};

template <int kMode, bool kHead, bool kTrain>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int c0, const ChunkArgs& a, float (&hacc)[4]) {
  float x[32];
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float y[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                  __uint_as_float(v[j + 3])};
    if (kMode == 0) {
      const float4 b = *reinterpret_cast<const float4*>(a.bias + c0 + j);
      y[0] = fmaxf(y[0] + b.x, a.lb); y[1] = fmaxf(y[1] + b.y, a.lb);
      y[2] = fmaxf(y[2] + b.z, a.lb); y[3] = fmaxf(y[3] + b.w, a.lb);
      if (kHead) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < a.hn) {
            const float4 w = *reinterpret_cast<const float4*>(a.hw + c * a.hk + c0 + j);
            hacc[c] = fmaf(y[0], w.x, fmaf(y[1], w.y, fmaf(y[2], w.z, fmaf(y[3], w.w, hacc[c]))));
          }
      }
      if (kTrain) {
#pragma unroll
        for (int q = 0; q < 4; ++q) bits |= (y[q] > 0.f ? 1u : 0u) << (j + q);
      }
    } else {
      if (kHead) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < a.hn) {
            const float4 w = *reinterpret_cast<const float4*>(a.hw + c * a.hk + c0 + j);
            const float d = a.dr[(a.hcol + c) & 3];
            y[0] = fmaf(d, w.x, y[0]); y[1] = fmaf(d, w.y, y[1]); y[2] = fmaf(d, w.z, y[2]); y[3] = fmaf(d, w.w, y[3]);
          }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) y[q] = ((a.mword_in >> (j + q)) & 1u) ? y[q] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) x[j + q] = y[q];
  }
  if (kMode == 0 && kTrain && a.mword_out) *a.mword_out = bits;
  if (kTrain && a.stg_row) {
    const float u = kMode == 1 ? a.unscale : 1.f;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(a.stg_row + ((((c0 >> 2) + q) ^ a.row7) << 4)) =
          kMode == 1 ? make_float4(x[4 * q] * u, x[4 * q + 1] * u, x[4 * q + 2] * u, x[4 * q + 3] * u)
                     : make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  }
  if (a.has_next) {
    // fp16 x 2: two K-adjacent values per tensor-memory column
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split_f16x2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
    tmem_st16(a.tmem_hi, hi);
    tmem_st16(a.tmem_lo, lo);
  }
}

template <int kMode>
__device__ __forceinline__ void epilogue_chunk_dispatch(bool head, bool train, const uint32_t (&v)[32], int c0,
                                                        const ChunkArgs& a, float (&hacc)[4]) {
  if (head) {
    if (train) epilogue_chunk<kMode, true, true>(v, c0, a, hacc);
    else epilogue_chunk<kMode, true, false>(v, c0, a, hacc);
  } else {
    if (train) epilogue_chunk<kMode, false, true>(v, c0, a, hacc);
    else epilogue_chunk<kMode, false, false>(v, c0, a, hacc);
  }
}

// kMode 0: forward, 1: dgrad
template <int kMode>
__global__ void __launch_bounds__(kThreadsTc, 1)
mlp_chain_tc_kernel(const __grid_constant__ Plan p, const float* __restrict__ blob, const float* __restrict__ rays,
                    int ray_stride, const float* __restrict__ z, int64_t P, int S, int64_t n_tiles,
                    float* __restrict__ raw,          // fwd: out [P][4];   dgrad: d_raw in (read only)
                    float* __restrict__ stash,        // fwd: out or NULL;  dgrad: in
                    float* __restrict__ gstash,       // dgrad: out
                    long long* __restrict__ prof, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1 KB aligned base
  float* s_bias = reinterpret_cast<float*>(sm + Smem::bias);
  float* s_headw = reinterpret_cast<float*>(sm + Smem::headw);
  float* s_viewb = reinterpret_cast<float*>(sm + Smem::viewb);
  float* s_encd = reinterpret_cast<float*>(sm + Smem::encd);
  float* s_hpart = reinterpret_cast<float*>(sm + Smem::hpart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Smem::bars);
  uint64_t* bar_full = bars;                      // [kMaxStages]  weights landed
  uint64_t* bar_empty = bars + kMaxStages;        // [kMaxStages]  stage consumed by the MMAs
  // A operand of the next layer: columns [0,64) ready AND the accumulator fully drained into registers (bar_a1),
  // columns [64,128) ready (bar_a2).  The next layer's first 8 k-steps only need the former, so they run while the
  // epilogue is still working on the second half of its columns.
  uint64_t* bar_a1 = bars + 2 * kMaxStages;
  uint64_t* bar_a2 = bars + 2 * kMaxStages + 1;
  uint64_t* bar_acc = bars + 2 * kMaxStages + 2;  // accumulator of the current layer is complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 3);
  const bool training = kMode == 1 || stash != nullptr;
  const uint32_t n_stages = kMaxStages;
  uint8_t* staging = sm + Smem::stage;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_a1, kEpiThreads);
    mbar_init(bar_a2, kEpiThreads);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // biases + head weights: once per CTA
  if (kMode == 0)
    for (int gi = 0; gi < p.n_gemm; ++gi)
      for (int i = tid; i < p.g[gi].n; i += kThreadsTc) s_bias[gi * 128 + i] = blob[p.g[gi].b_off + i];
  const int hw1 = p.h[0].n_out * p.h[0].k;
  const int hw2 = p.n_head > 1 ? p.h[1].n_out * p.h[1].k : 0;
  for (int i = tid; i < hw1; i += kThreadsTc) s_headw[i] = blob[p.h[0].w_off + i];
  for (int i = tid; i < hw2; i += kThreadsTc) s_headw[hw1 + i] = blob[p.h[1].w_off + i];
  float* s_headb = s_headw + ((hw1 + hw2 + 3) & ~3);
  if (tid < 4) s_headb[tid] = blob[p.h[0].b_off + tid];
  if (tid >= 4 && tid < 8) s_headb[tid] = p.n_head > 1 ? blob[p.h[1].b_off + tid - 4] : 0.f;
  for (int i = tid; i < kMaxRaysPerTile * 32; i += kThreadsTc) s_encd[i] = 0.f;  // padding channels stay zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 9) {
    // ===================== weight producer =====================
    if (lane == 0) {
      Pipe pp;
      const uint64_t pol = l2_policy_evict_last();
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int step = 0; step < p.n_gemm; ++step) {
          // forward: layer `step`;  dgrad: layers in reverse, operand of the CONSUMER of layer t
          const int t = kMode == 0 ? step : p.n_gemm - 1 - step;
          const int s = kMode == 0 ? t : consumer_of(p, t);
          if (s < 0) continue;
          const GemmLayer& g = p.g[s];
          const uint32_t kbytes = kMode == 0 ? 64u * g.n : 64u * g.k_h;  // one k-step (hi + lo)
          const uint8_t* src = reinterpret_cast<const uint8_t*>(blob + (kMode == 0 ? g.tc_off : g.tcd_off));
          const int ksteps = kMode == 0 ? (g.k_tc + 15) >> 4 : g.n >> 4;  // fp16: K = 16 per step
          for (int ks = 0; ks < ksteps; ks += kStepsPerStage) {
            const uint32_t bytes = (uint32_t)min(kStepsPerStage, ksteps - ks) * kbytes;
            mbar_wait(&bar_empty[pp.stage], pp.phase ^ 1);
            if (dbg & 1) {
              mbar_arrive(&bar_full[pp.stage]);
            } else {
              mbar_arrive_expect_tx(&bar_full[pp.stage], bytes);
              bulk_g2s_hint(sm + Smem::ring + pp.stage * kStageBytes, src + (size_t)ks * kbytes, bytes,
                            &bar_full[pp.stage], pol);
            }
            pp.advance(n_stages);
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      Pipe pp;
      uint32_t a_phase = 0;
      const uint32_t e_hi = smem_u32(sm + Smem::e_hi);
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int step = 0; step < p.n_gemm; ++step) {
          const int t = kMode == 0 ? step : p.n_gemm - 1 - step;
          const int s = kMode == 0 ? t : consumer_of(p, t);
          if (s < 0) continue;
          const GemmLayer& g = p.g[s];
          const int n_mma = kMode == 0 ? g.n : g.k_h;
          const uint32_t idesc = make_idesc_f16(n_mma);
          const uint32_t slab_b = 16u * n_mma;  // bytes of one weight slab
          const int ksteps = kMode == 0 ? (g.k_tc + 15) >> 4 : g.n >> 4;
          const int ksteps_h = kMode == 0 ? (g.k_h >> 4) : ksteps;  // k-steps whose A operand is in tensor memory
          constexpr int kHalfSteps = 4;                             // k-steps covered by A columns [0, 64)
          mbar_wait(bar_a1, a_phase);
          tc_fence_after();
          bool second = false;  // bar_a2 of this layer consumed?
          for (int ks0 = 0; ks0 < ksteps; ks0 += kStepsPerStage) {
            if (!second && ks0 + kStepsPerStage > kHalfSteps) {  // this stage touches A columns >= 64 (or the encodings)
              mbar_wait(bar_a2, a_phase);
              tc_fence_after();
              second = true;
            }
            mbar_wait(&bar_full[pp.stage], pp.phase);
            tc_fence_after();
            const uint32_t wb0 = smem_u32(sm + Smem::ring + pp.stage * kStageBytes);
            if (elect_one()) {
#pragma unroll
            for (int h = 0; h < kStepsPerStage; ++h) {
              const int ks = ks0 + h;
              if (ks < ksteps && !(dbg & 2)) {
                const uint32_t wb = wb0 + h * 4 * slab_b;
                const uint64_t b_hi = make_desc(wb, slab_b, 128);
                const uint64_t b_lo = make_desc(wb + 2 * slab_b, slab_b, 128);
                const uint32_t acc0 = ks > 0 ? 1u : 0u;
                const uint32_t a_hi = tmem + kColAhi + 8 * ks, a_lo = tmem + kColAlo + 8 * ks;
                if (ks < ksteps_h) {
                  mma_ts_f16(tmem + kColAcc, a_hi, b_hi, idesc, acc0);
                  mma_ts_f16(tmem + kColAcc2, a_lo, b_hi, idesc, acc0);
                  mma_ts_f16(tmem + kColAcc2, a_hi, b_lo, idesc, 1u);
                } else {  // encodings: both halves are shared-memory operands (8 slabs hi, 8 slabs lo)
                  const uint32_t off = (uint32_t)(ks - ksteps_h) * 2 * kSlabBytes;
                  const uint64_t e_hi_d = make_desc(e_hi + off, kSlabBytes, 128);
                  const uint64_t e_lo_d = make_desc(e_hi + 8 * kSlabBytes + off, kSlabBytes, 128);
                  mma_ss_f16(tmem + kColAcc, e_hi_d, b_hi, idesc, acc0);
                  mma_ss_f16(tmem + kColAcc2, e_lo_d, b_hi, idesc, acc0);
                  mma_ss_f16(tmem + kColAcc2, e_hi_d, b_lo, idesc, 1u);
                }
              }
            }
            mma_commit(&bar_empty[pp.stage]);  // frees the ring stage once these MMAs have read it
            }
            __syncwarp();
            pp.advance(n_stages);
          }
          if (!second) mbar_wait(bar_a2, a_phase);  // keep the phases aligned for short layers
          a_phase ^= 1;
          if (elect_one()) mma_commit(bar_acc);  // accumulator of this layer complete
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== prologue / epilogue warps =====================
    const int row = tid & 127, half = tid >> 7;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    float* tbuf = reinterpret_cast<float*>(sm + Smem::tbuf) + warp * 1024;
    uint32_t acc_phase = 0;
    uint8_t* e_hi = sm + Smem::e_hi;
    long long t_pro = 0, t_wait = 0, t_epi = 0, t_all = clock64(), t_ld = 0, t_ch = 0, t_st = 0, t_hd = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      long long t0 = clock64();
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t p0 = tile * kTileRows;
      int64_t pt = p0 + row;
      const bool valid = pt < P;
      if (!valid) pt = P - 1;
      const int64_t wrow0 = p0 + (warp & 3) * 32;  // first point of this warp's 32-row block
      const int rows_valid = (int)(P - wrow0 < 32 ? (P - wrow0 < 0 ? 0 : P - wrow0) : 32);
      int ray_slot = 0;
      float dr[4] = {0.f, 0.f, 0.f, 0.f};
      float row_unscale = 1.f;

      if (kMode == 0) {
        const int64_t ray = pt / S;
        const int64_t first_ray = p0 / S;
        const int64_t last_pt = (p0 + kTileRows - 1 < P) ? p0 + kTileRows - 1 : P - 1;
        const int n_rays_tile = (int)(last_pt / S - first_ray) + 1;
        ray_slot = (int)(ray - first_ray);
        // ---- prologue: encodings of this row -> E_hi / E_lo (canonical K-major slabs); the two halves split the frequencies
        {
          const float* rr = rays + ray * ray_stride;
          const float zz = z[pt];
          // the encoding also goes to the stash for the backward: through the coalesced path below when the
          // padded width is the usual 64, else element by element
          float* sx = (stash && valid && p.dim_xyz_pad != 64)
                          ? stash + (size_t)P * p.enc_cum[0] + (size_t)pt * p.dim_xyz_pad : nullptr;
          const int nf = p.n_freq_xyz, mid = nf >> 1;
          const int f0 = half ? mid : 0, f1 = half ? nf : mid;
          const int base = p.inc_xyz ? 3 : 0;
          auto put = [&](int k, float v) {
            if (sx) sx[k] = v;
            uint32_t hi, lo;
            split_f16x2(v, 0.f, hi, lo);  // this element in the low halves
            const int off = (k >> 3) * kSlabBytes + row * 16 + (k & 7) * 2;
            *reinterpret_cast<uint16_t*>(e_hi + off) = (uint16_t)hi;
            *reinterpret_cast<uint16_t*>(e_hi + 8 * kSlabBytes + off) = (uint16_t)lo;
          };
          for (int c = 0; c < 3; ++c) {
            const float x = __fadd_rn(rr[c], __fmul_rn(rr[3 + c], zz));  // pts = ro + rd * z (train_utils.py:67)
            if (p.inc_xyz && half == 0) put(c, x);
            for (int f = f0; f < f1; ++f) {
              float sn, cs;
              sincosf(__fmul_rn(x, p.freq_xyz[f]), &sn, &cs);
              put(base + 6 * f + c, sn);
              put(base + 6 * f + 3 + c, cs);
            }
          }
          if (half == 1) {  // zero padding: the stash row is dim_xyz_pad wide, the operand one fp16 k-step (16) granular
            for (int k = p.dim_xyz; k < p.dim_xyz_pad; ++k) put(k, 0.f);
            sx = nullptr;
            for (int k = p.dim_xyz_pad; k < ((p.dim_xyz_pad + 15) & ~15); ++k) put(k, 0.f);
          }
        }
        // ---- per-ray direction term of layers_dir[0]: vb[ray][n] = sum_k enc_dir(ray)[k] * W[n][H + k]
        if (p.use_viewdirs) {
          if (tid < n_rays_tile * 3) {
            const int j = tid / 3, c = tid - 3 * j;
            const float v = rays[(first_ray + j) * ray_stride + 8 + c];
            encode_coord(v, c, p.inc_dir, 0, p.n_freq_dir, p.freq_dir, s_encd + j * 32);
          }
          epi_bar256();
          const GemmLayer& gd = p.g[p.n_gemm - 1];
          const float* wv = blob + gd.wt_off + (size_t)gd.k_h * gd.n;  // rows k_h.. of Wt[k][n]
          for (int i = tid; i < n_rays_tile * gd.n; i += kEpiThreads) {
            const int j = i / gd.n, n = i - j * gd.n;
            float a = 0.f;
            for (int k = 0; k < p.dim_dir; ++k) a = fmaf(s_encd[j * 32 + k], wv[k * gd.n + n], a);
            s_viewb[j * 64 + n] = a + s_bias[(p.n_gemm - 1) * 128 + n];  // per-ray bias of layers_dir[0]
          }
        }
        tc_fence_before();
        fence_proxy_async();  // E was written through the generic proxy; the MMAs read it via the async proxy
        if (training && tid == 0) bulk_wait_read();  // transpose tiles below overlap the staging tile
        epi_bar256();         // also publishes s_viewb
        if (stash) {
          if (p.dim_xyz_pad == 64) {
            // this thread's row, channels [32*half, 32*half + 32): hi + lo is the value the forward contracted with
            float x[32];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 h8 = *reinterpret_cast<const uint4*>(e_hi + (4 * half + q) * kSlabBytes + row * 16);
              const uint4 l8 = *reinterpret_cast<const uint4*>(e_hi + (8 + 4 * half + q) * kSlabBytes + row * 16);
              const uint32_t hh[4] = {h8.x, h8.y, h8.z, h8.w}, ll[4] = {l8.x, l8.y, l8.z, l8.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                x[8 * q + 2 * i] = fmaf(f16_lo_to_f32(ll[i]), kLoInv, f16_lo_to_f32(hh[i]));
                x[8 * q + 2 * i + 1] = fmaf(f16_hi_to_f32(ll[i]), kLoInv, f16_hi_to_f32(hh[i]));
              }
            }
            store_tile_coalesced(tbuf, x, stash + (size_t)P * p.enc_cum[0] + (size_t)wrow0 * 64 + 32 * half, 64, lane,
                                 rows_valid);
          }
          if (valid && p.use_viewdirs && half == 0) {
            float4* sd = reinterpret_cast<float4*>(stash + (size_t)P * p.enc_cum[1] + (size_t)pt * p.dim_dir_pad);
            const float4* se = reinterpret_cast<const float4*>(s_encd + ray_slot * 32);
            for (int k = 0; k < (p.dim_dir_pad >> 2); ++k) __stcs(sd + k, se[k]);
          }
        }
        mbar_arrive(bar_a1);
        mbar_arrive(bar_a2);
      } else {
        const float4 d4 = valid ? reinterpret_cast<const float4*>(raw)[pt] : make_float4(0.f, 0.f, 0.f, 0.f);
        // this row's power-of-two scale: 2^(127 - e) with e the biased exponent of max |d_raw[row]| (1 for zero rows)
        const float m = fmaxf(fmaxf(fabsf(d4.x), fabsf(d4.y)), fmaxf(fabsf(d4.z), fabsf(d4.w)));
        const uint32_t e = (__float_as_uint(m) >> 23) & 0xFFu;
        const bool scaled = e >= 1u && e <= 253u;
        const float sc = scaled ? __uint_as_float((254u - e) << 23) : 1.f;
        row_unscale = scaled ? __uint_as_float(e << 23) : 1.f;
        dr[0] = d4.x * sc; dr[1] = d4.y * sc; dr[2] = d4.z * sc; dr[3] = d4.w * sc;
      }
      t_pro += clock64() - t0;

      // ---- layers ----
      for (int step = 0; step < p.n_gemm; ++step) {
        const int t = kMode == 0 ? step : p.n_gemm - 1 - step;
        const GemmLayer& g = p.g[t];
        const int s_cons = kMode == 0 ? -1 : consumer_of(p, t);
        const bool has_mma = kMode == 0 ? true : s_cons >= 0;
        const bool has_next = kMode == 0 ? (t + 1 < p.n_gemm) : (t > 0);
        int hsel = -1;
        if (p.h[0].src == t) hsel = 0;
        if (p.n_head > 1 && p.h[1].src == t) hsel = 1;
        const float* hw = hsel == 1 ? s_headw + hw1 : s_headw;
        const int hk = hsel >= 0 ? p.h[hsel].k : 0, hn = hsel >= 0 ? p.h[hsel].n_out : 0;
        const int hcol = hsel >= 0 ? p.h[hsel].out_col : 0;
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
        const bool is_dir = kMode == 0 && p.use_viewdirs && t == p.n_gemm - 1;
        // training side outputs / inputs of this layer
        uint32_t* mask_row = nullptr;
        if (kMode == 0 ? (stash != nullptr) : (g.relu != 0))
          mask_row = reinterpret_cast<uint32_t*>(stash) + (size_t)P * (p.mask_base + g.mask_cum) + (size_t)pt * (g.n >> 5);

        // dgrad: fetch this row's ReLU mask words now so that their latency hides behind the MMA wait
        uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (kMode == 1 && g.relu) {
          mw[0] = valid ? __ldg(mask_row + half) : 0u;                       // columns [32*half, +32)
          mw[1] = (valid && g.n == 128) ? __ldg(mask_row + 2 + half) : 0u;     // columns [64 + 32*half, +32)
        }
        t0 = clock64();
        if (training) {  // the previous layer's bulk store must have finished reading the staging tile; checked
          if (tid == 0) bulk_wait_read();  // BEFORE the accumulator wait, so this barrier hides under the MMAs
          epi_bar256();
        }
        if (has_mma) {
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
        }
        const long long t1 = clock64();
        t_wait += t1 - t0;

        const long long t2 = clock64();
        t_ld += t2 - t1;

        {
          ChunkArgs ca;
          const bool train = kMode == 0 ? (stash != nullptr) : true;
          ca.bias = (kMode == 0) ? (is_dir ? s_viewb + ray_slot * 64 : s_bias + t * 128) : nullptr;
          ca.lb = g.relu ? 0.f : -3.4e38f;
          ca.hw = hw; ca.hk = hk; ca.hn = hn; ca.hcol = hcol;
          ca.dr[0] = dr[0]; ca.dr[1] = dr[1]; ca.dr[2] = dr[2]; ca.dr[3] = dr[3];
          ca.has_next = has_next;
          ca.row7 = row & 7;
          ca.unscale = row_unscale;
          ca.stg_row = train ? staging + (size_t)row * g.n * 4 : nullptr;
          // column chunks of this thread: first [32*half, +32), second [64 + 32*half, +32) (128-wide layers only).
          // Both are pulled out of the accumulator up front; after the first chunk's A columns are stored the
          // next layer may start its first 8 k-steps (bar_a1), the second chunk follows under that shadow.
          const int nch = g.n >> 6;  // 2 for 128-wide layers, 1 for 64-wide ones
          const int c0a = 32 * half, c0b = 64 + 32 * half;
          uint32_t v0[32], v1[32];
          if (has_mma) {
            // hi*hi accumulator + 2^-11 x cross-term accumulator of the same columns (16 at a time: registers)
            auto fold = [&](uint32_t (&v)[32], int c0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t w[16];
                tmem_ld16(tmem + lane_base + kColAcc2 + c0 + 16 * h, w);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  v[16 * h + j] = __float_as_uint(fmaf(__uint_as_float(w[j]), kLoInv, __uint_as_float(v[16 * h + j])));
              }
            };
            tmem_ld32(tmem + lane_base + kColAcc + c0a, v0);
            if (nch == 2) tmem_ld32(tmem + lane_base + kColAcc + c0b, v1);
            fold(v0, c0a);
            if (nch == 2) fold(v1, c0b);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v0[j] = v1[j] = 0u;
          }
          ca.mword_in = mw[0];
          ca.mword_out = (kMode == 0 && mask_row && valid) ? mask_row + (c0a >> 5) : nullptr;
          constexpr int kPack = 2;  // two fp16 per tensor-memory column
          ca.tmem_hi = tmem + lane_base + kColAhi + c0a / kPack;
          ca.tmem_lo = tmem + lane_base + kColAlo + c0a / kPack;
          epilogue_chunk_dispatch<kMode>(hsel >= 0, train, v0, c0a, ca, hacc);
          if (has_next) {
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_a1);
          }
          if (nch == 2) {
            ca.mword_in = mw[1];
            ca.mword_out = (kMode == 0 && mask_row && valid) ? mask_row + (c0b >> 5) : nullptr;
            ca.tmem_hi = tmem + lane_base + kColAhi + c0b / kPack;
            ca.tmem_lo = tmem + lane_base + kColAlo + c0b / kPack;
            epilogue_chunk_dispatch<kMode>(hsel >= 0, train, v1, c0b, ca, hacc);
          }
        }
        const long long t3 = clock64();
        t_ch += t3 - t2;

        if (kMode == 0 && hsel >= 0 && half == 1)
          *reinterpret_cast<float4*>(s_hpart + (hsel * 128 + row) * 4) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
        if (has_next) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(bar_a2);
        } else {
          tc_fence_before();
        }
        const long long t4 = clock64();
        t_st += t4 - t3;
        if (training) fence_proxy_async();  // staging tile written through the generic proxy, read by the bulk copy
        if (training || (kMode == 0 && hsel >= 0)) epi_bar256();  // staging tile complete / head partials in s_hpart
        if (training && tid == 0) {
          const int64_t rows_tile = P - p0 < kTileRows ? P - p0 : kTileRows;
          float* dst = (kMode == 0 ? stash : gstash) + (size_t)P * g.cum_n + (size_t)p0 * g.n;
          bulk_s2g(dst, staging, (uint32_t)(rows_tile * g.n * 4));
        }
        if (kMode == 0 && hsel >= 0) {
          if (half == 0 && valid) {
            const float4 o = *reinterpret_cast<const float4*>(s_hpart + (hsel * 128 + row) * 4);
            const float tot[4] = {hacc[0] + o.x + s_headb[hsel * 4 + 0], hacc[1] + o.y + s_headb[hsel * 4 + 1],
                                  hacc[2] + o.z + s_headb[hsel * 4 + 2], hacc[3] + o.w + s_headb[hsel * 4 + 3]};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < hn) raw[pt * 4 + hcol + c] = tot[c];
          }
        }
        t_epi += clock64() - t1;
        t_hd += clock64() - t4;
      }
    }
    if (prof && tid == 0) {
      prof[blockIdx.x * 8 + 0] = t_pro;
      prof[blockIdx.x * 8 + 1] = t_wait;
      prof[blockIdx.x * 8 + 2] = t_epi;
      prof[blockIdx.x * 8 + 3] = clock64() - t_all;
      prof[blockIdx.x * 8 + 4] = t_ld;
      prof[blockIdx.x * 8 + 5] = t_ch;
      prof[blockIdx.x * 8 + 6] = t_st;
      prof[blockIdx.x * 8 + 7] = t_hd;
    }
  }

  if (tid == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

static int tc_supported(const Plan& p, int n_samples, const char* what) {
  if (p.hidden != 128) {
    set_error("%s impl=1 (tcgen05): hidden_size %d not supported (128 only); use impl=0", what, p.hidden);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (p.dim_xyz_pad > 64 || p.dim_dir > 32) {
    set_error("%s impl=1 (tcgen05): encodings wider than 64 (xyz) / 32 (dir) not supported; use impl=0", what);
    return NERFB200_ERR_UNSUPPORTED;
  }
  if (n_samples > 0 && (kTileRows + n_samples - 1) / n_samples + 1 > kMaxRaysPerTile) {
    set_error("%s impl=1 (tcgen05): fewer than 16 samples per ray not supported; use impl=0", what);
    return NERFB200_ERR_UNSUPPORTED;
  }
  return NERFB200_OK;
}

template <int kMode>
static int launch_chain(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z, int64_t P,
                        int n_samples, float* raw, float* stash, float* gstash, cudaStream_t s, const char* what) {
  const int64_t tiles = (P + kTileRows - 1) / kTileRows;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < sms ? tiles : sms);
  const size_t bytes = Smem::total + 1024;
  auto kern = mlp_chain_tc_kernel<kMode>;
  int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what);
  if (rc) return rc;
  kern<<<grid, kThreadsTc, bytes, s>>>(p, blob, rays, ray_stride, z, P, n_samples, tiles, raw, stash, gstash,
                                       g_tc_prof, g_tc_flags);
  count_launch();
  return check_cuda(cudaGetLastError(), what);
}

int launch_mlp_fwd_tc(const Plan& p, const float* blob, const float* rays, int ray_stride, const float* z,
                      int64_t n_rays, int n_samples, float* raw, float* stash, cudaStream_t s) {
  int rc = tc_supported(p, n_samples, "mlp_fwd");
  if (rc) return rc;
  return launch_chain<0>(p, blob, rays, ray_stride, z, n_rays * n_samples, n_samples, raw, stash, nullptr, s,
                         "mlp_fwd_tc launch");
}

int launch_dgrad_tc(const Plan& p, const float* blob, const float* d_raw, const float* stash, float* gstash, int64_t P,
                    cudaStream_t s) {
  int rc = tc_supported(p, 0, "dgrad");
  if (rc) return rc;
  return launch_chain<1>(p, blob, nullptr, 0, nullptr, P, 1, const_cast<float*>(d_raw), const_cast<float*>(stash),
                         gstash, s, "dgrad_tc launch");
}

}  // namespace nerfb200
