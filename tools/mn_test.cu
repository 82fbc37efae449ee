// mn_test.cu -- checks the MN-major ("transposed") shared-memory operand view of tcgen05.mma.kind::f16 that the
// fused backward kernel relies on (dev tool, not product):
//   tile T[pb][fb][8 points][8 features] fp16 (one 128-byte core matrix per (pb, fb))
//   K-major view  (M/N = points,   K = features): SBO = pb stride, LBO = fb stride
//   MN-major view (M/N = features, K = points)  : SBO = fb stride, LBO = pb stride
// Test 1: D[n][k] = sum_p G[p][n] X[p][k]   (A = G, B = X, both MN-major, M = N = 128, K = 128 points)
// Test 2: same with N = 64 (X tile 64 features wide)
// Test 3: D[p][k] = sum_f G[p][f] W[k][f]   (A = G K-major from the SAME tile, B = W K-major)  -- sanity
// Test 4: tcgen05.mma scale-input-d: D = A*B + D * 2^-11
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nerf_pytorch_b200/csrc -o tools/bin/mn_test tools/mn_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "tc_common.cuh"

using namespace nerfb200::tc;

__device__ __forceinline__ uint32_t idesc_f16(int n, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_scaled(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc)
               : "memory");
}

// smem: G tile (32 KB, F = 16), X tile (32 KB), W tile K-major [fb over k-features? ] see host
__global__ void __launch_bounds__(128, 1) mn_kernel(const __half* g, const __half* x, const __half* w, float* out, int test) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sG = sm;
  uint8_t* sX = sm + 32768;
  uint8_t* sW = sm + 65536;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 98304);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 16; i += 128) {
    reinterpret_cast<uint4*>(sG)[i] = reinterpret_cast<const uint4*>(g)[i];
    reinterpret_cast<uint4*>(sX)[i] = reinterpret_cast<const uint4*>(x)[i];
    reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(w)[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 0) {
    if (elect_one()) {
      if (test == 1 || test == 2 || test == 4) {
        const int n_mma = test == 2 ? 64 : 128;
        const int FX = n_mma / 8;  // fb blocks of the X tile
        const uint32_t id = idesc_f16(n_mma, 1, 1);
        for (int ks = 0; ks < 8; ++ks) {
          // A = G^T: M = features (fb stride 128 B = SBO), K = points (pb stride 16 * 128 = LBO); k-step = 2 pb
          const uint64_t a = make_desc(smem_u32(sG) + ks * 2 * 2048, /*lbo=*/2048, /*sbo=*/128);
          const uint64_t b = make_desc(smem_u32(sX) + ks * 2 * FX * 128, /*lbo=*/FX * 128, /*sbo=*/128);
          mma_ss_f16(tmem, a, b, id, ks > 0 ? 1u : 0u);
        }
        if (test == 4) {  // D = A0*B0 + D * 2^-11
          const uint64_t a = make_desc(smem_u32(sG), 2048, 128);
          const uint64_t b = make_desc(smem_u32(sX), 16 * 128, 128);
          mma_ss_scaled(tmem, a, b, id);
        }
      } else {
        // K-major views of the same tiles: A = G (M = points: SBO = pb stride 2048; K = features: LBO = 128),
        // B = W (N = k rows: W tile stored [kb][fb][8 k][8 f]: SBO = 2048, LBO = 128), K = 128 features = 8 k-steps of 2 fb
        const uint32_t id = idesc_f16(128, 0, 0);
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a = make_desc(smem_u32(sG) + ks * 256, 128, 2048);
          const uint64_t b = make_desc(smem_u32(sW) + ks * 256, 128, 2048);
          mma_ss_f16(tmem, a, b, id, ks > 0 ? 1u : 0u);
        }
      }
      mma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + lane_base + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

static size_t tile_off(int p, int f, int F) { return ((size_t)(p >> 3) * F + (f >> 3)) * 64 + (p & 7) * 8 + (f & 7); }  // in halves

int main() {
  std::vector<float> G(128 * 128), X(128 * 128), W(128 * 128), X64(128 * 64);
  srand(7);
  for (auto& v : G) v = (float)(rand() % 9 - 4);
  for (auto& v : X) v = (float)(rand() % 7 - 3) * 0.5f;
  for (auto& v : W) v = (float)(rand() % 5 - 2);
  for (int p = 0; p < 128; ++p)
    for (int k = 0; k < 64; ++k) X64[p * 64 + k] = X[p * 128 + k];
  std::vector<__half> tG(128 * 128), tX(128 * 128), tW(128 * 128), tX64(128 * 128);
  for (int p = 0; p < 128; ++p)
    for (int f = 0; f < 128; ++f) {
      tG[tile_off(p, f, 16)] = __float2half(G[p * 128 + f]);
      tX[tile_off(p, f, 16)] = __float2half(X[p * 128 + f]);
      tW[tile_off(p, f, 16)] = __float2half(W[p * 128 + f]);  // W[k][f]: rows k play the role of "points"
      if (f < 64) tX64[tile_off(p, f, 8)] = __float2half(X[p * 128 + f]);
    }
  __half *dG, *dX, *dW, *dX64;
  float* dO;
  cudaMalloc(&dG, 32768); cudaMalloc(&dX, 32768); cudaMalloc(&dW, 32768); cudaMalloc(&dX64, 32768);
  cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dG, tG.data(), 32768, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, tX.data(), 32768, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, tW.data(), 32768, cudaMemcpyHostToDevice);
  cudaMemcpy(dX64, tX64.data(), 32768, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 1024);
  std::vector<float> out(128 * 128);
  for (int test = 1; test <= 4; ++test) {
    cudaMemset(dO, 0, 128 * 128 * 4);
    mn_kernel<<<1, 128, 100 * 1024 + 1024>>>(dG, test == 2 ? dX64 : dX, dW, dO, test);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dO, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    const int ncols = test == 2 ? 64 : 128;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < ncols; ++c) {
        double ref = 0;
        if (test == 3) {
          for (int f = 0; f < 128; ++f) ref += (double)G[r * 128 + f] * W[c * 128 + f];
        } else {
          for (int p = 0; p < 128; ++p) ref += (double)G[p * 128 + r] * X[p * 128 + c];
          if (test == 4) {
            double extra = 0;
            for (int p = 0; p < 16; ++p) extra += (double)G[p * 128 + r] * X[p * 128 + c];
            ref = ref / 2048.0 + extra;
          }
        }
        const double err = fabs(ref - out[r * 128 + c]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) ++bad;
      }
    printf("test %d: %s  max err %.4g  bad %d / %d  (out[0][0..3] = %g %g %g %g)\n", test, cudaGetErrorString(e), maxerr, bad,
           128 * ncols, out[0], out[1], out[2], out[3]);
  }
  return 0;
}
