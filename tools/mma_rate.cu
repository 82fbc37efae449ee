// mma_rate.cu -- cycles per tcgen05.mma.kind::f16 (M = 128, K = 16) with a LEAN issue loop (descriptors precomputed,
// the start-address field advanced with one 32-bit add, k-loop unrolled): the tensor pipe's real rate by operand
// source / major-ness / N, and how much a fat issue loop (64-bit descriptor arithmetic per instruction) costs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nerf_pytorch_b200/csrc -o tools/bin/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include "tc_common.cuh"
using namespace nerfb200::tc;

// kMode 0: SS K-major x K-major   1: SS MN x MN   2: TS x K-major   3: pairs (N=128 into D0, N=16 into D1), MN x MN
template <int kMode, int kN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 196608);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 196608 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 65536);
    const uint64_t a_k = make_desc(a0, 128, 2048), a_mn = make_desc(a0, 2048, 128);
    const uint64_t b_k = make_desc(b0, 16 * kN, 128), b_mn = make_desc(b0, (kN / 8) * 128, 128);
    const uint64_t b16 = make_desc(b0 + 40960, 256, 128);
    const uint32_t id_k = make_idesc_f16(kN), id_mn = make_idesc_f16_mn(kN, 1, 1), id16 = make_idesc_f16_mn(16, 1, 1);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (kMode == 0) mma_ss_f16(tmem, desc_adv(a_k, ks * 256), desc_adv(b_k, ks * 32 * kN), id_k, 1u);
          if (kMode == 1) mma_ss_f16(tmem, desc_adv(a_mn, ks * 4096), desc_adv(b_mn, ks * 2 * (kN / 8) * 128), id_mn, 1u);
          if (kMode == 2) mma_ts_f16(tmem, tmem + 256 + 8 * ks, desc_adv(b_k, ks * 32 * kN), id_k, 1u);
          if (kMode == 3) {
            mma_ss_f16(tmem, desc_adv(a_mn, ks * 4096), desc_adv(b_mn, ks * 2 * (kN / 8) * 128), id_mn, 1u);
            mma_ss_f16(tmem + 128, desc_adv(a_mn, ks * 4096), desc_adv(b16, ks * 512), id16, 1u);
          }
        }
      }
      mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int kMode, int kN>
void run(long long* d, const char* name) {
  auto k = rate_kernel<kMode, kN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 128, 200 * 1024>>>(d, 8);
  k<<<1, 128, 200 * 1024>>>(d, 512);
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d : %6.1f cycles per k-step  [%s]\n", name, kN, (double)h / (512 * 8), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 256);
  run<0, 128>(d, "SS K-major x K-major");   run<0, 64>(d, "SS K-major x K-major");   run<0, 16>(d, "SS K-major x K-major");
  run<1, 128>(d, "SS MN-major x MN-major"); run<1, 64>(d, "SS MN-major x MN-major"); run<1, 16>(d, "SS MN-major x MN-major");
  run<2, 128>(d, "TS x K-major");           run<2, 64>(d, "TS x K-major");           run<2, 16>(d, "TS x K-major");
  run<3, 128>(d, "pair: N=128 (D0) + N=16 (D1), MN x MN");
  return 0;
}
