// mma_rate.cu -- how many cycles does one tcgen05.mma take on this B200?  (no memory traffic: operands are whatever
// sits in shared / tensor memory).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dualmessagepassing_b200/csrc
#include <cstdio>
#include "tc_common.cuh"
namespace dmp { void set_error(const char*, ...) {} }
using namespace dmp::gemm;

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 tf32 SS, 1 tf32 TS, 2 bf16 SS, 3 bf16 TS
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sbar = base + 96 * 1024, slot = sbar + 16;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 96 * 256; i += 128) reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { mbar_init(sbar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot_ptr;
  constexpr bool bf = MODE >= 2;
  const uint32_t idesc = bf ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24)) : make_idesc(128, N);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint64_t da = make_smem_desc(base + (j & 3) * 32);
        const uint64_t db = make_smem_desc(base + 32 * 1024 + (j & 3) * 32);
        const uint32_t d = tm + (uint32_t)((j & 1) * N) % 256;
        if (MODE == 0) umma_tf32(d, da, db, idesc, 1u);
        if (MODE == 1) umma_tf32_ts(d, tm + 256 + (j & 3) * 8, db, idesc, 1u);
        if (MODE == 2) umma_f16_ss(d, da, db, idesc, 1u);
        if (MODE == 3) umma_f16_ts(d, tm + 256 + (j & 3) * 8, db, idesc, 1u);
      }
    }
    umma_commit(sbar);
    mbar_wait(sbar, 0);
    t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int MODE, int N>
void run(const char* name, long long* d_out) {
  const int iters = 4000, smem = 100 * 1024 + 2048;
  cudaFuncSetAttribute(rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<MODE, N><<<148, 128, smem>>>(100, d_out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<MODE, N><<<148, 128, smem>>>(iters, d_out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double n_mma = 16.0 * iters;
  const int kk = MODE >= 2 ? 16 : 8;
  printf("%-10s N=%3d: %7.1f cycles/MMA  (%.0f MAC/clk/SM)  %.3f ms  -> %.0f TFLOP/s chip  [%s]\n", name, N, cyc / n_mma,
         128.0 * N * kk / (cyc / n_mma), ms, 2.0 * 128 * N * kk * n_mma * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  run<0, 64>("tf32 SS", d_out); run<0, 128>("tf32 SS", d_out); run<0, 256>("tf32 SS", d_out);
  run<1, 64>("tf32 TS", d_out); run<1, 128>("tf32 TS", d_out); run<1, 256>("tf32 TS", d_out);
  run<2, 64>("bf16 SS", d_out); run<2, 128>("bf16 SS", d_out); run<2, 256>("bf16 SS", d_out);
  run<3, 64>("bf16 TS", d_out); run<3, 128>("bf16 TS", d_out); run<3, 256>("bf16 TS", d_out);
  return 0;
}
