// corun.cu -- what a persistent one-CTA-per-SM projection kernel loses when another kernel (an NCCL collective) holds
// a few SMs: times dmp_gemm_tf32x3_dual (store form, 5 M x 128 x 128, through the C ABI of libdmp_b200.so) alone and
// next to a dummy kernel that keeps OCC SMs busy (64 KB of shared memory each, so no projection CTA fits beside it).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/corun scripts/micro/corun.cu \
//        -Ldualmessagepassing_b200 -ldmp_b200 -Xlinker -rpath -Xlinker $PWD/dualmessagepassing_b200
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/dmp_b200.h"

__global__ void occupy(long long ns) {
  extern __shared__ float hold[];
  hold[threadIdx.x] = 1.0f;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > ns) break;
  }
}

int main(int argc, char** argv) {
  const int64_t M = argc > 1 ? atoll(argv[1]) : 5000000;
  const int H = 128;
  float *A, *D, *W1, *W2, *c;
  cudaMalloc(&A, M * H * 4); cudaMalloc(&D, M * H * 4); cudaMalloc(&W1, H * H * 4); cudaMalloc(&W2, H * H * 4);
  cudaMalloc(&c, M * 4);
  cudaMemset(A, 0, M * H * 4); cudaMemset(W1, 0, H * H * 4); cudaMemset(W2, 0, H * H * 4); cudaMemset(c, 0, M * 4);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaFuncSetAttribute(occupy, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int occ : {0, 8, 16, 32}) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaDeviceSynchronize();
        if (occ) occupy<<<occ, 128, 65536, s2>>>(2500000);            // 2.5 ms: longer than the projection alone
        cudaEventRecord(e0, s1);
        int rc = dmp_gemm_tf32x3_dual(A, H, W1, W2, H, c, D, H, nullptr, 0, M, H, H, mode, s1);
        if (rc) { printf("error %d: %s\n", rc, dmp_last_error()); return 1; }
        cudaEventRecord(e1, s1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      printf("dual %s, %lld rows, %2d SMs held by another kernel: %.3f ms\n", mode ? "accumulate" : "store", (long long)M, occ, best);
    }
  }
  return 0;
}
