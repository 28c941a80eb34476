// Micro-benchmark: bytes per cycle per SM of tcgen05.st (registers -> tensor memory, 32x32b.x32 = 4 KB per warp
// instruction) and of st.shared.v4 (512 B per warp instruction), with W warps per CTA, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ecs-yolo_b200/csrc tools/microbench/sttm_rate.cu -o sttm_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ecsy_common.cuh"
void ecsy_set_error(const char*, ...) {}
int ecsy_num_sms() { return 148; }
using namespace ecsy;

template <int MODE>   // 0: tcgen05.st x32 + wait::st each time, 1: tcgen05.st x32, one wait at the end, 2: st.shared.v4
__global__ void __launch_bounds__(1024, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) tmem_alloc<512>(&tbase);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  const uint32_t taddr = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)((warp >> 2) & 7) * 32u;
  const uint32_t saddr = smem_u32(raw) + (uint32_t)warp * 4096u + (uint32_t)lane * 16u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { tmem_st_32x32(taddr, v); tmem_st_wait(); }
    if (MODE == 1) tmem_st_32x32(taddr, v);
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr + j * 512), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                     "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
    }
    v[it & 31] += 1;
  }
  if (MODE <= 1) tmem_st_wait();
  __syncthreads();
  const long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tb); }
}

template <int MODE>
void run(const char* name, int warps, long long* d_out) {
  const int iters = 4000, smem = 32 * 4096;
  cudaFuncSetAttribute(k_rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_rate<MODE><<<148, warps * 32, smem>>>(100, d_out);
  k_rate<MODE><<<148, warps * 32, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * warps * 4096.0;
  printf("%-44s warps=%2d : %8.1f B/clk/SM   %7.1f cycles per 16 KB operand tile  (%s)\n", name, warps, bytes / cyc, 16384.0 * cyc / bytes,
         cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  for (int w : {4, 8, 16, 32}) run<0>("tcgen05.st 32x32b.x32 + wait::st per store", w, d_out);
  for (int w : {4, 8, 16, 32}) run<1>("tcgen05.st 32x32b.x32, one wait at the end", w, d_out);
  for (int w : {4, 8, 16, 32}) run<2>("st.shared.v4 (8 per iteration)", w, d_out);
  return 0;
}
