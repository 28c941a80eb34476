// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, K = 16) on one CTA per SM as a function of
// (M, N), operand-A source (shared memory descriptor vs tensor memory) and the number of accumulators the
// instruction stream alternates between.  Operands are zeros: only the issue / execution rate matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ecs-yolo_b200/csrc tools/microbench/mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ecsy_common.cuh"
void ecsy_set_error(const char*, ...) {}
int ecsy_num_sms() { return 148; }
using namespace ecsy;

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

template <int M, int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t* a_s = smem;                 // 128 rows x 128 B
  uint8_t* b_s = smem + 128 * 128;     // 256 rows x 128 B
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint64_t da = umma_desc_sw128(smem_u32(a_s)), db = umma_desc_sw128(smem_u32(b_s));
    constexpr uint32_t idesc = idesc_mn(M, N);
    // accumulators at columns 0, N, 2N, ... (NACC * N <= 256); TS operand at column 256
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          if (TS) umma_f16_ts(tb + a * N, tb + 256 + k * 8, db + (uint64_t)(k * 2), idesc, 1u);
          else umma_f16(tb + a * N, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tb); }
}

// The issuer loop of the production kernels: whole warp converged, a runtime stage index, descriptors rebuilt per K block,
// elect_one() around the four MMAs, one tcgen05.commit per K block (optionally a satisfied mbarrier wait as well).
template <int N, bool TS, bool WAIT>
__global__ void __launch_bounds__(128, 1) k_issuer(int iters, int stages, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2, ready[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&bar2, 1u << 20);
    for (int s = 0; s < 8; ++s) mbar_init(&ready[s], 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tbase;
  if (threadIdx.x < 8) mbar_arrive(&ready[threadIdx.x]);   // phase 0 of every `ready` barrier is complete
  __syncthreads();
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = idesc_mn(128, N);
    uint32_t stage = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (WAIT) mbar_wait(&ready[stage], 0);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t b_addr = smem_u32(smem + 128 * 128) + (stage & 1) * 0u;
        const uint32_t a_tmem = tb + 256 + stage * 32;
        const uint64_t da = umma_desc_sw128(smem_u32(smem));
        const uint64_t db = umma_desc_sw128(b_addr);
        uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) umma_f16_ts(tb, a_tmem + k * 8, db + (uint64_t)(k * 2), idesc, acc);
          else umma_f16(tb, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
          acc = 1u;
        }
        umma_commit(&bar2);
      }
      __syncwarp();
      if (++stage == (uint32_t)stages) stage = 0;
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tb); }
}

template <int N, bool TS, bool WAIT>
void run_issuer(const char* name, long long* d_out) {
  const int iters = 2000, smem = 1024 + (128 + 256) * 128;
  cudaFuncSetAttribute(k_issuer<N, TS, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_issuer<N, TS, WAIT><<<148, 128, smem>>>(200, 8, d_out);
  k_issuer<N, TS, WAIT><<<148, 128, smem>>>(iters, 8, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d A=%s : %7.1f cycles per K block (4 MMAs + commit)  (%s)\n", name, N, TS ? "tmem" : "smem", (double)cyc / iters,
         cudaGetErrorString(e));
}

template <int M, int N, bool TS, int NACC>
void run(const char* name, long long* d_out) {
  const int iters = 2000;
  const int smem = 1024 + (128 + 256) * 128;
  cudaFuncSetAttribute(k_rate<M, N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_rate<M, N, TS, NACC><<<148, 128, smem>>>(200, d_out);   // warm-up
  k_rate<M, N, TS, NACC><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double per = (double)cyc / (iters * 4.0 * NACC);
  const double flop_clk = 2.0 * M * N * 16 / per;
  printf("%-34s M=%3d N=%3d A=%s acc=%d : %7.1f cycles/MMA  %7.0f FLOP/clk/SM  (%s)\n", name, M, N, TS ? "tmem" : "smem", NACC, per,
         flop_clk, cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<128, 64, false, 1>("SS", d_out);   run<128, 128, false, 1>("SS", d_out);  run<128, 256, false, 1>("SS", d_out);
  run<128, 64, true, 1>("TS", d_out);    run<128, 128, true, 1>("TS", d_out);   run<128, 256, true, 1>("TS", d_out);
  run<128, 64, false, 2>("SS two accumulators", d_out);  run<128, 64, false, 4>("SS four accumulators", d_out);
  run<128, 64, true, 2>("TS two accumulators", d_out);   run<128, 64, true, 4>("TS four accumulators", d_out);
  run<128, 128, true, 2>("TS two accumulators", d_out);
  run<64, 64, false, 1>("SS M=64", d_out);  run<64, 128, false, 1>("SS M=64", d_out);  run<64, 256, false, 1>("SS M=64", d_out);
  run<64, 256, false, 1>("SS M=64 N=256", d_out);
  run_issuer<64, true, false>("production issuer loop", d_out);   run_issuer<64, true, true>("production issuer loop + mbarrier wait", d_out);
  run_issuer<128, true, false>("production issuer loop", d_out);  run_issuer<128, true, true>("production issuer loop + mbarrier wait", d_out);
  run_issuer<64, false, true>("production issuer loop + mbarrier wait", d_out);
  run_issuer<256, false, true>("production issuer loop + mbarrier wait", d_out);
  return 0;
}
