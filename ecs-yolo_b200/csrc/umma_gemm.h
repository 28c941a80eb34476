// Internal (non-ABI) interface of the tcgen05 GEMM engine, shared between translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

int ecsy_pick_bn(int cout, int splits);
int ecsy_tensor_map_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMap* out);
int ecsy_umma_spike_conv(const uint32_t* bits, const void* w_packed, int splits, float* out, const float* scale,
                         const float* shift, const float* residual, int64_t res_imgs, int imgs, int H, int W, int Cin,
                         int Cout, int k, int stride, int pad, cudaStream_t st, int ts = 0);
int ecsy_pick_bn_ts(int cout);
int ecsy_umma_dense(const void* a_hi, const void* a_lo, int64_t M, int K, const void* w_packed, int splits, float* out,
                    int Cout, const float* scale, const float* shift, const float* residual, int64_t res_rows,
                    cudaStream_t st, int out_half = 0);

// elementwise.cu launchers used by lif.cu
int ecsy_launch_lif_first(const float* x, const float* scale, const float* shift, float* mem, uint32_t* bits,
                          int64_t pixels, int C, float thresh, cudaStream_t st);
struct EcsStep {
  const void* spread;      // [M][C] pw(dw(s_t)) without bias (GEMM output; fp32, or fp16 when `half_state`)
  const float* pw_b;       // [C]
  const float* x_next;     // [M][C] input current of step t+1
  const float* in_scale;   // optional folded tdBN on x
  const float* in_shift;
  const float* mem_in;     // membrane of step t
  float* mem_out;          // membrane of step t+1 (may alias mem_in; NULL = last step, not stored)
  void* ecs;               // e_{t-1} in (unless first), e_t out (if store_ecs); fp32 or fp16 (`half_state`)
  float* ecs_save;         // optional copy of e_t kept for the backward pass
  const uint32_t* bits_t;  // spikes of step t
  uint32_t* bits_next;     // spikes of step t+1
  int first, store_ecs;
  int half_state;          // fast mode: spread and the ECS trace are stored as fp16 (8 B/elem-step less traffic)
  float thresh, decay, alpha, beta, kappa;
};
int ecsy_launch_ecs_step(const EcsStep& s, int64_t pixels, int C, cudaStream_t st);
int ecsy_umma_ecs_step(const void* a_hi, int64_t M, int C, const void* pw_packed, const EcsStep& s, cudaStream_t st);
int ecsy_launch_spread_dw(const uint32_t* bits, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi,
                          __nv_bfloat16* a_lo, int N, int H, int W, int C, cudaStream_t st);

int ecsy_umma_xty(const void* p_hi, const void* p_lo, const void* q_hi, const void* q_lo, int64_t rows, int Ca, int Cb,
                  float alpha, float* out, cudaStream_t st);
int ecsy_umma_conv_bf16(const void* a_hi, const void* a_lo, const void* w_packed, int splits, float* out,
                        const float* scale, const float* shift, const float* residual, int64_t res_imgs, int imgs, int H,
                        int W, int Cin, int Cout, int k, int pad, cudaStream_t st);
int ecsy_umma_conv_gather(const float* x, int64_t x_imgs, const void* w_packed, float* out, const float* scale,
                          const float* shift, int imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                          cudaStream_t st);
int ecsy_launch_f32_to_bf16(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t st);
int ecsy_umma_spike_wgrad(const void* gy_hi, const void* gy_lo, const uint32_t* bits, float* dw, int imgs, int H, int W,
                          int Cin, int Cout, int k, int stride, int pad, cudaStream_t st);
int ecsy_umma_dw_gemm(const uint32_t* bits, const float* dw_w, const float* dw_b, const void* pw_packed, int splits,
                      float* out, int out_half, int N, int H, int W, int C, cudaStream_t st);
