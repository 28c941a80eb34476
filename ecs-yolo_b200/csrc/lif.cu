// ECS-LIF forward (mem_update, models/common.py:236-309) as a per-timestep pipeline:
//   step 0        : k_lif_first            mem_0 = x_0, s_0 = mem_0 > thresh          (streaming)
//   step t -> t+1 : k_spread_dw            A = dw3x3(s_t) + b as bf16 hi[/lo]          (streaming)
//                   k_umma_gemm (tcgen05)  S = A * Wpw^T                               (tensor cores)
//                   k_ecs_step             e_t = alpha*(S + b) + kappa*e_{t-1};  f_t = beta*tanh(e_t);
//                                          mem_{t+1} = mem_t*decay*(1-s_t) + x_{t+1} + f_t; s_{t+1}   (streaming)
// Spikes leave as bit-packed words; membrane / ECS state live in the caller-provided workspace.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" size_t ecsy_lif_ecs_ws_bytes(int T, int64_t N, int H, int W, int C, int splits) {
  const size_t mc = static_cast<size_t>(N) * H * W * C;
  size_t b = 512;
  b += 3 * align256(mc * 4);                    // mem, ecs, spread
  b += static_cast<size_t>(splits) * align256(mc * 2);  // dw output planes
  (void)T;
  return b;
}

extern "C" int ecsy_lif_ecs_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                                const float* dw_w, const float* dw_b, const void* pw_packed, const float* pw_b,
                                int splits, uint32_t* spikes, float* mem_save, float* ecs_save, int T, int64_t N, int H, int W, int C,
                                float thresh, float decay, float alpha, float beta, float kappa, void* ws,
                                size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(x && spikes && T >= 1 && N > 0 && H > 0 && W > 0, "lif_ecs_fwd: bad arguments");
  ECSY_CHECK_ARG(C % 64 == 0, "lif_ecs_fwd: C=%d must be a multiple of 64", C);
  ECSY_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "lif_ecs_fwd: scale/shift pair");
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "lif_ecs_fwd: splits must be 1 or 2");
  ECSY_CHECK_ARG(T == 1 || (dw_w && dw_b && pw_packed && pw_b), "lif_ecs_fwd: spread weights missing");
  const int64_t M = N * H * W;
  ECSY_CHECK_ARG(N * H < (1LL << 31) && M * (C / 8) < (1LL << 40), "lif_ecs_fwd: tensor too large");
  const size_t mc = static_cast<size_t>(M) * C;
  const size_t need = ecsy_lif_ecs_ws_bytes(T, N, H, W, C, splits);
  if (T > 1 && (ws == nullptr || ws_bytes < need)) {
    ecsy_set_error("lif_ecs_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  float* mem = reinterpret_cast<float*>(p); p += align256(mc * 4);
  float* ecs = reinterpret_cast<float*>(p); p += align256(mc * 4);
  float* spread = reinterpret_cast<float*>(p); p += align256(mc * 4);
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(p); p += align256(mc * 2);
  __nv_bfloat16* a_lo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(p) : nullptr;
  const int64_t words = M * (C / 32);

  // mem_0 == x_0 when no affine is pending on the input current (inference: tdBN is folded into the producing conv), so
  // step 0 only thresholds and the first ECS step reads its membrane straight from x_0: no 4 B/elem copy.
  const bool alias0 = in_scale == nullptr && mem_save == nullptr;
  float* mem0 = mem_save ? mem_save : ((T > 1 && !alias0) ? mem : nullptr);
  int rc = ecsy_launch_lif_first(x, in_scale, in_shift, mem0, spikes, M, C, thresh, st);
  if (rc) return rc;
  // Image chunks inside a step (inference path, opt-in): the depth-wise output A (bf16) and the point-wise output
  // `spread` are produced and consumed within one chunk, in chunk-sized buffers reused for every chunk, so these 8 B per
  // element-step can stay in the 126 MB L2.  MEASURED SLOWER (resnet34, batch 64: lif 32.7 ms unchunked, 37.1 / 42.7 /
  // 51.1 ms at 96 / 48 / 24 MB chunks): the extra launches' fixed costs (persistent GEMM set-up, grid tails) outweigh the
  // saved HBM traffic, so the default is off.  ECSY_LIF_CHUNK_MB = A + spread bytes per chunk (0 = whole batch per launch).
  static const int chunk_mb = getenv("ECSY_LIF_CHUNK_MB") ? atoi(getenv("ECSY_LIF_CHUNK_MB")) : 0;
  const int half_state = splits == 1 ? 1 : 0;  // fast mode: spread output and ECS trace stored as fp16
  const size_t per_img = static_cast<size_t>(H) * W * C;
  int64_t n_chunk = N;
  static const bool fused_dw = getenv("ECSY_FUSED_DW") != nullptr && getenv("ECSY_FUSED_DW")[0] == '1';
  if (chunk_mb > 0 && mem_save == nullptr && !fused_dw) {
    const size_t bytes_per_img = per_img * ((half_state ? 2 : 4) + 2 * (size_t)splits);
    n_chunk = static_cast<int64_t>((static_cast<size_t>(chunk_mb) << 20) / bytes_per_img);
    if (n_chunk < 1) n_chunk = 1;
    if (n_chunk > N) n_chunk = N;
  }
  for (int t = 0; t + 1 < T; ++t) {
    const bool more = t + 2 < T;
    for (int64_t n0 = 0; n0 < N; n0 += n_chunk) {
      const int64_t nc = (N - n0 < n_chunk) ? (N - n0) : n_chunk;
      const int64_t Mc = nc * H * W;
      const size_t eo = static_cast<size_t>(n0) * per_img;          // element offset of the chunk in full-batch tensors
      const int64_t wo = n0 * H * W * (C / 32);                     // word offset in a step's spike tensor
      // chunked: A / spread live at the start of their buffers for every chunk (reuse keeps them hot in L2)
      float* spread_c = spread;
      __nv_bfloat16* a_hi_c = a_hi;
      __nv_bfloat16* a_lo_c = a_lo;
      // fast precision, opt-in (ECSY_ECS_GEMM=1): the ECS step as the EPILOGUE of the point-wise GEMM (k_ecs_gemm: `spread`
      // never reaches HBM, -4 B per element-step, same spikes bit for bit -- tests/test_gpu_ops.py).  MEASURED SLOWER
      // (resnet34, batch 64: lif_ecs 29.8 ms two kernels, 32.5 ms fused): the step's x / membrane / trace loads land in
      // the registers of 8 epilogue warps (45 KB in flight per SM at 168 registers per thread), half of what the
      // stand-alone streaming kernel keeps in flight with 2048 threads per SM, so the fused epilogue runs below the HBM rate
      // and the saved bytes do not pay for it.
      static const bool ecs_gemm_env = getenv("ECSY_ECS_GEMM") != nullptr && getenv("ECSY_ECS_GEMM")[0] == '1';
      const bool ecs_gemm = ecs_gemm_env && half_state && !fused_dw && Mc >= 128;
      if (fused_dw) {
        // experimental: depth-wise spread computed by the GEMM's producer warps (no A round trip through HBM);
        // measured slower than the two-kernel path at 1 CTA/SM (producer address math + 8 warps of ALU work)
        rc = ecsy_umma_dw_gemm(spikes + t * words, dw_w, dw_b, pw_packed, splits, spread, half_state, (int)N, H, W, C, st);
        if (rc) return rc;
      } else {
        rc = ecsy_launch_spread_dw(spikes + t * words + wo, dw_w, dw_b, a_hi_c, a_lo_c, (int)nc, H, W, C, st);
        if (rc) return rc;
        if (!ecs_gemm) {
          rc = ecsy_umma_dense(a_hi_c, a_lo_c, Mc, C, pw_packed, splits, spread_c, C, nullptr, nullptr, nullptr, 0, st, half_state);
          if (rc) return rc;
        }
      }
      EcsStep s{};
      s.spread = spread_c; s.pw_b = pw_b;
      s.x_next = x + (t + 1) * x_tstride + eo;
      s.in_scale = in_scale; s.in_shift = in_shift;
      if (mem_save) {  // keep every membrane for the backward pass: read step t, write step t+1
        s.mem_in = mem_save + (size_t)t * mc + eo;
        s.mem_out = mem_save + (size_t)(t + 1) * mc + eo;
      } else {
        s.mem_in = (t == 0 && alias0) ? x + eo : mem + eo;
        s.mem_out = more ? mem + eo : nullptr;
      }
      // the ECS trace is fp16 in fast mode: its element offset is in units of the stored type
      s.ecs = half_state ? reinterpret_cast<float*>(reinterpret_cast<__half*>(ecs) + eo) : ecs + eo;
      s.store_ecs = more ? 1 : 0;
      s.ecs_save = ecs_save ? ecs_save + (size_t)t * mc + eo : nullptr;
      s.bits_t = spikes + t * words + wo;
      s.bits_next = spikes + (t + 1) * words + wo;
      s.first = (t == 0) ? 1 : 0;
      s.half_state = half_state;
      s.thresh = thresh; s.decay = decay; s.alpha = alpha; s.beta = beta; s.kappa = kappa;
      rc = ecs_gemm ? ecsy_umma_ecs_step(a_hi_c, Mc, C, pw_packed, s, st) : ecsy_launch_ecs_step(s, Mc, C, st);
      if (rc) return rc;
    }
  }
  return ECSY_OK;
}
