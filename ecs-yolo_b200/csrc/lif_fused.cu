// Fused ECS-LIF forward (mem_update, models/common.py:236-309) for C = 64 channels: ALL T timesteps of a spatial
// tile run inside one CTA, membrane potential and ECS trace never leave the SM.
//
//   tile      : 22 x 22 region of one image = output tile of (22 - 2(T-1))^2 pixels + a halo of T-1 pixels (the
//               ECS spread is a 3x3 stencil per step, so the valid region shrinks by one pixel per step).
//               Region pixels are operand rows r = ry*22 + rx (484 of 512 rows = 4 tensor-core M tiles).
//   state     : TENSOR MEMORY, pixel = lane, channel = column:  columns [k*128, k*128+64) of M tile k hold the
//               ECS accumulator, columns [k*128+64, k*128+128) the membrane potential (512 columns in total).
//   spread    : the depth-wise 3x3 and the point-wise 1x1 of `spread` are folded into ONE 3x3 convolution of the
//               binary spikes with W_eff[co][ci][tap] = pw[co][ci] * dw[ci][tap] (bf16, resident in shared memory,
//               73.7 KB) and a constant bconst = pw * b_dw + b_pw, i.e. exactly a spike convolution:
//               the step's spikes are expanded ONCE per pixel into a {0, 2.0} bf16 row (one shift + one mask per
//               channel pair, the packing of ecsy_pack_spike_conv_weight) of a shared-memory array in region-row
//               order, and the 9 taps are 9 UMMA descriptors whose start row is shifted by (ky-1)*22 + (kx-1)
//               rows (128-byte-swizzle base offset) -- no im2col, no per-tap expansion.
//               The tensor core ACCUMULATES onto the trace: TMEM holds kappa*E_{t-1}, the MMAs add conv(s_t), the
//               update reads E_t = D + bconst (e_t = alpha*E_t, models/common.py:263-267) and writes kappa*E_t back.
//   update    : 1024 threads, two per region pixel (32 channels each), 8 channels at a time: tcgen05.ld of the
//               trace and the membrane, one 32-byte sector of x_t from HBM, f = beta*tanh(e), mem' = (mem*decay)*
//               (1-s) + x + f (reference order), spike = mem' > thresh, tcgen05.st of the state.  Spike words of
//               the output tile go to HBM bit-packed; nothing else is written.
// HBM traffic per element-step: 4 B of x (x the halo overlap, mostly L2 hits) + 1/8 B of spikes, instead of
// 24 B for the per-timestep pipeline (lif.cu).  Fast mode only (bf16 folded weights).
//
// STATUS (round 1, B200): parity-green (tests/test_gpu_ops.py::test_lif_ecs_fused) but NOT yet faster than the
// pipeline: 64 ch @160^2, batch 64: 2.6-3.4 ms vs 1.94 ms.  Ablation (ECSY_LIF_DBG, profiles/r01_lif_fused_ablation.txt):
// without the x loads the kernel takes 1.29 ms, without x loads and MMAs 1.04 ms -- the row-per-lane x loads (each
// LDG.128 touches 32 different 128-byte lines) cost 60 % of the time.  Therefore opt-in (ECSY_LIF_FUSED=1);
// next: stage x through shared memory with TMA (needs the second spike buffer's 70 KB), or fp16 activations.
#include <stdlib.h>

#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

using namespace ecsy;

namespace {

constexpr int kRx = 22;                                 // region width in pixels (operand-row pitch)
constexpr int kRy = 21;                                 // region height: 462 rows, so the warps of rows 480..511 are free
constexpr int kRegion = kRx * kRy;                      // 462 operand rows in use
constexpr int kRows = 512;                              // 4 M tiles
constexpr int kGuard = 24;                              // zero rows before / after: taps reach +-23 rows
constexpr int kSpikeBytes = (kGuard + kRows + kGuard) * 128;   // 71680 = 70 KB per buffer
constexpr int kWBytes = 9 * 64 * 128;                   // 73728: nine [64 co x 64 ci] bf16 tap tiles
constexpr int kFThreads = 1024;
constexpr int kIssuerWarp = 31;                         // rows 480..511 hold no pixel: warp 31 issues the MMAs, warp 27 idles

struct FusedCtl {
  uint64_t w_full;
  uint64_t mma_done[4];   // the MMAs of one step on M tile k have completed (tcgen05.commit)
  uint64_t ready[4];      // the update warps of M tile k have stored the step's spike rows and state
  uint32_t tmem_base;
  uint32_t pad;
};

struct FusedArgs {
  const float* x;
  int64_t x_tstride;
  const float* in_scale;
  const float* in_shift;
  const float* bconst;
  uint32_t* spikes;
  int T, N, H, W;
  int tiles_y, tiles_x, ot_y, ot_x, halo;
  float thresh, decay, alpha, beta, kappa;
  int dbg;   // ablation switches for profiling (ECSY_LIF_DBG): 1 = no x loads, 2 = no MMAs (results are then wrong)
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Number of update warps that arrive on ready[k]: 8 per M tile, 6 on the last one (warps 27 and 31 hold no rows).
__device__ __forceinline__ uint32_t ready_count(int k) { return k == 3 ? 6u : 8u; }

__global__ void __launch_bounds__(kFThreads, 1)
k_lif_ecs_fused64(const __grid_constant__ CUtensorMap tm_w, const FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                                // 9 tap tiles
  uint8_t* s_smem = smem + kWBytes;                      // two spike-row buffers
  float* s_bconst = reinterpret_cast<float*>(s_smem + 2 * kSpikeBytes);
  float* s_scale = s_bconst + 64;
  float* s_shift = s_scale + 64;
  FusedCtl* ctl = reinterpret_cast<FusedCtl*>(s_shift + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int km = warp >> 3;          // M tile of this thread's pixel
  const int hf = (warp >> 2) & 1;    // channel half: channels [32*hf, 32*hf + 32)
  const int q = warp & 3;            // TMEM lane quarter (hardware: warp id % 4)
  const int row0 = km * 128 + q * 32;
  const int row = row0 + lane;
  const int ry = row / kRx, rx = row - ry * kRx;
  const bool in_region = row < kRegion;
  const bool affine = g.in_scale != nullptr;
  const bool update_warp = row0 < kRegion;    // false for warps 27 and 31

  // ---- one-time setup: zero the spike arrays (guard / padding rows stay zero), weights, TMEM, constants ----
  for (int i = threadIdx.x; i < 2 * kSpikeBytes / 16; i += kFThreads)
    reinterpret_cast<uint4*>(s_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x < 64) {
    s_bconst[threadIdx.x] = g.bconst[threadIdx.x];
    s_scale[threadIdx.x] = affine ? g.in_scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = affine ? g.in_shift[threadIdx.x] : 0.f;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_w);
    mbar_init(&ctl->w_full, 1);
    for (int k = 0; k < 4; ++k) {
      mbar_init(&ctl->mma_done[k], 1);
      mbar_init(&ctl->ready[k], ready_count(k));
    }
    mbar_fence_init();
    mbar_arrive_expect_tx(&ctl->w_full, (uint32_t)kWBytes);
    for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_smem + tap * 8192, &tm_w, &ctl->w_full, tap * 64, 0);
  }
  if (warp == 0) tmem_alloc<512>(&ctl->tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t s_base = smem_u32(s_smem);
  const int T = g.T, halo = g.halo;
  const int tiles_per_img = g.tiles_y * g.tiles_x;
  const int n_tiles = g.N * tiles_per_img;

  if (warp == kIssuerWarp) {
    // =============================== MMA issuer ===============================
    // Step t, M tile k: D_k (+)= sum over the 9 taps of A(rows 128k + shift .. +127 of the step's spike array) * W_eff[tap].
    // The array is written with the 128-byte swizzle of ABSOLUTE row addresses, so a descriptor may start at any row
    // (the hardware swizzle is a function of the address bits; measured: no base-offset field needed).
    // M tile k needs the spike rows of tiles k-1, k, k+1 (taps reach +-23 rows): the issuer waits for their `ready`.
    mbar_wait(&ctl->w_full, 0);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint64_t db0 = umma_desc_sw128(smem_u32(w_smem));
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int t = 0; t + 1 < T; ++t, ++use) {
        const uint32_t a_base = s_base + (uint32_t)(t & 1) * (uint32_t)kSpikeBytes + (uint32_t)kGuard * 128u;
        for (int k = 0; k < 4; ++k) {
          if (k == 0) { mbar_wait(&ctl->ready[0], use & 1u); mbar_wait(&ctl->ready[1], use & 1u); }
          else if (k < 3) mbar_wait(&ctl->ready[k + 1], use & 1u);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + (uint32_t)k * 128u;
            const uint64_t da0 = umma_desc_sw128(a_base + (uint32_t)(k * 128) * 128u);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int shift = (tap / 3 - 1) * kRx + (tap % 3 - 1);        // rows; 8 descriptor units (16 B) per row
              const uint64_t da = da0 + (uint64_t)(int64_t)(shift * 8);
              const uint64_t db = db0 + (uint64_t)(tap * 512);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (!(g.dbg & 2))
                  umma_f16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                           (t > 0 || tap > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&ctl->mma_done[k]);
          }
          __syncwarp();
        }
      }
    }
  } else if (update_warp) {
    // =============================== state update warps ===============================
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)km * 128u + (uint32_t)hf * 32u;
    const uint32_t my_srow = (uint32_t)(kGuard + row) * 128u;
    const uint32_t sw = (uint32_t)row & 7u;     // kGuard*128 is a multiple of 1024: swizzle phase = row & 7
    const uint32_t c_bconst = smem_u32(s_bconst) + (uint32_t)hf * 128u;
    const uint32_t c_scale = smem_u32(s_scale) + (uint32_t)hf * 128u;
    const uint32_t c_shift = smem_u32(s_shift) + (uint32_t)hf * 128u;
    const int64_t bits_tstride = (int64_t)g.N * g.H * g.W * 2;
    // pixel of this thread in a tile
    auto locate = [&](int tile, bool& inside, bool& in_out, int64_t& pix) {
      const int n = tile / tiles_per_img;
      const int trem = tile - n * tiles_per_img;
      const int ty = trem / g.tiles_x, tx = trem - ty * g.tiles_x;
      const int py = ty * g.ot_y - halo + ry, px = tx * g.ot_x - halo + rx;
      inside = in_region && py >= 0 && py < g.H && px >= 0 && px < g.W;
      in_out = inside && ry >= halo && ry < halo + g.ot_y && rx >= halo && rx < halo + g.ot_x;
      pix = inside ? ((int64_t)n * g.H + py) * g.W + px : 0;
    };
    // The input current of a step is requested a FULL STEP AHEAD into registers (32 channels = one 128-byte line per
    // thread) and stays in flight across the wait on the tensor core, where nothing else is live.
    float4 xr[8];
    auto load_x = [&](const float* p, bool ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        xr[i] = (ok && !(g.dbg & 1)) ? ldg_stream(reinterpret_cast<const float4*>(p) + i) : make_float4(0.1f, 0.6f, 0.f, 0.7f);
    };
    bool inside = false, in_out = false;
    int64_t pix = 0;
    if ((int)blockIdx.x < n_tiles) {
      locate((int)blockIdx.x, inside, in_out, pix);
      load_x(g.x + pix * 64 + hf * 32, inside);
    }
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const float* xp = g.x + pix * 64 + hf * 32;
      uint32_t* bp = g.spikes + pix * 2 + hf;
      const bool cur_inside = inside, cur_in_out = in_out;
      // every MMA of the previous tile has completed before its spike rows are overwritten (commits complete in
      // issue order and this warp has seen its own M tile's last phase, so barrier 3 is at most one phase behind)
      if (tile_it > 0) mbar_wait(&ctl->mma_done[3], (tile_it * (uint32_t)(T - 1) - 1u) & 1u);
      uint32_t s_prev = 0;

      for (int t = 0; t < T; ++t) {
        // rows of region lines [t, kRy - t) carry valid state at step t; a warp whose 32 rows lie outside skips the work
        const bool active = (row0 + 31 >= t * kRx) && (row0 < (kRy - t) * kRx);
        const bool next_active = (row0 + 31 >= (t + 1) * kRx) && (row0 < (kRy - t - 1) * kRx);
        uint32_t bits = 0;
        // Handshake of EVERY update warp (also the ones that skip the work): one `ready` arrival per completed
        // `mma_done` phase, so no warp can run a phase ahead of the barriers it shares.
        if (t > 0) {
          mbar_wait(&ctl->mma_done[km], (tile_it * (uint32_t)(T - 1) + (uint32_t)(t - 1)) & 1u);
          tc_fence_after_sync();
        }
        if (active) {
          // pass A: p = (mem * decay) * (1 - s) + x   (reference order, models/common.py:306-309), in place of x
          float pv[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float xv[8] = {xr[2 * j].x, xr[2 * j].y, xr[2 * j].z, xr[2 * j].w,
                           xr[2 * j + 1].x, xr[2 * j + 1].y, xr[2 * j + 1].z, xr[2 * j + 1].w};
            if (affine) {
              const float4 s0 = lds128f(c_scale + j * 32), s1 = lds128f(c_scale + j * 32 + 16);
              const float4 h0 = lds128f(c_shift + j * 32), h1 = lds128f(c_shift + j * 32 + 16);
              const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) xv[i] = add_rn(mul_rn(xv[i], sc[i]), sh[i]);
            }
            if (t == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pv[8 * j + i] = xv[i];
            } else {
              uint32_t mo[8];
              tmem_ld_32x8(t_lane + 64 + 8 * j, mo);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float keep = ((s_prev >> (8 * j + i)) & 1u) ? 0.f : 1.f;
                pv[8 * j + i] = add_rn(mul_rn(mul_rn(__uint_as_float(mo[i]), g.decay), keep), xv[i]);
              }
            }
          }
          // pass B: mem' = p + beta*tanh(alpha*E), E = D + bconst; kappa*E back to the accumulator; threshold
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (t > 0) {
              uint32_t dv[8];
              tmem_ld_32x8(t_lane + 8 * j, dv);
              const float4 b0 = lds128f(c_bconst + j * 32), b1 = lds128f(c_bconst + j * 32 + 16);
              const float bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float E = add_rn(__uint_as_float(dv[i]), bc[i]);
                dv[i] = __float_as_uint(mul_rn(g.kappa, E));
                pv[8 * j + i] = add_rn(pv[8 * j + i], mul_rn(g.beta, tanh_approx(mul_rn(g.alpha, E))));
              }
              if (t + 1 < T) tmem_st_32x8(t_lane + 8 * j, dv);   // this step's MMAs accumulate conv(s_t) onto kappa*E_{t-1}
            }
            if (t + 1 < T) {
              const uint32_t mv[8] = {__float_as_uint(pv[8 * j]), __float_as_uint(pv[8 * j + 1]),
                                      __float_as_uint(pv[8 * j + 2]), __float_as_uint(pv[8 * j + 3]),
                                      __float_as_uint(pv[8 * j + 4]), __float_as_uint(pv[8 * j + 5]),
                                      __float_as_uint(pv[8 * j + 6]), __float_as_uint(pv[8 * j + 7])};
              tmem_st_32x8(t_lane + 64 + 8 * j, mv);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) bits |= (pv[8 * j + i] > g.thresh ? 1u : 0u) << (8 * j + i);
          }
          if (!cur_inside) bits = 0;   // zero padding of the spread convolution / unused rows
          if (cur_in_out) bp[(int64_t)t * bits_tstride] = bits;
        }
        // the next input current goes in flight now (the registers of this step's are free): x of step t+1, or of
        // step 0 of this CTA's next tile; it lands while the warp stores its spikes and waits for the tensor core
        if (t + 1 < T) {
          if (next_active) load_x(xp + (int64_t)(t + 1) * g.x_tstride, cur_inside);
        } else if (tile + (int)gridDim.x < n_tiles) {
          locate(tile + (int)gridDim.x, inside, in_out, pix);
          load_x(g.x + pix * 64 + hf * 32, inside);
        }
        if (t + 1 < T) {
          if (active) {
            // {0, 2.0} bf16 row of this thread's 32 channels: word jj = channels (jj, jj + 16) of the half
            const uint32_t dst = s_base + (uint32_t)(t & 1) * (uint32_t)kSpikeBytes + my_srow;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint32_t w4[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int jj = c4 * 4 + u;
                w4[u] = (jj < 15 ? (bits << (14 - jj < 0 ? 0 : 14 - jj)) : (bits >> 1)) & 0x40004000u;
              }
              sts128u(dst + ((((uint32_t)(hf * 4 + c4)) ^ sw) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
            s_prev = bits;
            tmem_st_wait();
            fence_proxy_async_smem();
            tc_fence_before_sync();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->ready[km]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

extern "C" int ecsy_lif_ecs_fused_supported(int T, int C) { return (C == 64 && T >= 2 && T <= 8) ? 1 : 0; }

extern "C" int ecsy_lif_ecs_fused_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                                      const void* w_eff_ts, const float* bconst, uint32_t* spikes, int T, int64_t N, int H,
                                      int W, int C, float thresh, float decay, float alpha, float beta, float kappa,
                                      void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(x && spikes && w_eff_ts && bconst && N > 0 && H > 0 && W > 0, "lif_ecs_fused_fwd: bad arguments");
  ECSY_CHECK_ARG(ecsy_lif_ecs_fused_supported(T, C), "lif_ecs_fused_fwd: unsupported T=%d / C=%d (C == 64, 2 <= T <= 8)", T, C);
  ECSY_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "lif_ecs_fused_fwd: scale/shift pair");
  FusedArgs g{};
  g.x = x; g.x_tstride = x_tstride; g.in_scale = in_scale; g.in_shift = in_shift; g.bconst = bconst; g.spikes = spikes;
  g.T = T; g.N = (int)N; g.H = H; g.W = W;
  g.halo = T - 1;
  g.ot_y = kRy - 2 * g.halo;
  g.ot_x = kRx - 2 * g.halo;
  g.tiles_y = (H + g.ot_y - 1) / g.ot_y;
  g.tiles_x = (W + g.ot_x - 1) / g.ot_x;
  g.thresh = thresh; g.decay = decay; g.alpha = alpha; g.beta = beta; g.kappa = kappa;
  g.dbg = getenv("ECSY_LIF_DBG") ? atoi(getenv("ECSY_LIF_DBG")) : 0;
  ECSY_CHECK_ARG(N * g.tiles_y * g.tiles_x < (1LL << 31), "lif_ecs_fused_fwd: too many tiles");
  CUtensorMap tw;
  int rc = ecsy_tensor_map_bf16(w_eff_ts, 64, 9 * 64, 64, &tw);
  if (rc) return rc;
  const int smem = 1024 + kWBytes + 2 * kSpikeBytes + 3 * 64 * 4 + (int)sizeof(FusedCtl) + 64;
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(k_lif_ecs_fused64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  int64_t tiles = N * g.tiles_y * g.tiles_x;
  int grid = ecsy_num_sms();
  if (tiles < grid) grid = (int)tiles;
  k_lif_ecs_fused64<<<grid, kFThreads, smem, st>>>(tw, g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
