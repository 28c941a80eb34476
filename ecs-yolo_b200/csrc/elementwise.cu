// Streaming (HBM-bound) kernels of the spiking hot path: layout conversion, spike bit-packing,
// the first LIF step, tdBN statistics, affine+add, max-pool, nearest upsample / concat, T-fusion
// and the Stack-A / Stack-B head decodes.  All activations are NHWC fp32 ("[imgs][H][W][C]"),
// spikes are one bit per channel packed along C ("[imgs][H][W][C/32]" uint32, bit c&31 of word c>>5).
#include <cuda_fp16.h>
#include <algorithm>
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work, int per_block, int max_blocks) {
  int64_t g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC (edge of the model only).  32x32 smem transpose per image over (C, HW).
// ------------------------------------------------------------------------------------------
__global__ void k_transpose_inner(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  // per image: in [R][Cc] -> out [Cc][R]
  __shared__ float tile[32][33];
  const int64_t img = blockIdx.z;
  const float* src = in + img * (int64_t)R * Cc;
  float* dst = out + img * (int64_t)R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = src[(int64_t)r * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) dst[(int64_t)c * R + r] = tile[threadIdx.x][i];
  }
}

// Few channels (the RGB / event-frame input, C <= 4): one thread per pixel reads its C plane values (coalesced
// across the warp) and writes C consecutive floats -- the 32x32 tile transpose wastes 29 of 32 tile rows here.
template <int C>
__global__ void k_nchw_to_nhwc_small(const float* __restrict__ in, float* __restrict__ out, int64_t imgs, int64_t hw) {
  const int64_t total = imgs * hw;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t img = i / hw, p = i - img * hw;
    const float* src = in + img * C * hw + p;
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = ecsy::ldg_stream_f(src + c * hw);
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * C + c] = v[c];
  }
}

// ------------------------------------------------------------------------------------------
// spikes pack / unpack.  One warp packs 32 channels of one pixel per ballot.
// ------------------------------------------------------------------------------------------
__global__ void k_pack(const float* __restrict__ x, uint32_t* __restrict__ bits, int64_t words, float thresh) {
  // word w covers elements [32w, 32w+32); lane l reads element 32w+l
  const int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < words; w += nwarps) {
    float v = x[w * 32 + lane];
    uint32_t b = __ballot_sync(0xffffffffu, v > thresh);
    if (lane == 0) bits[w] = b;
  }
}

__global__ void k_unpack(const uint32_t* __restrict__ bits, float* __restrict__ x, int64_t words) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < words; w += nwarps) {
    uint32_t b = bits[w];
    x[w * 32 + lane] = (b >> lane) & 1u ? 1.0f : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------
// LIF step 0 (models/common.py:271,275): mem_0 = x_0 (+ optional per-channel affine = folded tdBN),
// spike_0 = mem_0 > thresh.  Each thread owns 4 channels of one pixel; 8 lanes assemble one word.
// ------------------------------------------------------------------------------------------
__global__ void k_lif_first(const float* __restrict__ x, const float* __restrict__ scale,
                            const float* __restrict__ shift, float* __restrict__ mem,
                            uint32_t* __restrict__ bits, int64_t n4, int C, float thresh) {
  // n4 = pixels*C/4 is a multiple of 8 (C % 32 == 0), so the 8 lanes that assemble one word are
  // always active together; the shuffles below only exchange data inside such 8-lane groups.
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4_round = (n4 + 31) & ~int64_t(31);
  const bool small = n4_round < (int64_t(1) << 32);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4_round; i += stride) {
    const bool ok = i < n4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = ecsy::ldg_stream(reinterpret_cast<const float4*>(x) + i);
    if (scale != nullptr) {
      int c = static_cast<int>(ecsy::mod_u(i, (uint32_t)(C >> 2), small) * 4u);
      float4 s = *reinterpret_cast<const float4*>(scale + c);
      float4 b = *reinterpret_cast<const float4*>(shift + c);
      v.x = ecsy::add_rn(ecsy::mul_rn(v.x, s.x), b.x);
      v.y = ecsy::add_rn(ecsy::mul_rn(v.y, s.y), b.y);
      v.z = ecsy::add_rn(ecsy::mul_rn(v.z, s.z), b.z);
      v.w = ecsy::add_rn(ecsy::mul_rn(v.w, s.w), b.w);
    }
    if (ok && mem != nullptr) reinterpret_cast<float4*>(mem)[i] = v;
    uint32_t nib = (v.x > thresh ? 1u : 0u) | (v.y > thresh ? 2u : 0u) | (v.z > thresh ? 4u : 0u) |
                   (v.w > thresh ? 8u : 0u);
    const int sub = threadIdx.x & 7;
    uint32_t w = nib << (4 * sub);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    if (ok && sub == 0) bits[i >> 3] = w;
  }
}

// ECS-LIF step t -> t+1 (models/common.py:263-281, 306-309), streaming half: consumes the point-wise
// spread GEMM output and produces e_t, mem_{t+1}, s_{t+1}:
//   e_t     = alpha*(spread + b) + kappa*e_{t-1}
//   f_t     = beta*tanh(e_t)
//   mem_t+1 = mem_t*decay*(1 - s_t) + x_{t+1} + f_t ;  s_{t+1} = mem_{t+1} > thresh
// Every product / sum is rounded separately, in the reference's evaluation order.
__device__ __forceinline__ float4 ld4h(const void* base, int64_t i) {
  const uint2 r = reinterpret_cast<const uint2*>(base)[i];
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4h(void* base, int64_t i, float a, float b, float c, float d) {
  __half2 h01 = __floats2half2_rn(a, b), h23 = __floats2half2_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&h01);
  pk.y = *reinterpret_cast<uint32_t*>(&h23);
  reinterpret_cast<uint2*>(base)[i] = pk;
}

// Fast-precision tanh: the hardware approximation (one MUFU op, absolute error ~2^-11).  Measured on the BASELINE plans
// (tests/test_gpu_baseline_cfgs.py, fast precision, every neuron fed the oracle's input): replacing it by the 1e-7-accurate
// 1 - 2 / (exp(2v) + 1) (two MUFU ops + three FP32 ops) did not change the spike agreement beyond the fifth digit but made
// the ECS step kernels, which are closer to the MUFU / issue limit than to the HBM limit, 20 % slower (lif_ecs 32.6 ->
// 40.1 ms per resnet34 batch-64 step).  ECSY_ACCURATE_TANH selects the accurate form at compile time.
__device__ __forceinline__ float tanh_fast(float v) {
#ifdef ECSY_ACCURATE_TANH
  const float y = __expf(2.f * v);
  return 1.f - __fdividef(2.f, y + 1.f);
#else
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#endif
}

// Parity-precision tanh: 1 - 2 / (exp(2v) + 1) with ex2.approx / rcp.approx -- ABSOLUTE error ~1e-7, which is what the
// membrane sees (f = beta * tanh(e)); five instructions against ~25 for tanhf.  Saturates correctly (exp -> inf: 1, -> 0: -1).
__device__ __forceinline__ float tanh_acc(float v) {
  const float y = __expf(2.f * v);
  return 1.f - __fdividef(2.f, y + 1.f);
}

template <bool HALF>
__global__ void k_ecs_step(const EcsStep p, int64_t n4, int64_t n4_round, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool small = n4_round < (int64_t(1) << 32);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4_round; i += stride) {
    const bool ok = i < n4;
    uint32_t nib = 0;
    const int sub = threadIdx.x & 7;
    if (ok) {
      const int c = static_cast<int>(ecsy::mod_u(i, (uint32_t)(C >> 2), small) * 4u);
      float4 xv = ecsy::ldg_stream(reinterpret_cast<const float4*>(p.x_next) + i);
      if (p.in_scale != nullptr) {
        const float4 s = *reinterpret_cast<const float4*>(p.in_scale + c);
        const float4 b = *reinterpret_cast<const float4*>(p.in_shift + c);
        xv.x = ecsy::add_rn(ecsy::mul_rn(xv.x, s.x), b.x);
        xv.y = ecsy::add_rn(ecsy::mul_rn(xv.y, s.y), b.y);
        xv.z = ecsy::add_rn(ecsy::mul_rn(xv.z, s.z), b.z);
        xv.w = ecsy::add_rn(ecsy::mul_rn(xv.w, s.w), b.w);
      }
      const float4 sv = HALF ? ld4h(p.spread, i) : ecsy::ldg_stream(reinterpret_cast<const float4*>(p.spread) + i);
      const float4 mv = reinterpret_cast<const float4*>(p.mem_in)[i];
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!p.first) ev = HALF ? ld4h(p.ecs, i) : reinterpret_cast<const float4*>(p.ecs)[i];
      const float4 pb = *reinterpret_cast<const float4*>(p.pw_b + c);
      const uint32_t pw = p.bits_t[i >> 3] >> (4 * sub);
      const float xin[4] = {xv.x, xv.y, xv.z, xv.w}, sp[4] = {sv.x, sv.y, sv.z, sv.w};
      const float mo[4] = {mv.x, mv.y, mv.z, mv.w}, eo[4] = {ev.x, ev.y, ev.z, ev.w};
      const float bb[4] = {pb.x, pb.y, pb.z, pb.w};
      float mn[4], en[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s_acc = ecsy::add_rn(sp[k], bb[k]);
        en[k] = ecsy::add_rn(ecsy::mul_rn(p.alpha, s_acc), ecsy::mul_rn(p.kappa, eo[k]));
        // fast mode (fp16 state): hardware tanh approximation (2^-11), parity mode: ex2 / rcp form (1e-7 absolute)
        const float fecs = ecsy::mul_rn(p.beta, HALF ? tanh_fast(en[k]) : tanh_acc(en[k]));
        const float keep = ((pw >> k) & 1u) ? 0.f : 1.f;
        mn[k] = ecsy::add_rn(ecsy::add_rn(ecsy::mul_rn(ecsy::mul_rn(mo[k], p.decay), keep), xin[k]), fecs);
        nib |= (mn[k] > p.thresh ? 1u : 0u) << k;
      }
      if (p.mem_out != nullptr) reinterpret_cast<float4*>(p.mem_out)[i] = make_float4(mn[0], mn[1], mn[2], mn[3]);
      if (p.store_ecs) {
        if (HALF) st4h(p.ecs, i, en[0], en[1], en[2], en[3]);
        else reinterpret_cast<float4*>(p.ecs)[i] = make_float4(en[0], en[1], en[2], en[3]);
      }
      if (p.ecs_save != nullptr) reinterpret_cast<float4*>(p.ecs_save)[i] = make_float4(en[0], en[1], en[2], en[3]);
    }
    uint32_t w = nib << (4 * sub);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    if (ok && sub == 0) p.bits_next[i >> 3] = w;
  }
}

// ------------------------------------------------------------------------------------------
// ECS spread, depthwise half (models/common.py:289-294 spread[0]): A[p][c] = b[c] + sum_tap
// bit(p+tap, c) * w[tap][c], written as bf16 hi (+ lo residual) rows for the point-wise GEMM.
// A thread owns 8 channels (one spike byte per pixel) with its 72 weights + 8 biases in registers and
// slides a 3x3 byte window along a run of kDwRun pixels of one image row (3 byte loads per pixel);
// zero padding = zero bytes.  fp32 accumulation in tap order (row-major 3x3).
// ------------------------------------------------------------------------------------------
constexpr int kDwRun = 16;

template <bool LO>
__global__ void __launch_bounds__(256)
k_spread_dw(const uint32_t* __restrict__ bits, const float* __restrict__ dw_w /*[9][C]*/,
            const float* __restrict__ dw_b, __nv_bfloat16* __restrict__ a_hi, __nv_bfloat16* __restrict__ a_lo,
            int N, int H, int W, int C) {
  const int c8 = C >> 3;
  const int nseg = (W + kDwRun - 1) / kDwRun;
  const int64_t items = (int64_t)N * H * nseg * c8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // a multiple of c8 (host guarantees)
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int cg = static_cast<int>(i0 % c8);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(bits) + cg;
  float w[72], bias[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = *reinterpret_cast<const float4*>(dw_w + t * C + cg * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(dw_w + t * C + cg * 8 + 4);
    w[t * 8 + 0] = w0.x; w[t * 8 + 1] = w0.y; w[t * 8 + 2] = w0.z; w[t * 8 + 3] = w0.w;
    w[t * 8 + 4] = w1.x; w[t * 8 + 5] = w1.y; w[t * 8 + 6] = w1.z; w[t * 8 + 7] = w1.w;
  }
  {
    const float4 b0 = *reinterpret_cast<const float4*>(dw_b + cg * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(dw_b + cg * 8 + 4);
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  for (int64_t i = i0; i < items; i += stride) {
    int64_t run = i / c8;
    const int seg = static_cast<int>(run % nseg);
    run /= nseg;
    const int h = static_cast<int>(run % H);
    const int64_t img = run / H;
    const int x0 = seg * kDwRun;
    const int x1 = min(x0 + kDwRun, W);
    const uint8_t* rowp[3];
    bool rok[3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      rok[ky] = hh >= 0 && hh < H;
      rowp[ky] = bytes + ((img * H + (rok[ky] ? hh : h)) * W) * (int64_t)c8;
    }
    uint32_t win[3][3];  // [ky][kx]: columns x-1, x, x+1
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      win[ky][0] = 0;
      win[ky][1] = (rok[ky] && x0 - 1 >= 0) ? rowp[ky][(int64_t)(x0 - 1) * c8] : 0u;
      win[ky][2] = rok[ky] ? rowp[ky][(int64_t)x0 * c8] : 0u;
    }
    int64_t o = ((img * H + h) * W + x0) * (int64_t)c8 + cg;
    const uint8_t* nextp[3];   // byte of column x + 1 in the three rows: running pointers instead of a 64-bit multiply per load
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) nextp[ky] = rowp[ky] + (int64_t)(x0 + 1) * c8;
    for (int x = x0; x < x1; ++x) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[ky][0] = win[ky][1];
        win[ky][1] = win[ky][2];
        win[ky][2] = (rok[ky] && x + 1 < W) ? *nextp[ky] : 0u;
        nextp[ky] += c8;
      }
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = bias[k];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t m = win[ky][kx];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (m & (1u << k)) acc[k] += w[(ky * 3 + kx) * 8 + k];
        }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 hp = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);   // one packed cvt
        hi[q] = *reinterpret_cast<const uint32_t*>(&hp);
        if (LO) {   // the residual plane exists in parity precision only: fast precision used to compute and drop it
          const float2 hf = __bfloat1622float2(hp);
          const __nv_bfloat162 lp = __floats2bfloat162_rn(acc[2 * q] - hf.x, acc[2 * q + 1] - hf.y);
          lo[q] = *reinterpret_cast<const uint32_t*>(&lp);
        }
      }
      reinterpret_cast<uint4*>(a_hi)[o] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (LO) reinterpret_cast<uint4*>(a_lo)[o] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      o += c8;
    }
  }
}

// Version 2: the bits of a tile (3 rows x (PX + 2) pixels x C/8 bytes) are staged in shared memory with
// coalesced 8-byte loads, so the per-pixel byte reads of the sliding window become LDS.U8 (v1 read them with
// one-byte global loads one pixel ahead and was latency bound: 32 % issue utilisation, 0.8 TB/s).  A block owns
// `segs` runs of R consecutive pixels of one image row for all C/8 channel groups (blockDim = segs * C/8);
// same accumulation order as v1, bit-identical output.
template <int R, bool LO>
__global__ void __launch_bounds__(256, 2)
k_spread_dw2(const uint32_t* __restrict__ bits, const float* __restrict__ dw_w /*[9][C]*/,
             const float* __restrict__ dw_b, __nv_bfloat16* __restrict__ a_hi, __nv_bfloat16* __restrict__ a_lo,
             int N, int H, int W, int C, int segs, int tiles_per_row) {
  extern __shared__ __align__(16) uint8_t dw_sm[];   // [3][PX + 2][c8]
  const int c8 = C >> 3;
  const int PX = segs * R;
  const int rowbytes = (PX + 2) * c8;
  const int chunks_row = rowbytes >> 3;              // c8 % 8 == 0
  const int cg = threadIdx.x % c8, seg = threadIdx.x / c8;
  const uint8_t* gbytes = reinterpret_cast<const uint8_t*>(bits);
  float w[72], bias[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = *reinterpret_cast<const float4*>(dw_w + t * C + cg * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(dw_w + t * C + cg * 8 + 4);
    w[t * 8 + 0] = w0.x; w[t * 8 + 1] = w0.y; w[t * 8 + 2] = w0.z; w[t * 8 + 3] = w0.w;
    w[t * 8 + 4] = w1.x; w[t * 8 + 5] = w1.y; w[t * 8 + 6] = w1.z; w[t * 8 + 7] = w1.w;
  }
  {
    const float4 b0 = *reinterpret_cast<const float4*>(dw_b + cg * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(dw_b + cg * 8 + 4);
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  const int64_t tiles = (int64_t)N * H * tiles_per_row;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tx = static_cast<int>(tile % tiles_per_row);
    const int64_t r = tile / tiles_per_row;
    const int h = static_cast<int>(r % H);
    const int64_t img = r / H;
    const int x0 = tx * PX;
    __syncthreads();   // the previous tile's window reads are done
    const int q = c8 >> 3;                           // 8-byte chunks per pixel
    const int qshift = (q & (q - 1)) == 0 ? __ffs(q) - 1 : -1;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      const bool rok = hh >= 0 && hh < H;
      const uint8_t* grow = gbytes + ((img * H + (rok ? hh : h)) * W + (x0 - 1)) * (int64_t)c8;
      for (int j = threadIdx.x; j < chunks_row; j += blockDim.x) {
        const int px = qshift >= 0 ? (j >> qshift) : j / q;   // staged pixel 0 .. PX+1  <->  image column x0 - 1 + px
        const int xx = x0 - 1 + px;
        uint2 v = make_uint2(0u, 0u);
        if (rok && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const uint2*>(grow) + j);
        reinterpret_cast<uint2*>(dw_sm + ky * rowbytes)[j] = v;
      }
    }
    __syncthreads();
    const int p0 = seg * R;
    const uint8_t* sp = dw_sm + p0 * c8 + cg;        // staged pixel p0 = image column x0 + p0 - 1
    uint32_t win[3][3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      win[ky][0] = 0;
      win[ky][1] = sp[ky * rowbytes];
      win[ky][2] = sp[ky * rowbytes + c8];
    }
    int64_t o = ((img * H + h) * W + x0 + p0) * (int64_t)c8 + cg;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      sp += c8;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[ky][0] = win[ky][1];
        win[ky][1] = win[ky][2];
        win[ky][2] = sp[ky * rowbytes + c8];
      }
      if (x0 + p0 + j < W) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = bias[k];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t m = win[ky][kx];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (m & (1u << k)) acc[k] += w[(ky * 3 + kx) * 8 + k];
          }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 hp = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);   // one packed cvt
          hi[q] = *reinterpret_cast<const uint32_t*>(&hp);
          if (LO) {
            const float2 hf = __bfloat1622float2(hp);
            const __nv_bfloat162 lp = __floats2bfloat162_rn(acc[2 * q] - hf.x, acc[2 * q + 1] - hf.y);
            lo[q] = *reinterpret_cast<const uint32_t*>(&lp);
          }
        }
        reinterpret_cast<uint4*>(a_hi)[o] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (LO) reinterpret_cast<uint4*>(a_lo)[o] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      o += c8;
    }
  }
}

// ------------------------------------------------------------------------------------------
// SiLU "analog spike" neuron: mem_update(act=True) inside class Conv (models/common.py:362-375, 263-281).
// The neuron output is silu(mem) (real valued) and feeds the spread convs and the reset term.
// `inplace` reproduces the reference models, whose initialize_weights() switches nn.SiLU to in-place
// (utils/torch_utils.py:165-166) so that mem_old = mem.clone() captures silu(mem).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

__global__ void k_silu_first(const float* __restrict__ x, const float* __restrict__ scale,
                             const float* __restrict__ shift, float* __restrict__ out, float* __restrict__ mem,
                             float* __restrict__ mem_save, int64_t n4, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ecsy::ldg_stream(reinterpret_cast<const float4*>(x) + i);
    if (scale != nullptr) {
      const int c = static_cast<int>((i * 4) % C);
      const float4 s = *reinterpret_cast<const float4*>(scale + c);
      const float4 b = *reinterpret_cast<const float4*>(shift + c);
      v.x = ecsy::add_rn(ecsy::mul_rn(v.x, s.x), b.x);
      v.y = ecsy::add_rn(ecsy::mul_rn(v.y, s.y), b.y);
      v.z = ecsy::add_rn(ecsy::mul_rn(v.z, s.z), b.z);
      v.w = ecsy::add_rn(ecsy::mul_rn(v.w, s.w), b.w);
    }
    if (mem != nullptr) reinterpret_cast<float4*>(mem)[i] = v;
    if (mem_save != nullptr) reinterpret_cast<float4*>(mem_save)[i] = v;
    reinterpret_cast<float4*>(out)[i] = make_float4(silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
  }
}

struct SiluStep {
  const float* spread;
  const float* pw_b;
  const float* x_next;
  const float* in_scale;
  const float* in_shift;
  const float* mem_old;   // mem_{t} (or silu(mem_t) when in-place) ; may alias out_prev
  float* mem_out;         // raw membrane of step t+1 (NULL when in-place: out_next doubles as mem_old)
  float* ecs;
  const float* out_prev;  // silu(mem_t)
  float* out_next;        // silu(mem_{t+1})
  float* mem_save;        // optional: raw membrane of step t+1 (backward recompute)
  float* ecs_save;        // optional: e_t (backward recompute)
  int first, store_ecs;
  float decay, alpha, beta, kappa;
};

__global__ void k_silu_step(const SiluStep p, int64_t n4, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = static_cast<int>((i * 4) % C);
    float4 xv = ecsy::ldg_stream(reinterpret_cast<const float4*>(p.x_next) + i);
    if (p.in_scale != nullptr) {
      const float4 s = *reinterpret_cast<const float4*>(p.in_scale + c);
      const float4 b = *reinterpret_cast<const float4*>(p.in_shift + c);
      xv.x = ecsy::add_rn(ecsy::mul_rn(xv.x, s.x), b.x);
      xv.y = ecsy::add_rn(ecsy::mul_rn(xv.y, s.y), b.y);
      xv.z = ecsy::add_rn(ecsy::mul_rn(xv.z, s.z), b.z);
      xv.w = ecsy::add_rn(ecsy::mul_rn(xv.w, s.w), b.w);
    }
    const float4 sv = reinterpret_cast<const float4*>(p.spread)[i];
    const float4 mv = reinterpret_cast<const float4*>(p.mem_old)[i];
    const float4 ov = reinterpret_cast<const float4*>(p.out_prev)[i];
    float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!p.first) ev = reinterpret_cast<const float4*>(p.ecs)[i];
    const float4 pb = *reinterpret_cast<const float4*>(p.pw_b + c);
    const float xin[4] = {xv.x, xv.y, xv.z, xv.w}, sp[4] = {sv.x, sv.y, sv.z, sv.w};
    const float mo[4] = {mv.x, mv.y, mv.z, mv.w}, eo[4] = {ev.x, ev.y, ev.z, ev.w};
    const float so[4] = {ov.x, ov.y, ov.z, ov.w}, bb[4] = {pb.x, pb.y, pb.z, pb.w};
    float mn[4], en[4], on[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s_acc = ecsy::add_rn(sp[k], bb[k]);
      en[k] = ecsy::add_rn(ecsy::mul_rn(p.alpha, s_acc), ecsy::mul_rn(p.kappa, eo[k]));
      const float fecs = ecsy::mul_rn(p.beta, tanhf(en[k]));
      const float keep = 1.0f - so[k];
      mn[k] = ecsy::add_rn(ecsy::add_rn(ecsy::mul_rn(ecsy::mul_rn(mo[k], p.decay), keep), xin[k]), fecs);
      on[k] = silu_f(mn[k]);
    }
    if (p.mem_out != nullptr) reinterpret_cast<float4*>(p.mem_out)[i] = make_float4(mn[0], mn[1], mn[2], mn[3]);
    if (p.store_ecs) reinterpret_cast<float4*>(p.ecs)[i] = make_float4(en[0], en[1], en[2], en[3]);
    if (p.mem_save != nullptr) reinterpret_cast<float4*>(p.mem_save)[i] = make_float4(mn[0], mn[1], mn[2], mn[3]);
    if (p.ecs_save != nullptr) reinterpret_cast<float4*>(p.ecs_save)[i] = make_float4(en[0], en[1], en[2], en[3]);
    reinterpret_cast<float4*>(p.out_next)[i] = make_float4(on[0], on[1], on[2], on[3]);
  }
}

// Depthwise 3x3 (pad 1, bias) of a REAL-valued [N][H][W][C] tensor -> bf16 hi (+lo) GEMM operand rows.
__global__ void k_dw_real(const float* __restrict__ s, const float* __restrict__ dw_w /*[9][C]*/,
                          const float* __restrict__ dw_b, __nv_bfloat16* __restrict__ a_hi,
                          __nv_bfloat16* __restrict__ a_lo, int N, int H, int W, int C) {
  const int c8 = C >> 3;
  const int64_t total = (int64_t)N * H * W * c8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = static_cast<int>(i % c8);
    const int64_t pix = i / c8;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    float acc[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(dw_b + j * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(dw_b + j * 8 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww < 0 || ww >= W) continue;
        const float* sp = s + (pix + (int64_t)(ky - 1) * W + (kx - 1)) * C + j * 8;
        const float* wt = dw_w + (ky * 3 + kx) * C + j * 8;
        const float4 s0 = *reinterpret_cast<const float4*>(sp), s1 = *reinterpret_cast<const float4*>(sp + 4);
        const float4 w0 = *reinterpret_cast<const float4*>(wt), w1 = *reinterpret_cast<const float4*>(wt + 4);
        acc[0] = fmaf(s0.x, w0.x, acc[0]); acc[1] = fmaf(s0.y, w0.y, acc[1]);
        acc[2] = fmaf(s0.z, w0.z, acc[2]); acc[3] = fmaf(s0.w, w0.w, acc[3]);
        acc[4] = fmaf(s1.x, w1.x, acc[4]); acc[5] = fmaf(s1.y, w1.y, acc[5]);
        acc[6] = fmaf(s1.z, w1.z, acc[6]); acc[7] = fmaf(s1.w, w1.w, acc[7]);
      }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(acc[2 * q]), h1 = __float2bfloat16_rn(acc[2 * q + 1]);
      hi[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(acc[2 * q] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(acc[2 * q + 1] - __bfloat162float(h1));
      lo[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    reinterpret_cast<uint4*>(a_hi)[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (a_lo != nullptr) reinterpret_cast<uint4*>(a_lo)[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------
// tdBN statistics (models/common.py:674-679 -> BatchNorm3d over N,T,H,W): per-channel mean and
// biased variance of x[rows][C].  Stage 1: each block reduces a row range to partial (sum, sumsq)
// in double; stage 2: one thread per channel combines.  Deterministic (no atomics).
// ------------------------------------------------------------------------------------------
__global__ void k_bn_partial(const float* __restrict__ x, int64_t rows, int C, int64_t rows_per_block,
                             double* __restrict__ part /*[blocks][2][C]*/) {
  extern __shared__ double sred[];  // [nty][2][C]
  const int c4 = C >> 2;
  const int tpr = c4;                // threads per row (C <= 1024 -> <= 256)
  const int nty = blockDim.x / tpr;  // rows in flight
  const int tq = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int64_t r0 = blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (ty < nty) {
    float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int64_t r = r0 + ty; r < r1; r += nty) {
      float4 v = ecsy::ldg_stream(reinterpret_cast<const float4*>(x + r * C) + tq);
      fs[0] += v.x; fs[1] += v.y; fs[2] += v.z; fs[3] += v.w;
      fq[0] += v.x * v.x; fq[1] += v.y * v.y; fq[2] += v.z * v.z; fq[3] += v.w * v.w;
      if (++cnt == 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] += fs[k]; q[k] += fq[k]; fs[k] = 0; fq[k] = 0; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] += fs[k]; q[k] += fq[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sred[(ty * 2 + 0) * C + tq * 4 + k] = s[k];
      sred[(ty * 2 + 1) * C + tq * 4 + k] = q[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    double a = 0;
    for (int y = 0; y < nty; ++y) a += sred[y * 2 * C + c];
    part[(int64_t)blockIdx.x * 2 * C + c] = a;
  }
}

__global__ void k_bn_final(const double* __restrict__ part, int blocks, int C, double inv_count,
                           float* __restrict__ mean, float* __restrict__ var) {
  // one warp per channel: lanes stride over the per-block partials, fixed-order shuffle reduction
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int b = lane; b < blocks; b += 32) {
    s += part[(int64_t)b * 2 * C + c];
    q += part[(int64_t)b * 2 * C + C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double m = s * inv_count;
    double v = q * inv_count - m * m;
    if (v < 0) v = 0;
    mean[c] = static_cast<float>(m);
    var[c] = static_cast<float>(v);
  }
}

// ------------------------------------------------------------------------------------------
// out = a*sa + ba (+ b*sb + bb): block output in training mode, where the two tdBN normalisations
// feeding the membrane shortcut are applied on the fly (models/common.py:1216).  `*_mod`: number of
// distinct images in the source (N for a T-broadcast tensor, imgs otherwise).
// ------------------------------------------------------------------------------------------
__global__ void k_affine_add(const float* __restrict__ a, int64_t a_mod, const float* __restrict__ sa,
                             const float* __restrict__ ba, const float* __restrict__ b, int64_t b_mod,
                             const float* __restrict__ sb, const float* __restrict__ bb,
                             float* __restrict__ out, int64_t imgs, int64_t hwc4, int C) {
  // blockIdx.y walks the images, blockIdx.x / the grid-stride loop one image's float4 items: the channel index
  // advances incrementally and no per-item division is left (the kernel was ALU-bound on index arithmetic).
  const int stride = (int)(gridDim.x * blockDim.x);
  const int i0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int c0 = (int)(((int64_t)i0 * 4) % C);
  const int cstep = (int)(((int64_t)stride * 4) % C);
  for (int64_t img = blockIdx.y; img < imgs; img += gridDim.y) {
    const float4* pa = reinterpret_cast<const float4*>(a) + (img % a_mod) * hwc4;
    const float4* pb = b != nullptr ? reinterpret_cast<const float4*>(b) + (img % b_mod) * hwc4 : nullptr;
    float4* po = reinterpret_cast<float4*>(out) + img * hwc4;
    int c = c0;
    for (int64_t i = i0; i < hwc4; i += stride) {
      float4 v = pa[i];
      if (sa != nullptr) {
        float4 s = *reinterpret_cast<const float4*>(sa + c), t = *reinterpret_cast<const float4*>(ba + c);
        v.x = v.x * s.x + t.x; v.y = v.y * s.y + t.y; v.z = v.z * s.z + t.z; v.w = v.w * s.w + t.w;
      }
      if (pb != nullptr) {
        float4 u = pb[i];
        if (sb != nullptr) {
          float4 s = *reinterpret_cast<const float4*>(sb + c), t = *reinterpret_cast<const float4*>(bb + c);
          u.x = u.x * s.x + t.x; u.y = u.y * s.y + t.y; u.z = u.z * s.z + t.z; u.w = u.w * s.w + t.w;
        }
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      po[i] = v;
      c += cstep;
      if (c >= C) c -= C;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Resample into a channel slice of the output: s x s max-pool (MaxPool3d((1,s,s)), common.py:1209),
// nearest up-sampling by u (Sample, common.py:856-868) or plain copy (Concat, common.py:1764),
// optionally with a per-channel affine on the source (a pending tdBN normalisation).
// ------------------------------------------------------------------------------------------
__global__ void k_resample(const float* __restrict__ in, int64_t in_mod, const float* __restrict__ sc,
                           const float* __restrict__ sh, float* __restrict__ out, int64_t imgs, int Hi,
                           int Wi, int C, int Ho, int Wo, int Ctot, int coff, int pool, int up) {
  const int c4 = C >> 2;
  const int64_t total = imgs * Ho * Wo * c4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool small = total < (int64_t(1) << 32);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int q = static_cast<int>(ecsy::mod_u(i, (uint32_t)c4, small));
    int64_t p = ecsy::div_u(i, (uint32_t)c4, small);
    const int wo = static_cast<int>(ecsy::mod_u(p, (uint32_t)Wo, small));
    p = ecsy::div_u(p, (uint32_t)Wo, small);
    const int ho = static_cast<int>(ecsy::mod_u(p, (uint32_t)Ho, small));
    const int64_t img = ecsy::div_u(p, (uint32_t)Ho, small);
    const float* src = in + ((int64_t)ecsy::mod_u(img, (uint32_t)in_mod, small) * Hi * Wi) * (int64_t)C + q * 4;
    float4 v;
    if (pool > 1) {
      v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      for (int dy = 0; dy < pool; ++dy)
        for (int dx = 0; dx < pool; ++dx) {
          float4 u = *reinterpret_cast<const float4*>(src + ((int64_t)(ho * pool + dy) * Wi + (wo * pool + dx)) * C);
          v.x = fmaxf(v.x, u.x); v.y = fmaxf(v.y, u.y); v.z = fmaxf(v.z, u.z); v.w = fmaxf(v.w, u.w);
        }
    } else {
      v = *reinterpret_cast<const float4*>(src + ((int64_t)(ho / up) * Wi + (wo / up)) * C);
    }
    if (sc != nullptr) {
      // affine AFTER the max is only valid for non-negative scales; callers pass sc only when pool==1
      float4 s = *reinterpret_cast<const float4*>(sc + q * 4), t = *reinterpret_cast<const float4*>(sh + q * 4);
      v.x = v.x * s.x + t.x; v.y = v.y * s.y + t.y; v.z = v.z * s.z + t.z; v.w = v.w * s.w + t.w;
    }
    *reinterpret_cast<float4*>(out + (((img * Ho + ho) * Wo + wo) * (int64_t)Ctot + coff + q * 4)) = v;
  }
}

// ------------------------------------------------------------------------------------------
// Weighted reduction over T: out[n][...] = (sum_t w[t] * x[t][n][...]) / div.  Conv_7 (the learned
// T->1 Conv3d, common.py:549-562) commutes with Detect's per-step 1x1 conv, so Stack A reduces the
// real features first; DDetect's mean over T (yolo_snn.py:115-116) uses w = 1, div = T.
// ------------------------------------------------------------------------------------------
__global__ void k_tsum(const float* __restrict__ x, const float* __restrict__ w, float div,
                       float* __restrict__ out, int T, int64_t per_t4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_t4; i += stride) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      float4 v = reinterpret_cast<const float4*>(x)[t * per_t4 + i];
      const float wt = w ? w[t] : 1.0f;
      a.x += wt * v.x; a.y += wt * v.y; a.z += wt * v.z; a.w += wt * v.w;
    }
    a.x /= div; a.y /= div; a.z /= div; a.w /= div;
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// Stack-A Detect decode (models/yolo.py:110-146).  y: [N][H][W][na*no] (channel = a*no + o).
// raw: [N][na][H][W][no]; z (eval only): [N][rows_total][no] at row offset row_off.
// ------------------------------------------------------------------------------------------
__global__ void k_detect_decode(const float* __restrict__ y, float* __restrict__ raw, float* __restrict__ z,
                                const float* __restrict__ anchors /*[na][2], grid units*/, float stride_px,
                                int N, int H, int W, int na, int no, int64_t rows_total, int64_t row_off) {
  const int64_t total = (int64_t)N * H * W * na * no;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    // iterate in raw-output order so stores coalesce
    int o = static_cast<int>(i % no);
    int64_t p = i / no;
    const int w = static_cast<int>(p % W); p /= W;
    const int h = static_cast<int>(p % H); p /= H;
    const int a = static_cast<int>(p % na);
    const int64_t n = p / na;
    const float v = y[(((n * H + h) * W + w) * na + a) * (int64_t)no + o];
    raw[i] = v;
    if (z != nullptr) {
      float sg = 1.0f / (1.0f + expf(-v));
      float r;
      if (o == 0) r = (sg * 2.0f - 0.5f + static_cast<float>(w)) * stride_px;
      else if (o == 1) r = (sg * 2.0f - 0.5f + static_cast<float>(h)) * stride_px;
      else if (o == 2 || o == 3) {
        float t = sg * 2.0f;
        r = t * t * (anchors[a * 2 + (o - 2)] * stride_px);
      } else r = sg;
      z[(n * rows_total + row_off + ((int64_t)a * H + h) * W + w) * no + o] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stack-B DDetect decode (models/yolo_snn.py:118-127, common.py:312-323, anchor_generator.py:8-32).
// box: [N][H][W][64] (4 sides x 16 bins), cls: [N][H][W][nc], both already averaged over T.
// xs: [N][64+nc][H][W] (reference-shaped raw level output);
// y (eval): [N][4+nc][A_total] at anchor offset a_off.
// ------------------------------------------------------------------------------------------
__global__ void k_ddetect_decode(const float* __restrict__ box, const float* __restrict__ cls,
                                 float* __restrict__ xs, float* __restrict__ yout, float stride_px, int N,
                                 int H, int W, int nc, int64_t a_total, int64_t a_off) {
  const int no = 64 + nc;
  const int64_t total = (int64_t)N * H * W;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int64_t n = i / ((int64_t)W * H);
    const float* b = box + i * 64;
    const float* c = cls + i * nc;
    float d[4];
    for (int s = 0; s < 4; ++s) {
      float mx = -INFINITY;
      for (int k = 0; k < 16; ++k) mx = fmaxf(mx, b[s * 16 + k]);
      float den = 0.f, num = 0.f;
      for (int k = 0; k < 16; ++k) {
        float e = expf(b[s * 16 + k] - mx);
        den += e;
        num += e * static_cast<float>(k);
      }
      d[s] = num / den;
    }
    float* xo = xs + (n * no) * (int64_t)H * W + (int64_t)h * W + w;
    for (int k = 0; k < 64; ++k) xo[(int64_t)k * H * W] = b[k];
    for (int k = 0; k < nc; ++k) xo[(int64_t)(64 + k) * H * W] = c[k];
    if (yout != nullptr) {
      const float ax = static_cast<float>(w) + 0.5f, ay = static_cast<float>(h) + 0.5f;
      const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
      float* yo = yout + n * (4 + nc) * a_total + a_off + (int64_t)h * W + w;
      yo[0] = (x1 + x2) / 2.0f * stride_px;
      yo[a_total] = (y1 + y2) / 2.0f * stride_px;
      yo[2 * a_total] = (x2 - x1) * stride_px;
      yo[3 * a_total] = (y2 - y1) * stride_px;
      for (int k = 0; k < nc; ++k) yo[(int64_t)(4 + k) * a_total] = 1.0f / (1.0f + expf(-c[k]));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward helpers of the resampling ops: max-pool (gradient to the first maximum of each window, as
// nn.MaxPool3d), nearest up-sampling (sum over each s x s block) and channel slicing (Concat backward).
// ------------------------------------------------------------------------------------------
__global__ void k_maxpool_bwd(const float* __restrict__ x, int64_t x_imgs, const float* __restrict__ gp,
                              float* __restrict__ gx, int64_t imgs, int Ho, int Wo, int C, int gC, int gcoff, int s) {
  const int c4 = C >> 2;
  const int64_t total = imgs * Ho * Wo * c4;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int Hi = Ho * s, Wi = Wo * s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int q = static_cast<int>(i % c4);
    int64_t p = i / c4;
    const int wo = static_cast<int>(p % Wo); p /= Wo;
    const int ho = static_cast<int>(p % Ho);
    const int64_t img = p / Ho;
    const float* src = x + ((img % x_imgs) * Hi * Wi) * (int64_t)C + q * 4;
    const float4 g = *reinterpret_cast<const float4*>(gp + (((img * Ho + ho) * Wo + wo) * (int64_t)gC + gcoff + q * 4));
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) {
        const float4 u = *reinterpret_cast<const float4*>(src + ((int64_t)(ho * s + dy) * Wi + (wo * s + dx)) * C);
        const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (uu[k] > best[k]) { best[k] = uu[k]; arg[k] = dy * s + dx; }
      }
    const float gg[4] = {g.x, g.y, g.z, g.w};
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) {
        float4 o;
        o.x = arg[0] == dy * s + dx ? gg[0] : 0.f;
        o.y = arg[1] == dy * s + dx ? gg[1] : 0.f;
        o.z = arg[2] == dy * s + dx ? gg[2] : 0.f;
        o.w = arg[3] == dy * s + dx ? gg[3] : 0.f;
        *reinterpret_cast<float4*>(gx + (((img * Hi + ho * s + dy) * Wi + wo * s + dx) * (int64_t)C + q * 4)) = o;
      }
  }
}

// out[img][ho][wo][c] = sum over the s x s block of in[img][ho*s+dy][wo*s+dx][coff + c]  (s == 1: channel slice)
__global__ void k_sumpool_slice(const float* __restrict__ in, float* __restrict__ out, int64_t imgs, int Ho, int Wo,
                                int C, int inC, int coff, int s) {
  const int c4 = C >> 2;
  const int64_t total = imgs * Ho * Wo * c4;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int Hi = Ho * s, Wi = Wo * s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int q = static_cast<int>(i % c4);
    int64_t p = i / c4;
    const int wo = static_cast<int>(p % Wo); p /= Wo;
    const int ho = static_cast<int>(p % Ho);
    const int64_t img = p / Ho;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) {
        const float4 u = *reinterpret_cast<const float4*>(
            in + (((img * Hi + ho * s + dy) * Wi + wo * s + dx) * (int64_t)inC + coff + q * 4));
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ecsy_nchw_to_nhwc_f32(const float* in, float* out, int64_t imgs, int C, int H, int W, void* stream) {
  ECSY_CHECK_ARG(in && out && imgs > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad arguments");
  if (C <= 4) {
    const int64_t hw = (int64_t)H * W;
    const int grid = grid_for(imgs * hw, kThreads, ecsy_num_sms() * 8);
    switch (C) {
      case 1: k_nchw_to_nhwc_small<1><<<grid, kThreads, 0, STREAM(stream)>>>(in, out, imgs, hw); break;
      case 2: k_nchw_to_nhwc_small<2><<<grid, kThreads, 0, STREAM(stream)>>>(in, out, imgs, hw); break;
      case 3: k_nchw_to_nhwc_small<3><<<grid, kThreads, 0, STREAM(stream)>>>(in, out, imgs, hw); break;
      default: k_nchw_to_nhwc_small<4><<<grid, kThreads, 0, STREAM(stream)>>>(in, out, imgs, hw); break;
    }
    ECSY_LAUNCH_CHECK();
    return ECSY_OK;
  }
  ECSY_CHECK_ARG(imgs <= 65535, "nchw_to_nhwc: more than 65535 images per call");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, static_cast<unsigned>(imgs)), block(32, 8);
  k_transpose_inner<<<grid, block, 0, STREAM(stream)>>>(in, out, C, H * W);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_nhwc_to_nchw_f32(const float* in, float* out, int64_t imgs, int C, int H, int W, void* stream) {
  ECSY_CHECK_ARG(in && out && imgs > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad arguments");
  ECSY_CHECK_ARG(imgs <= 65535, "nhwc_to_nchw: more than 65535 images per call");
  dim3 grid((C + 31) / 32, (H * W + 31) / 32, static_cast<unsigned>(imgs)), block(32, 8);
  k_transpose_inner<<<grid, block, 0, STREAM(stream)>>>(in, out, H * W, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_spikes_pack(const float* x_nhwc, uint32_t* bits, int64_t pixels, int C, float thresh, void* stream) {
  ECSY_CHECK_ARG(x_nhwc && bits && pixels > 0 && C > 0 && C % 32 == 0, "spikes_pack: C must be a multiple of 32");
  const int64_t words = pixels * (C / 32);
  k_pack<<<grid_for(words, kThreads / 32, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(x_nhwc, bits, words, thresh);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_spikes_unpack(const uint32_t* bits, float* x_nhwc, int64_t pixels, int C, void* stream) {
  ECSY_CHECK_ARG(x_nhwc && bits && pixels > 0 && C > 0 && C % 32 == 0, "spikes_unpack: C must be a multiple of 32");
  const int64_t words = pixels * (C / 32);
  k_unpack<<<grid_for(words, kThreads / 32, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(bits, x_nhwc, words);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// Internal launchers shared with lif.cu -------------------------------------------------------
int ecsy_launch_lif_first(const float* x, const float* scale, const float* shift, float* mem, uint32_t* bits,
                          int64_t pixels, int C, float thresh, cudaStream_t st) {
  const int64_t n4 = pixels * C / 4;
  k_lif_first<<<grid_for(n4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(x, scale, shift, mem, bits, n4, C, thresh);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

int ecsy_launch_ecs_step(const EcsStep& p, int64_t pixels, int C, cudaStream_t st) {
  const int64_t n4 = pixels * C / 4;
  const int64_t n4_round = (n4 + 31) & ~int64_t(31);
  if (p.half_state)
    k_ecs_step<true><<<grid_for(n4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(p, n4, n4_round, C);
  else
    k_ecs_step<false><<<grid_for(n4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(p, n4, n4_round, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

int ecsy_launch_dw_real(const float* s, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                        int N, int H, int W, int C, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W * (C / 8);
  k_dw_real<<<grid_for(total, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(s, dw_w, dw_b, a_hi, a_lo, N, H, W, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

template <int R>
static void launch_dw2(const uint32_t* bits, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi,
                       __nv_bfloat16* a_lo, int N, int H, int W, int C, int segs, cudaStream_t st) {
  const int c8 = C / 8, PX = segs * R;
  const int tiles_per_row = (W + PX - 1) / PX;
  const int64_t tiles = (int64_t)N * H * tiles_per_row;
  const size_t smem = (size_t)3 * (PX + 2) * c8;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)ecsy_num_sms() * 8);
  if (a_lo)
    k_spread_dw2<R, true><<<grid, segs * c8, smem, st>>>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, tiles_per_row);
  else
    k_spread_dw2<R, false><<<grid, segs * c8, smem, st>>>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, tiles_per_row);
}

// version: 0 = measured default (ECSY_DW_V=1|2 overrides), 1 = global byte loads, 2 = shared-memory staged tiles
int ecsy_launch_spread_dw_v(const uint32_t* bits, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi,
                            __nv_bfloat16* a_lo, int N, int H, int W, int C, int version, cudaStream_t st) {
  const int c8 = C / 8;
  if (c8 < 1 || c8 > 256 || C % 64 != 0) {
    ecsy_set_error("spread_dw: C=%d out of range", C);
    return ECSY_ERR_ARG;
  }
  if (version == 0) {
    // measured (profiles/r01_dw_microbench.txt): both versions are issue bound on the bit tests + predicated adds; the
    // staged tiles win from 256 channels up (fewer, longer byte rows per pixel), the direct loads below that
    static const char* env = getenv("ECSY_DW_V");
    version = (env != nullptr && (env[0] == '1' || env[0] == '2')) ? env[0] - '0' : (c8 >= 32 ? 2 : 1);
  }
  if (version == 2) {
    // runs of R pixels per thread; pick the R whose tiles waste the fewest columns (640-pixel inputs: W = 20 * 2^n -> R = 5)
    int segs = 256 / c8 > 0 ? 256 / c8 : 1;
    static const int kR[4] = {5, 4, 8, 6};
    int bestR = 5;
    int64_t best = -1;
    for (int i = 0; i < 4; ++i) {
      const int R = kR[i];
      const int sg = std::max(1, std::min(segs, (W + R - 1) / R));
      const int PX = sg * R;
      const int64_t padded = (int64_t)((W + PX - 1) / PX) * PX;
      if (best < 0 || padded < best) { best = padded; bestR = R; }
    }
    segs = std::max(1, std::min(segs, (W + bestR - 1) / bestR));
    switch (bestR) {
      case 4: launch_dw2<4>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, st); break;
      case 5: launch_dw2<5>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, st); break;
      case 6: launch_dw2<6>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, st); break;
      default: launch_dw2<8>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, segs, st); break;
    }
    ECSY_LAUNCH_CHECK();
    return ECSY_OK;
  }
  // block size: a multiple of c8 (every thread keeps one channel group for all of its work items)
  int lcm = c8;
  while (lcm % 32 != 0) lcm += c8;
  const int bd = lcm <= 256 ? (256 / lcm) * lcm : (256 / c8) * c8;
  const int nseg = (W + kDwRun - 1) / kDwRun;
  const int64_t items = (int64_t)N * H * nseg * c8;
  if (a_lo)
    k_spread_dw<true><<<grid_for(items, bd, ecsy_num_sms() * 4), bd, 0, st>>>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C);
  else
    k_spread_dw<false><<<grid_for(items, bd, ecsy_num_sms() * 4), bd, 0, st>>>(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

int ecsy_launch_spread_dw(const uint32_t* bits, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi,
                          __nv_bfloat16* a_lo, int N, int H, int W, int C, cudaStream_t st) {
  return ecsy_launch_spread_dw_v(bits, dw_w, dw_b, a_hi, a_lo, N, H, W, C, 0, st);
}

// Depth-wise half of the ECS spread on its own (models/common.py:289-294, spread[0]): bits [N,H,W,C/32] ->
// bf16 rows a_hi (+ a_lo residual plane or NULL) [N*H*W, C].  `version` selects the kernel (0 = default).
extern "C" int ecsy_spread_dw(const uint32_t* bits, const float* dw_w, const float* dw_b, void* a_hi, void* a_lo,
                              int64_t N, int H, int W, int C, int version, void* stream) {
  ECSY_CHECK_ARG(bits && dw_w && dw_b && a_hi && N > 0 && N < (1LL << 31) && H > 0 && W > 0, "spread_dw: bad arguments");
  ECSY_CHECK_ARG(version >= 0 && version <= 2, "spread_dw: version");
  return ecsy_launch_spread_dw_v(bits, dw_w, dw_b, reinterpret_cast<__nv_bfloat16*>(a_hi),
                                 reinterpret_cast<__nv_bfloat16*>(a_lo), (int)N, H, W, C, version,
                                 reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t ecsy_tdbn_stats_ws_bytes(int64_t rows, int C) {
  int64_t blocks = (rows + 255) / 256;
  const int64_t cap = (int64_t)ecsy_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<size_t>(blocks) * 2 * C * sizeof(double);
}

extern "C" int ecsy_tdbn_stats(const float* x, int64_t rows, int C, float* mean, float* var_biased, void* ws,
                               size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(x && mean && var_biased && rows > 0, "tdbn_stats: bad arguments");
  ECSY_CHECK_ARG(C % 4 == 0 && C >= 4 && C <= 1024, "tdbn_stats: C=%d must be a multiple of 4 in [4,1024]", C);
  const size_t need = ecsy_tdbn_stats_ws_bytes(rows, C);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("tdbn_stats: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  const int blocks = static_cast<int>(need / (2 * C * sizeof(double)));
  const int64_t rpb = (rows + blocks - 1) / blocks;
  const int nty = kThreads / (C / 4);
  const size_t smem = static_cast<size_t>(nty) * 2 * C * sizeof(double);
  ECSY_CUDA(cudaFuncSetAttribute(k_bn_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  k_bn_partial<<<blocks, kThreads, smem, STREAM(stream)>>>(x, rows, C, rpb, static_cast<double*>(ws));
  ECSY_LAUNCH_CHECK();
  k_bn_final<<<(C * 32 + 255) / 256, 256, 0, STREAM(stream)>>>(static_cast<const double*>(ws), blocks, C,
                                                           1.0 / static_cast<double>(rows), mean, var_biased);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// ---- tdBN per-channel glue in ONE launch each (the training step spent ~1000 launches of torch element-wise kernels on
// [C]-sized vectors here: ~5 ms of a 115 ms resnet18 step) ------------------------------------------------------------
namespace {
__global__ void k_tdbn_finish(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ rmean, float* __restrict__ rvar,
                              long long* __restrict__ nbt, float m, float unbias, float eps, int updates,
                              float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ rstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += updates;
  if (c >= C) return;
  const float mu = mean[c], v = var[c];
  if (rmean != nullptr) {
    float rm = rmean[c], rv = rvar[c];
    for (int u = 0; u < updates; ++u) {   // running = running * (1 - m) + batch * m   (nn.BatchNorm3d, unbiased variance)
      rm = ecsy::add_rn(ecsy::mul_rn(rm, 1.0f - m), ecsy::mul_rn(mu, m));
      rv = ecsy::add_rn(ecsy::mul_rn(rv, 1.0f - m), ecsy::mul_rn(v, m * unbias));
    }
    rmean[c] = rm;
    rvar[c] = rv;
  }
  const float r = rsqrtf(ecsy::add_rn(v, eps));
  const float sc = ecsy::mul_rn(w[c], r);
  rstd[c] = r;
  scale[c] = sc;
  shift[c] = ecsy::add_rn(b[c], -ecsy::mul_rn(mu, sc));
}

__global__ void k_tdbn_bwd_coef(const float* __restrict__ sg, const float* __restrict__ sgy, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ w, float n, float tfac,
                                float* __restrict__ A, float* __restrict__ B, float* __restrict__ Cc, float* __restrict__ gw,
                                int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = rstd[c], mu = mean[c], s = sg[c];
  const float sgx = ecsy::mul_rn(r, ecsy::add_rn(sgy[c], -ecsy::mul_rn(mu, s)));   // sum g * x_hat
  const float a = ecsy::mul_rn(w[c], r);
  const float bq = __fdiv_rn(ecsy::mul_rn(ecsy::mul_rn(-a, r), sgx), n);
  const float cq = ecsy::add_rn(__fdiv_rn(ecsy::mul_rn(-a, s), n), -ecsy::mul_rn(bq, mu));
  A[c] = a;
  B[c] = ecsy::mul_rn(bq, tfac);
  Cc[c] = ecsy::mul_rn(cq, tfac);
  gw[c] = sgx;
}
}  // namespace

// Train-mode tdBN after ecsy_tdbn_stats (models/common.py:668-700 through nn.BatchNorm3d): running statistics updated
// `updates` times with momentum m (running_* / num_batches_tracked may be NULL: track_running_stats off), and the affine
// the consumers apply: rstd = rsqrt(var + eps), scale = weight * rstd, shift = bias - mean * scale.
extern "C" int ecsy_tdbn_finish(const float* mean, const float* var_biased, const float* weight, const float* bias,
                                float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                                float unbias, float eps, int updates, float* scale, float* shift, float* rstd, int C,
                                void* stream) {
  ECSY_CHECK_ARG(mean && var_biased && weight && bias && scale && shift && rstd && C > 0 && updates >= 0,
                 "tdbn_finish: bad arguments");
  ECSY_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "tdbn_finish: running statistics come in pairs");
  k_tdbn_finish<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mean, var_biased, weight, bias, running_mean, running_var, num_batches_tracked, momentum, unbias, eps, updates, scale,
      shift, rstd, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// tdBN backward coefficients from the batch sums sg = sum g, sgy = sum g * y (ecsy_colsum2): g_y = A g + B y + C per channel
// with A = w rstd, B = -A rstd sgx / n, C = -A sg / n - B mean (B, C times tfac = T / Tp for a T-broadcast y), and the
// weight gradient sgx = rstd (sgy - mean sg); the bias gradient is sg itself.
extern "C" int ecsy_tdbn_bwd_coef(const float* sg, const float* sgy, const float* mean, const float* rstd,
                                  const float* weight, float n, float tfac, float* A, float* B, float* Cc, float* g_weight,
                                  int C, void* stream) {
  ECSY_CHECK_ARG(sg && sgy && mean && rstd && weight && A && B && Cc && g_weight && C > 0 && n > 0.f,
                 "tdbn_bwd_coef: bad arguments");
  k_tdbn_bwd_coef<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sg, sgy, mean, rstd, weight, n, tfac, A,
                                                                                      B, Cc, g_weight, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_affine_add(const float* a, int64_t a_imgs, const float* sa, const float* ba, const float* b,
                               int64_t b_imgs, const float* sb, const float* bb, float* out, int64_t imgs,
                               int64_t hw, int C, void* stream) {
  ECSY_CHECK_ARG(a && out && imgs > 0 && hw > 0 && C % 4 == 0, "affine_add: bad arguments");
  ECSY_CHECK_ARG(a_imgs > 0 && imgs % a_imgs == 0 && (!b || (b_imgs > 0 && imgs % b_imgs == 0)),
                 "affine_add: source image counts must divide imgs");
  ECSY_CHECK_ARG((sa == nullptr) == (ba == nullptr) && (sb == nullptr) == (bb == nullptr), "affine_add: scale/shift pairs");
  const int64_t hwc4 = hw * C / 4;
  const int64_t want = (int64_t)ecsy_num_sms() * 8;                      // blocks in flight
  const int64_t per_img = (hwc4 + kThreads - 1) / kThreads;
  int64_t gx = (want + imgs - 1) / imgs;
  if (gx > per_img) gx = per_img;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)(imgs < 65535 ? imgs : 65535));
  k_affine_add<<<grid, kThreads, 0, STREAM(stream)>>>(a, a_imgs, sa, ba, b, b ? b_imgs : 1, sb, bb, out, imgs, hwc4, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_resample(const float* in, int64_t in_imgs, const float* scale, const float* shift, float* out,
                             int64_t imgs, int Hi, int Wi, int C, int Ctot, int coff, int pool, int up, void* stream) {
  ECSY_CHECK_ARG(in && out && imgs > 0 && in_imgs > 0 && imgs % in_imgs == 0, "resample: bad image counts");
  ECSY_CHECK_ARG(C % 4 == 0 && Ctot % 4 == 0 && coff % 4 == 0 && coff + C <= Ctot, "resample: channel slice");
  ECSY_CHECK_ARG(pool >= 1 && up >= 1 && !(pool > 1 && up > 1), "resample: pool/up");
  ECSY_CHECK_ARG(!(pool > 1 && scale), "resample: affine with max-pool is not supported");
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "resample: scale/shift pair");
  const int Ho = pool > 1 ? Hi / pool : Hi * up, Wo = pool > 1 ? Wi / pool : Wi * up;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0, "resample: empty output");
  const int64_t total = imgs * Ho * Wo * (C / 4);
  k_resample<<<grid_for(total, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
      in, in_imgs, scale, shift, out, imgs, Hi, Wi, C, Ho, Wo, Ctot, coff, pool, up);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_tsum(const float* x, const float* w, float div, float* out, int T, int64_t per_t, void* stream) {
  ECSY_CHECK_ARG(x && out && T > 0 && per_t > 0 && per_t % 4 == 0 && div != 0.f, "tsum: bad arguments");
  k_tsum<<<grid_for(per_t / 4, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(x, w, div, out, T, per_t / 4);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_detect_decode(const float* y, float* raw, float* z, const float* anchors, float stride_px, int N,
                                  int H, int W, int na, int no, int64_t rows_total, int64_t row_off, void* stream) {
  ECSY_CHECK_ARG(y && raw && N > 0 && H > 0 && W > 0 && na > 0 && no > 4, "detect_decode: bad arguments");
  ECSY_CHECK_ARG(!z || (anchors && row_off + (int64_t)na * H * W <= rows_total), "detect_decode: z slice");
  const int64_t total = (int64_t)N * H * W * na * no;
  k_detect_decode<<<grid_for(total, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
      y, raw, z, anchors, stride_px, N, H, W, na, no, rows_total, row_off);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_ddetect_decode(const float* box, const float* cls, float* xs, float* y, float stride_px, int N,
                                   int H, int W, int nc, int64_t a_total, int64_t a_off, void* stream) {
  ECSY_CHECK_ARG(box && cls && xs && N > 0 && H > 0 && W > 0 && nc > 0, "ddetect_decode: bad arguments");
  ECSY_CHECK_ARG(!y || a_off + (int64_t)H * W <= a_total, "ddetect_decode: y slice");
  const int64_t total = (int64_t)N * H * W;
  k_ddetect_decode<<<grid_for(total, 128, ecsy_num_sms() * 8), 128, 0, STREAM(stream)>>>(box, cls, xs, y, stride_px, N, H,
                                                                                      W, nc, a_total, a_off);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// ---- mem_update(act=True).forward (models/common.py:252-283 with self.act, class Conv :362-375) ----
static inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" size_t ecsy_lif_silu_ws_bytes(int T, int64_t N, int H, int W, int C, int splits) {
  const size_t mc = static_cast<size_t>(N) * H * W * C;
  (void)T;
  return 512 + 3 * al256(mc * 4) + static_cast<size_t>(splits) * al256(mc * 2);
}

extern "C" int ecsy_lif_silu_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                                 const float* dw_w, const float* dw_b, const void* pw_packed, const float* pw_b,
                                 int splits, float* out, float* mem_save, float* ecs_save, int inplace, int T, int64_t N,
                                 int H, int W, int C, float decay, float alpha, float beta, float kappa, void* ws,
                                 size_t ws_bytes, void* stream) {
  cudaStream_t st = STREAM(stream);
  ECSY_CHECK_ARG(x && out && T >= 1 && N > 0 && H > 0 && W > 0, "lif_silu_fwd: bad arguments");
  ECSY_CHECK_ARG(C % 64 == 0, "lif_silu_fwd: C=%d must be a multiple of 64", C);
  ECSY_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "lif_silu_fwd: scale/shift pair");
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "lif_silu_fwd: splits must be 1 or 2");
  ECSY_CHECK_ARG(T == 1 || (dw_w && dw_b && pw_packed && pw_b), "lif_silu_fwd: spread weights missing");
  const int64_t M = N * H * W;
  const size_t mc = static_cast<size_t>(M) * C;
  const size_t need = ecsy_lif_silu_ws_bytes(T, N, H, W, C, splits);
  if (T > 1 && (ws == nullptr || ws_bytes < need)) {
    ecsy_set_error("lif_silu_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  float* mem = reinterpret_cast<float*>(p); p += al256(mc * 4);
  float* ecs = reinterpret_cast<float*>(p); p += al256(mc * 4);
  float* spread = reinterpret_cast<float*>(p); p += al256(mc * 4);
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2);
  __nv_bfloat16* a_lo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(p) : nullptr;
  const int64_t n4 = M * C / 4;
  const int grid = grid_for(n4, kThreads, ecsy_num_sms() * 8);
  k_silu_first<<<grid, kThreads, 0, st>>>(x, in_scale, in_shift, out, (T > 1 && !inplace) ? mem : nullptr, mem_save, n4, C);
  ECSY_LAUNCH_CHECK();
  for (int t = 0; t + 1 < T; ++t) {
    const float* out_t = out + (size_t)t * mc;
    k_dw_real<<<grid_for(M * (C / 8), kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(out_t, dw_w, dw_b, a_hi, a_lo,
                                                                                     (int)N, H, W, C);
    ECSY_LAUNCH_CHECK();
    int rc = ecsy_umma_dense(a_hi, a_lo, M, C, pw_packed, splits, spread, C, nullptr, nullptr, nullptr, 0, st);
    if (rc) return rc;
    SiluStep s{};
    s.spread = spread; s.pw_b = pw_b;
    s.x_next = x + (t + 1) * x_tstride;
    s.in_scale = in_scale; s.in_shift = in_shift;
    const bool more = t + 2 < T;
    s.mem_old = inplace ? out_t : mem;
    s.mem_out = (!inplace && more) ? mem : nullptr;
    s.ecs = ecs; s.store_ecs = more ? 1 : 0;
    s.out_prev = out_t;
    s.out_next = out + (size_t)(t + 1) * mc;
    s.mem_save = mem_save ? mem_save + (size_t)(t + 1) * mc : nullptr;
    s.ecs_save = ecs_save ? ecs_save + (size_t)t * mc : nullptr;
    s.first = t == 0 ? 1 : 0;
    s.decay = decay; s.alpha = alpha; s.beta = beta; s.kappa = kappa;
    k_silu_step<<<grid, kThreads, 0, st>>>(s, n4, C);
    ECSY_LAUNCH_CHECK();
  }
  return ECSY_OK;
}

extern "C" int ecsy_maxpool_bwd(const float* x, int64_t x_imgs, const float* g_pooled, float* gx, int64_t imgs, int Ho,
                                int Wo, int C, int gC, int gcoff, int s, void* stream) {
  ECSY_CHECK_ARG(x && g_pooled && gx && imgs > 0 && x_imgs > 0 && imgs % x_imgs == 0, "maxpool_bwd: bad arguments");
  ECSY_CHECK_ARG(C % 4 == 0 && gC % 4 == 0 && gcoff % 4 == 0 && gcoff + C <= gC && s >= 1, "maxpool_bwd: channels");
  const int64_t total = imgs * Ho * Wo * (C / 4);
  k_maxpool_bwd<<<grid_for(total, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
      x, x_imgs, g_pooled, gx, imgs, Ho, Wo, C, gC, gcoff, s);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_sumpool_slice(const float* in, float* out, int64_t imgs, int Ho, int Wo, int C, int inC, int coff,
                                  int s, void* stream) {
  ECSY_CHECK_ARG(in && out && imgs > 0 && s >= 1, "sumpool_slice: bad arguments");
  ECSY_CHECK_ARG(C % 4 == 0 && inC % 4 == 0 && coff % 4 == 0 && coff + C <= inC, "sumpool_slice: channels");
  const int64_t total = imgs * Ho * Wo * (C / 4);
  k_sumpool_slice<<<grid_for(total, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(in, out, imgs, Ho, Wo, C,
                                                                                               inC, coff, s);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
