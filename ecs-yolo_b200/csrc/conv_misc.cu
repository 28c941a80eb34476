// Weight packing, im2col for real-valued inputs, the generic SIMT convolution (odd shapes:
// grouped convs, Cout not a multiple of 64) and the C-ABI convolution entry points.
#include <stdlib.h>
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

// w [Co][Ci][kh][kw] fp32 -> out [splits][Co][Kpad] bf16 with k = (ky*kw + kx)*Ci + ci (hi plane,
// then the bf16 residual plane): hi + lo carries 16 mantissa bits of every weight.
__global__ void k_pack_weight(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Co, int Ci, int kh,
                              int kw, int Kpad, int splits) {
  const int64_t total = (int64_t)Co * Kpad;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int K = kh * kw * Ci;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int k = static_cast<int>(i % Kpad);
    const int co = static_cast<int>(i / Kpad);
    float v = 0.f;
    if (k < K) {
      const int tap = k / Ci, ci = k - tap * Ci;
      v = w[((int64_t)co * Ci + ci) * (kh * kw) + tap];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = hi;
    if (splits == 2) out[total + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// Same for the tensor-memory ("TS") spike convolution: the spike expanders emit a spike as the single bit 0x4000
// (= 2.0 in bf16) and pack channels (j, j+16) of every 32-channel half into one 32-bit word, so the weights are
// stored pre-multiplied by 0.5 (exact) with K element e of a 64-channel slab holding channel
// (e>>5)*32 + (e&1)*16 + ((e&31)>>1).
__global__ void k_pack_weight_ts(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Co, int Ci, int kh,
                                 int kw, int splits) {
  const int K = kh * kw * Ci;
  const int64_t total = (int64_t)Co * K;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int k = static_cast<int>(i % K);
    const int co = static_cast<int>(i / K);
    const int tap = k / Ci, pos = k - tap * Ci;
    const int e = pos & 63;
    const int ci = (pos & ~63) + ((e >> 5) << 5) + ((e & 1) << 4) + ((e & 31) >> 1);
    const float v = 0.5f * w[((int64_t)co * Ci + ci) * (kh * kw) + tap];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = hi;
    if (splits == 2) out[total + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// A[m][k] (bf16 hi [+ lo]) for a real-valued NHWC input; m = output pixel, k = (ky*kw+kx)*Ci + ci,
// zero padded to Kpad.  Each thread writes 8 consecutive k (16 bytes per plane).
// blockDim is a multiple of Kpad/8: a thread keeps ONE 8-wide k chunk (its eight (offset, ky, kx) entries live in registers)
// and walks output pixels with (wo, ho, img) tracked incrementally -- the first version decoded every 8 elements with four
// 64-bit divisions and eight table look-ups (1.9 ms for the stem of a batch-32 training step at 0.7 TB/s).
__global__ void k_im2col(const float* __restrict__ x, int64_t x_imgs, __nv_bfloat16* __restrict__ a_hi,
         __nv_bfloat16* __restrict__ a_lo, int64_t imgs, int H, int W, int Ci, int Ho, int Wo, int kh,
         int kw, int stride, int pad, int Kpad) {
  const int K = kh * kw * Ci;
  const int k8 = Kpad >> 3;
  const int j = threadIdx.x % k8, ty = threadIdx.x / k8, nty = blockDim.x / k8;
  int off[8], dy[8], dx[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = j * 8 + u;
    if (k < K) {
      const int tap = k / Ci, ci = k - tap * Ci;
      dy[u] = tap / kw;
      dx[u] = tap - dy[u] * kw;
      off[u] = (dy[u] * W + dx[u]) * Ci + ci;
    } else {
      dy[u] = -(1 << 20);   // padding of K: never inside the image
      dx[u] = 0;
      off[u] = 0;
    }
  }
  const int64_t rows = imgs * Ho * Wo;
  const int64_t rstep = (int64_t)gridDim.x * nty;
  int64_t m = (int64_t)blockIdx.x * nty + ty;
  int wo = static_cast<int>(m % Wo), ho = static_cast<int>((m / Wo) % Ho);
  int64_t img = m / ((int64_t)Wo * Ho);
  const int dwo = static_cast<int>(rstep % Wo), dho = static_cast<int>((rstep / Wo) % Ho);
  const int64_t dimg = rstep / ((int64_t)Wo * Ho);
  for (; m < rows; m += rstep) {
    const int hi0 = ho * stride - pad, wi0 = wo * stride - pad;
    const float* src = x + (img % x_imgs) * (int64_t)H * W * Ci + ((int64_t)hi0 * W + wi0) * Ci;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int hi = hi0 + dy[u], wi = wi0 + dx[u];
      v[u] = ((unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W) ? __ldg(src + off[u]) : 0.f;
    }
    uint32_t hi4[4], lo4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * q]), h1 = __float2bfloat16_rn(v[2 * q + 1]);
      hi4[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * q] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * q + 1] - __bfloat162float(h1));
      lo4[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const int64_t i = m * k8 + j;
    reinterpret_cast<uint4*>(a_hi)[i] = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    if (a_lo != nullptr) reinterpret_cast<uint4*>(a_lo)[i] = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
    wo += dwo;
    ho += dho;
    img += dimg;
    if (wo >= Wo) { wo -= Wo; ++ho; }
    if (ho >= Ho) { ho -= Ho; ++img; }
  }
}

static void launch_im2col(const float* x, int64_t x_imgs, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, int64_t imgs, int H, int W,
                          int Ci, int Ho, int Wo, int k, int stride, int pad, int Kpad, cudaStream_t st) {
  const int k8 = Kpad / 8;
  const int bd = k8 <= 256 ? (256 / k8) * k8 : k8;   // Kpad <= 8192: at most 1024 threads
  const int nty = bd / k8;
  const int64_t rows = imgs * Ho * Wo;
  int64_t blocks = (rows + nty - 1) / nty;
  const int64_t cap = (int64_t)ecsy_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_im2col<<<(int)blocks, bd, 0, st>>>(x, x_imgs, a_hi, a_lo, imgs, H, W, Ci, Ho, Wo, k, k, stride, pad, Kpad);
}

// fp32 -> bf16 hi (+ lo residual) planes
__global__ void k_f32_to_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                              int64_t n4) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += gs) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    reinterpret_cast<uint2*>(hi)[i] =
        make_uint2((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16),
                   (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16));
    if (lo != nullptr) {
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
      const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2));
      const __nv_bfloat16 l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
      reinterpret_cast<uint2*>(lo)[i] =
          make_uint2((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16),
                     (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16));
    }
  }
}

// Zero-insertion ("dilation") of an output gradient for the input-gradient of a strided conv:
// up[img][h'][w'][c] = gy[img][h'/s][w'/s][c] if both divide and are in range, else 0  -> bf16 hi (+ lo).
__global__ void k_dilate_to_bf16(const float* __restrict__ gy, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int64_t imgs, int Ho, int Wo, int C, int Hu, int Wu,
                                 int s) {
  const int c4 = C >> 2;
  const int64_t total = imgs * Hu * Wu * c4;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int q = static_cast<int>(i % c4);
    int64_t p = i / c4;
    const int wu = static_cast<int>(p % Wu); p /= Wu;
    const int hu = static_cast<int>(p % Hu);
    const int64_t img = p / Hu;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hu % s == 0 && wu % s == 0 && hu / s < Ho && wu / s < Wo)
      v = reinterpret_cast<const float4*>(gy + (((img * Ho + hu / s) * Wo + wu / s) * (int64_t)C))[q];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    reinterpret_cast<uint2*>(hi)[i] =
        make_uint2((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16),
                   (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16));
    if (lo != nullptr) {
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
      const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2));
      const __nv_bfloat16 l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
      reinterpret_cast<uint2*>(lo)[i] =
          make_uint2((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16),
                     (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16));
    }
  }
}

// Generic direct convolution, fp32 FMA: one thread per (output pixel, output channel).
// w: [kh][kw][Ci/g][Co] fp32.  (Snn_Conv2d on real inputs with odd shapes: Detect.m 1x1 + bias,
// models/yolo.py:73; grouped DDetect convs, models/yolo_snn.py:100-107.)
__global__ void k_simt_conv(const float* __restrict__ x, int64_t x_imgs, const float* __restrict__ w,
                            const float* __restrict__ bias, float bias_mul, const float* __restrict__ scale,
                            const float* __restrict__ shift, float* __restrict__ out, int64_t imgs, int H, int W,
                            int Ci, int Ho, int Wo, int Co, int kh, int kw, int stride, int pad, int groups) {
  const int64_t total = imgs * Ho * Wo * Co;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int cig = Ci / groups, cog = Co / groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int co = static_cast<int>(i % Co);
    int64_t m = i / Co;
    const int wo = static_cast<int>(m % Wo);
    const int ho = static_cast<int>((m / Wo) % Ho);
    const int64_t img = m / ((int64_t)Wo * Ho);
    const float* src = x + (img % x_imgs) * (int64_t)H * W * Ci + (co / cog) * cig;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int hi = ho * stride - pad + ky;
      if (hi < 0 || hi >= H) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int wi = wo * stride - pad + kx;
        if (wi < 0 || wi >= W) continue;
        const float* xp = src + ((int64_t)hi * W + wi) * Ci;
        const float* wp = w + ((int64_t)(ky * kw + kx) * cig) * Co + co;
        for (int ci = 0; ci < cig; ++ci) acc = fmaf(xp[ci], wp[(int64_t)ci * Co], acc);
      }
    }
    if (bias != nullptr) acc += bias[co] * bias_mul;
    if (scale != nullptr) acc = fmaf(acc, scale[co], shift[co]);
    out[i] = acc;
  }
}

// Few-input-channel direct convolution in plain fp32 (the parity-precision stem: Conv_1 on the 3-channel image,
// models/common.py:409-425; 123 GFLOP at batch 64 -- k_simt_conv, one thread per output element, ran it at 1.9 TFLOP/s =
// 66 of the parity step's 150 ms).  A thread owns one output pixel and 32 output channels: its 32 accumulators live in
// registers, the weights [kh*kw*Ci][Co] sit in shared memory and are read as warp-wide broadcasts (every lane of a warp
// works on the same channels), the input value of a tap is one cached global load per thread.  Same accumulation order
// as k_simt_conv (ky, kx, ci; fmaf), so the two kernels agree bit for bit.
__global__ void __launch_bounds__(256)
k_stem_conv_f32(const float* __restrict__ x, int64_t x_imgs, const float* __restrict__ w, const float* __restrict__ scale,
                const float* __restrict__ shift, float* __restrict__ out, int64_t imgs, int H, int W, int Ci, int Ho, int Wo,
                int Co, int kh, int kw, int stride, int pad) {
  extern __shared__ float s_w[];                   // [kh*kw*Ci][Co]
  const int K = kh * kw * Ci;
  for (int i = threadIdx.x; i < K * Co / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(w)[i];
  __syncthreads();
  const int cgroups = Co / 32;                     // channel groups of 32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t pixels = imgs * Ho * Wo;
  const int64_t units = ((pixels + 31) / 32) * cgroups;          // (32 consecutive pixels) x (32 channels) per warp
  for (int64_t u = (int64_t)blockIdx.x * nwarps + warp; u < units; u += (int64_t)gridDim.x * nwarps) {
    const int cg = static_cast<int>(u % cgroups);
    const int64_t m = (u / cgroups) * 32 + lane;
    const bool ok = m < pixels;
    const int64_t mm = ok ? m : pixels - 1;
    const int wo = static_cast<int>(mm % Wo);
    const int ho = static_cast<int>((mm / Wo) % Ho);
    const int64_t img = mm / ((int64_t)Wo * Ho);
    const float* src = x + (img % x_imgs) * (int64_t)H * W * Ci;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int hi = ho * stride - pad + ky;
      const bool hok = hi >= 0 && hi < H;
      for (int kx = 0; kx < kw; ++kx) {
        const int wi = wo * stride - pad + kx;
        const bool inb = hok && wi >= 0 && wi < W;
        const float* xp = src + ((int64_t)(inb ? hi : 0) * W + (inb ? wi : 0)) * Ci;
        const float* wp = s_w + (size_t)((ky * kw + kx) * Ci) * Co + cg * 32;
        for (int ci = 0; ci < Ci; ++ci) {
          // a tap outside the image is SKIPPED by k_simt_conv; adding fmaf(0, w, acc) == acc keeps the bits identical
          const float xv = inb ? __ldg(xp + ci) : 0.f;
          const float4* w4 = reinterpret_cast<const float4*>(wp + (size_t)ci * Co);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 wv = w4[c4];
            acc[4 * c4 + 0] = fmaf(xv, wv.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(xv, wv.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(xv, wv.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(xv, wv.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
    if (ok) {
      float* op = out + m * Co + cg * 32;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 v = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
        if (scale != nullptr) {
          const float4 sc = *reinterpret_cast<const float4*>(scale + cg * 32 + 4 * c4);
          const float4 sh = *reinterpret_cast<const float4*>(shift + cg * 32 + 4 * c4);
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        reinterpret_cast<float4*>(op)[c4] = v;
      }
    }
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

int ecsy_launch_f32_to_bf16(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t st) {
  k_f32_to_bf16<<<grid_for(n / 4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(x, hi, lo, n / 4);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_pack_conv_weight(const float* w, void* out_bf16, int Co, int Ci, int kh, int kw, int Kpad,
                                     int splits, void* stream) {
  ECSY_CHECK_ARG(w && out_bf16 && Co > 0 && Ci > 0 && kh > 0 && kw > 0, "pack_conv_weight: bad arguments");
  ECSY_CHECK_ARG(Kpad % 64 == 0 && Kpad >= kh * kw * Ci && (splits == 1 || splits == 2), "pack_conv_weight: Kpad/splits");
  k_pack_weight<<<grid_for((int64_t)Co * Kpad, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
      w, static_cast<__nv_bfloat16*>(out_bf16), Co, Ci, kh, kw, Kpad, splits);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_pack_spike_conv_weight(const float* w, void* out_bf16, int Co, int Ci, int kh, int kw, int splits,
                                           void* stream) {
  ECSY_CHECK_ARG(w && out_bf16 && Co > 0 && Ci > 0 && kh > 0 && kw > 0, "pack_spike_conv_weight: bad arguments");
  ECSY_CHECK_ARG(Ci % 64 == 0 && (splits == 1 || splits == 2), "pack_spike_conv_weight: Ci %% 64 / splits");
  k_pack_weight_ts<<<grid_for((int64_t)Co * kh * kw * Ci, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
      w, static_cast<__nv_bfloat16*>(out_bf16), Co, Ci, kh, kw, splits);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

extern "C" int ecsy_spike_conv_ts_supported(int Cin, int Cout) {
  return (Cin % 64 == 0 && Cin >= 64 && ecsy_pick_bn_ts(Cout) != 0) ? 1 : 0;
}

// Measured on B200 (profiles/r01_conv_microbench_*): with single-plane weights, wide layers (Cout % 256 == 0 and
// Cin >= 256) run faster with 256-column tiles and the operand in shared memory (1.3-1.4 PFLOP/s) than with the
// 128-column tiles the tensor-memory path is limited to (1.1 PFLOP/s); everything else prefers tensor memory.
extern "C" int ecsy_spike_conv_prefers_ts(int Cin, int Cout, int splits) {
  if (!ecsy_spike_conv_ts_supported(Cin, Cout)) return 0;
  if (splits == 1 && Cout % 256 == 0 && Cin >= 256) return 0;
  return 1;
}

extern "C" int ecsy_spike_conv_ts_fwd(const uint32_t* spikes, const void* w_ts, int splits, float* out,
                                      const float* scale, const float* shift, const float* residual, int64_t res_imgs,
                                      int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                      void* stream) {
  ECSY_CHECK_ARG(spikes && w_ts && out && imgs > 0 && H > 0 && W > 0, "spike_conv_ts_fwd: bad arguments");
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "spike_conv_ts_fwd: scale/shift pair");
  ECSY_CHECK_ARG(!residual || (res_imgs > 0 && imgs % res_imgs == 0), "spike_conv_ts_fwd: residual image count");
  ECSY_CHECK_ARG(imgs < (1 << 24), "spike_conv_ts_fwd: too many images");
  return ecsy_umma_spike_conv(spikes, w_ts, splits, out, scale, shift, residual, res_imgs, (int)imgs, H, W, Cin,
                              Cout, k, stride, pad, STREAM(stream), 1);
}

// Shared-memory-operand kernel (256-column tiles) on weights in the tensor-memory layout: same {0, 2.0} pair expansion
// (2 ALU ops per word instead of 4.5), used for the wide layers where ecsy_spike_conv_prefers_ts says no.
extern "C" int ecsy_spike_conv_pair_fwd(const uint32_t* spikes, const void* w_ts, int splits, float* out,
                                        const float* scale, const float* shift, const float* residual, int64_t res_imgs,
                                        int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                        void* stream) {
  ECSY_CHECK_ARG(spikes && w_ts && out && imgs > 0 && H > 0 && W > 0, "spike_conv_pair_fwd: bad arguments");
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "spike_conv_pair_fwd: scale/shift pair");
  ECSY_CHECK_ARG(!residual || (res_imgs > 0 && imgs % res_imgs == 0), "spike_conv_pair_fwd: residual image count");
  ECSY_CHECK_ARG(imgs < (1 << 24), "spike_conv_pair_fwd: too many images");
  return ecsy_umma_spike_conv(spikes, w_ts, splits, out, scale, shift, residual, res_imgs, (int)imgs, H, W, Cin,
                              Cout, k, stride, pad, STREAM(stream), 3);
}

extern "C" int ecsy_spike_conv_fwd(const uint32_t* spikes, const void* w_packed, int splits, float* out,
                                   const float* scale, const float* shift, const float* residual, int64_t res_imgs,
                                   int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                   void* stream) {
  ECSY_CHECK_ARG(spikes && w_packed && out && imgs > 0 && H > 0 && W > 0, "spike_conv_fwd: bad arguments");
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "spike_conv_fwd: scale/shift pair");
  ECSY_CHECK_ARG(!residual || (res_imgs > 0 && imgs % res_imgs == 0), "spike_conv_fwd: residual image count");
  ECSY_CHECK_ARG(imgs < (1 << 24), "spike_conv_fwd: too many images");
  return ecsy_umma_spike_conv(spikes, w_packed, splits, out, scale, shift, residual, res_imgs, (int)imgs, H, W, Cin,
                              Cout, k, stride, pad, STREAM(stream));
}

static bool use_gather_conv(int Cin, int Cout, int k, int groups, int splits) {
  static const bool gather = getenv("ECSY_CONV_GATHER") == nullptr || getenv("ECSY_CONV_GATHER")[0] != '0';
  return gather && splits == 1 && groups == 1 && Cout % 64 == 0 && Cin < 64 && k * k * Cin <= 1024;
}

extern "C" size_t ecsy_real_conv_ws_bytes(int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                          int groups, int splits) {
  if (groups != 1 || Cout % 64 != 0) return 0;
  if (use_gather_conv(Cin, Cout, k, groups, splits)) return 0;   // operand tiles gathered on the fly
  if (Cin % 64 == 0 && stride == 1)  // implicit GEMM over bf16 planes of the input, no im2col buffer
    return static_cast<size_t>(imgs) * H * W * Cin * 2 * splits + 1024;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t Kpad = ((int64_t)k * k * Cin + 63) / 64 * 64;
  return static_cast<size_t>(imgs * Ho * Wo * Kpad * 2 * splits + 512);
}

// Real-valued-input convolution.  groups == 1 and Cout % 64 == 0: im2col (bf16 hi[/lo]) + tcgen05 GEMM with
// `w_packed` ([splits][Cout][Kpad] bf16); otherwise the SIMT kernel with `w_simt` ([kh][kw][Ci/g][Co] fp32).
extern "C" int ecsy_real_conv_fwd(const float* x, int64_t x_imgs, const void* w_packed, const float* w_simt, int splits,
                                  const float* bias, float bias_mul, const float* scale, const float* shift, float* out,
                                  int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad, int groups,
                                  void* ws, size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(x && out && imgs > 0 && x_imgs > 0 && imgs % x_imgs == 0, "real_conv_fwd: bad arguments");
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "real_conv_fwd: scale/shift pair");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0, "real_conv_fwd: empty output");
  if (groups == 1 && Cout % 64 == 0 && w_packed != nullptr && bias == nullptr && Cin % 64 == 0 && stride == 1) {
    const size_t need = ecsy_real_conv_ws_bytes(imgs, H, W, Cin, Cout, k, stride, pad, groups, splits);
    if (ws == nullptr || ws_bytes < need) {
      ecsy_set_error("real_conv_fwd: workspace %zu < %zu bytes", ws_bytes, need);
      return ECSY_ERR_WS;
    }
    ECSY_CHECK_ARG(x_imgs == imgs, "real_conv_fwd: broadcast inputs are convolved once by the caller");
    const int64_t n = imgs * H * W * (int64_t)Cin;
    uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* a_lo = splits == 2 ? a_hi + n : nullptr;
    int rc = ecsy_launch_f32_to_bf16(x, a_hi, a_lo, n, STREAM(stream));
    if (rc) return rc;
    return ecsy_umma_conv_bf16(a_hi, a_lo, w_packed, splits, out, scale, shift, nullptr, 0, (int)imgs, H, W, Cin, Cout, k,
                               pad, STREAM(stream));
  }
  if (use_gather_conv(Cin, Cout, k, groups, splits) && w_packed != nullptr && bias == nullptr)
    // the stem: operand tiles gathered on the fly, no im2col matrix in HBM
    return ecsy_umma_conv_gather(x, x_imgs, w_packed, out, scale, shift, (int)imgs, H, W, Cin, Cout, k, stride, pad,
                                 STREAM(stream));
  if (groups == 1 && Cout % 64 == 0 && w_packed != nullptr && bias == nullptr) {
    const int Kpad = (k * k * Cin + 63) / 64 * 64;
    const int64_t M = imgs * Ho * Wo;
    const size_t need = ecsy_real_conv_ws_bytes(imgs, H, W, Cin, Cout, k, stride, pad, groups, splits);
    if (ws == nullptr || ws_bytes < need) {
      ecsy_set_error("real_conv_fwd: workspace %zu < %zu bytes", ws_bytes, need);
      return ECSY_ERR_WS;
    }
    uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* a_lo = splits == 2 ? a_hi + M * Kpad : nullptr;
    ECSY_CHECK_ARG(Kpad <= 8 * 1024, "real_conv_fwd: K=%d too large for the im2col kernel", Kpad);
    launch_im2col(x, x_imgs, a_hi, a_lo, imgs, H, W, Cin, Ho, Wo, k, stride, pad, Kpad, STREAM(stream));
    ECSY_LAUNCH_CHECK();
    return ecsy_umma_dense(a_hi, a_lo, M, Kpad, w_packed, splits, out, Cout, scale, shift, nullptr, 0, STREAM(stream));
  }
  ECSY_CHECK_ARG(w_simt != nullptr, "real_conv_fwd: this shape needs the SIMT weight layout");
  ECSY_CHECK_ARG(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "real_conv_fwd: groups");
  if (groups == 1 && bias == nullptr && Cin <= 8 && Cout % 32 == 0 && (size_t)k * k * Cin * Cout * 4 <= 96 * 1024) {
    // few input channels, many outputs (the stem in parity precision): register-tiled fp32 kernel
    const size_t smem = (size_t)k * k * Cin * Cout * 4;
    static size_t attr = 0;
    if (smem > attr) {
      ECSY_CUDA(cudaFuncSetAttribute(k_stem_conv_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    const int64_t units = ((imgs * Ho * Wo + 31) / 32) * (Cout / 32);
    int64_t grid = (units + 7) / 8;
    if (grid > (int64_t)ecsy_num_sms() * 4) grid = (int64_t)ecsy_num_sms() * 4;
    k_stem_conv_f32<<<(int)grid, 256, smem, STREAM(stream)>>>(x, x_imgs, w_simt, scale, shift, out, imgs, H, W, Cin, Ho, Wo,
                                                              Cout, k, k, stride, pad);
    ECSY_LAUNCH_CHECK();
    return ECSY_OK;
  }
  const int64_t total = imgs * Ho * Wo * Cout;
  k_simt_conv<<<grid_for(total, kThreads, ecsy_num_sms() * 16), kThreads, 0, STREAM(stream)>>>(
      x, x_imgs, w_simt, bias, bias_mul, scale, shift, out, imgs, H, W, Cin, Ho, Wo, Cout, k, k, stride, pad, groups);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// ---- backward of Snn_Conv2d (autograd of F.conv2d, models/common.py:623) ------------------------------------
static inline size_t al256c(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" size_t ecsy_conv_dgrad_ws_bytes(int64_t imgs, int H, int W, int Cout, int k, int stride, int pad,
                                           int splits) {
  const int Hu = H - k + 1 + 2 * pad, Wu = W - k + 1 + 2 * pad;
  (void)stride;
  return 512 + static_cast<size_t>(splits) * al256c(static_cast<size_t>(imgs) * Hu * Wu * Cout * 2);
}

// Input gradient: gx[imgs][H][W][Cin] = conv_transpose(gy[imgs][Ho][Wo][Cout], W).  wT_packed is
// ecsy_pack_conv_weight of the flipped, transposed weight W'[ci][co][ky][kx] = W[co][ci][k-1-ky][k-1-kx].
extern "C" int ecsy_conv_dgrad(const float* gy, const void* wT_packed, int splits, float* gx, int64_t imgs, int H, int W,
                               int Cin, int Cout, int k, int stride, int pad, void* ws, size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(gy && wT_packed && gx && imgs > 0, "conv_dgrad: bad arguments");
  ECSY_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0, "conv_dgrad: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "conv_dgrad: splits");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int Hu = H - k + 1 + 2 * pad, Wu = W - k + 1 + 2 * pad;   // zero-inserted gradient extent
  ECSY_CHECK_ARG(Hu >= (Ho - 1) * stride + 1 && Wu >= (Wo - 1) * stride + 1, "conv_dgrad: geometry");
  const size_t need = ecsy_conv_dgrad_ws_bytes(imgs, H, W, Cout, k, stride, pad, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("conv_dgrad: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  const size_t n = static_cast<size_t>(imgs) * Hu * Wu * Cout;
  uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* lo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(base + al256c(n * 2)) : nullptr;
  if (stride == 1 && Hu == Ho && Wu == Wo) {
    int rc = ecsy_launch_f32_to_bf16(gy, hi, lo, (int64_t)n, STREAM(stream));
    if (rc) return rc;
  } else {
    k_dilate_to_bf16<<<grid_for((int64_t)n / 4, kThreads, ecsy_num_sms() * 8), kThreads, 0, STREAM(stream)>>>(
        gy, hi, lo, imgs, Ho, Wo, Cout, Hu, Wu, stride);
    ECSY_LAUNCH_CHECK();
  }
  return ecsy_umma_conv_bf16(hi, lo, wT_packed, splits, gx, nullptr, nullptr, nullptr, 0, (int)imgs, Hu, Wu, Cout, Cin, k,
                             k - 1 - pad, STREAM(stream));
}

extern "C" size_t ecsy_spike_conv_wgrad_ws_bytes(int64_t imgs, int Ho, int Wo, int Cout, int splits) {
  return 512 + static_cast<size_t>(splits) * al256c(static_cast<size_t>(imgs) * Ho * Wo * Cout * 2);
}

// Weight gradient of a spike conv: dw[Cout][(ky*kw+kx)*Cin + ci] += sum_pixels gy * spikes (accumulated).
extern "C" int ecsy_spike_conv_wgrad(const float* gy, const uint32_t* spikes, float* dw, int64_t imgs, int H, int W,
                                     int Cin, int Cout, int k, int stride, int pad, int splits, void* ws,
                                     size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(gy && spikes && dw && imgs > 0, "spike_conv_wgrad: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t need = ecsy_spike_conv_wgrad_ws_bytes(imgs, Ho, Wo, Cout, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("spike_conv_wgrad: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  const size_t n = static_cast<size_t>(imgs) * Ho * Wo * Cout;
  uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* lo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(base + al256c(n * 2)) : nullptr;
  int rc = ecsy_launch_f32_to_bf16(gy, hi, lo, (int64_t)n, STREAM(stream));
  if (rc) return rc;
  return ecsy_umma_spike_wgrad(hi, lo, spikes, dw, (int)imgs, H, W, Cin, Cout, k, stride, pad, STREAM(stream));
}

// Output gradient of a conv in tensor-core form, produced ONCE for both the weight- and the input-gradient pass:
// v = A[c]*g + B[c]*y + Cv[c] (the tdBN backward folded in: g is the gradient w.r.t. the normalised output, y the raw
// conv output; A == NULL: v = g), written as bf16 hi (+ lo) planes [imgs][Hu][Wu][C], zero-inserted when the conv was
// strided (s > 1 or Hu != Ho).  Replaces affine_add (fp32 g_y round trip) + two separate fp32 -> bf16 passes.
__global__ void k_gy_to_bf16(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ A,
                             const float* __restrict__ B, const float* __restrict__ Cv, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, int64_t imgs, int Ho, int Wo, int C, int Hu, int Wu, int s) {
  const int c4 = C >> 2;
  const int64_t total = imgs * Hu * Wu * c4;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const bool small = total < (int64_t(1) << 32);
  const bool plain = s == 1 && Hu == Ho && Wu == Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += gs) {
    const int q = static_cast<int>(ecsy::mod_u(i, (uint32_t)c4, small));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t src = i;
    bool have = true;
    if (!plain) {
      int64_t p = ecsy::div_u(i, (uint32_t)c4, small);
      const int wu = static_cast<int>(ecsy::mod_u(p, (uint32_t)Wu, small));
      p = ecsy::div_u(p, (uint32_t)Wu, small);
      const int hu = static_cast<int>(ecsy::mod_u(p, (uint32_t)Hu, small));
      const int64_t img = ecsy::div_u(p, (uint32_t)Hu, small);
      have = hu % s == 0 && wu % s == 0 && hu / s < Ho && wu / s < Wo;
      src = ((img * Ho + hu / s) * Wo + wu / s) * (int64_t)c4 + q;
    }
    if (have) {
      v = reinterpret_cast<const float4*>(g)[src];
      if (A != nullptr) {
        const float4 yv = reinterpret_cast<const float4*>(y)[src];
        const float4 a = *reinterpret_cast<const float4*>(A + q * 4);
        const float4 b = *reinterpret_cast<const float4*>(B + q * 4);
        const float4 c = *reinterpret_cast<const float4*>(Cv + q * 4);
        v.x = fmaf(a.x, v.x, fmaf(b.x, yv.x, c.x)); v.y = fmaf(a.y, v.y, fmaf(b.y, yv.y, c.y));
        v.z = fmaf(a.z, v.z, fmaf(b.z, yv.z, c.z)); v.w = fmaf(a.w, v.w, fmaf(b.w, yv.w, c.w));
      }
    }
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if (lo != nullptr) {
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
  }
}

// The un-strided case (every 3x3 / stride-1 conv of the backbone) of k_gy_to_bf16: blockDim is a multiple of C/4, so a thread
// keeps one channel quad -- coefficients in registers, no index division -- and has two rows of loads in flight.
__global__ void __launch_bounds__(256)
k_gy_to_bf16_plain(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ A,
                   const float* __restrict__ B, const float* __restrict__ Cv, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo, int64_t rows, int C) {
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4, ty = threadIdx.x / c4, nty = blockDim.x / c4;
  float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f), c = b;
  if (A != nullptr) {
    a = *reinterpret_cast<const float4*>(A + tq * 4);
    b = *reinterpret_cast<const float4*>(B + tq * 4);
    c = *reinterpret_cast<const float4*>(Cv + tq * 4);
  }
  const int64_t rstep = (int64_t)gridDim.x * nty;
  auto emit = [&](int64_t i, float4 v, float4 yv) {
    if (A != nullptr) {
      v.x = fmaf(a.x, v.x, fmaf(b.x, yv.x, c.x)); v.y = fmaf(a.y, v.y, fmaf(b.y, yv.y, c.y));
      v.z = fmaf(a.z, v.z, fmaf(b.z, yv.z, c.z)); v.w = fmaf(a.w, v.w, fmaf(b.w, yv.w, c.w));
    }
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if (lo != nullptr) {
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
  };
  int64_t r = (int64_t)blockIdx.x * nty + ty;
  for (; r + rstep < rows; r += 2 * rstep) {
    const int64_t i0 = r * c4 + tq, i1 = (r + rstep) * c4 + tq;
    const float4 v0 = ecsy::ldg_stream(reinterpret_cast<const float4*>(g) + i0);
    const float4 v1 = ecsy::ldg_stream(reinterpret_cast<const float4*>(g) + i1);
    float4 y0 = v0, y1 = v1;
    if (A != nullptr) {
      y0 = ecsy::ldg_stream(reinterpret_cast<const float4*>(y) + i0);
      y1 = ecsy::ldg_stream(reinterpret_cast<const float4*>(y) + i1);
    }
    emit(i0, v0, y0);
    emit(i1, v1, y1);
  }
  if (r < rows) {
    const int64_t i0 = r * c4 + tq;
    const float4 v0 = ecsy::ldg_stream(reinterpret_cast<const float4*>(g) + i0);
    float4 y0 = v0;
    if (A != nullptr) y0 = ecsy::ldg_stream(reinterpret_cast<const float4*>(y) + i0);
    emit(i0, v0, y0);
  }
}

static void launch_gy_plain(const float* g, const float* y, const float* A, const float* B, const float* Cv, __nv_bfloat16* hi,
                            __nv_bfloat16* lo, int64_t rows, int C, cudaStream_t st) {
  const int c4 = C / 4;
  const int bd = (256 / c4) * c4;
  const int nty = bd / c4;
  int64_t blocks = (rows + 2 * nty - 1) / (2 * nty);
  const int64_t cap = (int64_t)ecsy_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_gy_to_bf16_plain<<<(int)blocks, bd, 0, st>>>(g, y, A, B, Cv, hi, lo, rows, C);
}

// Zero-inserted copy of the bf16 gradient planes for the dgrad of a strided conv: dst [imgs][Hu][Wu][C] (already zeroed) gets
// src [imgs][Ho][Wo][C] at (ho * s, wo * s).  The planes were formed once by k_gy_to_bf16_plain (tdBN backward folded in): the
// strided case used to recompute A g + B y + C per UP-SAMPLED element with three 64-bit divisions each (0.9 ms per launch).
__global__ void __launch_bounds__(256)
k_bf16_scatter_up(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t items, int Ho, int Wo, int Hu, int Wu, int c8,
                  int s) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool small = items < (int64_t(1) << 32);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += stride) {
    const int q = static_cast<int>(ecsy::mod_u(i, (uint32_t)c8, small));
    int64_t p = ecsy::div_u(i, (uint32_t)c8, small);
    const int wo = static_cast<int>(ecsy::mod_u(p, (uint32_t)Wo, small));
    p = ecsy::div_u(p, (uint32_t)Wo, small);
    const int ho = static_cast<int>(ecsy::mod_u(p, (uint32_t)Ho, small));
    const int64_t img = ecsy::div_u(p, (uint32_t)Ho, small);
    dst[((img * Hu + (int64_t)ho * s) * Wu + (int64_t)wo * s) * c8 + q] = src[i];
  }
}

static bool dgrad_plain(int H, int W, int k, int stride, int pad) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  return stride == 1 && H - k + 1 + 2 * pad == Ho && W - k + 1 + 2 * pad == Wo;
}

extern "C" size_t ecsy_spike_conv_bwd_ws_bytes(int64_t imgs, int H, int W, int Cout, int k, int stride, int pad,
                                               int splits) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int Hu = H - k + 1 + 2 * pad, Wu = W - k + 1 + 2 * pad;
  size_t b = 1024 + static_cast<size_t>(splits) * al256c(static_cast<size_t>(imgs) * Ho * Wo * Cout * 2);
  if (!dgrad_plain(H, W, k, stride, pad)) b += static_cast<size_t>(splits) * al256c(static_cast<size_t>(imgs) * Hu * Wu * Cout * 2);
  return b;
}

// Backward of Snn_Conv2d on spikes followed by tdBN (training), one call: the output gradient is formed once as bf16
// planes (tdBN backward folded in, see k_gy_to_bf16) and feeds both the weight gradient (accumulated into dw
// [Cout][k*k*Cin]) and the input gradient gx [imgs][H][W][Cin].
extern "C" int ecsy_spike_conv_bwd(const float* g, const float* y, const float* A, const float* B, const float* Cv,
                                   const uint32_t* spikes, const void* wT_packed, int splits, float* gx, float* dw,
                                   int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad, void* ws,
                                   size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(g && spikes && wT_packed && gx && dw && imgs > 0, "spike_conv_bwd: bad arguments");
  ECSY_CHECK_ARG((A == nullptr) == (B == nullptr) && (A == nullptr) == (Cv == nullptr) && (A == nullptr || y != nullptr),
                 "spike_conv_bwd: the tdBN-backward coefficients A, B, C and y come together");
  ECSY_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0, "spike_conv_bwd: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "spike_conv_bwd: splits");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int Hu = H - k + 1 + 2 * pad, Wu = W - k + 1 + 2 * pad;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0 && Hu >= (Ho - 1) * stride + 1 && Wu >= (Wo - 1) * stride + 1, "spike_conv_bwd: geometry");
  const size_t need = ecsy_spike_conv_bwd_ws_bytes(imgs, H, W, Cout, k, stride, pad, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("spike_conv_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  cudaStream_t st = STREAM(stream);
  const size_t n = static_cast<size_t>(imgs) * Ho * Wo * Cout;
  uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base); base += al256c(n * 2);
  __nv_bfloat16* lo = nullptr;
  if (splits == 2) { lo = reinterpret_cast<__nv_bfloat16*>(base); base += al256c(n * 2); }
  if (Cout <= 1024)
    launch_gy_plain(g, y, A, B, Cv, hi, lo, imgs * Ho * Wo, Cout, st);
  else
    k_gy_to_bf16<<<grid_for((int64_t)n / 4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(g, y, A, B, Cv, hi, lo, imgs, Ho,
                                                                                           Wo, Cout, Ho, Wo, 1);
  ECSY_LAUNCH_CHECK();
  int rc = ecsy_umma_spike_wgrad(hi, lo, spikes, dw, (int)imgs, H, W, Cin, Cout, k, stride, pad, st);
  if (rc) return rc;
  __nv_bfloat16 *dhi = hi, *dlo = lo;
  if (!dgrad_plain(H, W, k, stride, pad)) {
    const size_t nu = static_cast<size_t>(imgs) * Hu * Wu * Cout;
    dhi = reinterpret_cast<__nv_bfloat16*>(base); base += al256c(nu * 2);
    dlo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(base) : nullptr;
    if (Cout <= 1024) {
      // zero planes + scatter of the values formed above (2 B per element read, full 16-byte rows written)
      const int64_t items = (int64_t)(n / 8);
      const int sgrid = grid_for(items, kThreads, ecsy_num_sms() * 8);
      ECSY_CUDA(cudaMemsetAsync(dhi, 0, nu * 2, st));
      k_bf16_scatter_up<<<sgrid, kThreads, 0, st>>>(reinterpret_cast<const uint4*>(hi), reinterpret_cast<uint4*>(dhi), items, Ho,
                                                    Wo, Hu, Wu, Cout / 8, stride);
      ECSY_LAUNCH_CHECK();
      if (dlo != nullptr) {
        ECSY_CUDA(cudaMemsetAsync(dlo, 0, nu * 2, st));
        k_bf16_scatter_up<<<sgrid, kThreads, 0, st>>>(reinterpret_cast<const uint4*>(lo), reinterpret_cast<uint4*>(dlo), items,
                                                      Ho, Wo, Hu, Wu, Cout / 8, stride);
        ECSY_LAUNCH_CHECK();
      }
    } else {
      k_gy_to_bf16<<<grid_for((int64_t)nu / 4, kThreads, ecsy_num_sms() * 8), kThreads, 0, st>>>(g, y, A, B, Cv, dhi, dlo, imgs,
                                                                                              Ho, Wo, Cout, Hu, Wu, stride);
      ECSY_LAUNCH_CHECK();
    }
  }
  return ecsy_umma_conv_bf16(dhi, dlo, wT_packed, splits, gx, nullptr, nullptr, nullptr, 0, (int)imgs, Hu, Wu, Cout, Cin, k,
                             k - 1 - pad, st);
}

extern "C" size_t ecsy_real_conv_wgrad_ws_bytes(int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                                int splits) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t M = static_cast<size_t>(imgs) * Ho * Wo;
  const size_t Kpad = (static_cast<size_t>(k) * k * Cin + 63) / 64 * 64;
  return 1024 + static_cast<size_t>(splits) * (al256c(M * Kpad * 2) + al256c(M * Cout * 2));
}

// Weight gradient of a real-input conv (stem Conv_1, class Conv): dw[Cout][Kpad] += gy^T * im2col(x).
extern "C" int ecsy_real_conv_wgrad(const float* gy, const float* x, int64_t x_imgs, float* dw, int64_t imgs, int H,
                                    int W, int Cin, int Cout, int k, int stride, int pad, int splits, void* ws,
                                    size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(gy && x && dw && imgs > 0 && x_imgs > 0 && imgs % x_imgs == 0, "real_conv_wgrad: bad arguments");
  ECSY_CHECK_ARG(Cout % 64 == 0, "real_conv_wgrad: Cout=%d must be a multiple of 64", Cout);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int Kpad = (k * k * Cin + 63) / 64 * 64;
  const int64_t M = imgs * Ho * Wo;
  const size_t need = ecsy_real_conv_wgrad_ws_bytes(imgs, H, W, Cin, Cout, k, stride, pad, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("real_conv_wgrad: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  ECSY_CHECK_ARG(Kpad * 8 <= 48 * 1024, "real_conv_wgrad: K=%d too large for the im2col table", Kpad);
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256c((size_t)M * Kpad * 2);
  __nv_bfloat16* a_lo = nullptr;
  if (splits == 2) { a_lo = reinterpret_cast<__nv_bfloat16*>(p); p += al256c((size_t)M * Kpad * 2); }
  __nv_bfloat16* g_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256c((size_t)M * Cout * 2);
  __nv_bfloat16* g_lo = splits == 2 ? reinterpret_cast<__nv_bfloat16*>(p) : nullptr;
  launch_im2col(x, x_imgs, a_hi, a_lo, imgs, H, W, Cin, Ho, Wo, k, stride, pad, Kpad, STREAM(stream));
  ECSY_LAUNCH_CHECK();
  int rc = ecsy_launch_f32_to_bf16(gy, g_hi, g_lo, M * Cout, STREAM(stream));
  if (rc) return rc;
  return ecsy_umma_xty(g_hi, g_lo, a_hi, a_lo, M, Cout, Kpad, 1.0f, dw, STREAM(stream));
}
