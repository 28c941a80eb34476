// Surrogate-gradient BPTT through the ECS-LIF neuron (backward of mem_update, models/common.py:236-309 with
// ActFun.backward :66-79; derivation in SURVEY.md Appendix B).  Membranes m_t and ECS traces e_t are NOT kept
// from the training forward: the caller re-runs ecsy_lif_ecs_fwd with mem_save / ecs_save right before this
// call (recompute instead of store), so only the layer input and the spike bits live across the step.
//
// Reverse scan t = T-1 .. 0 with carries gm = dL/dm_{t+1}, ge = dL/de_{t+1}:
//   ge_t = gm*beta*(1 - tanh(e_t)^2) + kappa*ge                                  k_lif_bwd_pre      (streaming)
//   G1   = ge_t * Wpw                      (point-wise spread, transposed)        k_umma_gemm        (tcgen05)
//   dWpw += alpha * ge_t^T * dw(s_t)       (contraction over pixels)              k_umma_xty         (tcgen05, MN-major)
//   db_pw, db_dw, dWdw += alpha * per-channel pixel sums                          k_lif_bwd_reduce   (streaming)
//   gs_t = gout_t + alpha * dw^T(G1);  gm_t = gs_t*sigma'(m_t) + gm*decay*(1 - s_t);  gx_t = gm_t
//                                                                                 k_lif_bwd_post     (streaming)
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPartBlocksPerSm = 6;   // grid of the streaming backward kernels that leave per-block gradient partials

inline int grid_for(int64_t work, int per_block, int max_blocks) {
  int64_t g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
         ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// 1 - tanh^2(v) = r (2 - r) with r = 2 / (exp(2v) + 1) = 1 - tanh(v): ex2.approx + rcp.approx + three FP32 ops, no
// cancellation near saturation (tanh -> +-1: r -> 0 or 2, the product -> 0), absolute error ~1e-7.  tanhf + 1 - th*th was ~25
// instructions per element, a quarter of k_lif_bwd_post's instruction stream (the kernel is issue bound).
__device__ __forceinline__ float sech2_fast(float v) {
  const float r = __fdividef(2.f, __expf(2.f * v) + 1.f);
  return r * (2.f - r);
}

// ge_t = gm_next*beta*(1-tanh^2(e_t)) + kappa*ge_next  -> fp32 carry (in place) + bf16 hi/lo GEMM operand
// With `part` (spiking neuron): also the per-block partial of sum_p ge_t[p][c] (row 0 of part[block][11][C], the point-wise
// bias gradient) -- blockDim is a multiple of C/4, so a thread keeps one channel quad for its whole grid-stride loop.
__global__ void k_lif_bwd_pre(const float* __restrict__ gm_next, const float* __restrict__ ecs_t,
                              float* __restrict__ ge /*in: ge_next (if has_next), out: ge_t*/, int has_next,
                              __nv_bfloat16* __restrict__ ge_hi, __nv_bfloat16* __restrict__ ge_lo, int64_t n4,
                              float beta, float kappa, float* __restrict__ part, int C) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(gm_next)[i];
    const float4 e = reinterpret_cast<const float4*>(ecs_t)[i];
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_next) n = reinterpret_cast<const float4*>(ge)[i];
    float4 o;
    o.x = g.x * beta * sech2_fast(e.x) + kappa * n.x;
    o.y = g.y * beta * sech2_fast(e.y) + kappa * n.y;
    o.z = g.z * beta * sech2_fast(e.z) + kappa * n.z;
    o.w = g.w * beta * sech2_fast(e.w) + kappa * n.w;
    reinterpret_cast<float4*>(ge)[i] = o;
    sum[0] += o.x; sum[1] += o.y; sum[2] += o.z; sum[3] += o.w;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(o.z), h3 = __float2bfloat16_rn(o.w);
    reinterpret_cast<uint2*>(ge_hi)[i] =
        make_uint2((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16),
                   (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16));
    if (ge_lo != nullptr)
      reinterpret_cast<uint2*>(ge_lo)[i] =
          make_uint2(pack_bf16x2(o.x - __bfloat162float(h0), o.y - __bfloat162float(h1)),
                     pack_bf16x2(o.z - __bfloat162float(h2), o.w - __bfloat162float(h3)));
  }
  if (part == nullptr) return;
  extern __shared__ float sred[];  // [nty][C]
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4, ty = threadIdx.x / c4, nty = blockDim.x / c4;
#pragma unroll
  for (int k = 0; k < 4; ++k) sred[ty * C + tq * 4 + k] = sum[k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < C; idx += blockDim.x) {
    double t = 0;
    for (int y = 0; y < nty; ++y) t += sred[y * C + idx];
    part[(size_t)blockIdx.x * 11 * C + idx] = static_cast<float>(t);
  }
}

// Per-channel pixel sums for the spread parameter gradients: acc[0][c] = sum ge, acc[1][c] = sum G1,
// acc[2+tap][c] = sum_p G1[p][c] * s_t[p + off(tap)][c].  Per-block partial sums [grid][11][C], reduced in double by the final kernel.
__global__ void k_lif_bwd_reduce(const float* __restrict__ ge, const float* __restrict__ g1,
                                 const uint32_t* __restrict__ bits, float* __restrict__ part, int N, int H, int W,
                                 int C) {
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4;         // channel quad (blockDim is a multiple of c4)
  const int ty = threadIdx.x / c4, nty = blockDim.x / c4;
  const int64_t pixels = (int64_t)N * H * W;
  float s[11][4];
#pragma unroll
  for (int a = 0; a < 11; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[a][k] = 0.f;
  const int64_t per_block = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per_block;
  const int64_t p1 = p0 + per_block < pixels ? p0 + per_block : pixels;
  const int Cw = C >> 5;
  // (h, w) of the thread's pixel advance incrementally (64-bit div/mod per pixel was most of this kernel's time)
  int w = 0, h = 0;
  if (p0 + ty < p1) {
    const int64_t pp = p0 + ty;
    w = static_cast<int>(pp % W);
    h = static_cast<int>((pp / W) % H);
  }
  const int dw_ = nty % W, dh_ = (nty / W) % H;
  for (int64_t p = p0 + ty; p < p1; p += nty) {
    const float4 a = reinterpret_cast<const float4*>(ge + p * C)[tq];
    const float4 b = reinterpret_cast<const float4*>(g1 + p * C)[tq];
    s[0][0] += a.x; s[0][1] += a.y; s[0][2] += a.z; s[0][3] += a.w;
    s[1][0] += b.x; s[1][1] += b.y; s[1][2] += b.z; s[1][3] += b.w;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww < 0 || ww >= W) continue;
        const uint32_t word = bits[(p + (int64_t)(ky - 1) * W + (kx - 1)) * Cw + (tq >> 3)];
        const uint32_t nib = (word >> (4 * (tq & 7))) & 0xFu;
        const int t = 2 + ky * 3 + kx;
        if (nib & 1u) s[t][0] += b.x;
        if (nib & 2u) s[t][1] += b.y;
        if (nib & 4u) s[t][2] += b.z;
        if (nib & 8u) s[t][3] += b.w;
      }
    }
    w += dw_;
    h += dh_;
    if (w >= W) { w -= W; ++h; }
    if (h >= H) h -= H;
  }
  // block reduction over the ty rows in shared memory, then one double atomic per (quantity, channel)
  extern __shared__ float sred[];  // [nty][11][C]
#pragma unroll
  for (int a = 0; a < 11; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) sred[(ty * 11 + a) * C + tq * 4 + k] = s[a][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 11 * C; idx += blockDim.x) {
    double t = 0;
    for (int y = 0; y < nty; ++y) t += sred[y * 11 * C + idx];
    part[(size_t)blockIdx.x * 11 * C + idx] = static_cast<float>(t);   // per-block partial: no atomics (592 blocks x 11C
                                                                        // double atomics on 11C addresses serialised in L2)
  }
}

// outputs += alpha * acc ; layouts: g_pw_b [C], g_dw_b [C], g_dw_w [9][C].  A block reduces 32 consecutive outputs: 8 slices
// of the per-block partials are summed in parallel (coalesced 128-byte rows, independent loads) and combined in a fixed
// order -- deterministic, and ~8x shorter than one thread walking all ~600 partials of its output.
__global__ void __launch_bounds__(256)
k_lif_bwd_reduce_final(const float* __restrict__ part, const float* __restrict__ part0, int nblocks0, int nblocks,
                       float* __restrict__ g_pw_b, float* __restrict__ g_dw_b, float* __restrict__ g_dw_w, int C, float alpha) {
  // row 0 (sum ge) holds nblocks0 partials -- in part0 [nblocks0][C] when the ge came out of k_lif_bwd_post(t+1), else in
  // row 0 of part (k_lif_bwd_pre) --, rows 1..10 nblocks (grid of k_lif_bwd_post)
  __shared__ double sm[8][32];
  const int ib = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ib;
  const int n_out = 11 * C;
  double t = 0;
  if (i < n_out) {
    if (i < C && part0 != nullptr) {
      for (int b = sl; b < nblocks0; b += 8) t += part0[(size_t)b * C + i];
    } else {
      const int nb = i < C ? nblocks0 : nblocks;
      for (int b = sl; b < nb; b += 8) t += part[(size_t)b * n_out + i];
    }
  }
  sm[sl][ib] = t;
  __syncthreads();
  if (sl != 0 || i >= n_out) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) t += sm[k][ib];
  const int a = i / C, c = i - a * C;
  const float v = alpha * static_cast<float>(t);
  if (a == 0) g_pw_b[c] += v;
  else if (a == 1) g_dw_b[c] += v;
  else g_dw_w[(a - 2) * C + c] += v;
}

// k_lif_bwd_post(t) can also emit the ECS-trace gradient of the step BEFORE it (what k_lif_bwd_pre(t-1) computed from the
// gx[t] this kernel has just formed): ge_{t-1} = gx_t * beta * (1 - tanh(e_{t-1})^2) + kappa * ge_t, its bf16 GEMM operand and
// its per-channel sum -- gx[t] is not re-read, one launch less per timestep.
struct PostEmit {
  const float* ecs_prev;   // e_{t-1}; NULL = do not emit
  float* ge;               // in: ge_t (if ge_has_next), out: ge_{t-1}
  __nv_bfloat16* ge_hi;
  __nv_bfloat16* ge_lo;    // parity precision: residual plane
  float* part0;            // per-block partial of sum_p ge_{t-1}[p][c]: [grid][C]
  int ge_has_next;
  float beta, kappa;
};

// gs = gout + alpha*dw^T(G1); gm = gs*sigma'(m_t) + gm_next*decay*(1-s_t); writes gm carry and gx[t].
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_lif_bwd_post(const float* __restrict__ gout, const float* __restrict__ g1 /*null at t=T-1*/,
                               const float* __restrict__ dw_w, const float* __restrict__ mem_t,
                               const uint32_t* __restrict__ bits_t, const float* __restrict__ gm /*dL/dm_{t+1} = gx[t+1]*/,
                               int has_next, float* __restrict__ gx, int N, int H, int W, int C, float thresh,
                               float lens, float decay, float alpha, float* __restrict__ part, const PostEmit em) {
  // blockDim is a multiple of C/4: a thread keeps ONE channel quad and walks pixels p = block*nty + ty + k*grid*nty with
  // (h, w) tracked incrementally -- no division, no 64-bit multiply per element (the index arithmetic of the first version
  // was ~580 instructions per float4 and the kernel was issue bound at a third of the HBM rate).
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4, ty = threadIdx.x / c4, nty = blockDim.x / c4;
  const int64_t pixels = (int64_t)N * H * W;
  const int64_t pstep = (int64_t)gridDim.x * nty;
  const float inv = 1.0f / (2.0f * lens);
  // spread parameter gradients (with `part`): the depth-wise dgrad below already gathers G1 at the nine neighbours
  // p - off(tap) of pixel p, and
  //   dWdw[tap][c] = sum_p' G1[p'][c] * s_t[p' + off(tap)][c] = sum_p s_t[p][c] * G1[p - off(tap)][c]
  // is the same nine values gated by the spike of the centre pixel: rows 2..10 of part[block][11][C]; row 1 = sum_p G1[p][c].
  float aw[9][4], ab[4] = {0.f, 0.f, 0.f, 0.f}, s0[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) aw[a][k] = 0.f;
  int64_t p = (int64_t)blockIdx.x * nty + ty;
  int w = static_cast<int>(p % W), h = static_cast<int>((p / W) % H);
  const int dw_ = static_cast<int>(pstep % W), dh_ = static_cast<int>((pstep / W) % H);
  const int64_t estep = pstep * C;
  const int WC = W * C;
  int64_t e = p * C + tq * 4;
  const int Cw = C >> 5, sh = 4 * (tq & 7);
  const uint32_t* bp = bits_t + p * Cw + (tq >> 3);
  const int64_t bstep = pstep * Cw;
  const float* wq = dw_w + tq * 4;
  const float* g1p = g1 != nullptr ? g1 + e : nullptr;
  for (; p < pixels; p += pstep, e += estep, bp += bstep, g1p += estep) {
    // every load of the iteration is issued before the first use (clamped addresses + masks instead of branches around
    // the border taps)
    float4 gs = *reinterpret_cast<const float4*>(gout + e);
    const float4 m = *reinterpret_cast<const float4*>(mem_t + e);
    float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t nib = 0;
    if (has_next) {
      gn = *reinterpret_cast<const float4*>(gm + e);
      nib = (*bp >> sh) & 0xFu;
    }
    float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), nv = ev;   // trace / carry of the emitted ge: in flight with the rest
    if (em.ecs_prev != nullptr) {
      ev = *reinterpret_cast<const float4*>(em.ecs_prev + e);
      if (em.ge_has_next) nv = *reinterpret_cast<const float4*>(em.ge + e);
    }
    if (g1 != nullptr) {
      float4 g[9];
      if ((unsigned)(h - 1) < (unsigned)(H - 2) && (unsigned)(w - 1) < (unsigned)(W - 2)) {
        // interior pixel (97 % at 160 x 160): nine loads at fixed offsets from the running pointer, no tests
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            g[ky * 3 + kx] = *reinterpret_cast<const float4*>(g1p + ((1 - ky) * WC + (1 - kx) * C));   // s[q] feeds out[q - off(tap)]
      } else {
        uint32_t okm = 0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int hh = h - (ky - 1), ww = w - (kx - 1);
            const bool ok = (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
            const int off = ok ? (1 - ky) * WC + (1 - kx) * C : 0;   // 32-bit offset from the running pointer
            g[ky * 3 + kx] = *reinterpret_cast<const float4*>(g1p + off);
            okm |= (ok ? 1u : 0u) << (ky * 3 + kx);
          }
        }
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)   // image border: the taps outside contribute nothing
          if (!((okm >> tp) & 1u)) g[tp] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float s0 = (nib & 1u) ? 1.f : 0.f, s1 = (nib & 2u) ? 1.f : 0.f;
      const float s2 = (nib & 4u) ? 1.f : 0.f, s3 = (nib & 8u) ? 1.f : 0.f;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const float4 wt = __ldg(reinterpret_cast<const float4*>(wq + tp * C));
        acc.x = fmaf(g[tp].x, wt.x, acc.x); acc.y = fmaf(g[tp].y, wt.y, acc.y);
        acc.z = fmaf(g[tp].z, wt.z, acc.z); acc.w = fmaf(g[tp].w, wt.w, acc.w);
        if (part != nullptr) {
          // g * 1 or g * 0 (+-0 leaves the sum unchanged): same values as a predicated add
          aw[tp][0] = fmaf(g[tp].x, s0, aw[tp][0]); aw[tp][1] = fmaf(g[tp].y, s1, aw[tp][1]);
          aw[tp][2] = fmaf(g[tp].z, s2, aw[tp][2]); aw[tp][3] = fmaf(g[tp].w, s3, aw[tp][3]);
          if (tp == 4) { ab[0] += g[tp].x; ab[1] += g[tp].y; ab[2] += g[tp].z; ab[3] += g[tp].w; }
        }
      }
      gs.x = fmaf(alpha, acc.x, gs.x); gs.y = fmaf(alpha, acc.y, gs.y);
      gs.z = fmaf(alpha, acc.z, gs.z); gs.w = fmaf(alpha, acc.w, gs.w);
    }
    float4 o;
    o.x = fabsf(m.x - thresh) < lens ? gs.x * inv : 0.f;
    o.y = fabsf(m.y - thresh) < lens ? gs.y * inv : 0.f;
    o.z = fabsf(m.z - thresh) < lens ? gs.z * inv : 0.f;
    o.w = fabsf(m.w - thresh) < lens ? gs.w * inv : 0.f;
    if (has_next) {
      o.x += (nib & 1u) ? 0.f : gn.x * decay;
      o.y += (nib & 2u) ? 0.f : gn.y * decay;
      o.z += (nib & 4u) ? 0.f : gn.z * decay;
      o.w += (nib & 8u) ? 0.f : gn.w * decay;
    }
    *reinterpret_cast<float4*>(gx + e) = o;   // gx[t] IS the carry dL/dm_t of the next (earlier) step: no separate copy
    if (em.ecs_prev != nullptr) {
      float4 q4;
      q4.x = o.x * em.beta * sech2_fast(ev.x) + em.kappa * nv.x;
      q4.y = o.y * em.beta * sech2_fast(ev.y) + em.kappa * nv.y;
      q4.z = o.z * em.beta * sech2_fast(ev.z) + em.kappa * nv.z;
      q4.w = o.w * em.beta * sech2_fast(ev.w) + em.kappa * nv.w;
      *reinterpret_cast<float4*>(em.ge + e) = q4;
      s0[0] += q4.x; s0[1] += q4.y; s0[2] += q4.z; s0[3] += q4.w;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(q4.x), h1 = __float2bfloat16_rn(q4.y);
      const __nv_bfloat16 h2 = __float2bfloat16_rn(q4.z), h3 = __float2bfloat16_rn(q4.w);
      *reinterpret_cast<uint2*>(em.ge_hi + e) =
          make_uint2((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16),
                     (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16));
      if (em.ge_lo != nullptr)
        *reinterpret_cast<uint2*>(em.ge_lo + e) =
            make_uint2(pack_bf16x2(q4.x - __bfloat162float(h0), q4.y - __bfloat162float(h1)),
                       pack_bf16x2(q4.z - __bfloat162float(h2), q4.w - __bfloat162float(h3)));
    }
    w += dw_;
    h += dh_;
    if (w >= W) { w -= W; ++h; }
    if (h >= H) h -= H;
  }
  extern __shared__ float sred[];  // [nty][10][C] (spread sums) / [nty][C] (sum of the emitted ge)
  if (em.ecs_prev != nullptr) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sred[ty * C + tq * 4 + k] = s0[k];
    __syncthreads();
    for (int idx = threadIdx.x; idx < C; idx += blockDim.x) {
      double t = 0;
      for (int y = 0; y < nty; ++y) t += sred[y * C + idx];
      em.part0[(size_t)blockIdx.x * C + idx] = static_cast<float>(t);
    }
    __syncthreads();
  }
  if (part == nullptr) return;
  // block reduction over the threads of a channel quad in shared memory (double), one per-block partial per (row, channel)
#pragma unroll
  for (int k = 0; k < 4; ++k) sred[(ty * 10 + 0) * C + tq * 4 + k] = ab[k];
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) sred[(ty * 10 + 1 + a) * C + tq * 4 + k] = aw[a][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 10 * C; idx += blockDim.x) {
    double t = 0;
    for (int y = 0; y < nty; ++y) t += sred[y * 10 * C + idx];
    part[(size_t)blockIdx.x * 11 * C + C + idx] = static_cast<float>(t);
  }
}

// ---- SiLU ("analog spike") neuron backward: mem_update(act=True) with the in-place SiLU of the reference
// models (mem_old = silu(mem)):  m_t = d*o_{t-1}*(1 - stopgrad(o_{t-1})) + x_t + f_{t-1},  o_t = silu(m_t).
//   go_t = gout_t + alpha*dw^T(G1) + gm*decay*(1 - o_t);  gm_t = go_t * silu'(m_t)
__global__ void k_silu_bwd_reduce(const float* __restrict__ ge, const float* __restrict__ g1,
                                  const float* __restrict__ o, float* __restrict__ part, int N, int H, int W, int C) {
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4;
  const int ty = threadIdx.x / c4, nty = blockDim.x / c4;
  const int64_t pixels = (int64_t)N * H * W;
  float s[11][4];
#pragma unroll
  for (int a = 0; a < 11; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[a][k] = 0.f;
  const int64_t per_block = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per_block;
  const int64_t p1 = p0 + per_block < pixels ? p0 + per_block : pixels;
  for (int64_t p = p0 + ty; p < p1; p += nty) {
    const float4 a = reinterpret_cast<const float4*>(ge + p * C)[tq];
    const float4 b = reinterpret_cast<const float4*>(g1 + p * C)[tq];
    s[0][0] += a.x; s[0][1] += a.y; s[0][2] += a.z; s[0][3] += a.w;
    s[1][0] += b.x; s[1][1] += b.y; s[1][2] += b.z; s[1][3] += b.w;
    const bool small = pixels < (int64_t(1) << 32);
    const int w = static_cast<int>(ecsy::mod_u(p, (uint32_t)W, small));
    const int h = static_cast<int>(ecsy::mod_u(ecsy::div_u(p, (uint32_t)W, small), (uint32_t)H, small));
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww < 0 || ww >= W) continue;
        const float4 ov = reinterpret_cast<const float4*>(o + (p + (int64_t)(ky - 1) * W + (kx - 1)) * C)[tq];
        const int t = 2 + ky * 3 + kx;
        s[t][0] += b.x * ov.x; s[t][1] += b.y * ov.y; s[t][2] += b.z * ov.z; s[t][3] += b.w * ov.w;
      }
    }
  }
  extern __shared__ float sred[];
#pragma unroll
  for (int a = 0; a < 11; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) sred[(ty * 11 + a) * C + tq * 4 + k] = s[a][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 11 * C; idx += blockDim.x) {
    double t = 0;
    for (int y = 0; y < nty; ++y) t += sred[y * 11 * C + idx];
    part[(size_t)blockIdx.x * 11 * C + idx] = static_cast<float>(t);   // per-block partial: no atomics (592 blocks x 11C
                                                                        // double atomics on 11C addresses serialised in L2)
  }
}

__device__ __forceinline__ float silu_grad(float m) {
  const float sg = 1.0f / (1.0f + expf(-m));
  return sg * (1.0f + m * (1.0f - sg));
}

__global__ void k_silu_bwd_post(const float* __restrict__ gout, const float* __restrict__ g1,
                                const float* __restrict__ dw_w, const float* __restrict__ mem_t,
                                const float* __restrict__ o_t, float* __restrict__ gm, int has_next,
                                float* __restrict__ gx, int N, int H, int W, int C, float decay, float alpha) {
  const int c4 = C >> 2;
  const int64_t total = (int64_t)N * H * W * c4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool small = total < (int64_t(1) << 32);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int q = static_cast<int>(ecsy::mod_u(i, (uint32_t)c4, small));
    const int64_t p = ecsy::div_u(i, (uint32_t)c4, small);
    float4 go = reinterpret_cast<const float4*>(gout)[i];
    if (g1 != nullptr) {
      const int w = static_cast<int>(ecsy::mod_u(p, (uint32_t)W, small));
      const int h = static_cast<int>(ecsy::mod_u(ecsy::div_u(p, (uint32_t)W, small), (uint32_t)H, small));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int hh = h - (ky - 1);
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ww = w - (kx - 1);
          if (ww < 0 || ww >= W) continue;
          const float4 g = reinterpret_cast<const float4*>(g1 + (p - (int64_t)(ky - 1) * W - (kx - 1)) * C)[q];
          const float4 wt = *reinterpret_cast<const float4*>(dw_w + (ky * 3 + kx) * C + q * 4);
          acc.x = fmaf(g.x, wt.x, acc.x); acc.y = fmaf(g.y, wt.y, acc.y);
          acc.z = fmaf(g.z, wt.z, acc.z); acc.w = fmaf(g.w, wt.w, acc.w);
        }
      }
      go.x = fmaf(alpha, acc.x, go.x); go.y = fmaf(alpha, acc.y, go.y);
      go.z = fmaf(alpha, acc.z, go.z); go.w = fmaf(alpha, acc.w, go.w);
    }
    if (has_next) {
      const float4 gn = reinterpret_cast<const float4*>(gm)[i];
      const float4 ov = reinterpret_cast<const float4*>(o_t)[i];
      go.x += gn.x * decay * (1.0f - ov.x); go.y += gn.y * decay * (1.0f - ov.y);
      go.z += gn.z * decay * (1.0f - ov.z); go.w += gn.w * decay * (1.0f - ov.w);
    }
    const float4 m = reinterpret_cast<const float4*>(mem_t)[i];
    const float4 o = make_float4(go.x * silu_grad(m.x), go.y * silu_grad(m.y), go.z * silu_grad(m.z), go.w * silu_grad(m.w));
    reinterpret_cast<float4*>(gm)[i] = o;
    reinterpret_cast<float4*>(gx)[i] = o;
  }
}

// per-channel sums over rows: sum_g[c] = sum g, sum_gx[c] = sum g*x  (tdBN / folded-affine backward)
__global__ void k_colsum2(const float* __restrict__ g, const float* __restrict__ x, int64_t rows, int64_t x_rows,
                          int C, double* __restrict__ acc /*[2][C]*/) {
  const int c4 = C >> 2;
  const int tq = threadIdx.x % c4, ty = threadIdx.x / c4, nty = blockDim.x / c4;
  const int64_t per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per_block;
  const int64_t r1 = r0 + per_block < rows ? r0 + per_block : rows;
  double sg[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};
  float fg[4] = {0, 0, 0, 0}, fx[4] = {0, 0, 0, 0};
  int cnt = 0;
  // row of x (a T-broadcast x repeats every x_rows rows) tracked incrementally: no 64-bit modulo per row
  int64_t xr = r0 + ty < r1 ? (r0 + ty) % x_rows : 0;
  const int64_t xstep = nty % x_rows;
  auto add_row = [&](const float4& a, const float4& b) {
    fg[0] += a.x; fg[1] += a.y; fg[2] += a.z; fg[3] += a.w;
    fx[0] += a.x * b.x; fx[1] += a.y * b.y; fx[2] += a.z * b.z; fx[3] += a.w * b.w;
    if (++cnt == 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { sg[k] += fg[k]; sx[k] += fx[k]; fg[k] = 0; fx[k] = 0; }
      cnt = 0;
    }
  };
  int64_t r = r0 + ty;
  for (; r + nty < r1; r += 2 * nty) {   // two rows of loads in flight per thread (same summation order as one at a time)
    int64_t xr2 = xr + xstep;
    if (xr2 >= x_rows) xr2 -= x_rows;
    const float4 a0 = ecsy::ldg_stream(reinterpret_cast<const float4*>(g + r * C) + tq);
    const float4 a1 = ecsy::ldg_stream(reinterpret_cast<const float4*>(g + (r + nty) * C) + tq);
    const float4 b0 = reinterpret_cast<const float4*>(x + xr * C)[tq];
    const float4 b1 = reinterpret_cast<const float4*>(x + xr2 * C)[tq];
    xr = xr2 + xstep;
    if (xr >= x_rows) xr -= x_rows;
    add_row(a0, b0);
    add_row(a1, b1);
  }
  if (r < r1) {
    const float4 a = reinterpret_cast<const float4*>(g + r * C)[tq];
    const float4 b = reinterpret_cast<const float4*>(x + xr * C)[tq];
    add_row(a, b);
  }
  extern __shared__ double dred[];  // [nty][2][C]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dred[(ty * 2 + 0) * C + tq * 4 + k] = sg[k] + fg[k];
    dred[(ty * 2 + 1) * C + tq * 4 + k] = sx[k] + fx[k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    double t = 0;
    for (int y = 0; y < nty; ++y) t += dred[y * 2 * C + idx];
    atomicAdd(acc + idx, t);
  }
}

__global__ void k_d2f(const double* __restrict__ a, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = static_cast<float>(a[i]);
}

}  // namespace

static inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" size_t ecsy_lif_ecs_bwd_ws_bytes(int T, int64_t N, int H, int W, int C, int splits) {
  const size_t mc = static_cast<size_t>(N) * H * W * C;
  (void)T;
  // gm, ge, G1 (fp32) + ge planes + dw planes (bf16) + reduction scratch
  return 512 + 3 * al256(mc * 4) + 2 * static_cast<size_t>(splits) * al256(mc * 2) +
         al256((size_t)ecsy_num_sms() * 8 * 13 * C * 4);   // per-block partial sums of the parameter gradients (+ 2 x sum ge)
}

extern "C" int ecsy_lif_ecs_bwd(const float* gout, const uint32_t* spikes, const float* mem, const float* ecs,
                                const float* dw_w, const float* dw_b, const void* pwT_packed, int splits, float* gx,
                                float* g_dw_w, float* g_dw_b, float* g_pw_w, float* g_pw_b, int T, int64_t N, int H,
                                int W, int C, float thresh, float lens, float decay, float alpha, float beta,
                                float kappa, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(gout && spikes && mem && gx && T >= 1 && N > 0 && H > 0 && W > 0, "lif_ecs_bwd: bad arguments");
  ECSY_CHECK_ARG(C % 64 == 0 && C <= 1024, "lif_ecs_bwd: C=%d must be a multiple of 64, <= 1024", C);
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "lif_ecs_bwd: splits must be 1 or 2");
  ECSY_CHECK_ARG(T == 1 || (ecs && dw_w && dw_b && pwT_packed && g_dw_w && g_dw_b && g_pw_w && g_pw_b),
                 "lif_ecs_bwd: spread tensors missing");
  const int64_t M = N * H * W;
  const size_t mc = static_cast<size_t>(M) * C;
  const size_t need = ecsy_lif_ecs_bwd_ws_bytes(T, N, H, W, C, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("lif_ecs_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  p += al256(mc * 4);   // (the membrane-gradient carry lived here; gx[t+1] is the carry now -- layout kept for the SiLU path)
  float* ge = reinterpret_cast<float*>(p); p += al256(mc * 4);
  float* g1 = reinterpret_cast<float*>(p); p += al256(mc * 4);
  __nv_bfloat16* ge_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2);
  __nv_bfloat16* ge_lo = nullptr;
  if (splits == 2) { ge_lo = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2); }
  __nv_bfloat16* d_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2);
  __nv_bfloat16* d_lo = nullptr;
  if (splits == 2) { d_lo = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2); }
  float* part = reinterpret_cast<float*>(p);
  const int64_t words = M * (C / 32);
  const int64_t n4 = M * C / 4;
  const int c4 = C / 4;
  const int rbd = (256 / c4) * c4;  // block size a multiple of c4 (C <= 1024 -> c4 <= 256): one channel quad per thread
  // the parameter-gradient sums ride in the streaming kernels (pre: sum ge; post: sum G1 and the spike-gated taps): both
  // run on the same grid and leave part[block][11][C]; the old separate pass (k_lif_bwd_reduce: 16 of 142 ms of a
  // resnet18 step, profiles/r02_launch_summary_train_r18_b32.txt) re-read ge, G1 and nine spike words per element
  const int egrid = grid_for(n4, rbd, ecsy_num_sms() * kPartBlocksPerSm);
  // k_lif_bwd_post<2> is resident at two blocks per SM (128 registers): one wave of 2 x SMs blocks -- no tail, and a third of
  // the per-block partials for k_lif_bwd_reduce_final to read
  static const int post_bps = getenv("ECSY_POST_BPS") ? atoi(getenv("ECSY_POST_BPS")) : 2;
  const int pgrid = grid_for(n4, rbd, ecsy_num_sms() * post_bps);

  // ECSY_LIF_BWD_V=0: the separate reduction pass of round 1 (k_lif_bwd_reduce); =1: ge from k_lif_bwd_pre (one more launch and
  // a re-read of gx per timestep); default: k_lif_bwd_post(t) emits ge_{t-1} itself.  Kept for A/B measurements.
  static const int variant = getenv("ECSY_LIF_BWD_V") ? atoi(getenv("ECSY_LIF_BWD_V")) : 2;
  const bool fused = variant != 0;
  const bool emit = variant >= 2;
  float* part0 = part + (size_t)ecsy_num_sms() * 8 * 11 * C;   // [2][pgrid][C] behind the [grid][11][C] partials
  for (int t = T - 1; t >= 0; --t) {
    const bool spread = t <= T - 2;
    const bool has_next = t < T - 1;
    if (spread) {
      if (!emit) {
        k_lif_bwd_pre<<<egrid, rbd, (size_t)(rbd / c4) * C * sizeof(float), st>>>(
            gx + (size_t)(t + 1) * mc, ecs + (size_t)t * mc, ge, t < T - 2 ? 1 : 0, ge_hi, ge_lo, n4, beta, kappa, part, C);
        ECSY_LAUNCH_CHECK();
      }
      int rc = ecsy_umma_dense(ge_hi, ge_lo, M, C, pwT_packed, splits, g1, C, nullptr, nullptr, nullptr, 0, st);
      if (rc) return rc;
      rc = ecsy_launch_spread_dw(spikes + t * words, dw_w, dw_b, d_hi, d_lo, (int)N, H, W, C, st);
      if (rc) return rc;
      rc = ecsy_umma_xty(ge_hi, ge_lo, d_hi, d_lo, M, C, C, alpha, g_pw_w, st);
      if (rc) return rc;
    }
    if (spread && !fused) {
      const int rgrid = grid_for(M, 64, ecsy_num_sms() * 4);
      k_lif_bwd_reduce<<<rgrid, rbd, (size_t)(rbd / c4) * 11 * C * sizeof(float), st>>>(ge, g1, spikes + t * words, part,
                                                                                       (int)N, H, W, C);
      ECSY_LAUNCH_CHECK();
      k_lif_bwd_reduce_final<<<(11 * C + 31) / 32, 256, 0, st>>>(part, nullptr, rgrid, rgrid, g_pw_b, g_dw_b, g_dw_w, C, alpha);
      ECSY_LAUNCH_CHECK();
    }
    PostEmit em{};
    if (emit && t >= 1) {   // this step's gx[t] -> ge_{t-1}, consumed by the GEMM / xty of the next loop iteration
      em.ecs_prev = ecs + (size_t)(t - 1) * mc;
      em.ge = ge; em.ge_hi = ge_hi; em.ge_lo = ge_lo;
      em.part0 = part0 + (size_t)((t - 1) & 1) * pgrid * C;
      em.ge_has_next = has_next ? 1 : 0;   // ge_t exists iff step t had a spread
      em.beta = beta; em.kappa = kappa;
    }
    size_t psm = (spread && fused) ? (size_t)(rbd / c4) * 10 * C * sizeof(float) : 0;
    if (em.ecs_prev != nullptr && psm < (size_t)(rbd / c4) * C * sizeof(float)) psm = (size_t)(rbd / c4) * C * sizeof(float);
    float* pp = (spread && fused) ? part : nullptr;
    k_lif_bwd_post<2><<<pgrid, rbd, psm, st>>>(   // <3> / <4> blocks per SM spill and ran 20-30 % slower
        gout + (size_t)t * mc, spread ? g1 : nullptr, dw_w, mem + (size_t)t * mc, spikes + t * words,
        has_next ? gx + (size_t)(t + 1) * mc : nullptr, has_next ? 1 : 0, gx + (size_t)t * mc, (int)N, H, W, C, thresh, lens,
        decay, alpha, pp, em);
    ECSY_LAUNCH_CHECK();
    if (spread && fused) {
      k_lif_bwd_reduce_final<<<(11 * C + 31) / 32, 256, 0, st>>>(part, emit ? part0 + (size_t)(t & 1) * pgrid * C : nullptr,
                                                                 emit ? pgrid : egrid, pgrid, g_pw_b, g_dw_b, g_dw_w, C, alpha);
      ECSY_LAUNCH_CHECK();
    }
  }
  return ECSY_OK;
}

// sum_g[c] = sum_rows g[r][c]; sum_gx[c] = sum_rows g[r][c]*x[r % x_rows][c]   (tdBN / affine backward)
extern "C" int ecsy_colsum2(const float* g, const float* x, int64_t rows, int64_t x_rows, int C, float* sum_g,
                            float* sum_gx, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(g && x && sum_g && sum_gx && rows > 0 && x_rows > 0, "colsum2: bad arguments");
  ECSY_CHECK_ARG(C % 4 == 0 && C <= 1024, "colsum2: C=%d", C);
  if (ws == nullptr || ws_bytes < 2 * (size_t)C * 8 + 256) {
    ecsy_set_error("colsum2: workspace too small");
    return ECSY_ERR_WS;
  }
  double* acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  ECSY_CUDA(cudaMemsetAsync(acc, 0, 2 * (size_t)C * sizeof(double), st));
  const int c4 = C / 4;
  const int bd = (256 / c4) * c4;
  k_colsum2<<<grid_for(rows, 64, ecsy_num_sms() * 4), bd, (size_t)(bd / c4) * 2 * C * sizeof(double), st>>>(g, x, rows,
                                                                                                        x_rows, C, acc);
  ECSY_LAUNCH_CHECK();
  k_d2f<<<(C + 255) / 256, 256, 0, st>>>(acc, sum_g, C);
  ECSY_LAUNCH_CHECK();
  k_d2f<<<(C + 255) / 256, 256, 0, st>>>(acc + C, sum_gx, C);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// Depth-wise 3x3 of a real tensor -> bf16 planes (elementwise.cu)
int ecsy_launch_dw_real(const float* s, const float* dw_w, const float* dw_b, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                        int N, int H, int W, int C, cudaStream_t st);

extern "C" int ecsy_lif_silu_bwd(const float* gout, const float* out, const float* mem, const float* ecs,
                                 const float* dw_w, const float* dw_b, const void* pwT_packed, int splits, float* gx,
                                 float* g_dw_w, float* g_dw_b, float* g_pw_w, float* g_pw_b, int T, int64_t N, int H,
                                 int W, int C, float decay, float alpha, float beta, float kappa, void* ws,
                                 size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(gout && out && mem && gx && T >= 1 && N > 0 && H > 0 && W > 0, "lif_silu_bwd: bad arguments");
  ECSY_CHECK_ARG(C % 64 == 0 && C <= 1024, "lif_silu_bwd: C=%d must be a multiple of 64, <= 1024", C);
  ECSY_CHECK_ARG(T == 1 || (ecs && dw_w && dw_b && pwT_packed && g_dw_w && g_dw_b && g_pw_w && g_pw_b),
                 "lif_silu_bwd: spread tensors missing");
  const int64_t M = N * H * W;
  const size_t mc = static_cast<size_t>(M) * C;
  const size_t need = ecsy_lif_ecs_bwd_ws_bytes(T, N, H, W, C, splits);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("lif_silu_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  float* gm = reinterpret_cast<float*>(p); p += al256(mc * 4);
  float* ge = reinterpret_cast<float*>(p); p += al256(mc * 4);
  float* g1 = reinterpret_cast<float*>(p); p += al256(mc * 4);
  __nv_bfloat16* ge_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2);
  __nv_bfloat16* ge_lo = nullptr;
  if (splits == 2) { ge_lo = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2); }
  __nv_bfloat16* d_hi = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2);
  __nv_bfloat16* d_lo = nullptr;
  if (splits == 2) { d_lo = reinterpret_cast<__nv_bfloat16*>(p); p += al256(mc * 2); }
  float* part = reinterpret_cast<float*>(p);
  const int64_t n4 = M * C / 4;
  const int egrid = grid_for(n4, kThreads, ecsy_num_sms() * 8);
  const int c4 = C / 4;
  const int rbd = (256 / c4) * c4;
  for (int t = T - 1; t >= 0; --t) {
    const bool spread = t <= T - 2;
    const bool has_next = t < T - 1;
    const float* o_t = out + (size_t)t * mc;
    if (spread) {
      k_lif_bwd_pre<<<egrid, kThreads, 0, st>>>(gx + (size_t)(t + 1) * mc, ecs + (size_t)t * mc, ge, t < T - 2 ? 1 : 0, ge_hi,
                                                ge_lo, n4, beta, kappa, nullptr, C);
      ECSY_LAUNCH_CHECK();
      int rc = ecsy_umma_dense(ge_hi, ge_lo, M, C, pwT_packed, splits, g1, C, nullptr, nullptr, nullptr, 0, st);
      if (rc) return rc;
      rc = ecsy_launch_dw_real(o_t, dw_w, dw_b, d_hi, d_lo, (int)N, H, W, C, st);
      if (rc) return rc;
      rc = ecsy_umma_xty(ge_hi, ge_lo, d_hi, d_lo, M, C, C, alpha, g_pw_w, st);
      if (rc) return rc;
      const int rgrid = grid_for(M, 64, ecsy_num_sms() * 4);
      k_silu_bwd_reduce<<<rgrid, rbd, (size_t)(rbd / c4) * 11 * C * sizeof(float), st>>>(ge, g1, o_t, part, (int)N, H, W, C);
      ECSY_LAUNCH_CHECK();
      k_lif_bwd_reduce_final<<<(11 * C + 31) / 32, 256, 0, st>>>(part, nullptr, rgrid, rgrid, g_pw_b, g_dw_b, g_dw_w, C, alpha);
      ECSY_LAUNCH_CHECK();
    }
    k_silu_bwd_post<<<egrid, kThreads, 0, st>>>(gout + (size_t)t * mc, spread ? g1 : nullptr, dw_w, mem + (size_t)t * mc,
                                                o_t, gm, has_next ? 1 : 0, gx + (size_t)t * mc, (int)N, H, W, C, decay,
                                                alpha);
    ECSY_LAUNCH_CHECK();
  }
  return ECSY_OK;
}
