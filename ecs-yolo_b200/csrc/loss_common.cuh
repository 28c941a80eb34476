// Pieces shared by the two training-loss translation units (loss.cu: Stack A, tal_loss.cu: Stack B): forward-mode dual
// numbers for box-IoU gradients, BCE-with-logits, a fixed-order block reduction.
#pragma once
#include "ecsy_common.cuh"

namespace ecsy_loss {

constexpr int kThreads = 256;

// ---- forward-mode dual numbers over the four box coordinates: value + d/d(px, py, pw, ph) -------------------
struct D4 {
  float v, d[4];
};
__device__ __forceinline__ D4 cst(float c) { return D4{c, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D4 var(float c, int i) {
  D4 r = cst(c);
  r.d[i] = 1.f;
  return r;
}
__device__ __forceinline__ D4 operator+(D4 a, D4 b) {
  return D4{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2], a.d[3] + b.d[3]}};
}
__device__ __forceinline__ D4 operator-(D4 a, D4 b) {
  return D4{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2], a.d[3] - b.d[3]}};
}
__device__ __forceinline__ D4 operator+(D4 a, float c) { a.v += c; return a; }
__device__ __forceinline__ D4 operator-(D4 a, float c) { a.v -= c; return a; }
__device__ __forceinline__ D4 rsub(float c, D4 a) { return D4{c - a.v, {-a.d[0], -a.d[1], -a.d[2], -a.d[3]}}; }
__device__ __forceinline__ D4 scale(D4 a, float c, float dc) {   // value c * a.v given, derivative factor dc
  return D4{c, {a.d[0] * dc, a.d[1] * dc, a.d[2] * dc, a.d[3] * dc}};
}
__device__ __forceinline__ D4 operator*(D4 a, float c) { return scale(a, a.v * c, c); }
__device__ __forceinline__ D4 operator*(D4 a, D4 b) {
  return D4{a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2],
                        a.d[3] * b.v + a.v * b.d[3]}};
}
__device__ __forceinline__ D4 operator/(D4 a, D4 b) {
  const float q = a.v / b.v, ib = 1.f / b.v;
  return D4{q, {(a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib,
                (a.d[3] - q * b.d[3]) * ib}};
}
// autograd conventions: ties of the binary min / max share the gradient, clamp(0) passes it at x >= 0, |x|' = sign(x)
__device__ __forceinline__ D4 mix(D4 a, D4 b) {
  return D4{a.v, {0.5f * (a.d[0] + b.d[0]), 0.5f * (a.d[1] + b.d[1]), 0.5f * (a.d[2] + b.d[2]), 0.5f * (a.d[3] + b.d[3])}};
}
__device__ __forceinline__ D4 dmax(D4 a, D4 b) { return a.v > b.v ? a : (a.v < b.v ? b : mix(a, b)); }
__device__ __forceinline__ D4 dmin(D4 a, D4 b) { return a.v < b.v ? a : (a.v > b.v ? b : mix(a, b)); }
__device__ __forceinline__ D4 clamp0(D4 a) { return a.v >= 0.f ? a : cst(0.f); }
__device__ __forceinline__ D4 dabs(D4 a) {
  const float s = a.v > 0.f ? 1.f : (a.v < 0.f ? -1.f : 0.f);
  return scale(a, fabsf(a.v), s);
}
__device__ __forceinline__ D4 dexp(D4 a) {
  const float e = expf(a.v);
  return scale(a, e, e);
}
__device__ __forceinline__ D4 dsqrt(D4 a) {
  const float r = sqrtf(a.v);
  return scale(a, r, 0.5f / r);
}
__device__ __forceinline__ D4 dpow4(D4 a) {
  const float a2 = a.v * a.v;
  return scale(a, a2 * a2, 4.f * a2 * a.v);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// F.binary_cross_entropy_with_logits(x, t, pos_weight = pw), element-wise: (1 - t) x - (1 + (pw - 1) t) logsigmoid(x)
__device__ __forceinline__ float bce(float x, float t, float pw, float& dx) {
  const float lw = (pw - 1.f) * t + 1.f;
  const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
  dx = (1.f - t) - lw * (1.f - sigmoidf_(x));
  return (1.f - t) * x - lw * ls;
}

// FocalLoss around an element BCE (utils/loss.py:80-106, utils/loss_tal.py:32-60; alpha 0.25):
// bce * (t*0.25 + (1-t)*0.75) * (1 - p_t)^gamma with p_t = t*p + (1-t)*(1-p); l / dx: the BCE value and derivative.
__device__ __forceinline__ float focal_wrap(float x, float t, float gamma, float l, float& dx) {
  const float p = sigmoidf_(x);
  const float base = 1.f - (t * p + (1.f - t) * (1.f - p));
  const float af = t * 0.25f + (1.f - t) * 0.75f;
  const float mod = powf(base, gamma);
  const float dmod = -gamma * powf(base, gamma - 1.f) * (2.f * t - 1.f) * p * (1.f - p);
  dx = af * (dx * mod + l * dmod);
  return l * af * mod;
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kThreads / 32; ++w) r += sh[w];
  __syncthreads();
  return r;   // valid in thread 0
}

}  // namespace ecsy_loss
