// Multi-tensor optimizer step: SGD with Nesterov momentum over all parameters plus the
// ModelEMA update over all floating-point state_dict entries in ONE launch, instead of ~4 small kernels per tensor from
// torch.optim.SGD (train.py:282-287, 576-580) and a Python loop over ~250-500 tensors in ModelEMA.update
// (utils/torch_utils.py:306-316).
//
// Work is described by device-resident tables the host builds once per model (the caller owns them): per tensor the
// addresses of value / gradient / momentum buffer / EMA copy (0 = absent), its length and parameter group; per chunk
// of kOptChunk elements the tensor it belongs to and its offset.  One CTA per chunk, float4 accesses.
//   SGD (torch/optim/sgd.py, dampening 0):  d = g + wd*p;  buf = first ? d : momentum*buf + d;
//                                           d = nesterov ? d + momentum*buf : buf;  p = p - lr*d
//       (fused multiply-adds, as torch's own CUDA and vectorised CPU kernels evaluate a + alpha*b)
//   EMA (utils/torch_utils.py:311-316):     e = (e*d) + ((1-d)*v), every product and the sum rounded separately
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"

namespace {

constexpr int kOptChunk = 4096;
constexpr int kOptThreads = 256;
constexpr int kMaxGroups = 8;

struct OptArgs {
  const uint64_t* val;      // [n] parameter / buffer address (fp32)
  const uint64_t* grad;     // [n] gradient address or 0 (no SGD: a buffer, or a parameter without grad)
  const uint64_t* mom;      // [n] momentum buffer address or 0
  const uint64_t* ema;      // [n] EMA copy address or 0
  const int64_t* numel;     // [n]
  const int32_t* group;     // [n] parameter group
  const int32_t* chunk_tensor;   // [chunks]
  const int64_t* chunk_off;      // [chunks]
  float lr[kMaxGroups], wd[kMaxGroups];
  float momentum, ema_d, ema_1md;
  int nesterov, first, do_ema;
};

__device__ __forceinline__ void sgd1(float& p, float g, float& b, float lr, float wd, float mom, bool nesterov, bool first,
                                     bool has_mom) {
  float d = wd != 0.f ? fmaf(wd, p, g) : g;
  if (has_mom) {
    b = first ? d : fmaf(mom, b, d);
    d = nesterov ? fmaf(mom, b, d) : b;
  }
  p = fmaf(-lr, d, p);
}

__global__ void __launch_bounds__(kOptThreads) k_sgd_ema(const OptArgs a) {
  const int t = a.chunk_tensor[blockIdx.x];
  const int64_t off = a.chunk_off[blockIdx.x];
  const int64_t n = a.numel[t];
  const int len = (int)min((int64_t)kOptChunk, n - off);
  float* p = reinterpret_cast<float*>(a.val[t]) + off;
  const float* g = a.grad[t] ? reinterpret_cast<const float*>(a.grad[t]) + off : nullptr;
  float* b = a.mom[t] ? reinterpret_cast<float*>(a.mom[t]) + off : nullptr;
  float* e = (a.do_ema && a.ema[t]) ? reinterpret_cast<float*>(a.ema[t]) + off : nullptr;
  const int gi = a.group[t];
  const float lr = a.lr[gi], wd = a.wd[gi], mom = a.momentum;
  const bool nest = a.nesterov != 0, first = a.first != 0, has_mom = b != nullptr && mom != 0.f;
  // chunk offsets are multiples of 4096 and torch allocations are 512-byte aligned: float4 is safe when every
  // base pointer is 16-byte aligned (checked here because a parameter may be a view into a flat buffer)
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b) |
                     reinterpret_cast<uintptr_t>(e)) & 15u) == 0;
  const int n4 = vec ? (len >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += kOptThreads) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    if (g != nullptr) {
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_mom && !first) bv = reinterpret_cast<float4*>(b)[i];
      sgd1(pv.x, gv.x, bv.x, lr, wd, mom, nest, first, has_mom);
      sgd1(pv.y, gv.y, bv.y, lr, wd, mom, nest, first, has_mom);
      sgd1(pv.z, gv.z, bv.z, lr, wd, mom, nest, first, has_mom);
      sgd1(pv.w, gv.w, bv.w, lr, wd, mom, nest, first, has_mom);
      if (has_mom) reinterpret_cast<float4*>(b)[i] = bv;
      reinterpret_cast<float4*>(p)[i] = pv;
    }
    if (e != nullptr) {
      float4 ev = reinterpret_cast<float4*>(e)[i];
      ev.x = __fadd_rn(__fmul_rn(ev.x, a.ema_d), __fmul_rn(a.ema_1md, pv.x));
      ev.y = __fadd_rn(__fmul_rn(ev.y, a.ema_d), __fmul_rn(a.ema_1md, pv.y));
      ev.z = __fadd_rn(__fmul_rn(ev.z, a.ema_d), __fmul_rn(a.ema_1md, pv.z));
      ev.w = __fadd_rn(__fmul_rn(ev.w, a.ema_d), __fmul_rn(a.ema_1md, pv.w));
      reinterpret_cast<float4*>(e)[i] = ev;
    }
  }
  for (int i = n4 * 4 + threadIdx.x; i < len; i += kOptThreads) {
    float pv = p[i];
    if (g != nullptr) {
      float bv = (has_mom && !first) ? b[i] : 0.f;
      sgd1(pv, g[i], bv, lr, wd, mom, nest, first, has_mom);
      if (has_mom) b[i] = bv;
      p[i] = pv;
    }
    if (e != nullptr) e[i] = __fadd_rn(__fmul_rn(e[i], a.ema_d), __fmul_rn(a.ema_1md, pv));
  }
}

}  // namespace

extern "C" int ecsy_optim_chunk(void) { return kOptChunk; }

extern "C" int ecsy_sgd_ema_step(const uint64_t* val, const uint64_t* grad, const uint64_t* mom, const uint64_t* ema,
                                 const int64_t* numel, const int32_t* group, int n_tensors, const int32_t* chunk_tensor,
                                 const int64_t* chunk_off, int64_t n_chunks, const float* lr, const float* weight_decay,
                                 int n_groups, float momentum, int nesterov, int first_step, int do_ema, float ema_d,
                                 float ema_one_minus_d, void* stream) {
  ECSY_CHECK_ARG(val && grad && mom && ema && numel && group && chunk_tensor && chunk_off, "sgd_ema_step: null table");
  ECSY_CHECK_ARG(n_tensors > 0 && n_chunks > 0 && n_chunks < (1LL << 31), "sgd_ema_step: empty / oversized work list");
  ECSY_CHECK_ARG(n_groups >= 1 && n_groups <= kMaxGroups && lr && weight_decay, "sgd_ema_step: 1..%d parameter groups",
                 kMaxGroups);
  ECSY_CHECK_ARG(momentum >= 0.f && (!nesterov || momentum > 0.f), "sgd_ema_step: Nesterov needs momentum > 0");
  OptArgs a{};
  a.val = val; a.grad = grad; a.mom = mom; a.ema = ema; a.numel = numel; a.group = group;
  a.chunk_tensor = chunk_tensor; a.chunk_off = chunk_off;
  for (int i = 0; i < n_groups; ++i) { a.lr[i] = lr[i]; a.wd[i] = weight_decay[i]; }   // host arrays, passed by value
  a.momentum = momentum; a.ema_d = ema_d; a.ema_1md = ema_one_minus_d;
  a.nesterov = nesterov; a.first = first_step; a.do_ema = do_ema;
  k_sgd_ema<<<(unsigned)n_chunks, kOptThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
