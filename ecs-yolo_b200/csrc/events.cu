// Gen1 event-camera input on the GPU: event bins -> ternary frames -> network-size frames.
// Reference (CPU, per sample, in the data loader): g1-resnet/utils/give_g1_data.py:550-565 `create_data` paints each of
// the T event bins into a grey (127) 304 x 240 frame, `img[i, y, x, :] = 255 * p`, the LAST event of a pixel winning;
// g1-resnet/utils/datasets_g1T.py:518-533 resizes every frame with cv2.resize (INTER_LINEAR, uint8) to the network
// size; g1-resnet/train_g1.py:298 divides by 255.
//   k_event_stamp  : one thread per event: atomicMax of (event index + 1) << 1 | p into a per-pixel stamp -- the
//                    largest stamp is the last event in sensor order, so the result is order-exact and deterministic.
//   k_event_frames : one thread per OUTPUT pixel: decodes the 2 x 2 source stamps (0 -> 127, else 255 * p), applies
//                    OpenCV's 11-bit fixed-point bilinear kernel (imgproc/src/resize.cpp: coefficients from
//                    float((d + 0.5) * scale - 0.5), x weights reset at the borders, y rows clamped;
//                    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2) and writes value / 255 to the three
//                    identical channels of the model's NHWC input [T][N][S][S][3].
// Integer work, bit-exact against create_data + cv2.resize (tests/golden/post_events.pt).
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"

namespace {

__global__ void k_event_stamp(const int32_t* __restrict__ ex, const int32_t* __restrict__ ey,
                              const int32_t* __restrict__ ep, const int32_t* __restrict__ ef, int64_t n_events,
                              uint32_t* __restrict__ stamp, int frames, int H, int W, int* __restrict__ oob) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_events; e += stride) {
    const int x = ex[e], y = ey[e], f = ef[e];
    if (x < 0 || x >= W || y < 0 || y >= H || f < 0 || f >= frames) {   // the reference asserts (give_g1_data.py:562-563)
      atomicAdd(oob, 1);
      continue;
    }
    const uint32_t v = ((uint32_t)(e + 1) << 1) | (ep[e] != 0 ? 1u : 0u);
    atomicMax(stamp + ((size_t)f * H + y) * W + x, v);
  }
}

__device__ __forceinline__ int decode(uint32_t s) { return s == 0u ? 127 : ((s & 1u) ? 255 : 0); }

struct Coef {
  int s;        // source index (before clamping for y)
  int a0, a1;   // 11-bit weights
};

__device__ __forceinline__ Coef coef(int d, double scale, int ssize, bool reset) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (reset) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  }
  Coef c;
  c.s = s;
  c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));   // cvRound: round half to even
  c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}

__global__ void k_event_frames(const uint32_t* __restrict__ stamp, float* __restrict__ out, int N, int T, int H, int W,
                               int Ho, int Wo, double scale_x, double scale_y) {
  const int64_t total = (int64_t)N * T * Ho * Wo;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int dx = (int)(i % Wo);
    int64_t r = i / Wo;
    const int dy = (int)(r % Ho);
    r /= Ho;
    const int n = (int)(r % N);      // output order [t][n][dy][dx]: consecutive threads write consecutive pixels
    const int t = (int)(r / N);
    const Coef cx = coef(dx, scale_x, W, true);
    const Coef cy = coef(dy, scale_y, H, false);
    const int x0 = cx.s, x1 = min(cx.s + 1, W - 1);
    const int y0 = min(max(cy.s, 0), H - 1), y1 = min(max(cy.s + 1, 0), H - 1);
    const uint32_t* fr = stamp + (size_t)(n * T + t) * H * W;     // frames are stored sample-major [n][t]
    const int h0 = decode(fr[(size_t)y0 * W + x0]) * cx.a0 + decode(fr[(size_t)y0 * W + x1]) * cx.a1;
    const int h1 = decode(fr[(size_t)y1 * W + x0]) * cx.a0 + decode(fr[(size_t)y1 * W + x1]) * cx.a1;
    const int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
    const float fv = __fdiv_rn((float)min(max(v, 0), 255), 255.f);
    float* o = out + i * 3;
    o[0] = fv; o[1] = fv; o[2] = fv;
  }
}

}  // namespace

extern "C" size_t ecsy_event_frames_ws_bytes(int64_t N, int T, int H, int W) {
  if (N <= 0 || T <= 0 || H <= 0 || W <= 0) return 0;
  return 512 + (size_t)N * T * H * W * sizeof(uint32_t);
}

extern "C" int ecsy_event_frames(const int32_t* ex, const int32_t* ey, const int32_t* ep, const int32_t* eframe,
                                 int64_t n_events, int64_t N, int T, int H, int W, int Ho, int Wo, float* out,
                                 int* oob_count, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(out && oob_count && N > 0 && T > 0 && H > 1 && W > 1 && Ho > 0 && Wo > 0, "event_frames: bad arguments");
  ECSY_CHECK_ARG(n_events >= 0 && n_events < (1LL << 31) - 1, "event_frames: at most 2^31 - 2 events per call");
  ECSY_CHECK_ARG(n_events == 0 || (ex && ey && ep && eframe), "event_frames: null event arrays");
  ECSY_CHECK_ARG(N * T < (1LL << 24), "event_frames: too many frames");
  const size_t need = ecsy_event_frames_ws_bytes(N, T, H, W);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("event_frames: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uint32_t* stamp = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  ECSY_CUDA(cudaMemsetAsync(stamp, 0, (size_t)N * T * H * W * sizeof(uint32_t), st));
  ECSY_CUDA(cudaMemsetAsync(oob_count, 0, sizeof(int), st));
  const int sms = ecsy_num_sms();
  if (n_events > 0) {
    const int64_t blocks = (n_events + 255) / 256;
    k_event_stamp<<<(unsigned)(blocks < sms * 16 ? blocks : sms * 16), 256, 0, st>>>(ex, ey, ep, eframe, n_events, stamp,
                                                                                   (int)(N * T), H, W, oob_count);
    ECSY_LAUNCH_CHECK();
  }
  const int64_t total = N * T * (int64_t)Ho * Wo;
  const int64_t blocks = (total + 255) / 256;
  // resize.cpp: inv_scale = dsize / ssize; scale = 1. / inv_scale (both in double)
  const double scale_x = 1.0 / ((double)Wo / (double)W), scale_y = 1.0 / ((double)Ho / (double)H);
  k_event_frames<<<(unsigned)(blocks < sms * 16 ? blocks : sms * 16), 256, 0, st>>>(stamp, out, (int)N, T, H, W, Ho, Wo,
                                                                                 scale_x, scale_y);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
