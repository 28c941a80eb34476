// Stack-B training loss on the device: `ComputeLoss.__call__` of utils/loss_tal.py:162-215
// with `TaskAlignedAssigner` (utils/tal/assigner.py:51-179), forward AND the gradient w.r.t. the raw DDetect outputs in
// one C-ABI call, no host synchronisation.
//
// The reference pads every image's labels to the longest list of the batch (a Python loop over images with one
// device-to-host synchronisation each, utils/loss_tal.py:142-155) and builds dense [batch, max labels, anchors]
// tensors through ~120 launches.  Here the labels are bucketed by image on the device (k_tal_sort) and every label q
// owns one row of [labels, anchors] tables:
//   k_tal_decode   DFL expectation -> predicted boxes (:154-160)
//   k_tal_metric   one CTA per label: CIoU overlaps, align = score^0.5 * overlap^6, centre-in-box mask, top-10 by ten
//                  rounds of a block-wide arg-max over the metrics staged in shared memory (assigner.py:98-147)
//   k_tal_resolve  per anchor: an anchor claimed by several labels goes to the highest overlap (assigner.py:25-48)
//   k_tal_norm     per label: max aligned metric / overlap over its positives (assigner.py:92-96)
//   k_tal_anchor   per anchor: normalised target score, partial sums of target_scores.sum()
//   k_tal_loss     per anchor: BCE class term, box term, DFL, and their gradients into the NCHW gradient planes
//   k_tal_final    composition (:194-214)
// Choices the reference leaves to torch (topk among equal metrics, argmax among equal overlaps): the lowest index.
// The box term: the reference asks bbox_iou for SIoU=True but utils/metrics2.py:282-311 only reaches its SIoU branch
// when CIoU or DIoU is also set, so what it evaluates -- and what is implemented here -- is GIoU.
#include "loss_common.cuh"
#include "../../include/ecsy.h"

namespace {

using namespace ecsy_loss;

constexpr int kMaxLevels = 5;
constexpr int kReg = 16;         // DDetect.reg_max (models/yolo_snn.py:95)
constexpr float kAssignEps = 1e-9f;
constexpr float kIouEps = 1e-7f;

struct TalArgs {
  const float* f[kMaxLevels];
  float* g[kMaxLevels];
  int ny[kMaxLevels], nx[kMaxLevels], a_base[kMaxLevels + 1];
  float stride[kMaxLevels];
  const float* targets;
  int nl, nc, no, A;
  int64_t N, nt;
  float img_w, img_h, cls_pw, gain_box, gain_cls, gain_dfl, fl_gamma;
  int topk;            // TaskAlignedAssigner hyper-parameters (utils/loss_tal.py:134-137: YOLOM / YOLOA / YOLOB)
  float alpha, beta;
  // workspace
  float* pbox;       // [N][A][4] xyxy, grid units
  int* img_off;      // [N + 1]
  float* gt_box;     // [nt][4] xyxy pixels, bucketed order
  int* gt_label;     // [nt]
  int* gt_valid;     // [nt]
  float* ov;         // [nt][A]
  float* align;      // [nt][A]
  unsigned char* pos;  // [nt][A]
  float* pos_align;  // [nt]
  float* pos_ov;     // [nt]
  int* a_gt;         // [N][A] label row of the anchor or -1
  float* a_norm;     // [N][A]
  double* part;      // [blocks][4]: target-score sum, cls, box, dfl (blocks of the per-anchor kernels)
  double* scal;      // [2]: target_scores_sum (max(., 1)), foreground count
  float* out;        // [6]: loss, box, cls, dfl, foreground anchors, target_scores_sum
};

__device__ __forceinline__ int level_of(const TalArgs& a, int an) {
  int l = 0;
  while (l + 1 < a.nl && an >= a.a_base[l + 1]) ++l;
  return l;
}

// ---- labels bucketed by image, in their original order (utils/loss_tal.py:142-155) --------------------------------
__global__ void __launch_bounds__(kThreads) k_tal_sort(const TalArgs a) {
  extern __shared__ int s_cnt[];   // [N + 1]
  for (int64_t b = threadIdx.x; b < a.N; b += kThreads) {
    int c = 0;
    for (int64_t j = 0; j < a.nt; ++j) c += ((int64_t)a.targets[j * 6] == b) && a.targets[j * 6] >= 0.f;
    s_cnt[b] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int64_t b = 0; b < a.N; ++b) {
      const int c = s_cnt[b];
      s_cnt[b] = run;
      a.img_off[b] = run;
      run += c;
    }
    s_cnt[a.N] = run;
    a.img_off[a.N] = run;
  }
  __syncthreads();
  for (int64_t b = threadIdx.x; b < a.N; b += kThreads) {
    int q = s_cnt[b];
    for (int64_t j = 0; j < a.nt; ++j) {
      const float* t = a.targets + j * 6;
      if ((int64_t)t[0] != b || t[0] < 0.f) continue;
      const float x = t[2] * a.img_w, y = t[3] * a.img_h, w = t[4] * a.img_w, h = t[5] * a.img_h;   // mul_(scale) (:154)
      const float x1 = x - w / 2, y1 = y - h / 2, x2 = x + w / 2, y2 = y + h / 2;                   // xywh2xyxy
      a.gt_box[q * 4 + 0] = x1; a.gt_box[q * 4 + 1] = y1; a.gt_box[q * 4 + 2] = x2; a.gt_box[q * 4 + 3] = y2;
      a.gt_label[q] = min(max((int)t[1], 0), a.nc - 1);
      a.gt_valid[q] = (x1 + y1 + x2 + y2) > 0.f;                                                   // mask_gt (:178)
      ++q;
    }
  }
}

// ---- DFL expectation of one side: softmax over 16 logits with channel stride cs ------------------------------------
__device__ __forceinline__ float dfl_expect(const float* p, int64_t cs, float& mx, float& sum) {
  mx = p[0];
#pragma unroll
  for (int k = 1; k < kReg; ++k) mx = fmaxf(mx, p[k * cs]);
  sum = 0.f;
  float e = 0.f;
#pragma unroll
  for (int k = 0; k < kReg; ++k) {
    const float w = expf(p[k * cs] - mx);
    sum += w;
    e += w * (float)k;
  }
  return e / sum;
}

__global__ void __launch_bounds__(kThreads) k_tal_decode(const TalArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.N * a.A) return;
  const int64_t b = idx / a.A;
  const int an = (int)(idx - b * a.A);
  const int l = level_of(a, an);
  const int pix = an - a.a_base[l];
  const int64_t cs = (int64_t)a.ny[l] * a.nx[l];
  const float* p = a.f[l] + b * a.no * cs + pix;
  const float ax = (float)(pix % a.nx[l]) + 0.5f, ay = (float)(pix / a.nx[l]) + 0.5f;
  float mx, sm;
  const float el = dfl_expect(p, cs, mx, sm), et = dfl_expect(p + kReg * cs, cs, mx, sm);
  const float er = dfl_expect(p + 2 * kReg * cs, cs, mx, sm), eb = dfl_expect(p + 3 * kReg * cs, cs, mx, sm);
  reinterpret_cast<float4*>(a.pbox)[idx] = make_float4(ax - el, ay - et, ax + er, ay + eb);   // dist2bbox, xyxy
}

// CIoU of box 1 (label) and box 2 (prediction), utils/metrics2.py:264-289 with xywh=False
__device__ float ciou(float4 g, float4 p) {
  const float w1 = g.z - g.x, h1 = g.w - g.y + kIouEps, w2 = p.z - p.x, h2 = p.w - p.y + kIouEps;
  const float inter = fmaxf(fminf(g.z, p.z) - fmaxf(g.x, p.x), 0.f) * fmaxf(fminf(g.w, p.w) - fmaxf(g.y, p.y), 0.f);
  const float uni = w1 * h1 + w2 * h2 - inter + kIouEps;
  const float iou = inter / uni;
  const float cw = fmaxf(g.z, p.z) - fminf(g.x, p.x), ch = fmaxf(g.w, p.w) - fminf(g.y, p.y);
  const float c2 = cw * cw + ch * ch + kIouEps;
  const float dx = p.x + p.z - g.x - g.z, dy = p.y + p.w - g.y - g.w;
  const float rho2 = (dx * dx + dy * dy) / 4;
  const float da = atanf(w2 / h2) - atanf(w1 / h1);
  const float v = 0.40528473456935116f * (da * da);       // 4 / pi^2
  const float alpha = v / (v - iou + (1.f + kIouEps));
  return iou - (rho2 / c2 + v * alpha);
}

// x ** e as torch.pow evaluates it for a float tensor and a Python scalar: sqrt for 0.5, products for 1 / 2 / 3, powf else
// (the defaults alpha = 0.5, beta = 6 take sqrtf and powf: the arithmetic the golden fixtures were matched with)
__device__ __forceinline__ float pow_like_torch(float x, float e) {
  if (e == 0.5f) return sqrtf(x);
  if (e == 1.f) return x;
  if (e == 2.f) return x * x;
  if (e == 3.f) return x * x * x;
  return powf(x, e);
}

// ---- one CTA per label: metrics over all anchors of its image + top-k ------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_tal_metric(const TalArgs a) {
  extern __shared__ float s_metric[];                       // [A] metric * in_gt, then -1 for taken entries
  unsigned char* s_in = reinterpret_cast<unsigned char*>(s_metric + a.A);   // [A] centre inside the label box
  __shared__ float s_v[kThreads / 32];
  __shared__ int s_i[kThreads / 32];
  const int q = blockIdx.x;
  if (q >= a.img_off[a.N]) return;                          // rows of labels whose image index is outside the batch
  int b = 0;
  while (q >= a.img_off[b + 1]) ++b;
  const float4 gt = reinterpret_cast<const float4*>(a.gt_box)[q];
  const int label = a.gt_label[q];
  const bool valid = a.gt_valid[q] != 0;
  for (int an = threadIdx.x; an < a.A; an += kThreads) {
    const int l = level_of(a, an);
    const int pix = an - a.a_base[l];
    const int64_t cs = (int64_t)a.ny[l] * a.nx[l];
    const float st = a.stride[l];
    const float x = a.f[l][((int64_t)b * a.no + 4 * kReg + label) * cs + pix];
    const float score = 1.f / (1.f + expf(-x));
    float4 pb = reinterpret_cast<const float4*>(a.pbox)[(int64_t)b * a.A + an];
    pb.x *= st; pb.y *= st; pb.z *= st; pb.w *= st;
    const float ov = fmaxf(ciou(gt, pb), 0.f);
    const float al = pow_like_torch(score, a.alpha) * pow_like_torch(ov, a.beta);   // (utils/tal/assigner.py:121)
    const float ax = ((float)(pix % a.nx[l]) + 0.5f) * st, ay = ((float)(pix / a.nx[l]) + 0.5f) * st;
    const bool in = fminf(fminf(ax - gt.x, ay - gt.y), fminf(gt.z - ax, gt.w - ay)) > kAssignEps;
    a.ov[(int64_t)q * a.A + an] = ov;
    a.align[(int64_t)q * a.A + an] = al;
    a.pos[(int64_t)q * a.A + an] = 0;
    s_metric[an] = in ? al : 0.f;
    s_in[an] = in;
  }
  __syncthreads();
  if (!valid) return;                                       // padded / degenerate labels never become positive
  for (int r = 0; r < a.topk && r < a.A; ++r) {
    float bv = -2.f;
    int bi = 0x7fffffff;
    for (int an = threadIdx.x; an < a.A; an += kThreads) {
      const float v = s_metric[an];
      if (v > bv) { bv = v; bi = an; }                      // ascending scan: the lowest index of equal values stays
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_down_sync(0xffffffffu, bv, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kThreads / 32; ++w)
        if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
      s_metric[bi] = -1.f;
      if (s_in[bi]) a.pos[(int64_t)q * a.A + bi] = 1;       // mask_topk * mask_in_gts * mask_gt (assigner.py:105)
    }
    __syncthreads();
  }
}

// ---- per anchor: several claims -> the label with the highest overlap (assigner.py:25-48) -------------------------
__global__ void __launch_bounds__(kThreads) k_tal_resolve(const TalArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.N * a.A) return;
  const int64_t b = idx / a.A;
  const int an = (int)(idx - b * a.A);
  const int q0 = a.img_off[b], q1 = a.img_off[b + 1];
  int cnt = 0, first = -1, best = q0;
  float best_ov = -1.f;
  for (int q = q0; q < q1; ++q) {
    if (a.pos[(int64_t)q * a.A + an]) {
      if (first < 0) first = q;
      ++cnt;
    }
    const float o = a.ov[(int64_t)q * a.A + an];
    if (o > best_ov) { best_ov = o; best = q; }
  }
  if (cnt > 1) {
    for (int q = q0; q < q1; ++q) a.pos[(int64_t)q * a.A + an] = (q == best);
    first = best;
  }
  a.a_gt[idx] = first;
}

// ---- per label: maxima over its positives (assigner.py:92-95) ---------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_tal_norm(const TalArgs a) {
  __shared__ float s_a[kThreads / 32], s_o[kThreads / 32];
  const int q = blockIdx.x;
  if (q >= a.img_off[a.N]) return;
  float ma = 0.f, mo = 0.f;
  for (int an = threadIdx.x; an < a.A; an += kThreads)
    if (a.pos[(int64_t)q * a.A + an]) {
      ma = fmaxf(ma, a.align[(int64_t)q * a.A + an]);
      mo = fmaxf(mo, a.ov[(int64_t)q * a.A + an]);
    }
  for (int o = 16; o > 0; o >>= 1) {
    ma = fmaxf(ma, __shfl_down_sync(0xffffffffu, ma, o));
    mo = fmaxf(mo, __shfl_down_sync(0xffffffffu, mo, o));
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = ma; s_o[threadIdx.x >> 5] = mo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) { ma = fmaxf(ma, s_a[w]); mo = fmaxf(mo, s_o[w]); }
    a.pos_align[q] = ma;
    a.pos_ov[q] = mo;
  }
}

// ---- per anchor: normalised target score (assigner.py:96-97), partial sums of target_scores.sum() ----------------
__global__ void __launch_bounds__(kThreads) k_tal_anchor(const TalArgs a) {
  __shared__ double sh[kThreads / 32];
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  double s = 0.0, n = 0.0;
  if (idx < a.N * a.A) {
    const int q = a.a_gt[idx];
    float norm = 0.f;
    if (q >= 0) {
      const int an = (int)(idx % a.A);
      norm = a.align[(int64_t)q * a.A + an] * a.pos_ov[q] / (a.pos_align[q] + kAssignEps);
      n = 1.0;
    }
    a.a_norm[idx] = norm;
    s = (double)norm;
  }
  s = block_sum(s, sh);
  n = block_sum(n, sh);
  if (threadIdx.x == 0) {
    a.part[4 * blockIdx.x] = s;
    a.part[4 * blockIdx.x + 1] = n;
  }
}

__global__ void __launch_bounds__(kThreads) k_tal_tss(const TalArgs a, int blocks) {
  __shared__ double sh[kThreads / 32];
  double s = 0.0, n = 0.0;
  for (int i = threadIdx.x; i < blocks; i += kThreads) {
    s += a.part[4 * i];
    n += a.part[4 * i + 1];
  }
  s = block_sum(s, sh);
  n = block_sum(n, sh);
  if (threadIdx.x == 0) {
    a.scal[0] = fmax((double)(float)s, 1.0);                 // max(target_scores.sum(), 1) (:192)
    a.scal[1] = n;
    a.out[5] = (float)s;
    a.out[4] = (float)n;
  }
}

// GIoU of the predicted box (duals over x1, y1, x2, y2) and the target box: utils/metrics2.py:264-281, 310-311
__device__ D4 giou(D4 x1, D4 y1, D4 x2, D4 y2, float4 t) {
  const D4 w1 = x2 - x1, h1 = y2 - y1 + kIouEps;
  const float w2 = t.z - t.x, h2 = t.w - t.y + kIouEps;
  const D4 inter = clamp0(dmin(x2, cst(t.z)) - dmax(x1, cst(t.x))) * clamp0(dmin(y2, cst(t.w)) - dmax(y1, cst(t.y)));
  const D4 uni = w1 * h1 + w2 * h2 - inter + kIouEps;
  const D4 iou = inter / uni;
  const D4 cw = dmax(x2, cst(t.z)) - dmin(x1, cst(t.x)), ch = dmax(y2, cst(t.w)) - dmin(y1, cst(t.y));
  const D4 area = cw * ch + kIouEps;
  return iou - (area - uni) / area;
}

// ---- per anchor: class BCE (:196), box term (:74-81), DFL (:84-103) and all gradients -----------------------------
__global__ void __launch_bounds__(kThreads) k_tal_loss(const TalArgs a) {
  __shared__ double sh[kThreads / 32];
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  double s_cls = 0.0, s_box = 0.0, s_dfl = 0.0;
  if (idx < a.N * a.A) {
    const int64_t b = idx / a.A;
    const int an = (int)(idx - b * a.A);
    const int l = level_of(a, an);
    const int pix = an - a.a_base[l];
    const int64_t cs = (int64_t)a.ny[l] * a.nx[l];
    const float* p = a.f[l] + b * a.no * cs + pix;
    float* g = a.g[l] ? a.g[l] + b * a.no * cs + pix : nullptr;
    const int q = a.a_gt[idx];
    const float norm = a.a_norm[idx];
    const float inv = (float)((double)a.N / a.scal[0]);      // d(loss) = gain * N / target_scores_sum * d(sum)
    const int label = q >= 0 ? a.gt_label[q] : -1;
    float acc = 0.f;
    for (int c = 0; c < a.nc; ++c) {
      float dx;
      const float xv = p[(4 * kReg + c) * cs], tv = c == label ? norm : 0.f;
      float l = bce(xv, tv, a.cls_pw, dx);
      if (a.fl_gamma > 0.f) l = focal_wrap(xv, tv, a.fl_gamma, l, dx);       // FocalLoss(BCEcls, g), utils/loss_tal.py:116-119
      acc += l;
      if (g) g[(4 * kReg + c) * cs] = a.gain_cls * inv * dx;
    }
    s_cls = (double)acc;
    if (q < 0 || norm == 0.f) {
      if (g)
        for (int c = 0; c < 4 * kReg; ++c) g[c * cs] = 0.f;
      // a positive with zero weight still adds exact zeros to the box / DFL sums in the reference
    } else {
      const float st = a.stride[l];
      const float ax = (float)(pix % a.nx[l]) + 0.5f, ay = (float)(pix / a.nx[l]) + 0.5f;
      float mx[4], sm[4], ex[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) ex[s] = dfl_expect(p + s * kReg * cs, cs, mx[s], sm[s]);
      float4 t = reinterpret_cast<const float4*>(a.gt_box)[q];
      t.x /= st; t.y /= st; t.z /= st; t.w /= st;             // target_bboxes /= stride_tensor (:191)
      const D4 r = giou(var(ax - ex[0], 0), var(ay - ex[1], 1), var(ax + ex[2], 2), var(ay + ex[3], 3), t);
      s_box = (double)((1.f - r.v) * norm);
      // d box / d expectation: x1 = ax - E_l, y1 = ay - E_t, x2 = ax + E_r, y2 = ay + E_b
      const float dE[4] = {r.d[0], r.d[1], -r.d[2], -r.d[3]};     // d(1 - giou) / dE_s
      const float tgt[4] = {ax - t.x, ay - t.y, t.z - ax, t.w - ay};   // bbox2dist (anchor_generator.py:37-40)
      float dfl = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float d = fminf(fmaxf(tgt[s], 0.f), (float)(kReg - 1) - 0.01f);
        const int tl = (int)d;
        const float wl = (float)(tl + 1) - d, wr = 1.f - wl;
        const float lse = mx[s] + logf(sm[s]);
        const float* ps = p + s * kReg * cs;
        dfl += (lse - ps[tl * cs]) * wl + (lse - ps[(tl + 1) * cs]) * wr;   // F.cross_entropy left / right (:97-102)
        if (g) {
          const float cb = a.gain_box * inv * norm * dE[s], cd = a.gain_dfl * inv * norm * 0.25f;
          for (int k = 0; k < kReg; ++k) {
            const float pk = expf(ps[k * cs] - mx[s]) / sm[s];
            const float onehot = (k == tl ? wl : 0.f) + (k == tl + 1 ? wr : 0.f);
            g[(s * kReg + k) * cs] = cb * pk * ((float)k - ex[s]) + cd * (pk - onehot);
          }
        }
      }
      s_dfl = (double)(dfl * 0.25f * norm);                   // .mean(-1) over the four sides (:103)
    }
  }
  s_cls = block_sum(s_cls, sh);
  s_box = block_sum(s_box, sh);
  s_dfl = block_sum(s_dfl, sh);
  if (threadIdx.x == 0) {
    a.part[4 * blockIdx.x + 1] = s_cls;
    a.part[4 * blockIdx.x + 2] = s_box;
    a.part[4 * blockIdx.x + 3] = s_dfl;
  }
}

__global__ void __launch_bounds__(kThreads) k_tal_final(const TalArgs a, int blocks) {
  __shared__ double sh[kThreads / 32];
  double c = 0.0, x = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < blocks; i += kThreads) {
    c += a.part[4 * i + 1];
    x += a.part[4 * i + 2];
    d += a.part[4 * i + 3];
  }
  c = block_sum(c, sh);
  x = block_sum(x, sh);
  d = block_sum(d, sh);
  if (threadIdx.x == 0) {
    const double tss = a.scal[0];
    const float lbox = (float)(x / tss) * a.gain_box, lcls = (float)(c / tss) * a.gain_cls,
                ldfl = (float)(d / tss) * a.gain_dfl;
    a.out[0] = (lbox + lcls + ldfl) * (float)a.N;            // loss.sum() * batch_size (:214)
    a.out[1] = lbox;
    a.out[2] = lcls;
    a.out[3] = ldfl;
  }
}

struct TalWs {
  size_t pbox, img_off, gt_box, gt_label, gt_valid, ov, align, pos, pos_align, pos_ov, a_gt, a_norm, part, scal, total;
};

inline size_t up(size_t x) { return (x + 255) & ~size_t(255); }

TalWs tal_ws(int64_t N, int64_t A, int64_t nt) {
  TalWs w{};
  size_t o = 0;
  const size_t na = (size_t)(N * A), blocks = (na + kThreads - 1) / kThreads;
  w.pbox = o;      o = up(o + na * 16);
  w.img_off = o;   o = up(o + (size_t)(N + 1) * 4);
  w.gt_box = o;    o = up(o + (size_t)nt * 16);
  w.gt_label = o;  o = up(o + (size_t)nt * 4);
  w.gt_valid = o;  o = up(o + (size_t)nt * 4);
  w.ov = o;        o = up(o + (size_t)nt * A * 4);
  w.align = o;     o = up(o + (size_t)nt * A * 4);
  w.pos = o;       o = up(o + (size_t)nt * A);
  w.pos_align = o; o = up(o + (size_t)nt * 4);
  w.pos_ov = o;    o = up(o + (size_t)nt * 4);
  w.a_gt = o;      o = up(o + na * 4);
  w.a_norm = o;    o = up(o + na * 4);
  w.part = o;      o = up(o + blocks * 32);
  w.scal = o;      o = up(o + 16);
  w.total = o;
  return w;
}

}  // namespace

extern "C" size_t ecsy_tal_loss_ws_bytes(int nl, int64_t N, int64_t nt, const int* ny, const int* nx) {
  if (nl < 1 || nl > kMaxLevels || !ny || !nx || N < 1 || nt < 0) return 0;
  int64_t A = 0;
  for (int l = 0; l < nl; ++l) A += (int64_t)ny[l] * nx[l];
  return tal_ws(N, A, nt).total;
}

extern "C" int ecsy_tal_loss(const float* const* feats, float* const* gfeats, const float* targets, int64_t nt, int nl,
                             int64_t N, int nc, const int* ny, const int* nx, const float* strides, float cls_pw,
                             float gain_box, float gain_cls, float gain_dfl, float fl_gamma, int topk, float alpha,
                             float beta, float* out, void* ws, size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(nl >= 1 && nl <= kMaxLevels, "tal_loss: 1..%d detection levels, got %d", kMaxLevels, nl);
  ECSY_CHECK_ARG(feats && ny && nx && strides && out, "tal_loss: null argument");
  ECSY_CHECK_ARG(N >= 1 && nc >= 1 && nt >= 0 && (nt == 0 || targets), "tal_loss: bad sizes N=%lld nc=%d nt=%lld",
                 (long long)N, nc, (long long)nt);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TalArgs a{};
  a.nl = nl; a.nc = nc; a.no = nc + 4 * kReg; a.N = N; a.nt = nt; a.targets = targets; a.out = out;
  a.cls_pw = cls_pw; a.gain_box = gain_box; a.gain_cls = gain_cls; a.gain_dfl = gain_dfl; a.fl_gamma = fl_gamma;
  ECSY_CHECK_ARG(topk >= 1 && alpha >= 0.f && beta >= 0.f, "tal_loss: assigner topk=%d alpha=%g beta=%g", topk, (double)alpha,
                 (double)beta);
  a.topk = topk; a.alpha = alpha; a.beta = beta;
  a.a_base[0] = 0;
  for (int l = 0; l < nl; ++l) {
    ECSY_CHECK_ARG(feats[l] && ny[l] >= 1 && nx[l] >= 1 && strides[l] > 0.f, "tal_loss: level %d: bad grid / stride", l);
    a.f[l] = feats[l]; a.g[l] = gfeats ? gfeats[l] : nullptr;
    a.ny[l] = ny[l]; a.nx[l] = nx[l]; a.stride[l] = strides[l];
    a.a_base[l + 1] = a.a_base[l] + ny[l] * nx[l];
  }
  a.A = a.a_base[nl];
  a.img_w = (float)nx[0] * strides[0];                      // feats[0].shape[2:] * stride[0] (:172)
  a.img_h = (float)ny[0] * strides[0];
  const size_t metric_smem = (size_t)a.A * 5;
  ECSY_CHECK_ARG(a.A >= topk && metric_smem <= 200 * 1024, "tal_loss: %d anchors per image unsupported (topk = %d)", a.A, topk);
  ECSY_CHECK_ARG(N * a.A < (1LL << 31) && nt * (int64_t)a.A < (1LL << 40) && N < 12000, "tal_loss: sizes too large");
  const TalWs w = tal_ws(N, a.A, nt);
  if (!ws || ws_bytes < w.total) {
    ecsy_set_error("tal_loss: workspace %zu < %zu bytes", ws_bytes, w.total);
    return ECSY_ERR_WS;
  }
  char* base = static_cast<char*>(ws);
  a.pbox = reinterpret_cast<float*>(base + w.pbox);
  a.img_off = reinterpret_cast<int*>(base + w.img_off);
  a.gt_box = reinterpret_cast<float*>(base + w.gt_box);
  a.gt_label = reinterpret_cast<int*>(base + w.gt_label);
  a.gt_valid = reinterpret_cast<int*>(base + w.gt_valid);
  a.ov = reinterpret_cast<float*>(base + w.ov);
  a.align = reinterpret_cast<float*>(base + w.align);
  a.pos = reinterpret_cast<unsigned char*>(base + w.pos);
  a.pos_align = reinterpret_cast<float*>(base + w.pos_align);
  a.pos_ov = reinterpret_cast<float*>(base + w.pos_ov);
  a.a_gt = reinterpret_cast<int*>(base + w.a_gt);
  a.a_norm = reinterpret_cast<float*>(base + w.a_norm);
  a.part = reinterpret_cast<double*>(base + w.part);
  a.scal = reinterpret_cast<double*>(base + w.scal);
  const int blocks = (int)((N * a.A + kThreads - 1) / kThreads);
  k_tal_sort<<<1, kThreads, (size_t)(N + 1) * 4, st>>>(a);
  ECSY_LAUNCH_CHECK();
  k_tal_decode<<<blocks, kThreads, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  if (nt > 0) {
    static bool attr_set = false;
    if (!attr_set && metric_smem > 48 * 1024) {
      ECSY_CUDA(cudaFuncSetAttribute(k_tal_metric, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    // labels whose image index is outside the batch are dropped by k_tal_sort: the grid covers nt rows and each CTA
    // checks its row against img_off[N]
    k_tal_metric<<<(unsigned)nt, kThreads, metric_smem, st>>>(a);
    ECSY_LAUNCH_CHECK();
  }
  k_tal_resolve<<<blocks, kThreads, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  if (nt > 0) {
    k_tal_norm<<<(unsigned)nt, kThreads, 0, st>>>(a);
    ECSY_LAUNCH_CHECK();
  }
  k_tal_anchor<<<blocks, kThreads, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  k_tal_tss<<<1, kThreads, 0, st>>>(a, blocks);
  ECSY_LAUNCH_CHECK();
  k_tal_loss<<<blocks, kThreads, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  k_tal_final<<<1, kThreads, 0, st>>>(a, blocks);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
