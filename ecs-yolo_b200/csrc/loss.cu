// Stack-A training loss on the device: `ComputeLoss.__call__` + `build_targets`
// (utils/loss.py:162-290) with the SIoU box term (utils/metrics.py:227-307, SIoU=True), forward AND the gradient
// w.r.t. the raw Detect outputs in one C-ABI call, no host synchronisation and no data-dependent shapes.
//
// The reference filters the na * nt (anchor, target) pairs with boolean masks, replicates them over up to five
// neighbouring cells and gathers / scatters with the resulting variable-length index lists -- each mask is a
// device-to-host synchronisation and the whole thing is ~150 small launches per step.  Here every one of the
// nl * 5 * na * nt CANDIDATES (level l, offset k, anchor a, target j) is a thread; its position
// k * na * nt + a * nt + j is the row it would have in the reference's lists (`t.repeat((5, 1, 1))[j]`, :258-266), so
//   * the `tobj[b, a, gj, gi] = iou` scatter with duplicate indices (:205) resolves to the LAST row, as the reference's
//     serial CPU index_put_ does, through an atomicMax of (row + 1) into a per-cell stamp;
//   * means over the n matched rows (:187-199, :214) are sums over valid candidates divided by a count that is reduced
//     on the device between two candidate passes.
// Kernels: k_loss_match (validity, cell, SIoU value + d SIoU / d raw box by forward-mode duals, stamp) ->
// k_loss_level_sums (n, sum(1 - iou) per level) -> k_loss_cand_grad (box + class gradients, class BCE sums) ->
// k_loss_obj per level (objectness BCE over every cell, its gradient) -> k_loss_final.  Reductions run in double in a
// fixed order (bit-reproducible loss); only gradients of cells hit by several candidates are accumulated with float
// atomics.  All streaming is a few MB: the point of the kernel is removing the host round trips, not bandwidth.
#include "loss_common.cuh"
#include "../../include/ecsy.h"

namespace {

using namespace ecsy_loss;

constexpr int kMaxLevels = 5;      // P3-P7, the length of the reference's balance list (utils/loss.py:156)
constexpr float kEps = 1e-7f;      // utils/metrics.py:228

struct LossArgs {
  const float* p[kMaxLevels];
  float* gp[kMaxLevels];
  int ny[kMaxLevels], nx[kMaxLevels];
  int64_t cell_base[kMaxLevels + 1];   // prefix sums of N * na * ny * nx
  int obj_block_base[kMaxLevels + 1];  // prefix sums of the objectness kernel's grid sizes
  float balance[kMaxLevels];
  const float* targets;   // [nt][6] = (image, class, cx, cy, w, h), normalised
  const float* anchors;   // [nl][na][2], grid units
  int nl, na, nc, no;
  int64_t N, nt, per_level;            // per_level = 5 * na * nt candidates
  float box, obj, cls, cls_pw, obj_pw, cp, cn, anchor_t, gr;
  float fl_gamma;      // > 0: FocalLoss around the BCE terms (utils/loss.py:80-106, alpha 0.25)
  float* slide_state;  // non-null: SlideLoss (utils/loss.py:38-76); caller-owned [4] = ema_cls, ema_obj, has_cls, has_obj
  float* thr;          // [nl][2] the SlideLoss ema in force for the class / objectness term of level l
  int* stamp;        // [cells]      row + 1 of the last candidate writing the cell, 0 = background
  float* c_iou;      // [nl * per_level]   SIoU of the candidate (valid ones)
  float* c_diou;     // [nl * per_level][4]  d SIoU / d raw (x, y, w, h) logits
  int* c_cell;       // [nl * per_level]   cell inside the level, -1 = not a match
  float* c_cls;      // [nl * per_level]   sum over classes of the candidate's class BCE
  double* obj_part;  // [obj blocks]  per-block sums of the objectness BCE
  double* lvl;       // [nl][2]  n, sum(1 - iou)
  float* out;        // [4 + nl]  loss, lbox, lobj, lcls, objectness BCE mean per level
};

// bbox_iou(pbox.T, tbox, x1y1x2y2=False, SIoU=True): utils/metrics.py:236-257, 286-307 (alpha = 1, not Focal)
__device__ D4 siou(D4 px, D4 py, D4 pw, D4 ph, float tx, float ty, float tw, float th) {
  const D4 px1 = px - pw * 0.5f, px2 = px + pw * 0.5f, py1 = py - ph * 0.5f, py2 = py + ph * 0.5f;
  const float tx1 = tx - tw / 2, tx2 = tx + tw / 2, ty1 = ty - th / 2, ty2 = ty + th / 2;
  const D4 inter = clamp0(dmin(px2, cst(tx2)) - dmax(px1, cst(tx1))) * clamp0(dmin(py2, cst(ty2)) - dmax(py1, cst(ty1)));
  const D4 w1 = px2 - px1, h1 = py2 - py1 + kEps;
  const float w2 = tx2 - tx1, h2 = ty2 - ty1 + kEps;
  const D4 uni = w1 * h1 + w2 * h2 - inter + kEps;
  const D4 iou = inter / uni;
  const D4 cw = dmax(px2, cst(tx2)) - dmin(px1, cst(tx1));
  const D4 ch = dmax(py2, cst(ty2)) - dmin(py1, cst(ty1));
  const D4 sx = (rsub(tx1 + tx2, px1) - px2) * 0.5f + kEps;
  const D4 sy = (rsub(ty1 + ty2, py1) - py2) * 0.5f + kEps;
  const D4 sigma = dsqrt(sx * sx + sy * sy);
  const D4 s1 = dabs(sx) / sigma, s2 = dabs(sy) / sigma;
  const D4 s = s1.v > 0.70710678f ? s2 : s1;
  // cos(2 asin(s) - pi/2): value through the same steps as the reference, derivative 2 sin(.) ... / sqrt(1 - s^2)
  const float ang = asinf(s.v) * 2.f - 1.5707963267948966f;
  const D4 angle = scale(s, cosf(ang), -sinf(ang) * 2.f / sqrtf(1.f - s.v * s.v));
  const D4 gamma = angle - 2.f;
  const D4 rx = sx / cw, ry = sy / ch;
  const D4 dist = rsub(2.f, dexp(gamma * (rx * rx))) - dexp(gamma * (ry * ry));
  const D4 ow = dabs(w1 - w2) / dmax(w1, cst(w2)), oh = dabs(h1 - h2) / dmax(h1, cst(h2));
  const D4 shape = dpow4(rsub(1.f, dexp(ow * -1.f))) + dpow4(rsub(1.f, dexp(oh * -1.f)));
  return iou - ((dist + shape) * 0.5f + kEps);
}

// ---- the wrapped criteria (utils/loss.py:145-152) ----------------------------------------------------------------
// SlideLoss.forward's modulating weight (:61-69): 1 up to ema - 0.1, exp(1 - ema) below ema, exp(1 - true) from ema on.
__device__ __forceinline__ float slide_weight(float t, float ema) {
  if (t >= ema) return expf(-(t - 1.f));
  if (t > ema - 0.1f) return (float)exp(1.0 - (double)ema);
  return 1.f;
}

// element loss and its derivative: plain BCE, BCE * slide weight (a constant w.r.t. the logit), or the focal form
// bce * (t*0.25 + (1-t)*0.75) * (1 - p_t)^gamma with p_t = t*p + (1-t)*(1-p)
__device__ __forceinline__ float criterion(const LossArgs& a, float x, float t, float pw, float ema, float& dx) {
  float l = bce(x, t, pw, dx);
  if (a.slide_state != nullptr) {
    const float w = slide_weight(t, ema);
    dx *= w;
    return l * w;
  }
  if (a.fl_gamma > 0.f) return focal_wrap(x, t, a.fl_gamma, l, dx);
  return l;
}

// SlideLoss keeps an exponential average of `auto_iou` per wrapped criterion, updated on EVERY call in level order
// (:50-59; the new value carries alpha = 0.999): class term only at levels with matches, objectness term always
// (auto_iou defaults to 0.5 at a level without matches, :219-222).  One thread: nl sequential updates.
__global__ void k_loss_slide_ema(const LossArgs a) {
  float ema_c = a.slide_state[0], ema_o = a.slide_state[1];
  bool has_c = a.slide_state[2] != 0.f, has_o = a.slide_state[3] != 0.f;
  for (int l = 0; l < a.nl; ++l) {
    const double n = a.nt > 0 ? a.lvl[2 * l] : 0.0;
    float au = 0.5f;
    if (n > 0.0) {
      au = fmaxf((float)(1.0 - a.lvl[2 * l + 1] / n), 0.2f);      // iou.mean(), floored at 0.2 (:48-49)
      if (a.nc > 1) {
        ema_c = has_c ? 0.999f * au + 0.001f * ema_c : au;
        has_c = true;
      }
    }
    ema_o = has_o ? 0.999f * au + 0.001f * ema_o : au;
    has_o = true;
    a.thr[2 * l] = ema_c;
    a.thr[2 * l + 1] = ema_o;
  }
  a.slide_state[0] = ema_c;
  a.slide_state[1] = ema_o;
  a.slide_state[2] = has_c ? 1.f : 0.f;
  a.slide_state[3] = has_o ? 1.f : 0.f;
}

// ---- pass 1 over the candidates: build_targets (:236-288) + box regression (:177-199) --------------------------
__global__ void __launch_bounds__(kThreads) k_loss_match(const LossArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.per_level * a.nl) return;
  const int l = (int)(idx / a.per_level);
  const int64_t row = idx - (int64_t)l * a.per_level;
  const int64_t pairs = (int64_t)a.na * a.nt;
  const int k = (int)(row / pairs);
  const int an = (int)((row - k * pairs) / a.nt);
  const int64_t j = row - k * pairs - (int64_t)an * a.nt;
  const float* t = a.targets + j * 6;
  const int nx = a.nx[l], ny = a.ny[l];
  const float fx = (float)nx, fy = (float)ny;
  const float gx = t[2] * fx, gy = t[3] * fy, gw = t[4] * fx, gh = t[5] * fy;     // targets * gain (:250-253)
  const float aw = a.anchors[(l * a.na + an) * 2], ah = a.anchors[(l * a.na + an) * 2 + 1];
  const float rw = gw / aw, rh = gh / ah;
  bool ok = fmaxf(fmaxf(rw, 1.f / rw), fmaxf(rh, 1.f / rh)) < a.anchor_t;        // (:256-257)
  float ox = 0.f, oy = 0.f;
  if (k == 1) { ok = ok && (fmodf(gx, 1.f) < 0.5f) && (gx > 1.f); ox = 0.5f; }
  if (k == 2) { ok = ok && (fmodf(gy, 1.f) < 0.5f) && (gy > 1.f); oy = 0.5f; }
  if (k == 3) { const float ix = fx - gx; ok = ok && (fmodf(ix, 1.f) < 0.5f) && (ix > 1.f); ox = -0.5f; }
  if (k == 4) { const float iy = fy - gy; ok = ok && (fmodf(iy, 1.f) < 0.5f) && (iy > 1.f); oy = -0.5f; }
  const int64_t b = (int64_t)t[0];
  ok = ok && b >= 0 && b < a.N;
  a.c_cls[idx] = 0.f;
  if (!ok) {
    a.c_cell[idx] = -1;
    a.c_iou[idx] = 0.f;
    return;
  }
  const int gi = min(max((int)(gx - ox), 0), nx - 1);      // .long() truncates; clamp_ acts on gij itself (:278-283)
  const int gj = min(max((int)(gy - oy), 0), ny - 1);
  const int64_t cell = ((b * a.na + an) * ny + gj) * nx + gi;
  const float* ps = a.p[l] + cell * a.no;
  const float s0 = sigmoidf_(ps[0]), s1 = sigmoidf_(ps[1]), s2 = sigmoidf_(ps[2]), s3 = sigmoidf_(ps[3]);
  const float pw2 = s2 * 2.f, ph2 = s3 * 2.f;
  const D4 r = siou(var(s0 * 2.f - 0.5f, 0), var(s1 * 2.f - 0.5f, 1), var(pw2 * pw2 * aw, 2), var(ph2 * ph2 * ah, 3),
                    gx - (float)gi, gy - (float)gj, gw, gh);
  a.c_cell[idx] = (int)cell;
  a.c_iou[idx] = r.v;
  float4 d;
  d.x = r.d[0] * 2.f * s0 * (1.f - s0);
  d.y = r.d[1] * 2.f * s1 * (1.f - s1);
  d.z = r.d[2] * 4.f * pw2 * s2 * (1.f - s2) * aw;
  d.w = r.d[3] * 4.f * ph2 * s3 * (1.f - s3) * ah;
  reinterpret_cast<float4*>(a.c_diou)[idx] = d;
  atomicMax(a.stamp + a.cell_base[l] + cell, (int)row + 1);
}

// ---- per level: n and sum(1 - iou) over the matches, fixed summation order --------------------------------------
__global__ void __launch_bounds__(kThreads) k_loss_level_sums(const LossArgs a) {
  __shared__ double sh[kThreads / 32];
  const int l = blockIdx.x;
  double n = 0.0, s = 0.0;
  for (int64_t r = threadIdx.x; r < a.per_level; r += kThreads) {
    const int64_t idx = (int64_t)l * a.per_level + r;
    if (a.c_cell[idx] >= 0) {
      n += 1.0;
      s += (double)(1.0f - a.c_iou[idx]);
    }
  }
  n = block_sum(n, sh);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    a.lvl[2 * l] = n;
    a.lvl[2 * l + 1] = s;
  }
}

// ---- pass 2 over the candidates: gradients of the box and class terms, class BCE sums (:208-214) ---------------
__global__ void __launch_bounds__(kThreads) k_loss_cand_grad(const LossArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.per_level * a.nl) return;
  const int cell = a.c_cell[idx];
  if (cell < 0) return;
  const int l = (int)(idx / a.per_level);
  const int64_t row = idx - (int64_t)l * a.per_level;
  const float n = (float)a.lvl[2 * l];
  const float bs = (float)a.N;
  float* g = a.gp[l] ? a.gp[l] + (int64_t)cell * a.no : nullptr;
  if (g) {
    const float4 d = reinterpret_cast<const float4*>(a.c_diou)[idx];
    const float w = -a.box * bs / n;                       // lbox += (1 - iou).mean(), times hyp['box'] * bs
    atomicAdd(g + 0, w * d.x);
    atomicAdd(g + 1, w * d.y);
    atomicAdd(g + 2, w * d.z);
    atomicAdd(g + 3, w * d.w);
  }
  if (a.nc > 1) {
    const int64_t j = row % a.nt;
    const int c = (int)a.targets[j * 6 + 1];
    const float* ps = a.p[l] + (int64_t)cell * a.no + 5;
    const float w = a.cls * bs / (n * (float)a.nc);
    const float ema = a.slide_state ? a.thr[2 * l] : 0.f;
    float sum = 0.f;
    for (int q = 0; q < a.nc; ++q) {
      float dx;
      sum += criterion(a, ps[q], q == c ? a.cp : a.cn, a.cls_pw, ema, dx);
      if (g) atomicAdd(g + 5 + q, w * dx);
    }
    a.c_cls[idx] = sum;
  }
}

// ---- objectness over every cell of one level (:219-223) -----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_loss_obj(const LossArgs a, int l) {
  __shared__ double sh[kThreads / 32];
  const int64_t cells = a.cell_base[l + 1] - a.cell_base[l];
  const float* p = a.p[l];
  float* gp = a.gp[l];
  const int* stamp = a.stamp + a.cell_base[l];
  const float* iou = a.c_iou + (int64_t)l * a.per_level;
  const float w = a.obj * (float)a.N * a.balance[l] / (float)cells;
  const float ema = a.slide_state ? a.thr[2 * l + 1] : 0.f;
  double s = 0.0;
  for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < cells; c += (int64_t)gridDim.x * kThreads) {
    const int st = stamp[c];
    const float t = st > 0 ? (1.f - a.gr) + a.gr * fmaxf(iou[st - 1], 0.f) : 0.f;   // iou.detach().clamp(0) (:201-205)
    float dx;
    s += (double)criterion(a, p[c * a.no + 4], t, a.obj_pw, ema, dx);
    if (gp) gp[c * a.no + 4] = w * dx;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) a.obj_part[a.obj_block_base[l] + blockIdx.x] = s;
}

// ---- composition (:224-231) ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_loss_final(const LossArgs a) {
  __shared__ double sh[kThreads / 32];
  __shared__ float s_obj[kMaxLevels], s_cls[kMaxLevels];
  for (int l = 0; l < a.nl; ++l) {
    double so = 0.0, sc = 0.0;
    for (int i = a.obj_block_base[l] + threadIdx.x; i < a.obj_block_base[l + 1]; i += kThreads) so += a.obj_part[i];
    if (a.nt > 0 && a.nc > 1)
      for (int64_t r = threadIdx.x; r < a.per_level; r += kThreads) sc += (double)a.c_cls[(int64_t)l * a.per_level + r];
    so = block_sum(so, sh);
    sc = block_sum(sc, sh);
    if (threadIdx.x == 0) {
      s_obj[l] = (float)(so / (double)(a.cell_base[l + 1] - a.cell_base[l]));
      const double n = a.nt > 0 ? a.lvl[2 * l] : 0.0;
      s_cls[l] = n > 0.0 ? (float)(sc / (n * (double)a.nc)) : 0.f;
    }
  }
  if (threadIdx.x == 0) {
    float lbox = 0.f, lobj = 0.f, lcls = 0.f;
    for (int l = 0; l < a.nl; ++l) {
      const double n = a.nt > 0 ? a.lvl[2 * l] : 0.0;
      if (n > 0.0) lbox += (float)(a.lvl[2 * l + 1] / n);
      lobj += s_obj[l] * a.balance[l];
      lcls += s_cls[l];
      a.out[4 + l] = s_obj[l];
    }
    lbox *= a.box;
    lobj *= a.obj;
    lcls *= a.cls;
    a.out[0] = (lbox + lobj + lcls) * (float)a.N;
    a.out[1] = lbox;
    a.out[2] = lobj;
    a.out[3] = lcls;
  }
}

struct WsLayout {
  size_t stamp, c_iou, c_diou, c_cell, c_cls, obj_part, lvl, thr, total;
};

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

WsLayout ws_layout(int nl, int64_t cells_total, int64_t cands, int obj_blocks) {
  WsLayout w{};
  size_t o = 0;
  w.stamp = o;    o = align_up(o + (size_t)cells_total * 4);
  w.c_iou = o;    o = align_up(o + (size_t)cands * 4);
  w.c_diou = o;   o = align_up(o + (size_t)cands * 16);
  w.c_cell = o;   o = align_up(o + (size_t)cands * 4);
  w.c_cls = o;    o = align_up(o + (size_t)cands * 4);
  w.obj_part = o; o = align_up(o + (size_t)obj_blocks * 8);
  w.lvl = o;      o = align_up(o + (size_t)nl * 16);
  w.thr = o;      o = align_up(o + (size_t)nl * 8);
  w.total = o;
  return w;
}

int obj_grid(int64_t cells) {
  const int64_t want = (cells + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)ecsy_num_sms() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" size_t ecsy_yolo_loss_ws_bytes(int nl, int64_t N, int na, int64_t nt, const int* ny, const int* nx) {
  if (nl < 1 || nl > kMaxLevels || !ny || !nx || N < 0 || na < 1 || nt < 0) return 0;
  int64_t cells = 0;
  int blocks = 0;
  for (int l = 0; l < nl; ++l) {
    const int64_t c = N * na * ny[l] * nx[l];
    cells += c;
    blocks += obj_grid(c);
  }
  return ws_layout(nl, cells, (int64_t)nl * 5 * na * nt, blocks).total;
}

extern "C" int ecsy_yolo_loss(const float* const* p, float* const* gp, const float* targets, int64_t nt,
                              const float* anchors, int nl, int64_t N, int na, int nc, const int* ny, const int* nx,
                              const float* balance, float box, float obj, float cls, float cls_pw, float obj_pw,
                              float cp, float cn, float anchor_t, float gr, float fl_gamma, float* slide_state,
                              float* out, void* ws, size_t ws_bytes, void* stream) {
  ECSY_CHECK_ARG(nl >= 1 && nl <= kMaxLevels, "yolo_loss: 1..%d detection levels, got %d", kMaxLevels, nl);
  ECSY_CHECK_ARG(p && ny && nx && balance && anchors && out, "yolo_loss: null argument");
  ECSY_CHECK_ARG(N >= 1 && na >= 1 && nc >= 1 && nt >= 0, "yolo_loss: bad sizes N=%lld na=%d nc=%d nt=%lld", (long long)N,
                 na, nc, (long long)nt);
  ECSY_CHECK_ARG(nt == 0 || targets, "yolo_loss: null targets");
  ECSY_CHECK_ARG(!(fl_gamma > 0.f && slide_state), "yolo_loss: FocalLoss and SlideLoss cannot be combined (nor can the "
                 "reference: FocalLoss.forward takes no auto_iou)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LossArgs a{};
  a.nl = nl; a.na = na; a.nc = nc; a.no = nc + 5; a.N = N; a.nt = nt; a.per_level = 5LL * na * nt;
  a.targets = targets; a.anchors = anchors;
  a.box = box; a.obj = obj; a.cls = cls; a.cls_pw = cls_pw; a.obj_pw = obj_pw; a.cp = cp; a.cn = cn;
  a.anchor_t = anchor_t; a.gr = gr; a.out = out; a.fl_gamma = fl_gamma; a.slide_state = slide_state;
  a.cell_base[0] = 0;
  a.obj_block_base[0] = 0;
  for (int l = 0; l < nl; ++l) {
    ECSY_CHECK_ARG(p[l] && ny[l] >= 1 && nx[l] >= 1, "yolo_loss: level %d: null output or empty grid", l);
    const int64_t c = N * na * ny[l] * nx[l];
    ECSY_CHECK_ARG(c < (1LL << 31), "yolo_loss: level %d has too many cells", l);
    a.p[l] = p[l]; a.gp[l] = gp ? gp[l] : nullptr;
    a.ny[l] = ny[l]; a.nx[l] = nx[l]; a.balance[l] = balance[l];
    a.cell_base[l + 1] = a.cell_base[l] + c;
    a.obj_block_base[l + 1] = a.obj_block_base[l] + obj_grid(c);
  }
  ECSY_CHECK_ARG(a.per_level < (1LL << 31) - 1, "yolo_loss: too many targets");
  const int64_t cands = a.per_level * nl;
  const WsLayout w = ws_layout(nl, a.cell_base[nl], cands, a.obj_block_base[nl]);
  if (!ws || ws_bytes < w.total) {
    ecsy_set_error("yolo_loss: workspace %zu < %zu bytes", ws_bytes, w.total);
    return ECSY_ERR_WS;
  }
  char* base = static_cast<char*>(ws);
  a.stamp = reinterpret_cast<int*>(base + w.stamp);
  a.c_iou = reinterpret_cast<float*>(base + w.c_iou);
  a.c_diou = reinterpret_cast<float*>(base + w.c_diou);
  a.c_cell = reinterpret_cast<int*>(base + w.c_cell);
  a.c_cls = reinterpret_cast<float*>(base + w.c_cls);
  a.obj_part = reinterpret_cast<double*>(base + w.obj_part);
  a.lvl = reinterpret_cast<double*>(base + w.lvl);
  a.thr = reinterpret_cast<float*>(base + w.thr);
  ECSY_CUDA(cudaMemsetAsync(a.stamp, 0, (size_t)a.cell_base[nl] * 4, st));
  for (int l = 0; l < nl; ++l)
    if (a.gp[l]) ECSY_CUDA(cudaMemsetAsync(a.gp[l], 0, (size_t)(a.cell_base[l + 1] - a.cell_base[l]) * a.no * 4, st));
  if (nt > 0) {
    const unsigned grid = (unsigned)((cands + kThreads - 1) / kThreads);
    k_loss_match<<<grid, kThreads, 0, st>>>(a);
    ECSY_LAUNCH_CHECK();
    k_loss_level_sums<<<nl, kThreads, 0, st>>>(a);
    ECSY_LAUNCH_CHECK();
  }
  if (slide_state != nullptr) {
    k_loss_slide_ema<<<1, 1, 0, st>>>(a);
    ECSY_LAUNCH_CHECK();
  }
  if (nt > 0) {
    const unsigned grid = (unsigned)((cands + kThreads - 1) / kThreads);
    k_loss_cand_grad<<<grid, kThreads, 0, st>>>(a);
    ECSY_LAUNCH_CHECK();
  }
  for (int l = 0; l < nl; ++l) {
    k_loss_obj<<<a.obj_block_base[l + 1] - a.obj_block_base[l], kThreads, 0, st>>>(a, l);
    ECSY_LAUNCH_CHECK();
  }
  k_loss_final<<<1, kThreads, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
