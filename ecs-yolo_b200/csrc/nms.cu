// Batched non-maximum suppression for the decoded Detect output (reference:
// utils/general.py:649-741 `non_max_suppression`, which loops over images in Python and calls torchvision.ops.nms).
//
//   k_nms_candidates : one thread per prediction row [x, y, w, h, obj, cls_0..cls_{nc-1}]: rows with obj > conf_thres,
//                      conf = obj * cls (general.py:691), best class (first maximum) or, with multi_label, every class
//                      above the threshold (general.py:697-702), optional class filter (:705-706).  A candidate is
//                      a 64-bit key (~score bits << 32 | row * nc + class): ascending key order = descending score,
//                      ties by prediction order -- the order of the stable sort inside torchvision's nms.
//   k_nms_select     : one CTA per image.  Up to 8192 candidates: one bitonic sort of the keys in shared memory.  More (the
//                      val setting conf_thres = 0.001 + multi_label: ~78 000 per image): the candidates are consumed in
//                      score order in BATCHES of 8192 -- a radix select (eight 8-bit passes over the L2-resident keys) finds
//                      the 8192-th smallest key above the previous batch, the batch is gathered and sorted in shared memory
//                      and scanned; the next batch is only formed if fewer than max_det boxes survived so far (the first
//                      max_det survivors of a greedy scan do not depend on anything after them, so a full sort of 131 072
//                      padded keys -- 11.9 ms per batch of 64 images, nearly all of the call -- is never needed).
//                      Truncation to max_nms (general.py:716-717), then the
//                      greedy scan in chunks of 256 candidates: every thread tests one candidate against the boxes
//                      kept so far, the chunk's own 256 x 256 suppression bit matrix resolves the order dependence
//                      inside the chunk, and the scan stops at max_det kept boxes (general.py:723-724: the first
//                      max_det survivors of a greedy scan do not depend on anything after them).
// Arithmetic follows the reference to the bit: xywh -> xyxy as x -+ w/2 (general.py:603-609), class offset
// cls * 4096 added to the corners (general.py:720-721), IoU = inter / (area_a + area_b - inter) with separately
// rounded products (torchvision/csrc/ops/cpu/nms_kernel.cpp), suppressed when IoU > iou_thres.
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"

namespace {

constexpr int kChunk = 256;         // candidates resolved per round of the greedy scan
constexpr int kSplit = 4;           // threads per candidate of a chunk: the kept list / the chunk's columns are split kSplit ways
constexpr int kSelThreads = kChunk * kSplit;
constexpr int kSortSmemKeys = 8192; // 64 KB of keys sorted in shared memory
constexpr float kMaxWh = 4096.f;    // general.py:667

struct NmsArgs {
  const float* pred;     // [N][R][5 + nc]
  const uint8_t* cls_ok; // optional [nc]: 1 = keep this class (general.py:705)
  unsigned long long* keys;  // [N][cap]
  int* counts;           // [N]
  float* out;            // [N][max_det][6]
  int* out_count;        // [N]
  int N, R, nc, cap, max_det, max_nms, agnostic, multi_label;
  float conf_thres;
  double iou_thres;
};

__device__ __forceinline__ unsigned long long make_key(float score, uint32_t idx) {
  return ((unsigned long long)(~__float_as_uint(score)) << 32) | idx;   // score > 0: raw bits are monotonic
}

__global__ void k_nms_candidates(const NmsArgs a) {
  const int img = blockIdx.y;
  const int no = 5 + a.nc;
  unsigned long long* keys = a.keys + (size_t)img * a.cap;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.R; r += gridDim.x * blockDim.x) {
    const float* p = a.pred + ((size_t)img * a.R + r) * no;
    const float obj = p[4];
    if (!(obj > a.conf_thres)) continue;
    if (a.multi_label) {
      for (int j = 0; j < a.nc; ++j) {
        const float conf = __fmul_rn(p[5 + j], obj);
        if (conf > a.conf_thres && (a.cls_ok == nullptr || a.cls_ok[j])) {
          const int slot = atomicAdd(a.counts + img, 1);
          keys[slot] = make_key(conf, (uint32_t)r * (uint32_t)a.nc + (uint32_t)j);
        }
      }
    } else {
      float best = __fmul_rn(p[5], obj);
      int bj = 0;
      for (int j = 1; j < a.nc; ++j) {
        const float conf = __fmul_rn(p[5 + j], obj);
        if (conf > best) { best = conf; bj = j; }
      }
      if (best > a.conf_thres && (a.cls_ok == nullptr || a.cls_ok[bj])) {
        const int slot = atomicAdd(a.counts + img, 1);
        keys[slot] = make_key(best, (uint32_t)r * (uint32_t)a.nc + (uint32_t)bj);
      }
    }
  }
}

// Strides 2^ls_start .. 1 of bitonic stage `size = 1 << lsize` on the CH keys of `s` (shared memory) whose first key has
// global index `base`: the direction of a compare-exchange depends on the GLOBAL index of its lower element.
__device__ void bitonic_chunk(unsigned long long* s, int base, int CH, int lsize, int ls_start) {
  const int size = 1 << lsize;
  for (int ls = ls_start; ls >= 0; --ls) {
    const int stride = 1 << ls;
    for (int i = threadIdx.x; i < (CH >> 1); i += blockDim.x) {
      const int lo = ((i >> ls) << (ls + 1)) | (i & (stride - 1));
      const int hi = lo + stride;
      const bool up = ((base + lo) & size) == 0;
      const unsigned long long x = s[lo], y = s[hi];
      if ((x > y) == up) { s[lo] = y; s[hi] = x; }
    }
    __syncthreads();
  }
}

// k-th smallest (k >= 1) of the keys in g[0, n) that are > prev (all of them when !have_prev): MSB-first radix select with
// 8-bit digits, one histogram pass over the keys per digit.  Keys are unique (the low word is the candidate index), so the
// result identifies exactly k keys in (prev, result].  Every thread of the CTA must call it; all return the same value.
__device__ unsigned long long radix_select_kth(const unsigned long long* __restrict__ g, int n, bool have_prev,
                                               unsigned long long prev, int k, uint32_t* s_hist /*[256]*/,
                                               unsigned long long* s_prefix, int* s_k) {
  unsigned long long prefix = 0, mask = 0;
  for (int d = 7; d >= 0; --d) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const int sh = 8 * d;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = g[i];
      if ((!have_prev || key > prev) && (key & mask) == prefix) atomicAdd(&s_hist[(uint32_t)(key >> sh) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kk = k, b = 0;
      for (; b < 255; ++b) {
        const int h = (int)s_hist[b];
        if (kk <= h) break;
        kk -= h;
      }
      *s_k = kk;
      *s_prefix = prefix | ((unsigned long long)b << sh);
    }
    __syncthreads();
    k = *s_k;
    prefix = *s_prefix;
    mask |= 0xffull << sh;
    __syncthreads();
  }
  return prefix;
}

struct Cand {
  float x1, y1, x2, y2;   // class-offset corners (what torchvision's nms sees)
  float area;
};

__device__ __forceinline__ bool iou_gt(const Cand& a, const Cand& b, double thr) {
  const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
  const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  // no overlap (nearly every pair: the class offset separates classes): inter = 0, the reference's quotient is 0 (or NaN for
  // two empty boxes) and never exceeds a threshold >= 0 -- same answer without the IEEE division and the double compare
  if (!(w > 0.f && h > 0.f)) return false;
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
  return (double)ovr > thr;
}

__global__ void __launch_bounds__(kSelThreads, 1) k_nms_select(const NmsArgs a) {
  extern __shared__ __align__(16) uint8_t nms_sm[];
  __shared__ Cand s_chunk[kChunk];
  __shared__ uint32_t s_sup[kChunk][kChunk / 32];   // s_sup[i]: candidates of the chunk that i suppresses
  __shared__ uint32_t s_alive[kChunk / 32];
  __shared__ int s_newk[kChunk];
  __shared__ int s_dead[kChunk];
  __shared__ int s_nnew, s_kept, s_cnt, s_k;
  __shared__ uint32_t s_hist[256];
  __shared__ unsigned long long s_prefix;
  const int img = blockIdx.x;
  const int no = 5 + a.nc;
  unsigned long long* gkeys = a.keys + (size_t)img * a.cap;
  int n = a.counts[img];
  // kept boxes live in dynamic shared memory: [max_det] Cand
  Cand* s_keep = reinterpret_cast<Cand*>(nms_sm);
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(nms_sm + (((size_t)a.max_det * sizeof(Cand) + 15) & ~size_t(15)));

  if (threadIdx.x == 0) s_kept = 0;
  if (n == 0) {
    if (threadIdx.x == 0) a.out_count[img] = 0;
    return;
  }
  const int n_lim = n > a.max_nms ? a.max_nms : n;   // general.py:716-717
  const bool small = n <= kSortSmemKeys;
  const int n_all = n;
  const unsigned long long* keys = s_keys;
  const float* pred = a.pred + (size_t)img * a.R * no;
  float* out = a.out + (size_t)img * a.max_det * 6;
  int processed = 0;
  bool have_prev = false;
  unsigned long long prev = 0;
  while (processed < n_lim) {
  // ---- the next batch of candidates in score order: s_keys[0, nb) ascending ----
  int nb, nsort;
  if (small) {
    nsort = n_all;
    nb = n_lim;
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) s_keys[i] = gkeys[i];
  } else {
    nb = nsort = (n_lim - processed) < kSortSmemKeys ? (n_lim - processed) : kSortSmemKeys;
    const unsigned long long kth = radix_select_kth(gkeys, n_all, have_prev, prev, nb, s_hist, &s_prefix, &s_k);
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_all; i += blockDim.x) {
      const unsigned long long key = gkeys[i];
      if ((!have_prev || key > prev) && key <= kth) s_keys[atomicAdd(&s_cnt, 1)] = key;
    }
    prev = kth;
    have_prev = true;
  }
  {
    int P = 1, lp = 0;
    while (P < nsort) { P <<= 1; ++lp; }
    for (int i = nsort + threadIdx.x; i < P; i += blockDim.x) s_keys[i] = ~0ull;
    __syncthreads();
    for (int lsize = 1; lsize <= lp; ++lsize) bitonic_chunk(s_keys, 0, P, lsize, lsize - 1);
  }
  n = nb;   // the scan below runs over this batch
  for (int base = 0; base < n; base += kChunk) {
    const int t = threadIdx.x % kChunk, part = threadIdx.x / kChunk;   // kSplit threads share candidate t
    const int c = base + t;
    const bool valid = c < n;
    Cand me{0.f, 0.f, 0.f, 0.f, 0.f};
    float bx1 = 0.f, by1 = 0.f, bx2 = 0.f, by2 = 0.f, conf = 0.f;
    int cls = 0;
    if (valid) {
      const unsigned long long key = keys[c];
      const uint32_t idx = (uint32_t)(key & 0xffffffffull);
      const int r = idx / a.nc;
      cls = idx - r * a.nc;
      const float* p = pred + (size_t)r * no;
      const float hw = __fmul_rn(p[2], 0.5f), hh = __fmul_rn(p[3], 0.5f);   // w / 2 is exact either way
      bx1 = __fsub_rn(p[0], hw); by1 = __fsub_rn(p[1], hh);
      bx2 = __fadd_rn(p[0], hw); by2 = __fadd_rn(p[1], hh);
      conf = __fmul_rn(p[5 + cls], p[4]);
      const float off = a.agnostic ? 0.f : __fmul_rn((float)cls, kMaxWh);
      me.x1 = __fadd_rn(bx1, off); me.y1 = __fadd_rn(by1, off);
      me.x2 = __fadd_rn(bx2, off); me.y2 = __fadd_rn(by2, off);
      me.area = __fmul_rn(__fsub_rn(me.x2, me.x1), __fsub_rn(me.y2, me.y1));
    }
    if (part == 0) {
      s_chunk[t] = me;
      s_dead[t] = 0;
    }
    __syncthreads();
    // (a) against the boxes kept by earlier chunks: the kSplit threads of a candidate take every kSplit-th kept box
    const int kept0 = s_kept;
    if (valid) {
      bool dead = false;
      for (int k = part; !dead && k < kept0; k += kSplit)
        if (iou_gt(s_keep[k], me, a.iou_thres)) dead = true;
      if (dead) s_dead[t] = 1;
    }
    __syncthreads();
    const bool alive = valid && s_dead[t] == 0;
    if (part == 0) {   // threads 0 .. kChunk-1 = whole warps
      const uint32_t ball = __ballot_sync(0xffffffffu, alive);
      if ((t & 31) == 0) s_alive[t >> 5] = ball;
    }
    // (b) the chunk's own suppression matrix: row t = later candidates of the chunk that t would suppress
    {
      // a candidate that is already dead never suppresses anything in the greedy scan: its row stays empty
      // thread `part` of the candidate fills the words of its kChunk / kSplit columns
      uint32_t word = 0;
      for (int u = part * (kChunk / kSplit); u < (part + 1) * (kChunk / kSplit); ++u) {
        const bool s = alive && u > t && (base + u) < n && iou_gt(me, s_chunk[u], a.iou_thres);
        word |= (s ? 1u : 0u) << (u & 31);
        if ((u & 31) == 31) { s_sup[t][u >> 5] = word; word = 0; }
      }
    }
    __syncthreads();
    // (c) sequential resolution by one thread: 256 steps over 8-word masks
    if (threadIdx.x == 0) {
      uint32_t al[kChunk / 32];
#pragma unroll
      for (int w = 0; w < kChunk / 32; ++w) al[w] = s_alive[w];
      int nn = 0, kept = s_kept;
      for (int i = 0; i < kChunk && kept + nn < a.max_det; ++i) {
        uint32_t bit = 0;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w)
          if ((i >> 5) == w) bit = (al[w] >> (i & 31)) & 1u;
        if (!bit) continue;
        s_newk[nn++] = i;
#pragma unroll
        for (int w = 0; w < kChunk / 32; ++w) al[w] &= ~s_sup[i][w];
      }
      s_nnew = nn;
    }
    __syncthreads();
    // (d) append the survivors (in score order) to the kept list and the output
    const int nn = s_nnew, kept = s_kept;
    for (int j = 0; part == 0 && j < nn; ++j) {
      if (s_newk[j] == t) {
        s_keep[kept + j] = me;
        float* o = out + (size_t)(kept + j) * 6;
        o[0] = bx1; o[1] = by1; o[2] = bx2; o[3] = by2; o[4] = conf; o[5] = (float)cls;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_kept = kept + nn;
    __syncthreads();
    if (s_kept >= a.max_det) break;
  }
  processed += nb;
  __syncthreads();
  if (s_kept >= a.max_det) break;
  }   // batches
  if (threadIdx.x == 0) a.out_count[img] = s_kept;
}

inline size_t al256n(size_t v) { return (v + 255) & ~size_t(255); }

int cap_for(int R, int nc, int multi_label) {
  const long long m = (long long)R * (multi_label ? nc : 1);
  long long cap = 1;
  while (cap < m) cap <<= 1;
  return (int)cap;
}

}  // namespace

extern "C" size_t ecsy_nms_ws_bytes(int64_t N, int R, int nc, int multi_label) {
  if (N <= 0 || R <= 0 || nc <= 0) return 0;
  return 512 + al256n((size_t)N * sizeof(int)) + (size_t)N * cap_for(R, nc, multi_label) * sizeof(unsigned long long);
}

extern "C" int ecsy_nms(const float* pred, int64_t N, int R, int nc, float conf_thres, double iou_thres, int agnostic,
                        int multi_label, const uint8_t* cls_ok, int max_det, int max_nms, float* out, int* out_count,
                        void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(pred && out && out_count && N > 0 && N < 65536 && R > 0 && nc > 0, "nms: bad arguments");
  ECSY_CHECK_ARG(conf_thres >= 0.f && conf_thres <= 1.f && iou_thres >= 0.0 && iou_thres <= 1.0,
                 "nms: thresholds must lie in [0, 1] (utils/general.py:661-662)");
  ECSY_CHECK_ARG(max_det > 0 && max_det <= 4096 && max_nms > 0, "nms: max_det in [1, 4096], max_nms > 0");
  ECSY_CHECK_ARG((long long)R * nc < (1LL << 31), "nms: too many (row, class) pairs");
  if (nc == 1) multi_label = 0;   // general.py:669
  const size_t need = ecsy_nms_ws_bytes(N, R, nc, multi_label);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("nms: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  NmsArgs a{};
  uintptr_t p = (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255);
  a.counts = reinterpret_cast<int*>(p); p += al256n((size_t)N * sizeof(int));
  a.keys = reinterpret_cast<unsigned long long*>(p);
  a.pred = pred; a.cls_ok = cls_ok; a.out = out; a.out_count = out_count;
  a.N = (int)N; a.R = R; a.nc = nc; a.cap = cap_for(R, nc, multi_label);
  a.max_det = max_det; a.max_nms = max_nms; a.agnostic = agnostic; a.multi_label = multi_label;
  a.conf_thres = conf_thres;
  a.iou_thres = iou_thres;   // a double, like the argument of torchvision.ops.nms
  ECSY_CUDA(cudaMemsetAsync(a.counts, 0, (size_t)N * sizeof(int), st));
  dim3 grid((unsigned)((R + 255) / 256), (unsigned)N);
  if (grid.x > 64) grid.x = 64;
  k_nms_candidates<<<grid, 256, 0, st>>>(a);
  ECSY_LAUNCH_CHECK();
  const size_t keep_bytes = (((size_t)max_det * sizeof(Cand) + 15) & ~size_t(15));
  const size_t smem = keep_bytes + (size_t)kSortSmemKeys * sizeof(unsigned long long);
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(k_nms_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  k_nms_select<<<(unsigned)N, kSelThreads, smem, st>>>(a);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
