// out[Ca][Cb] += alpha * sum_p P[p][ca] * Q[p][cb]   (contraction over the ROWS p of two row-major bf16
// matrices) on tcgen05: the weight-gradient shape of the ECS point-wise spread (and of 1x1 convs):
//   dW_pw[co][ci] = alpha * sum_pixels ge[p][co] * dw(s)[p][ci]        (SURVEY Appendix B).
// Both operands are loaded by TMA as {64 columns, 64 rows} boxes and handed to the tensor core as
// MN-major tiles (the row index is the MMA K dimension), so no transposed copies are materialised.
// Optional hi/lo bf16 planes for both operands: P_hi*Q_hi + P_lo*Q_hi + P_hi*Q_lo.
// One CTA owns one 128 x NB output tile and a slice of the rows; the accumulator stays in TMEM for the
// whole slice and is added to the fp32 output with red.global (order of CTAs is not deterministic).
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

using namespace ecsy;

namespace {

constexpr int kRows = 64;                // contraction rows per stage
constexpr int kBlk = kRows * 128;        // bytes of one {64 cols, 64 rows} bf16 box
constexpr int kXStages = 4;

struct XCtl {
  uint64_t full[kXStages];
  uint64_t empty[kXStages];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct XArgs {
  int64_t rows;       // contraction length
  int Ca, Cb;         // output rows / cols
  int nb;             // output columns handled per CTA (multiple of 64, <= 256)
  int row_splits;     // CTAs along the contraction
  int stages;         // smem ring depth (<= kXStages)
  float alpha;
  float* out;         // [Ca][Cb] fp32, accumulated with atomics
};

template <int SPLIT>
__global__ void __launch_bounds__(192, 1)
k_umma_xty(const __grid_constant__ CUtensorMap tm_p0, const __grid_constant__ CUtensorMap tm_p1,
           const __grid_constant__ CUtensorMap tm_q0, const __grid_constant__ CUtensorMap tm_q1, const XArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nqb = g.nb / 64;                                  // Q blocks per stage plane
  const int stage_bytes = SPLIT * (2 + nqb) * kBlk;           // P: two 64-col blocks (128 output rows)
  XCtl* ctl = reinterpret_cast<XCtl*>(smem + g.stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = blockIdx.y * 128;          // first output row (column of P)
  const int b0 = blockIdx.z * g.nb;         // first output column (column of Q)
  // rows of this CTA: contiguous slice in units of kRows
  const int64_t tiles = (g.rows + kRows - 1) / kRows;
  const int64_t per = (tiles + g.row_splits - 1) / g.row_splits;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = t_begin + per < tiles ? t_begin + per : tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<256>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 4) {
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      mbar_wait(&ctl->empty[stage], phase ^ 1);
      if (lane == 0) {
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        const int r0 = (int)((t_begin + i) * kRows);
        mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)stage_bytes);
#pragma unroll
        for (int sp = 0; sp < SPLIT; ++sp) {
          uint8_t* pl = st + sp * (2 + nqb) * kBlk;
          const CUtensorMap* tp = sp == 0 ? &tm_p0 : &tm_p1;
          const CUtensorMap* tq = sp == 0 ? &tm_q0 : &tm_q1;
          tma_load_2d(pl, tp, &ctl->full[stage], a0, r0);                 // columns past Ca: zero fill
          tma_load_2d(pl + kBlk, tp, &ctl->full[stage], a0 + 64, r0);
          for (int qb = 0; qb < nqb; ++qb)
            tma_load_2d(pl + (2 + qb) * kBlk, tq, &ctl->full[stage], b0 + qb * 64, r0);
        }
      }
      __syncwarp();
      if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    const uint32_t idesc = umma_idesc_bf16_mn(128, g.nb);
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      mbar_wait(&ctl->full[stage], phase);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
        uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
        for (int combo = 0; combo < 2 * SPLIT - 1; ++combo) {
          const int ps = (SPLIT == 2 && combo == 1) ? 1 : 0;
          const int qs = (SPLIT == 2 && combo == 2) ? 1 : 0;
          const uint32_t pa = base + ps * (2 + nqb) * kBlk;
          const uint32_t qa = base + qs * (2 + nqb) * kBlk + 2 * kBlk;
#pragma unroll
          for (int k = 0; k < kRows / 16; ++k) {
            // 16 contraction rows = 2048 bytes further along k
            const uint64_t da = umma_desc_sw128_mn(pa + k * 2048, kBlk);
            const uint64_t db = umma_desc_sw128_mn(qa + k * 2048, kBlk);
            umma_f16(tmem_base, da, db, idesc, acc);
            acc = 1u;
          }
        }
        umma_commit(&ctl->empty[stage]);
        if (i == ntiles - 1) umma_commit(&ctl->done);
      }
      __syncwarp();
      if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp < 4) {
    if (ntiles > 0) {
      mbar_wait(&ctl->done, 0);
      tc_fence_after_sync();
      const int ra = a0 + warp * 32 + lane;
      for (int c0 = 0; c0 < g.nb; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (ra < g.Ca) {
          // Cb is a multiple of 64 and nb divides it: the 32 columns are in range; 16-byte vector reductions (a quarter
          // of the L2 atomic operations -- the small layers of this kernel were bound by them)
          float* dst = g.out + (int64_t)ra * g.Cb + b0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                         "f"(g.alpha * __uint_as_float(v[j])), "f"(g.alpha * __uint_as_float(v[j + 1])),
                         "f"(g.alpha * __uint_as_float(v[j + 2])), "f"(g.alpha * __uint_as_float(v[j + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

// P: [rows][Ca], Q: [rows][Cb] bf16 row-major (hi planes, optional lo planes).  out[Ca][Cb] += alpha * P^T Q.
int ecsy_umma_xty(const void* p_hi, const void* p_lo, const void* q_hi, const void* q_lo, int64_t rows, int Ca, int Cb,
                  float alpha, float* out, cudaStream_t st) {
  ECSY_CHECK_ARG(p_hi && q_hi && out && rows > 0, "xty: bad arguments");
  ECSY_CHECK_ARG(Ca % 64 == 0 && Cb % 64 == 0, "xty: Ca=%d and Cb=%d must be multiples of 64", Ca, Cb);
  ECSY_CHECK_ARG((p_lo == nullptr) == (q_lo == nullptr), "xty: lo planes come in pairs");
  const int split = p_lo ? 2 : 1;
  int nb = split == 2 ? 128 : 256;
  while (Cb % nb != 0) nb >>= 1;
  CUtensorMap tp0, tp1{}, tq0, tq1{};
  int rc = ecsy_tensor_map_bf16(p_hi, (uint64_t)rows, (uint64_t)Ca, kRows, &tp0);
  if (rc) return rc;
  rc = ecsy_tensor_map_bf16(q_hi, (uint64_t)rows, (uint64_t)Cb, kRows, &tq0);
  if (rc) return rc;
  if (split == 2) {
    rc = ecsy_tensor_map_bf16(p_lo, (uint64_t)rows, (uint64_t)Ca, kRows, &tp1);
    if (rc) return rc;
    rc = ecsy_tensor_map_bf16(q_lo, (uint64_t)rows, (uint64_t)Cb, kRows, &tq1);
    if (rc) return rc;
  }
  XArgs g{};
  g.rows = rows; g.Ca = Ca; g.Cb = Cb; g.nb = nb; g.alpha = alpha; g.out = out;
  const int ya = (Ca + 127) / 128, zb = Cb / nb;
  const int64_t tiles = (rows + kRows - 1) / kRows;
  int64_t rs = (2 * (int64_t)ecsy_num_sms()) / (ya * zb);
  // every CTA ends with 128 x nb reductions into the output: give it at least 16 row tiles of contraction to amortise them
  if (rs > tiles / 16) rs = tiles / 16;
  if (rs < 1) rs = 1;
  g.row_splits = (int)rs;
  const int stage_bytes = split * (2 + nb / 64) * kBlk;
  int stages = (220 * 1024) / stage_bytes;
  if (stages > kXStages) stages = kXStages;
  g.stages = stages;
  const int smem = 1024 + stages * stage_bytes + (int)sizeof(XCtl) + 64;
  if (split == 1) {
    static bool done1 = false;
    if (!done1) { ECSY_CUDA(cudaFuncSetAttribute(k_umma_xty<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done1 = true; }
    k_umma_xty<1><<<dim3((unsigned)rs, ya, zb), 192, smem, st>>>(tp0, tp1, tq0, tq1, g);
  } else {
    static bool done2 = false;
    if (!done2) { ECSY_CUDA(cudaFuncSetAttribute(k_umma_xty<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done2 = true; }
    k_umma_xty<2><<<dim3((unsigned)rs, ya, zb), 192, smem, st>>>(tp0, tp1, tq0, tq1, g);
  }
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// Test / integration entry: fp32 inputs are split into bf16 planes by the caller (ecsy_f32_to_bf16_planes).
extern "C" int ecsy_xty_bf16(const void* p_hi, const void* p_lo, const void* q_hi, const void* q_lo, int64_t rows,
                             int Ca, int Cb, float alpha, float* out, void* stream) {
  return ecsy_umma_xty(p_hi, p_lo, q_hi, q_lo, rows, Ca, Cb, alpha, out, reinterpret_cast<cudaStream_t>(stream));
}
