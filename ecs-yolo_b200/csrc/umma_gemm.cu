// tcgen05 / TMEM / TMA implicit-GEMM engine of the spiking hot path (sm_100a).
//
// One persistent, warp-specialised kernel template computes D[128 x BN] tiles as
//     D = A[128 x K] * B[BN x K]^T      (bf16 operands, fp32 accumulation in TMEM)
// with
//   A_MODE = kASpikes : A rows are output pixels of a convolution over BIT-PACKED spikes
//                       ([imgs][H][W][Cin/32] uint32).  Expander warps read the 64 spike bits of every
//                       (pixel + tap, 64-channel slab) from the bit tensor (8 bytes per operand row and K
//                       block, L2-resident) and materialise the bf16 {0,1} operand tile directly in the
//                       128-byte-swizzled UMMA layout -- spikes never exist as bf16/fp32 in HBM.
//                       (Snn_Conv2d on spikes, models/common.py:593-624.)
//   A_MODE = kATma    : A is a row-major bf16 matrix [M][K] in HBM loaded by TMA (ECS point-wise
//                       spread, im2col'ed real-input convolutions).
//   A_MODE = kATma4   : A rows are output pixels of a stride-1 convolution over a REAL-valued bf16 NHWC tensor
//                       (hi [+ lo] planes): every (tap, 64-channel slab) K block is one 4-D TMA box
//                       {64 ch, tw, th, tn} shifted by the tap, out-of-bounds = zero padding.  No im2col
//                       buffer.  (Snn_Conv2d on real inputs, class Conv; conv input-gradients.)
//   A_MODE = kADw     : A rows are pixels of ONE timestep and A[p][c] = b[c] + sum_tap s(p+tap, c) * w[tap][c], the
//                       depth-wise half of the ECS spread (models/common.py:289-294), computed by the producer
//                       warps from the spike bits (8 channels per thread, weights in registers) straight into the
//                       swizzled operand tile -- the dw output never goes to HBM.  B = point-wise weights.
//   B                 : packed weights [B_SPLIT*Cout][K] bf16 (hi plane, optional lo residual plane),
//                       loaded by TMA.  With B_SPLIT=2 (and A_SPLIT=2 for real-valued A) the products
//                       A_hi*B_hi + A_lo*B_hi + A_hi*B_lo reproduce fp32 weights to ~2^-17.
//   EPI = kEpiConv    : y = acc*scale[c] + shift[c] (+ residual) -> fp32 NHWC   (tdBN folded)
//
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp id), 4 TMA producer, 5 MMA issuer + TMEM
// allocator, 6-13 spike expanders (kASpikes only).  smem ring of `stages` {A,B} slots guarded by
// full/empty mbarriers; two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of
// tile i+1.
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>

#include <cuda_fp16.h>
#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

using namespace ecsy;

namespace {

constexpr int kMaxStages = 8;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kEpiWarpFloats = 32 * 36 + 128;  // per epilogue warp: 32 x 36 float staging tile + 32 x 2 int64 row offsets
constexpr int kExpWarps = 8;               // spike expander warps, grouped per pipeline stage
constexpr int kSpikeThreads = 192 + kExpWarps * 32;

struct SharedCtl {
  uint64_t full_a[kMaxStages];
  uint64_t full_b[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// ---- epilogue parameter blocks ----
struct EpiConv {
  float* out;             // [rows][ldc]
  const float* scale;     // [Cout] or null
  const float* shift;     // [Cout] or null
  const float* residual;  // [res_rows][ldc] or null
  int64_t res_rows;       // rows in residual (T-broadcast sources repeat)
  int ldc;                // Cout
  int out_half;           // out points to __half (fp16 storage of the ECS spread in fast mode)
};

struct SpikeGeom {
  const uint32_t* bits;  // [imgs][H][W][Cw]
  int imgs, H, W, Cw;    // Cw = Cin/32
  int Ho, Wo;
  int kh, kw, stride, pad;
  int tn_b, th_b, tw_b;              // tile box (powers of two, product 128)
  int tn_sh, th_sh, tw_sh;           // log2 of the above
  int tiles_h, tiles_w;              // tiles per image dimension (tiles over imgs = m_tiles/(tiles_h*tiles_w))
  int Hp, Wp, PP;                    // patch rows/cols per image, patch pixels per tile
  int nslab;                         // Cin/64
};

struct GemmArgs {
  int m_tiles, n_tiles, kb_total, stages;
  uint32_t epi_off;  // byte offset of the epilogue staging buffers (4 warps x kEpiWarpFloats floats)
  int64_t M;       // valid rows (kATma, kADw) / unused (kASpikes)
  const float* dw_w;   // kADw: depth-wise weights [9][C]
  const float* dw_b;   // kADw: depth-wise bias [C]
  int wpg;             // expander warps per stage group (kExpWarps / stages)
  // kAGather: im2col on the fly from a real-valued fp32 NHWC tensor (small Cin: the stem)
  const float* gx;     // [gx_imgs][H][W][gcin]
  int64_t gx_imgs;     // images in gx (T-broadcast sources repeat)
  int gcin, gK;        // input channels, K = kh*kw*gcin (kb_total*64 >= gK, zero padded)
  uint32_t ktab_off;   // byte offset of the k -> (offset, ky, kx) table in shared memory
  int pair_expand;     // kASpikes: weights in the ecsy_pack_spike_conv_weight layout, spikes emitted as {0, 2.0} pairs
};

template <int EPI>
struct EpiSel;
template <>
struct EpiSel<0> { using type = EpiConv; };

constexpr int kATma = 0, kASpikes = 1, kATma4 = 2, kADw = 3, kASpikesT = 4, kAGather = 5;
constexpr int kEpiConv = 0;

__device__ __forceinline__ uint32_t bits2_to_bf16x2(uint32_t x) {
  // bit0 -> low bf16 (1.0 = 0x3F80), bit1 -> high bf16
  return ((x & 1u) | ((x & 2u) << 15)) * 0x3F80u;
}

template <int BN, int A_MODE, int A_SPLIT, int B_SPLIT, int EPI>
__global__ void __launch_bounds__((A_MODE == kASpikes || A_MODE == kADw || A_MODE == kASpikesT || A_MODE == kAGather) ? kSpikeThreads : 192, 1)
k_umma_gemm(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
            const __grid_constant__ CUtensorMap tm_b, const GemmArgs g, const SpikeGeom sg,
            const typename EpiSel<EPI>::type ep) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBTileBytes = BN * 128;
  constexpr bool kTS = A_MODE == kASpikesT;   // A operand lives in tensor memory (no smem A tile)
  constexpr int kASmemBytes = kTS ? 0 : A_SPLIT * kATileBytes;
  constexpr int kStageBytes = kASmemBytes + B_SPLIT * kBTileBytes;
  // two accumulator buffers (128, 256 or 512 columns); TS mode adds a ring of 32-column A stages behind them
  constexpr int kTmemCols = kTS ? 512 : 2 * BN;
  static_assert(!kTS || BN <= 128, "TS mode: accumulators + A ring must fit 512 TMEM columns");
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(smem + (size_t)g.stages * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    if (A_MODE == kATma || A_MODE == kATma4) {
      tma_prefetch_desc(&tm_a0);
      if (A_SPLIT == 2) tma_prefetch_desc(&tm_a1);
    }
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&ctl->full_a[s], (A_MODE == kADw || A_MODE == kAGather) ? kExpWarps : (kTS ? 4 : g.wpg));
      mbar_init(&ctl->full_b[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->tmem_full[b], 1);
      mbar_init(&ctl->tmem_empty[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(&ctl->tmem_base);
  if (A_MODE == kAGather) {   // k -> {float offset inside the receptive field, (ky << 16) | kx}; ky = 0xFFFF: zero padding of K
    int2* ktab = reinterpret_cast<int2*>(smem + g.ktab_off);
    for (int k = threadIdx.x; k < g.kb_total * 64; k += blockDim.x) {
      int2 e = make_int2(0, (int)0xFFFF0000u);
      if (k < g.gK) {
        const int tap = k / g.gcin, ci = k - tap * g.gcin;
        const int ky = tap / sg.kw, kx = tap - ky * sg.kw;
        e = make_int2((ky * sg.W + kx) * g.gcin + ci, (ky << 16) | kx);
      }
      ktab[k] = e;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 4) {
    // =============================== TMA producer ===============================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
      int img0 = 0, h0 = 0, w0 = 0;
      if (A_MODE == kATma4) {
        const int tiles_hw = sg.tiles_h * sg.tiles_w;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        img0 = tn * sg.tn_b; h0 = th * sg.th_b - sg.pad; w0 = tw * sg.tw_b - sg.pad;
      }
      int slab = 0, ky = 0, kx = 0;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* st = smem + (size_t)stage * kStageBytes;
          constexpr uint32_t tx = ((A_MODE == kATma || A_MODE == kATma4) ? A_SPLIT * kATileBytes : 0) + B_SPLIT * kBTileBytes;
          mbar_arrive_expect_tx(&ctl->full_b[stage], tx);
          if (A_MODE == kATma) {
            tma_load_2d(st, &tm_a0, &ctl->full_b[stage], kb * 64, m_tile * 128);
            if (A_SPLIT == 2) tma_load_2d(st + kATileBytes, &tm_a1, &ctl->full_b[stage], kb * 64, m_tile * 128);
          }
          if (A_MODE == kATma4) {
            tma_load_4d(st, &tm_a0, &ctl->full_b[stage], slab * 64, w0 + kx, h0 + ky, img0);
            if (A_SPLIT == 2)
              tma_load_4d(st + kATileBytes, &tm_a1, &ctl->full_b[stage], slab * 64, w0 + kx, h0 + ky, img0);
          }
#pragma unroll
          for (int bs = 0; bs < B_SPLIT; ++bs)
            tma_load_2d(st + kASmemBytes + bs * kBTileBytes, &tm_b, &ctl->full_b[stage], kb * 64,
                        bs * (g.n_tiles * BN) + n_tile * BN);
        }
        __syncwarp();
        if (A_MODE == kATma4) {
          if (++slab == sg.nslab) {
            slab = 0;
            if (++kx == sg.kw) { kx = 0; ++ky; }
          }
        }
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&ctl->tmem_empty[buf], bphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->full_b[stage], phase);
        if (A_MODE == kASpikes || A_MODE == kADw || A_MODE == kAGather || kTS) mbar_wait(&ctl->full_a[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint32_t b_addr = a_addr + kASmemBytes;
          uint32_t acc = kb > 0 ? 1u : 0u;
          if constexpr (kTS) {
            const uint32_t a_tmem = tmem_base + 2 * BN + stage * 32;
#pragma unroll
            for (int bs = 0; bs < B_SPLIT; ++bs) {
              const uint64_t db = umma_desc_sw128(b_addr + bs * kBTileBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16_ts(d_tmem, a_tmem + k * 8, db + (uint64_t)(k * 2), idesc, acc);
                acc = 1u;
              }
            }
          } else {
#pragma unroll
          for (int combo = 0; combo < A_SPLIT + B_SPLIT - 1; ++combo) {
            // combos: (A0,B0) [, (A1,B0)] [, (A0,B1)]  -- the lo*lo term is below fp32 resolution
            const int as = (A_SPLIT == 2 && combo == 1) ? 1 : 0;
            const int bs = (B_SPLIT == 2 && combo == A_SPLIT) ? 1 : 0;
            const uint64_t da = umma_desc_sw128(a_addr + as * kATileBytes);
            const uint64_t db = umma_desc_sw128(b_addr + bs * kBTileBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row (encoded >> 4)
              umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
              acc = 1u;
            }
          }
          }
          umma_commit(&ctl->empty[stage]);
          if (kb == g.kb_total - 1) umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 6 && A_MODE == kAGather) {
    // =============================== im2col-on-the-fly producers (real input, small Cin) ===============================
    // 256 threads build every [128 rows x 64 k] operand tile: thread = (row, half of the K block); the receptive
    // field of an output pixel is read straight from the fp32 NHWC tensor (L1/L2 resident: neighbouring pixels
    // overlap), converted to bf16 and written in the 128-byte-swizzled layout -- no im2col matrix in HBM.
    if constexpr (A_MODE == kAGather) {
      const int et = threadIdx.x - 192;
      const int row = et & 127, part = et >> 7;
      const int w_l = row & (sg.tw_b - 1);
      const int h_l = (row >> sg.tw_sh) & (sg.th_b - 1);
      const int n_l = row >> (sg.tw_sh + sg.th_sh);
      const int2* ktab = reinterpret_cast<const int2*>(smem + g.ktab_off);
      const int tiles_hw = sg.tiles_h * sg.tiles_w;
      const uint32_t sw = (uint32_t)row & 7u;
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / g.n_tiles;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        const int img = tn * sg.tn_b + n_l, ho = th * sg.th_b + h_l, wo = tw * sg.tw_b + w_l;
        const bool valid = img < sg.imgs && ho < sg.Ho && wo < sg.Wo;
        const int hi0 = ho * sg.stride - sg.pad, wi0 = wo * sg.stride - sg.pad;
        const float* px = g.gx + (((int64_t)(valid ? img % g.gx_imgs : 0) * sg.H + hi0) * sg.W + wi0) * g.gcin;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          const uint32_t dst = smem_u32(smem + (size_t)stage * kStageBytes) + (uint32_t)row * 128u;
          const int2* kt = ktab + kb * 64 + part * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float v[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int2 e = kt[c * 8 + u * 2 + h];
                const int hi = hi0 + (int)((uint32_t)e.y >> 16), wi = wi0 + (e.y & 0xFFFF);
                v[h] = (valid && (unsigned)hi < (unsigned)sg.H && (unsigned)wi < (unsigned)sg.W) ? __ldg(px + e.x) : 0.f;
              }
              const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[0], v[1]);
              w4[u] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)(part * 4 + c)) ^ sw) << 4)),
                         "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
          if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 6 && A_MODE == kADw) {
    // =============================== depth-wise spread producers ===============================
    if constexpr (A_MODE == kADw) {
      const int et = threadIdx.x - 192;   // 0 .. 255
      const int cg = et & 7;              // 8-channel group inside the 64-channel slab
      const int rl = et >> 3;             // row lane: this thread owns rows rl, rl+32, rl+64, rl+96
      const int H = sg.H, W = sg.W;
      const int64_t HW = (int64_t)H * W;
      const int c8 = sg.Cw * 4;           // bytes of spike bits per pixel
      const int C = sg.Cw * 32;
      const uint8_t* bytes = reinterpret_cast<const uint8_t*>(sg.bits);
      float wreg[72], breg[8];
      int cur_slab = -1;
      const int nkb = g.kb_total;         // = C / 64
      const uint32_t my_tiles = blockIdx.x < (uint32_t)total_tiles
                                    ? (uint32_t)(total_tiles - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0u;
      const uint32_t n_items = my_tiles * (uint32_t)nkb;

      // packed spike bytes of the 4 rows for the 9 taps of one (tile, slab): pk[tap] = row0 | row1<<8 | ...
      auto load_bytes = [&](uint32_t item, uint32_t (&pk)[9]) {
        const uint32_t ti = item / (uint32_t)nkb;
        const int slab = (int)(item - ti * nkb);
        const int tile = (int)blockIdx.x + (int)ti * (int)gridDim.x;
        const int m_tile = tile / g.n_tiles;
#pragma unroll
        for (int t = 0; t < 9; ++t) pk[t] = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t m = (int64_t)m_tile * 128 + rl + 32 * i;
          if (m < g.M) {
            const int64_t n = m / HW;
            const int rem = (int)(m - n * HW);
            const int h = rem / W, w = rem - h * W;
            const uint8_t* base = bytes + m * c8 + slab * 8 + cg;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int hh = h + ky - 1;
              if (hh < 0 || hh >= H) continue;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int ww = w + kx - 1;
                if (ww < 0 || ww >= W) continue;
                pk[ky * 3 + kx] |= (uint32_t)__ldg(base + ((int64_t)(ky - 1) * W + (kx - 1)) * c8) << (8 * i);
              }
            }
          }
        }
      };

      uint32_t stage = 0, phase = 0;
      uint32_t cur[9];
      if (n_items > 0) load_bytes(0, cur);
      for (uint32_t item = 0; item < n_items; ++item) {
        uint32_t nxt[9];
        if (item + 1 < n_items) load_bytes(item + 1, nxt);
        const int slab = (int)(item % (uint32_t)nkb);
        if (slab != cur_slab) {
          cur_slab = slab;
          const int c0 = slab * 64 + cg * 8;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(g.dw_w + t * C + c0));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(g.dw_w + t * C + c0 + 4));
            wreg[t * 8 + 0] = w0.x; wreg[t * 8 + 1] = w0.y; wreg[t * 8 + 2] = w0.z; wreg[t * 8 + 3] = w0.w;
            wreg[t * 8 + 4] = w1.x; wreg[t * 8 + 5] = w1.y; wreg[t * 8 + 6] = w1.z; wreg[t * 8 + 7] = w1.w;
          }
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.dw_b + c0));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.dw_b + c0 + 4));
          breg[0] = b0.x; breg[1] = b0.y; breg[2] = b0.z; breg[3] = b0.w;
          breg[4] = b1.x; breg[5] = b1.y; breg[6] = b1.z; breg[7] = b1.w;
        }
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        uint8_t* tile_hi = smem + (size_t)stage * kStageBytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float acc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = breg[k];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t m = (cur[t] >> (8 * i)) & 0xFFu;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (m & (1u << k)) acc[k] += wreg[t * 8 + k];
          }
          const int r = rl + 32 * i;
          const uint32_t off = (uint32_t)r * 128u + (((uint32_t)cg ^ ((uint32_t)r & 7u)) << 4);
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(acc[2 * q]), h1 = __float2bfloat16_rn(acc[2 * q + 1]);
            hi[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            if (A_SPLIT == 2) {
              const __nv_bfloat16 l0 = __float2bfloat16_rn(acc[2 * q] - __bfloat162float(h0));
              const __nv_bfloat16 l1 = __float2bfloat16_rn(acc[2 * q + 1] - __bfloat162float(h1));
              lo[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
          }
          *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (A_SPLIT == 2) *reinterpret_cast<uint4*>(tile_hi + kATileBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
#pragma unroll
        for (int t = 0; t < 9; ++t) cur[t] = nxt[t];
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 6 && kTS) {
    // =============================== spike expanders, A operand in tensor memory ===============================
    // A thread owns ONE operand row (= output pixel) of the tile: TMEM lane = row, and a warp may only touch
    // the lane quarter (warp id % 4).  Two warps share a quarter and alternate K blocks.  Per K block a thread
    // reads the 64 spike bits of (pixel + tap, 64-channel slab) from the L2-resident bit tensor (kPf blocks
    // ahead), turns them into 32 packed bf16 pairs with one shift + one mask each (a spike is the single bit
    // 0x4000 = 2.0; the host packs the weights pre-multiplied by 0.5 and in the matching channel order:
    // word j of a 32-channel half holds channels j and j+16) and stores them with one tcgen05.st -- the
    // operand never touches shared memory, so the MMA reads only B from smem.
    if constexpr (kTS) {
      constexpr int kPf = 4;
      const int q = warp & 3;
      const uint32_t par = (uint32_t)(warp - 6) >> 2;
      const int r = q * 32 + lane;
      const int w_l = r & (sg.tw_b - 1);
      const int h_l = (r >> sg.tw_sh) & (sg.th_b - 1);
      const int n_l = r >> (sg.tw_sh + sg.th_sh);
      const uint32_t S = (uint32_t)g.stages;
      const int kbt = g.kb_total;
      const int tiles_hw = sg.tiles_h * sg.tiles_w;
      const uint32_t my_tiles = blockIdx.x < (uint32_t)total_tiles
                                    ? (uint32_t)(total_tiles - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0u;
      const uint32_t c_end = my_tiles * (uint32_t)kbt;
      const uint32_t a_row = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * BN;
      const int64_t wstride = (int64_t)sg.Cw;

      // load cursor: K block `lc` of this CTA's sequence = (tile, ky, kx, slab)
      uint32_t lc = par;
      int l_tile = (int)blockIdx.x, l_kb = (int)par, l_ky = 0, l_kx = 0, l_slab = 0;
      int l_hi0 = 0, l_wi0 = 0;
      const uint32_t* l_img = nullptr;
      bool l_ok = false;
      auto set_tile = [&]() {
        const int m_tile = l_tile / g.n_tiles;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        const int img = tn * sg.tn_b + n_l;
        l_ok = img < sg.imgs;
        l_hi0 = (th * sg.th_b + h_l) * sg.stride - sg.pad;
        l_wi0 = (tw * sg.tw_b + w_l) * sg.stride - sg.pad;
        l_img = sg.bits + (int64_t)(l_ok ? img : 0) * sg.H * sg.W * wstride;
      };
      auto set_kb = [&]() {   // decode l_kb (only after a tile change or at start)
        const int tap = l_kb / sg.nslab;
        l_slab = l_kb - tap * sg.nslab;
        l_ky = tap / sg.kw;
        l_kx = tap - l_ky * sg.kw;
      };
      auto advance2 = [&]() {  // move the cursor two K blocks forward
        lc += 2;
        l_kb += 2;
        if (l_kb >= kbt) {
          do { l_kb -= kbt; l_tile += (int)gridDim.x; } while (l_kb >= kbt);
          set_tile();
          set_kb();
        } else {
          l_slab += 2;
          while (l_slab >= sg.nslab) {
            l_slab -= sg.nslab;
            if (++l_kx == sg.kw) { l_kx = 0; ++l_ky; }
          }
        }
      };
      auto load_cur = [&]() -> uint2 {
        uint2 v = make_uint2(0u, 0u);
        if (lc < c_end) {
          const int hi = l_hi0 + l_ky, wi = l_wi0 + l_kx;
          if (l_ok && hi >= 0 && hi < sg.H && wi >= 0 && wi < sg.W)
            v = __ldg(reinterpret_cast<const uint2*>(l_img + ((int64_t)hi * sg.W + wi) * wstride + l_slab * 2));
        }
        return v;
      };

      uint2 pf[kPf];
      while (l_kb >= kbt) { l_kb -= kbt; l_tile += (int)gridDim.x; }
      if (lc < c_end) { set_tile(); set_kb(); }
#pragma unroll
      for (int j = 0; j < kPf; ++j) {
        pf[j] = load_cur();
        if (lc < c_end) advance2();
      }
      uint32_t c = par;
      while (c < c_end) {
#pragma unroll
        for (int j = 0; j < kPf; ++j) {
          if (c < c_end) {
            const uint2 cur = pf[j];
            pf[j] = load_cur();
            if (lc < c_end) advance2();
            const uint32_t stage = c % S;
            uint32_t v[32];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              v[jj] = (jj < 15 ? (cur.x << (14 - jj < 0 ? 0 : 14 - jj)) : (cur.x >> 1)) & 0x40004000u;
              v[16 + jj] = (jj < 15 ? (cur.y << (14 - jj < 0 ? 0 : 14 - jj)) : (cur.y >> 1)) & 0x40004000u;
            }
            mbar_wait(&ctl->empty[stage], ((c / S) & 1) ^ 1);
            tc_fence_after_sync();
            tmem_st_32x32(a_row + stage * 32, v);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
            c += 2;
          }
        }
      }
    }
  } else if (warp >= 6) {
    // =============================== spike expanders ===============================
    // Stage group g (= pipeline stage g) owns every K block c with c % stages == g, so `stages` K blocks
    // are expanded concurrently and one block's latency chain (wait, load, STS, proxy fence, arrive) never
    // sits on the critical path.  Each lane materialises whole 128-byte rows: the 64 spike bits of
    // (pixel + tap, 64-channel slab) are read straight from the bit-packed tensor (L2-resident, 8 bytes per
    // row and K block, prefetched one K block ahead; zero padding = predicated-off load) and written as 64
    // {0,1} bf16 values in the 128-byte-swizzled K-major layout.
    if constexpr (A_MODE == kASpikes) {
      const int ew = warp - 6;                 // 0 .. 7
      const uint32_t S = (uint32_t)g.stages;
      const int grp = ew / g.wpg, sub = ew - grp * g.wpg;
      const int rows_per_warp = 128 / g.wpg;   // 128, 64 or 32
      const int nrow = rows_per_warp >> 5;     // rows per lane: 4, 2 or 1
      int n_l[4], h_l[4], w_l[4];
      uint32_t row_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = sub * rows_per_warp + i * 32 + lane;
        w_l[i] = r & (sg.tw_b - 1);
        h_l[i] = (r >> sg.tw_sh) & (sg.th_b - 1);
        n_l[i] = r >> (sg.tw_sh + sg.th_sh);
        row_off[i] = (uint32_t)r * 128u;
      }
      const uint32_t sw = (uint32_t)(lane & 7);  // r & 7 == lane & 7 for every row of this lane
      uint8_t* tile_a = smem + (size_t)grp * kStageBytes;
      const int tiles_hw = sg.tiles_h * sg.tiles_w;
      const uint32_t kbt = (uint32_t)g.kb_total;
      const uint32_t my_tiles = blockIdx.x < (uint32_t)total_tiles
                                    ? (uint32_t)(total_tiles - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0u;
      const uint32_t c_end = my_tiles * kbt;

      // load cursor of this stage group: K block c = (tile, kb) advances by S blocks; the tile is decoded only when it
      // changes (3 divisions per tile instead of 5 per block)
      int l_kb = grp, l_tile = (int)blockIdx.x, l_img0 = 0, l_h0 = 0, l_w0 = 0;
      auto decode_tile = [&]() {
        const int m_tile = l_tile / g.n_tiles;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        l_img0 = tn * sg.tn_b;
        l_h0 = th * sg.th_b * sg.stride - sg.pad;
        l_w0 = tw * sg.tw_b * sg.stride - sg.pad;
      };
      while (l_kb >= (int)kbt) { l_kb -= (int)kbt; l_tile += (int)gridDim.x; }
      decode_tile();
      auto load_rows = [&](uint2 (&wd)[4]) {   // rows of the cursor's block, then advance the cursor by S blocks
        const int tap = l_kb / sg.nslab, slab = l_kb - tap * sg.nslab;
        const int ky = tap / sg.kw, kx = tap - ky * sg.kw;
        const int hi0 = l_h0 + ky, wi0 = l_w0 + kx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          wd[i] = make_uint2(0u, 0u);
          if (i < nrow) {
            const int img = l_img0 + n_l[i], hi = hi0 + h_l[i] * sg.stride, wi = wi0 + w_l[i] * sg.stride;
            if (img < sg.imgs && hi >= 0 && hi < sg.H && wi >= 0 && wi < sg.W)
              wd[i] = __ldg(reinterpret_cast<const uint2*>(
                  sg.bits + (((int64_t)img * sg.H + hi) * sg.W + wi) * sg.Cw + slab * 2));
          }
        }
        l_kb += (int)S;
        if (l_kb >= (int)kbt) {
          do { l_kb -= (int)kbt; l_tile += (int)gridDim.x; } while (l_kb >= (int)kbt);
          decode_tile();
        }
      };

      uint32_t c = (uint32_t)grp;
      uint32_t wphase = 1;   // parity of the wait on empty[grp] for block c (first use of a fresh barrier passes)
      uint2 cur[4];
      if (c < c_end) load_rows(cur);
      while (c < c_end) {
        uint2 nxt[4];
        if (c + S < c_end) load_rows(nxt);
        mbar_wait(&ctl->empty[grp], wphase);
        wphase ^= 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < nrow) {
            uint8_t* row = tile_a + row_off[i];
            if (g.pair_expand) {
              // one shift + one mask per packed pair: word jj of a 32-channel half = channels (jj, jj + 16), spike = 2.0
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t x = j < 4 ? cur[i].x : cur[i].y;
                uint4 o;
                uint32_t w4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int jj = (j & 3) * 4 + u;
                  w4[u] = (jj < 15 ? (x << (14 - jj < 0 ? 0 : 14 - jj)) : (x >> 1)) & 0x40004000u;
                }
                o.x = w4[0]; o.y = w4[1]; o.z = w4[2]; o.w = w4[3];
                *reinterpret_cast<uint4*>(row + (((uint32_t)j ^ sw) << 4)) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t byte = ((j < 4 ? cur[i].x : cur[i].y) >> (8 * (j & 3))) & 0xFFu;
                uint4 o;
                o.x = bits2_to_bf16x2(byte);
                o.y = bits2_to_bf16x2(byte >> 2);
                o.z = bits2_to_bf16x2(byte >> 4);
                o.w = bits2_to_bf16x2(byte >> 6);
                *reinterpret_cast<uint4*>(row + (((uint32_t)j ^ sw) << 4)) = o;
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->full_a[grp]);
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
        c += S;
      }
    }
  } else {
    // =============================== epilogue (warps 0-3) ===============================
    const int row = warp * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      // output row of this thread
      int64_t pix;
      bool valid;
      if (A_MODE == kASpikes || A_MODE == kATma4 || A_MODE == kAGather || kTS) {
        const int tiles_hw = sg.tiles_h * sg.tiles_w;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        const int w_l = row & (sg.tw_b - 1);
        const int h_l = (row >> sg.tw_sh) & (sg.th_b - 1);
        const int n_l = row >> (sg.tw_sh + sg.th_sh);
        const int img = tn * sg.tn_b + n_l, ho = th * sg.th_b + h_l, wo = tw * sg.tw_b + w_l;
        valid = img < sg.imgs && ho < sg.Ho && wo < sg.Wo;
        pix = ((int64_t)img * sg.Ho + ho) * sg.Wo + wo;
      } else {
        pix = (int64_t)m_tile * 128 + row;
        valid = pix < g.M;
      }
      mbar_wait(&ctl->tmem_full[buf], bphase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * BN;
      // Coalesced epilogue: TMEM gives each lane one ROW (32 columns); the affine is applied there, the tile
      // slice is transposed through a padded per-warp staging buffer (pitch 36 floats: conflict-free 128-bit
      // accesses both ways), and residual read / store are issued with 8 lanes per row (full 128-byte lines
      // for fp32, 64-byte for fp16) instead of one lane per row (32 partially written sectors per request).
      float* stg = reinterpret_cast<float*>(smem + g.epi_off) + warp * kEpiWarpFloats;
      long long* rinfo = reinterpret_cast<long long*>(stg + 32 * 36);   // [32][2]: out offset (-1 = invalid), residual offset
      const EpiConv& e = ep;
      rinfo[lane * 2 + 0] = valid ? (long long)pix * e.ldc : -1LL;
      rinfo[lane * 2 + 1] = (valid && e.residual != nullptr) ? (long long)(pix % e.res_rows) * e.ldc : 0LL;
      __syncwarp();
      const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int n0 = n_tile * BN + c0;
        // the residual slice of this chunk is requested first: its loads are in flight while the accumulator is read and staged
        float4 r4[8];
        if (e.residual != nullptr) {
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int rr = itr * 4 + sub;
            r4[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rinfo[rr * 2] >= 0) r4[itr] = *reinterpret_cast<const float4*>(e.residual + rinfo[rr * 2 + 1] + n0 + c4);
          }
        }
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                 __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          if (e.scale != nullptr) {
            const float4 s = *reinterpret_cast<const float4*>(e.scale + n0 + 4 * q);
            const float4 b = *reinterpret_cast<const float4*>(e.shift + n0 + 4 * q);
            o.x = fmaf(o.x, s.x, b.x); o.y = fmaf(o.y, s.y, b.y);
            o.z = fmaf(o.z, s.z, b.z); o.w = fmaf(o.w, s.w, b.w);
          }
          *reinterpret_cast<float4*>(stg + lane * 36 + 4 * q) = o;
        }
        __syncwarp();
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + sub;                       // row of this warp's 32-row slice
          const long long ooff = rinfo[rr * 2];
          float4 o = *reinterpret_cast<const float4*>(stg + rr * 36 + c4);
          if (ooff >= 0) {
            if (e.residual != nullptr) { o.x += r4[itr].x; o.y += r4[itr].y; o.z += r4[itr].z; o.w += r4[itr].w; }
            if (e.out_half) {
              __half2 h01 = __floats2half2_rn(o.x, o.y), h23 = __floats2half2_rn(o.z, o.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h01);
              pk.y = *reinterpret_cast<uint32_t*>(&h23);
              *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(e.out) + ooff + n0 + c4) = pk;
            } else {
              *reinterpret_cast<float4*>(e.out + ooff + n0 + c4) = o;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      mbar_arrive(&ctl->tmem_empty[buf]);
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Production forward spike convolution: tensor-memory A operand + TMA-store epilogue.
//
// Same mainloop as k_umma_gemm<kASpikesT> (B by TMA, A expanded from spike bits straight into TMEM, two
// accumulator buffers), but the epilogue is built for throughput: TWO epilogue groups of 4 warps (group g owns
// accumulator buffer g, i.e. every other tile), each draining 32-column slices: tcgen05.ld (row per lane) ->
// folded tdBN affine from shared memory -> (+ residual slice that TMA loaded into the staging buffer) ->
// 128-byte-swizzled staging tile -> ONE cp.async.bulk.tensor store per slice.  No per-row address arithmetic,
// no partially written sectors; image / row / column edges are clipped by the tensor map.
// Warps: 0-3 epilogue group 0, 4-7 epilogue group 1, 8 TMA producer (weights), 9 MMA issuer + TMEM allocator,
// 10-17 spike expanders (lane quarter = warp % 4, two warps per quarter alternate K blocks).
// ---------------------------------------------------------------------------------------------
constexpr int kTsThreads = 576;     // V = 1: 8 expander warps, one K block per barrier round
constexpr int kTsThreads2 = 832;    // V = 2: 16 expander warps, two K blocks per barrier round
constexpr int kTsSlice = 128 * 128;   // staging slice: 128 rows x 32 fp32
constexpr int kTsMaxBuf = 3;

constexpr int kTsMaxStages = 12;
struct TsCtl {
  uint64_t full_a[kTsMaxStages];
  uint64_t full_b[kTsMaxStages];
  uint64_t empty[kTsMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[2][kTsMaxBuf];
  uint32_t tmem_base;
  uint32_t pad;
};

struct TsArgs {
  int m_tiles, n_tiles, kb_total, stages;
  int resident;    // stages == kb_total: every weight K block stays in shared memory, loaded once per CTA (no
                   // per-tile weight re-reads: 148 SMs hammering the same few hundred L2 lines was the limiter)
  int nbuf;        // staging buffers per epilogue group: 3 = residual prefetched one slice ahead, else 2
  int has_res;
  int res_imgs;    // images in the residual tensor (T-broadcast sources repeat)
  int cout;
  const float* scale;
  const float* shift;
  uint32_t stg_off, aff_off, ctl_off;   // byte offsets inside the 1024-aligned dynamic shared memory
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// V = 2 ("rounds"): the single MMA-issuer warp needs ~430 cycles of its own instructions per barrier round and an
// expander warp ~1000 cycles per K block (single-warp issue latency), against a 128 / 256-cycle MMA per K block at
// N = 64 / 128.  So a barrier round covers TWO K blocks (one wait / commit per 8 MMAs, A ring stages of 64 TMEM
// columns, weight stages of two tiles) and four expander warps share a lane quarter (warp e expands the K blocks
// kb = e mod 4 of every tile; a tile with an odd block count gets a phantom block that only synchronises).
template <int BN, int B_SPLIT, int V>
__global__ void __launch_bounds__(V == 2 ? kTsThreads2 : kTsThreads, 1)
k_spike_conv_ts(const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ CUtensorMap tm_res, const TsArgs g, const SpikeGeom sg) {
  static_assert(BN == 64 || BN == 128, "accumulators + A ring must fit 512 TMEM columns");
  constexpr int kNT = V == 2 ? kTsThreads2 : kTsThreads;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = B_SPLIT * kBTileBytes;
  constexpr int kSlices = BN / 32;
  TsCtl* ctl = reinterpret_cast<TsCtl*>(smem + g.ctl_off);
  float* s_scale = reinterpret_cast<float*>(smem + g.aff_off);
  float* s_shift = s_scale + g.cout;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;
  const int tiles_hw = sg.tiles_h * sg.tiles_w;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
    if (g.has_res) tma_prefetch_desc(&tm_res);
    for (int s = 0; s < kTsMaxStages; ++s) {
      mbar_init(&ctl->full_a[s], V == 2 ? 8 : 4);
      mbar_init(&ctl->full_b[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->tmem_full[b], 1);
      mbar_init(&ctl->tmem_empty[b], 128);
      for (int k = 0; k < kTsMaxBuf; ++k) mbar_init(&ctl->res_full[b][k], 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(&ctl->tmem_base);
  if (g.scale != nullptr)
    for (int i = threadIdx.x; i < g.cout; i += kNT) {
      s_scale[i] = g.scale[i];
      s_shift[i] = g.shift[i];
    }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 8 && V == 2) {
    // =============================== TMA producer (rounds of two K blocks) ===============================
    const int kbt = g.kb_total, R = (kbt + 1) >> 1;
    if (g.resident) {   // every weight K block has its own slot and barrier, loaded once
      if (lane == 0 && (int)blockIdx.x < total_tiles)
        for (int kb = 0; kb < kbt; ++kb) {
          uint8_t* st = smem + (size_t)kb * kStageBytes;
          mbar_arrive_expect_tx(&ctl->full_b[kb], (uint32_t)kStageBytes);
#pragma unroll
          for (int bs = 0; bs < B_SPLIT; ++bs)
            tma_load_2d(st + bs * kBTileBytes, &tm_b, &ctl->full_b[kb], kb * 64, bs * (g.n_tiles * BN));
        }
    } else {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % g.n_tiles;
        for (int r = 0; r < R; ++r) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          if (lane == 0) {
            const int nb = (2 * r + 1 < kbt) ? 2 : 1;
            uint8_t* st = smem + (size_t)stage * (2 * kStageBytes);
            mbar_arrive_expect_tx(&ctl->full_b[stage], (uint32_t)(nb * kStageBytes));
            for (int h = 0; h < nb; ++h)
#pragma unroll
              for (int bs = 0; bs < B_SPLIT; ++bs)
                tma_load_2d(st + h * kStageBytes + bs * kBTileBytes, &tm_b, &ctl->full_b[stage], (2 * r + h) * 64,
                            bs * (g.n_tiles * BN) + n_tile * BN);
          }
          __syncwarp();
          if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    // =============================== TMA producer: weight tiles ===============================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % g.n_tiles;
      if (g.resident && tile != (int)blockIdx.x) break;   // resident weights: loaded by the first tile only
      for (int kb = 0; kb < g.kb_total; ++kb) {
        if (!g.resident) mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* st = smem + (size_t)stage * kStageBytes;
          mbar_arrive_expect_tx(&ctl->full_b[stage], (uint32_t)kStageBytes);
#pragma unroll
          for (int bs = 0; bs < B_SPLIT; ++bs)
            tma_load_2d(st + bs * kBTileBytes, &tm_b, &ctl->full_b[stage], kb * 64, bs * (g.n_tiles * BN) + n_tile * BN);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9 && V == 2) {
    // =============================== MMA issuer (rounds of two K blocks) ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    const int kbt = g.kb_total, R = (kbt + 1) >> 1;
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&ctl->tmem_empty[buf], bphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int r = 0; r < R; ++r) {
        const int nb = (2 * r + 1 < kbt) ? 2 : 1;
        if (g.resident) {
          if (it == 0) {
            mbar_wait(&ctl->full_b[2 * r], 0);
            if (nb == 2) mbar_wait(&ctl->full_b[2 * r + 1], 0);
          }
        } else {
          mbar_wait(&ctl->full_b[stage], phase);
        }
        mbar_wait(&ctl->full_a[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t b_addr = smem_u32(smem) + (g.resident ? (uint32_t)(2 * r) * (uint32_t)kStageBytes
                                                               : stage * (uint32_t)(2 * kStageBytes));
          const uint32_t a_tmem = tmem_base + 2 * BN + stage * 64;
          uint32_t acc = r > 0 ? 1u : 0u;
          for (int h = 0; h < nb; ++h) {
#pragma unroll
            for (int bs = 0; bs < B_SPLIT; ++bs) {
              const uint64_t db = umma_desc_sw128(b_addr + h * kStageBytes + bs * kBTileBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16_ts(d_tmem, a_tmem + h * 32 + k * 8, db + (uint64_t)(k * 2), idesc, acc);
                acc = 1u;
              }
            }
          }
          umma_commit(&ctl->empty[stage]);
          if (r == R - 1) umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&ctl->tmem_empty[buf], bphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        if (!g.resident || it == 0) mbar_wait(&ctl->full_b[stage], phase);
        mbar_wait(&ctl->full_a[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t b_addr = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint32_t a_tmem = tmem_base + 2 * BN + stage * 32;
          uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int bs = 0; bs < B_SPLIT; ++bs) {
            const uint64_t db = umma_desc_sw128(b_addr + bs * kBTileBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_ts(d_tmem, a_tmem + k * 8, db + (uint64_t)(k * 2), idesc, acc);
              acc = 1u;
            }
          }
          umma_commit(&ctl->empty[stage]);
          if (kb == g.kb_total - 1) umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 10 && V == 2) {
    // =============================== spike expanders, four warps per lane quarter ===============================
    constexpr int kPf = 4;
    const int q = warp & 3;
    const int e = (warp - 10) >> 2;                 // this warp expands the (padded) K blocks kb = e mod 4 of every tile
    const int r = q * 32 + lane;
    const int w_l = r & (sg.tw_b - 1);
    const int h_l = (r >> sg.tw_sh) & (sg.th_b - 1);
    const int n_l = r >> (sg.tw_sh + sg.th_sh);
    const uint32_t S = (uint32_t)g.stages;
    const int kbt = g.kb_total, R = (kbt + 1) >> 1, KP = 2 * R;
    const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t a_row = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * BN;
    const int tapw = sg.Cw, roww = sg.W * sg.Cw;
    static const bool kFastTap = true;
    if (kFastTap && sg.nslab == 1 && KP <= 12 && g.n_tiles == 1) {
      // ---- one 64-channel slab (Cin = 64): K block = filter tap.  Tile-level code: this warp's taps kb = e, e+4, e+8 are
      // fixed for the whole kernel, a tile costs one pixel decode + up to three guarded 8-byte loads (requested one
      // tile ahead) -- no per-block cursor arithmetic (the kernel is bound by exactly that skeleton).
      if (e < KP && my_tiles > 0) {
        int tky[3], tkx[3], toff[3];
        bool town[3], treal[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int kb = e + 4 * i;
          town[i] = kb < KP;
          treal[i] = kb < kbt;
          tky[i] = kb / sg.kw;
          tkx[i] = kb - tky[i] * sg.kw;
          toff[i] = (tky[i] * sg.W + tkx[i]) * tapw;
        }
        int d_w, d_h, d_t, s_w, s_h, s_t;
        {
          int v = (int)blockIdx.x;
          d_w = v % sg.tiles_w; v /= sg.tiles_w;
          d_h = v % sg.tiles_h; d_t = v / sg.tiles_h;
          v = (int)gridDim.x;
          s_w = v % sg.tiles_w; v /= sg.tiles_w;
          s_h = v % sg.tiles_h; s_t = v / sg.tiles_h;
        }
        auto load_tile = [&](uint2 (&wd)[3]) {   // this thread's pixel of the tile at the cursor, then advance the cursor
          const int img = d_t * sg.tn_b + n_l;
          const int hi0 = (d_h * sg.th_b + h_l) * sg.stride - sg.pad;
          const int wi0 = (d_w * sg.tw_b + w_l) * sg.stride - sg.pad;
          const bool ok = img < sg.imgs;
          const uint32_t* base = sg.bits + ((int64_t)(ok ? img : 0) * sg.H + hi0) * roww + (int64_t)wi0 * tapw;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            wd[i] = make_uint2(0u, 0u);
            if (treal[i] && ok && (unsigned)(hi0 + tky[i]) < (unsigned)sg.H && (unsigned)(wi0 + tkx[i]) < (unsigned)sg.W)
              wd[i] = __ldg(reinterpret_cast<const uint2*>(base + toff[i]));
          }
          d_w += s_w;
          int cy = d_w >= sg.tiles_w ? 1 : 0;
          d_w -= cy ? sg.tiles_w : 0;
          d_h += s_h + cy;
          cy = d_h >= sg.tiles_h ? 1 : 0;
          d_h -= cy ? sg.tiles_h : 0;
          d_t += s_t + cy;
        };
        uint2 nxt[3];
        load_tile(nxt);
        uint32_t stage = (uint32_t)(e >> 1) % S, sphase = (((uint32_t)(e >> 1) / S) & 1u) ^ 1u;
        for (int ti = 0; ti < my_tiles; ++ti) {
          uint2 cur[3] = {nxt[0], nxt[1], nxt[2]};
          if (ti + 1 < my_tiles) load_tile(nxt);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            if (town[i]) {
              const int kb = e + 4 * i;
              mbar_wait(&ctl->empty[stage], sphase);
              if (treal[i]) {
                tc_fence_after_sync();
                const uint32_t dst = a_row + stage * 64 + (uint32_t)(kb & 1) * 32u;
#pragma unroll
                for (int hw = 0; hw < 2; ++hw) {
                  const uint32_t x = hw == 0 ? cur[i].x : cur[i].y;
                  uint32_t v[16];
#pragma unroll
                  for (int jj = 0; jj < 16; ++jj)
                    v[jj] = (jj < 15 ? (x << (14 - jj < 0 ? 0 : 14 - jj)) : (x >> 1)) & 0x40004000u;
                  tmem_st_32x16(dst + hw * 16, v);
                }
                tmem_st_wait();
                tc_fence_before_sync();
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
              const bool last = (i == 2) || !(e + 4 * (i + 1) < KP);     // last own block of this tile
              stage += last ? (uint32_t)(R - (kb >> 1) + (e >> 1)) : 2u;
              while (stage >= S) { stage -= S; sphase ^= 1; }
            }
          }
        }
      }
    } else if (e < KP && my_tiles > 0) {
      // ---- load cursor (tile digits in the mixed radix (n_tile, tile_w, tile_h, tile_n); no divisions per block) ----
      int l_ti = 0, l_kb = e, l_ky = 0, l_kx = 0, l_slab = e;
      int d_n, d_w, d_h, d_t, s_n, s_w, s_h, s_t;
      {
        int v = (int)blockIdx.x;
        d_n = v % g.n_tiles; v /= g.n_tiles;
        d_w = v % sg.tiles_w; v /= sg.tiles_w;
        d_h = v % sg.tiles_h; d_t = v / sg.tiles_h;
        v = (int)gridDim.x;
        s_n = v % g.n_tiles; v /= g.n_tiles;
        s_w = v % sg.tiles_w; v /= sg.tiles_w;
        s_h = v % sg.tiles_h; s_t = v / sg.tiles_h;
      }
      int l_hi0 = 0, l_wi0 = 0;
      const uint32_t* l_row0 = nullptr;
      bool l_ok = false;
      auto set_tile = [&]() {
        const int img = d_t * sg.tn_b + n_l;
        l_ok = img < sg.imgs;
        l_hi0 = (d_h * sg.th_b + h_l) * sg.stride - sg.pad;
        l_wi0 = (d_w * sg.tw_b + w_l) * sg.stride - sg.pad;
        l_row0 = sg.bits + ((int64_t)(l_ok ? img : 0) * sg.H + l_hi0) * roww + (int64_t)l_wi0 * tapw;
      };
      auto norm_slab = [&]() {
        while (l_slab >= sg.nslab) {
          l_slab -= sg.nslab;
          if (++l_kx == sg.kw) { l_kx = 0; ++l_ky; }
        }
      };
      auto l_advance = [&]() {
        l_kb += 4;
        if (l_kb >= KP) {
          l_kb = e; ++l_ti;
          d_n += s_n;
          int cy = d_n >= g.n_tiles ? 1 : 0;
          d_n -= cy ? g.n_tiles : 0;
          d_w += s_w + cy;
          cy = d_w >= sg.tiles_w ? 1 : 0;
          d_w -= cy ? sg.tiles_w : 0;
          d_h += s_h + cy;
          cy = d_h >= sg.tiles_h ? 1 : 0;
          d_h -= cy ? sg.tiles_h : 0;
          d_t += s_t + cy;
          set_tile();
          l_slab = e; l_ky = 0; l_kx = 0;
        } else {
          l_slab += 4;
        }
        norm_slab();
      };
      auto load_cur = [&]() -> uint2 {
        uint2 v = make_uint2(0u, 0u);
        if (l_ti < my_tiles && l_kb < kbt) {
          const int hi = l_hi0 + l_ky, wi = l_wi0 + l_kx;
          if (l_ok && (unsigned)hi < (unsigned)sg.H && (unsigned)wi < (unsigned)sg.W)
            v = __ldg(reinterpret_cast<const uint2*>(l_row0 + (l_ky * roww + l_kx * tapw + l_slab * 2)));
        }
        return v;
      };
      set_tile();
      norm_slab();
      uint2 pf[kPf];
#pragma unroll
      for (int j = 0; j < kPf; ++j) {
        pf[j] = load_cur();
        l_advance();
      }
      // ---- consume cursor: (tile, kb) -> round -> A stage / wait parity, advanced incrementally ----
      int c_ti = 0, c_kb = e;
      uint32_t stage = (uint32_t)(e >> 1) % S, sphase = (((uint32_t)(e >> 1) / S) & 1u) ^ 1u;
      while (c_ti < my_tiles) {
#pragma unroll
        for (int j = 0; j < kPf; ++j) {
          if (c_ti < my_tiles) {
            const uint2 cur = pf[j];
            pf[j] = load_cur();
            l_advance();
            const bool phantom = c_kb >= kbt;
            mbar_wait(&ctl->empty[stage], sphase);
            if (!phantom) {
              tc_fence_after_sync();
              const uint32_t dst = a_row + stage * 64 + (uint32_t)(c_kb & 1) * 32u;
#pragma unroll
              for (int hw = 0; hw < 2; ++hw) {   // 16 columns at a time: 832 threads leave 72 registers per thread
                const uint32_t x = hw == 0 ? cur.x : cur.y;
                uint32_t v[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                  v[jj] = (jj < 15 ? (x << (14 - jj < 0 ? 0 : 14 - jj)) : (x >> 1)) & 0x40004000u;
                tmem_st_32x16(dst + hw * 16, v);
              }
              tmem_st_wait();
              tc_fence_before_sync();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
            const int r_old = c_kb >> 1;
            c_kb += 4;
            uint32_t dr = 2;
            if (c_kb >= KP) { c_kb = e; ++c_ti; dr = (uint32_t)(R - r_old + (e >> 1)); }
            stage += dr;
            while (stage >= S) { stage -= S; sphase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 10) {
    // =============================== spike expanders (A operand -> tensor memory) ===============================
    // See k_umma_gemm<kASpikesT>: one operand row (= output pixel) per thread, 64 spike bits per K block read
    // kPf blocks ahead, one shift + one mask per packed bf16 pair (spike = 0x4000 = 2.0, weights carry the 0.5),
    // one tcgen05.st per K block.
    constexpr int kPf = 4;
    const int q = warp & 3;
    const uint32_t par = (uint32_t)(warp - 10) >> 2;
    const int r = q * 32 + lane;
    const int w_l = r & (sg.tw_b - 1);
    const int h_l = (r >> sg.tw_sh) & (sg.th_b - 1);
    const int n_l = r >> (sg.tw_sh + sg.th_sh);
    const uint32_t S = (uint32_t)g.stages;
    const int kbt = g.kb_total;
    const uint32_t my_tiles = blockIdx.x < (uint32_t)total_tiles
                                  ? (uint32_t)(total_tiles - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0u;
    const uint32_t c_end = my_tiles * (uint32_t)kbt;
    const uint32_t a_row = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * BN;
    const int64_t wstride = (int64_t)sg.Cw;

    // Load cursor: K block `lc` of this CTA's sequence = (tile, ky, kx, slab).  No divisions on the per-block path:
    // the tile index advances by gridDim.x in the mixed radix (n_tile, tile_w, tile_h, tile_n) with carries.
    uint32_t lc = par;
    int l_kb = (int)par, l_ky = 0, l_kx = 0, l_slab = 0;
    int d_n, d_w, d_h, d_t;              // digits of the cursor's tile
    int s_n, s_w, s_h, s_t;              // digits of the tile step (gridDim.x)
    {
      int v = (int)blockIdx.x;
      d_n = v % g.n_tiles; v /= g.n_tiles;
      d_w = v % sg.tiles_w; v /= sg.tiles_w;
      d_h = v % sg.tiles_h; d_t = v / sg.tiles_h;
      v = (int)gridDim.x;
      s_n = v % g.n_tiles; v /= g.n_tiles;
      s_w = v % sg.tiles_w; v /= sg.tiles_w;
      s_h = v % sg.tiles_h; s_t = v / sg.tiles_h;
    }
    int l_hi0 = 0, l_wi0 = 0;
    const uint32_t* l_row0 = nullptr;    // &bits[img][hi0][wi0][0] of this thread's pixel (may point outside: guarded)
    bool l_ok = false;
    const int tapw = (int)wstride;       // words per pixel
    const int roww = sg.W * tapw;        // words per image row
    auto set_tile = [&]() {
      const int img = d_t * sg.tn_b + n_l;
      l_ok = img < sg.imgs;
      l_hi0 = (d_h * sg.th_b + h_l) * sg.stride - sg.pad;
      l_wi0 = (d_w * sg.tw_b + w_l) * sg.stride - sg.pad;
      l_row0 = sg.bits + ((int64_t)(l_ok ? img : 0) * sg.H + l_hi0) * roww + (int64_t)l_wi0 * tapw;
    };
    auto next_tile = [&]() {
      d_n += s_n;
      int cy = d_n >= g.n_tiles ? 1 : 0;
      d_n -= cy ? g.n_tiles : 0;
      d_w += s_w + cy;
      cy = d_w >= sg.tiles_w ? 1 : 0;
      d_w -= cy ? sg.tiles_w : 0;
      d_h += s_h + cy;
      cy = d_h >= sg.tiles_h ? 1 : 0;
      d_h -= cy ? sg.tiles_h : 0;
      d_t += s_t + cy;
    };
    auto norm_kb = [&]() {   // (ky, kx, slab) of a small l_kb after a tile change
      l_slab = l_kb; l_ky = 0; l_kx = 0;
      while (l_slab >= sg.nslab) {
        l_slab -= sg.nslab;
        if (++l_kx == sg.kw) { l_kx = 0; ++l_ky; }
      }
    };
    auto advance2 = [&]() {
      lc += 2;
      l_kb += 2;
      if (l_kb >= kbt) {
        do { l_kb -= kbt; next_tile(); } while (l_kb >= kbt);
        set_tile();
        norm_kb();
      } else {
        l_slab += 2;
        while (l_slab >= sg.nslab) {
          l_slab -= sg.nslab;
          if (++l_kx == sg.kw) { l_kx = 0; ++l_ky; }
        }
      }
    };
    auto load_cur = [&]() -> uint2 {
      uint2 v = make_uint2(0u, 0u);
      if (lc < c_end) {
        const int hi = l_hi0 + l_ky, wi = l_wi0 + l_kx;
        if (l_ok && (unsigned)hi < (unsigned)sg.H && (unsigned)wi < (unsigned)sg.W)
          v = __ldg(reinterpret_cast<const uint2*>(l_row0 + (l_ky * roww + l_kx * tapw + l_slab * 2)));
      }
      return v;
    };

    uint2 pf[kPf];
    while (l_kb >= kbt) { l_kb -= kbt; next_tile(); }
    set_tile();
    norm_kb();
#pragma unroll
    for (int j = 0; j < kPf; ++j) {
      pf[j] = load_cur();
      if (lc < c_end) advance2();
    }
    uint32_t c = par;
    uint32_t stage = par % S, sphase = ((par / S) & 1) ^ 1;   // stage / wait parity of block c, advanced incrementally
    while (c < c_end) {
#pragma unroll
      for (int j = 0; j < kPf; ++j) {
        if (c < c_end) {
          const uint2 cur = pf[j];
          pf[j] = load_cur();
          if (lc < c_end) advance2();
          uint32_t v[32];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            v[jj] = (jj < 15 ? (cur.x << (14 - jj < 0 ? 0 : 14 - jj)) : (cur.x >> 1)) & 0x40004000u;
            v[16 + jj] = (jj < 15 ? (cur.y << (14 - jj < 0 ? 0 : 14 - jj)) : (cur.y >> 1)) & 0x40004000u;
          }
          mbar_wait(&ctl->empty[stage], sphase);
          tc_fence_after_sync();
          tmem_st_32x32(a_row + stage * 32, v);
          tmem_st_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->full_a[stage]);
          c += 2;
          stage += 2;
          if (stage >= S) { stage -= S; sphase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue groups (warps 0-3: buffer 0, warps 4-7: buffer 1) ===============================
    const uint32_t ge = (uint32_t)warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (q == 0 && lane == 0);
    const uint32_t nbuf = (uint32_t)g.nbuf;
    const bool prefetch = g.has_res && nbuf == 3;
    uint8_t* stg = smem + g.stg_off + (size_t)ge * nbuf * kTsSlice;
    const uint32_t my_row = smem_u32(stg) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)row & 7u;
    const uint32_t aff = smem_u32(s_scale);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ge * BN;

    auto coords = [&](int tile, int& n_base, int& w0, int& h0, int& img0) {
      const int m_tile = tile / g.n_tiles;
      n_base = (tile - m_tile * g.n_tiles) * BN;
      const int tn = m_tile / tiles_hw;
      const int rem = m_tile - tn * tiles_hw;
      const int th = rem / sg.tiles_w;
      img0 = tn * sg.tn_b;
      h0 = th * sg.th_b;
      w0 = (rem - th * sg.tiles_w) * sg.tw_b;
    };
    // leader only: make staging buffer (s % nbuf) reusable (its last TMA store has finished reading it) and start
    // the residual load of slice s into it
    auto prepare = [&](uint32_t s, int c0, int w0, int h0, int img0) {
      bulk_wait_group_read<1>();
      if (g.has_res) {
        const uint32_t b = s % nbuf;
        mbar_arrive_expect_tx(&ctl->res_full[ge][b], (uint32_t)kTsSlice);
        tma_load_4d(stg + (size_t)b * kTsSlice, &tm_res, &ctl->res_full[ge][b], c0, w0, h0, img0 % g.res_imgs);
      }
    };

    uint32_t sl = 0;   // slices this group has processed
    const int first = (int)blockIdx.x + (int)ge * (int)gridDim.x;
    const int tstep = 2 * (int)gridDim.x;
    if (prefetch && leader && first < total_tiles) {
      int nb, w0, h0, i0;
      coords(first, nb, w0, h0, i0);
      prepare(0, nb, w0, h0, i0);
    }
    uint32_t git = 0;   // tiles this group has processed: accumulator phase
    for (int tile = first; tile < total_tiles; tile += tstep, ++git) {
      int n_base, w0, h0, img0;
      coords(tile, n_base, w0, h0, img0);
#pragma unroll 1
      for (int j = 0; j < kSlices; ++j, ++sl) {
        const uint32_t b = sl % nbuf;
        const uint32_t buf_addr = my_row + b * (uint32_t)kTsSlice;
        if (leader) {
          if (!prefetch) {
            prepare(sl, n_base + j * 32, w0, h0, img0);
          } else if (j + 1 < kSlices) {
            prepare(sl + 1, n_base + (j + 1) * 32, w0, h0, img0);
          } else if (tile + tstep < total_tiles) {
            int nb2, w2, h2, i2;
            coords(tile + tstep, nb2, w2, h2, i2);
            prepare(sl + 1, nb2, w2, h2, i2);
          }
        }
        if (j == 0) {
          mbar_wait(&ctl->tmem_full[ge], git & 1);
          tc_fence_after_sync();
        }
        named_bar_sync(1 + (int)ge, 128);   // the staging buffer of this slice is free
        uint32_t v[32];
        tmem_ld_32x32(t_row + j * 32, v);
        tmem_ld_wait();
        if (j == kSlices - 1) {   // accumulator buffer drained: the MMA warp may start the group's next tile
          tc_fence_before_sync();
          mbar_arrive(&ctl->tmem_empty[ge]);
        }
        float4 o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          o[k] = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                             __uint_as_float(v[4 * k + 3]));
        if (g.scale != nullptr) {
          const uint32_t a0 = aff + (uint32_t)(n_base + j * 32) * 4u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 sc = lds128(a0 + k * 16), sh = lds128(a0 + (uint32_t)g.cout * 4u + k * 16);
            o[k].x = fmaf(o[k].x, sc.x, sh.x); o[k].y = fmaf(o[k].y, sc.y, sh.y);
            o[k].z = fmaf(o[k].z, sc.z, sh.z); o[k].w = fmaf(o[k].w, sc.w, sh.w);
          }
        }
        if (g.has_res) {
          mbar_wait(&ctl->res_full[ge][b], (sl / nbuf) & 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 rv = lds128(buf_addr + (((uint32_t)k ^ sw) << 4));
            o[k].x += rv.x; o[k].y += rv.y; o[k].z += rv.z; o[k].w += rv.w;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sts128(buf_addr + (((uint32_t)k ^ sw) << 4), o[k]);
        fence_proxy_async_smem();
        named_bar_sync(1 + (int)ge, 128);
        if (leader) {
          tma_store_4d(&tm_out, stg + (size_t)b * kTsSlice, n_base + j * 32, w0, h0, img0);
          bulk_commit_group();
        }
      }
    }
    if (leader) bulk_wait_group_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Dense GEMM with fp16 output and the TMA-store epilogue: the point-wise half of the ECS spread in fast mode,
//     S[M][C] (fp16) = A[M][C] (bf16, the depth-wise output) * Wpw[C][C]^T
// K = C is only 1..16 K blocks, so a tile is dominated by its 32 KB of HBM traffic and its epilogue, not by the MMAs:
// same two-epilogue-group / staging / cp.async.bulk.tensor-store structure as k_spike_conv_ts (64-column fp16
// slices = 128-byte rows), A and B by TMA, SS-mode MMA.  Warps: 0-3 epilogue group 0, 4-7 group 1, 8 TMA, 9 MMA.
// ---------------------------------------------------------------------------------------------
constexpr int kDtThreads = 320;

struct DtCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

struct DtArgs {
  int m_tiles, n_tiles, kb_total, stages;
  uint32_t stg_off, ctl_off;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// F32: fp32 output (the G1 = ge * Wpw GEMM of the LIF backward, fast precision): 32-column slices (128-byte rows of fp32).
template <int BN, bool F32 = false>
__global__ void __launch_bounds__(kDtThreads, 1)
k_dense_tma_h(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
              const __grid_constant__ CUtensorMap tm_out, const DtArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr int kTmemCols = 2 * BN;
  constexpr int kSliceCols = F32 ? 32 : 64;
  constexpr int kSlices = BN / kSliceCols;
  DtCtl* ctl = reinterpret_cast<DtCtl*>(smem + g.ctl_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->tmem_full[b], 1);
      mbar_init(&ctl->tmem_empty[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<kTmemCols>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 8) {
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* st = smem + (size_t)stage * kStageBytes;
          mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)kStageBytes);
          tma_load_2d(st, &tm_a, &ctl->full[stage], kb * 64, m_tile * 128);
          tma_load_2d(st + kATileBytes, &tm_b, &ctl->full[stage], kb * 64, n_tile * BN);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&ctl->tmem_empty[buf], bphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + kATileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ctl->empty[stage]);
          if (kb == g.kb_total - 1) umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- epilogue groups: TMEM -> fp16 -> swizzled staging rows (64 halves = 128 bytes) -> one TMA store per slice ----
    const uint32_t ge = (uint32_t)warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (q == 0 && lane == 0);
    uint8_t* stg = smem + g.stg_off + (size_t)ge * 2 * kATileBytes;
    const uint32_t my_row = smem_u32(stg) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)row & 7u;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ge * BN;
    uint32_t sl = 0, git = 0;
    for (int tile = (int)blockIdx.x + (int)ge * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, ++git) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
#pragma unroll 1
      for (int j = 0; j < kSlices; ++j, ++sl) {
        const uint32_t b = sl & 1u;
        if (leader) bulk_wait_group_read<1>();   // the store that last used this staging buffer has read it
        if (j == 0) {
          mbar_wait(&ctl->tmem_full[ge], git & 1);
          tc_fence_after_sync();
        }
        named_bar_sync(1 + (int)ge, 128);
        const uint32_t dst = my_row + b * (uint32_t)kATileBytes;
        if constexpr (F32) {
          uint32_t v0[32];
          tmem_ld_32x32(t_row + j * 32, v0);
          tmem_ld_wait();
          if (j == kSlices - 1) {
            tc_fence_before_sync();
            mbar_arrive(&ctl->tmem_empty[ge]);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)   // chunk c = 4 floats
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)c ^ sw) << 4)), "r"(v0[4 * c]),
                         "r"(v0[4 * c + 1]), "r"(v0[4 * c + 2]), "r"(v0[4 * c + 3]) : "memory");
        } else {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_row + j * 64, v0);
        tmem_ld_32x32(t_row + j * 64 + 32, v1);
        tmem_ld_wait();
        if (j == kSlices - 1) {
          tc_fence_before_sync();
          mbar_arrive(&ctl->tmem_empty[ge]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {   // chunk c = 8 halves
          const uint32_t* src = c < 4 ? &v0[8 * c] : &v1[8 * (c - 4)];
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const __half2 h = __floats2half2_rn(__uint_as_float(src[2 * u]), __uint_as_float(src[2 * u + 1]));
            w[u] = *reinterpret_cast<const uint32_t*>(&h);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)c ^ sw) << 4)), "r"(w[0]),
                       "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + (int)ge, 128);
        if (leader) {
          tma_store_2d(&tm_out, stg + (size_t)b * kATileBytes, n_tile * BN + j * kSliceCols, m_tile * 128);
          bulk_commit_group();
        }
      }
    }
    if (leader) bulk_wait_group_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Point-wise spread GEMM with the ECS step as its EPILOGUE (fast precision):
//     S = A[M][C] (bf16, the depth-wise output) * Wpw[C][C]^T   stays in tensor memory / shared memory, and the epilogue
//     runs   e_t = alpha (S + b) + kappa e_{t-1};  m_{t+1} = decay m_t (1 - s_t) + x_{t+1} + beta tanh(e_t);  s_{t+1} = m_{t+1} > th
// (models/common.py:263-281) on it: `spread` never exists in HBM (-4 B per element-step, one launch less per timestep).
// The arithmetic is k_ecs_step<true>'s, value for value -- S is rounded to fp16 exactly where the two-kernel path stored
// it -- so this kernel and the two-kernel path produce the SAME spikes (the BPTT recompute relies on that).
// Mainloop of k_dense_tma_h (A and B by TMA, SS-mode MMA, two accumulator buffers); warps 0-3 / 4-7: epilogue groups
// (group g owns buffer g = every other tile), 8 TMA, 9 MMA.  An epilogue warp owns 32 rows: tcgen05.ld (a row per lane)
// -> padded staging tile -> 16 lanes per row: x, membrane, trace and spike word of 2 rows x 64 channels are 2 x 256-byte
// runs per request, four requests of each in flight per lane.
// ---------------------------------------------------------------------------------------------
constexpr int kEgRow = 68;   // floats per staged row (64 + 4: conflict-free 128-bit accesses both ways)

struct EgArgs {
  int m_tiles, n_tiles, kb_total, stages;
  uint32_t stg_off, ctl_off;
  int64_t M;
  int C;
  EcsStep s;
};

template <int BN>
__global__ void __launch_bounds__(kDtThreads, 1)
k_ecs_gemm(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const EgArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr int kTmemCols = 2 * BN;
  constexpr int kSlices = BN / 64;
  DtCtl* ctl = reinterpret_cast<DtCtl*>(smem + g.ctl_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctl->tmem_full[b], 1);
      mbar_init(&ctl->tmem_empty[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<kTmemCols>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 8) {
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* st = smem + (size_t)stage * kStageBytes;
          mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)kStageBytes);
          tma_load_2d(st, &tm_a, &ctl->full[stage], kb * 64, m_tile * 128);
          tma_load_2d(st + kATileBytes, &tm_b, &ctl->full[stage], kb * 64, n_tile * BN);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, bphase = (it >> 1) & 1;
      mbar_wait(&ctl->tmem_empty[buf], bphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + kATileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ctl->empty[stage]);
          if (kb == g.kb_total - 1) umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- epilogue groups: TMEM -> staging tile -> ECS step on coalesced rows ----
    const uint32_t ge = (uint32_t)warp >> 2;
    const int q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem + g.stg_off) + warp * (32 * kEgRow);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ge * BN;
    const EcsStep& p = g.s;
    const int C = g.C;
    const int lr = lane >> 4, c4 = (lane & 15) * 4, sub = lane & 7;
    uint32_t git = 0;
    for (int tile = (int)blockIdx.x + (int)ge * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, ++git) {
      const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
      const int64_t pix0 = (int64_t)m_tile * 128 + q * 32;
      mbar_wait(&ctl->tmem_full[ge], git & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int j = 0; j < kSlices; ++j) {
        {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(t_row + j * 64, v0);
          tmem_ld_32x32(t_row + j * 64 + 32, v1);
          tmem_ld_wait();
          if (j == kSlices - 1) {
            tc_fence_before_sync();
            mbar_arrive(&ctl->tmem_empty[ge]);
          }
          float* my = stg + lane * kEgRow;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4*>(my + 4 * c) = make_uint4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
            *reinterpret_cast<uint4*>(my + 32 + 4 * c) = make_uint4(v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
          }
        }
        __syncwarp();
        const int ch = n_tile * BN + j * 64 + c4;
        const float4 pb = *reinterpret_cast<const float4*>(p.pw_b + ch);
        float4 isc = make_float4(1.f, 1.f, 1.f, 1.f), ish = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.in_scale != nullptr) {
          isc = *reinterpret_cast<const float4*>(p.in_scale + ch);
          ish = *reinterpret_cast<const float4*>(p.in_shift + ch);
        }
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += 4) {
          float4 xv[4], mv[4];
          uint2 eh[4];
          uint32_t wd[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = (i0 + u) * 2 + lr;
            const int64_t pix = pix0 + r;
            xv[u] = mv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            eh[u] = make_uint2(0u, 0u);
            wd[u] = 0u;
            if (pix < g.M) {
              const int64_t idx = pix * C + ch;
              xv[u] = ldg_stream(reinterpret_cast<const float4*>(p.x_next + idx));
              mv[u] = *reinterpret_cast<const float4*>(p.mem_in + idx);
              if (!p.first) eh[u] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.ecs) + idx);
              wd[u] = p.bits_t[idx >> 5];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = (i0 + u) * 2 + lr;
            const int64_t pix = pix0 + r;
            const bool ok = pix < g.M;
            const int64_t idx = pix * C + ch;
            const float4 sv4 = *reinterpret_cast<const float4*>(stg + r * kEgRow + c4);
            float xin[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
            if (p.in_scale != nullptr) {
              xin[0] = add_rn(mul_rn(xin[0], isc.x), ish.x); xin[1] = add_rn(mul_rn(xin[1], isc.y), ish.y);
              xin[2] = add_rn(mul_rn(xin[2], isc.z), ish.z); xin[3] = add_rn(mul_rn(xin[3], isc.w), ish.w);
            }
            // the two-kernel path stores S as fp16 between the GEMM and the step: same rounding here
            const float2 s01 = __half22float2(__floats2half2_rn(sv4.x, sv4.y)), s23 = __half22float2(__floats2half2_rn(sv4.z, sv4.w));
            const float sp[4] = {s01.x, s01.y, s23.x, s23.y};
            const float2 e01 = __half22float2(*reinterpret_cast<const __half2*>(&eh[u].x));
            const float2 e23 = __half22float2(*reinterpret_cast<const __half2*>(&eh[u].y));
            const float eo[4] = {e01.x, e01.y, e23.x, e23.y};
            const float mo[4] = {mv[u].x, mv[u].y, mv[u].z, mv[u].w};
            const float bb[4] = {pb.x, pb.y, pb.z, pb.w};
            const uint32_t pw = wd[u] >> (4 * sub);
            float mn[4], en[4];
            uint32_t nib = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float s_acc = add_rn(sp[k], bb[k]);
              en[k] = add_rn(mul_rn(p.alpha, s_acc), mul_rn(p.kappa, eo[k]));
              const float fecs = mul_rn(p.beta, tanh_hw(en[k]));
              const float keep = ((pw >> k) & 1u) ? 0.f : 1.f;
              mn[k] = add_rn(add_rn(mul_rn(mul_rn(mo[k], p.decay), keep), xin[k]), fecs);
              nib |= (mn[k] > p.thresh ? 1u : 0u) << k;
            }
            if (ok) {
              if (p.mem_out != nullptr) *reinterpret_cast<float4*>(p.mem_out + idx) = make_float4(mn[0], mn[1], mn[2], mn[3]);
              if (p.store_ecs) {
                const __half2 h01 = __floats2half2_rn(en[0], en[1]), h23 = __floats2half2_rn(en[2], en[3]);
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.ecs) + idx) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              }
              if (p.ecs_save != nullptr) *reinterpret_cast<float4*>(p.ecs_save + idx) = make_float4(en[0], en[1], en[2], en[3]);
            } else {
              nib = 0;
            }
            uint32_t w = nib << (4 * sub);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if (ok && sub == 0) p.bits_next[idx >> 5] = w;
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of the spike convolution:  dW[co][(tap, ci)] += sum_pixels gy[p][co] * s[p*stride + tap - pad][ci]
// The contraction runs over PIXELS, i.e. over the rows of both operands, so both are fed to the tensor
// core as MN-major tiles: A = gy tile [128 pixels x 128 co] (bf16 hi [+ lo], 4-D TMA boxes in the tile's
// pixel order), B = the SAME expanded spike tile the forward kernel builds ([128 pixels x 64 channels] for
// one (tap, slab)), reinterpreted.  A CTA owns one 128-co tile, `bpc` (tap, slab) blocks (bpc*64 TMEM
// columns) and a slice of the pixel tiles; the accumulators stay in TMEM until the slice is done and are
// then added to dW with red.global.
// ---------------------------------------------------------------------------------------------
constexpr int kWgStages = 4;

struct WgCtl {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kWgStages], b_empty[kWgStages];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct WgArgs {
  int m_tiles;      // pixel tiles
  int splits_m;     // CTAs along the pixel tiles
  int bpc;          // (tap, slab) blocks per CTA
  int Cout, K;      // dW is [Cout][K]
  float* dw;
};

template <int G_SPLIT>
__global__ void __launch_bounds__(kSpikeThreads, 1)
k_umma_wgrad(const __grid_constant__ CUtensorMap tm_g0, const __grid_constant__ CUtensorMap tm_g1, const WgArgs g,
             const SpikeGeom sg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABuf = G_SPLIT * 2 * kATileBytes;   // planes x two 64-co blocks x (128 rows x 128 B)
  uint8_t* a_smem = smem;                             // 2 buffers
  uint8_t* b_smem = smem + 2 * kABuf;                 // kWgStages x 16 KB
  WgCtl* ctl = reinterpret_cast<WgCtl*>(b_smem + kWgStages * kATileBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * 128;
  const int blk0 = blockIdx.z * g.bpc;
  const int per = (g.m_tiles + g.splits_m - 1) / g.splits_m;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(t_begin + per, g.m_tiles);
  const int ntiles = max(t_end - t_begin, 0);
  constexpr int kWpg = kExpWarps / kWgStages;  // expander warps per stage group

  if (warp == 4 && lane == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->a_full[b], 1); mbar_init(&ctl->a_empty[b], 1); }
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&ctl->b_full[s], kWpg); mbar_init(&ctl->b_empty[s], 1); }
    mbar_init(&ctl->done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const int tiles_hw = sg.tiles_h * sg.tiles_w;

  if (warp == 4) {
    // ---- TMA: gy tiles (two 64-co blocks per plane), one per pixel tile ----
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ab = i & 1, ph = (i >> 1) & 1;
      mbar_wait(&ctl->a_empty[ab], ph ^ 1);
      if (lane == 0) {
        const int m_tile = t_begin + i;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        uint8_t* dst = a_smem + ab * kABuf;
        mbar_arrive_expect_tx(&ctl->a_full[ab], (uint32_t)kABuf);
#pragma unroll
        for (int sp = 0; sp < G_SPLIT; ++sp) {
          const CUtensorMap* tm = sp == 0 ? &tm_g0 : &tm_g1;
          tma_load_4d(dst + sp * 2 * kATileBytes, tm, &ctl->a_full[ab], co0, tw * sg.tw_b, th * sg.th_b, tn * sg.tn_b);
          tma_load_4d(dst + sp * 2 * kATileBytes + kATileBytes, tm, &ctl->a_full[ab], co0 + 64, tw * sg.tw_b,
                      th * sg.th_b, tn * sg.tn_b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ---- MMA issuer ----
    const uint32_t idesc = umma_idesc_bf16_mn(128, 64);
    uint32_t c = 0;
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ab = i & 1, ph = (i >> 1) & 1;
      mbar_wait(&ctl->a_full[ab], ph);
      for (int b = 0; b < g.bpc; ++b, ++c) {
        const uint32_t stage = c % kWgStages, sph = (c / kWgStages) & 1;
        mbar_wait(&ctl->b_full[stage], sph);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(a_smem + ab * kABuf);
          const uint32_t b_addr = smem_u32(b_smem + stage * kATileBytes);
          uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
          for (int sp = 0; sp < G_SPLIT; ++sp) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 128 pixels = 8 K steps of 16 rows (2048 bytes each)
              const uint64_t da = umma_desc_sw128_mn(a_addr + sp * 2 * kATileBytes + k * 2048, kATileBytes);
              const uint64_t db = umma_desc_sw128_mn(b_addr + k * 2048, kATileBytes);
              umma_f16(tmem_base + b * 64, da, db, idesc, acc);
              acc = 1u;
            }
          }
          umma_commit(&ctl->b_empty[stage]);
          if (b == g.bpc - 1) {
            umma_commit(&ctl->a_empty[ab]);
            if (i == ntiles - 1) umma_commit(&ctl->done);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 6) {
    // ---- spike expanders (same tile the forward conv builds; stage group = stage) ----
    const int ew = warp - 6;
    const int grp = ew / kWpg, sub = ew - grp * kWpg;
    constexpr int rows_per_warp = 128 / kWpg;
    constexpr int nrow = rows_per_warp >> 5;
    int n_l[nrow], h_l[nrow], w_l[nrow];
    uint32_t row_off[nrow];
#pragma unroll
    for (int i = 0; i < nrow; ++i) {
      const int r = sub * rows_per_warp + i * 32 + lane;
      w_l[i] = r & (sg.tw_b - 1);
      h_l[i] = (r >> sg.tw_sh) & (sg.th_b - 1);
      n_l[i] = r >> (sg.tw_sh + sg.th_sh);
      row_off[i] = (uint32_t)r * 128u;
    }
    const uint32_t sw = (uint32_t)(lane & 7);
    uint8_t* tile_b = b_smem + (size_t)grp * kATileBytes;
    const uint32_t c_end = (uint32_t)ntiles * (uint32_t)g.bpc;

    auto load_rows = [&](uint32_t c, uint2 (&wd)[nrow]) {
      const uint32_t ti = c / (uint32_t)g.bpc;
      const int blk = blk0 + (int)(c - ti * g.bpc);
      const int m_tile = t_begin + (int)ti;
      const int tn = m_tile / tiles_hw;
      const int rem = m_tile - tn * tiles_hw;
      const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
      const int tap = blk / sg.nslab, slab = blk - tap * sg.nslab;
      const int ky = tap / sg.kw, kx = tap - ky * sg.kw;
      const int img0 = tn * sg.tn_b;
      const int hi0 = th * sg.th_b * sg.stride - sg.pad + ky;
      const int wi0 = tw * sg.tw_b * sg.stride - sg.pad + kx;
#pragma unroll
      for (int i = 0; i < nrow; ++i) {
        wd[i] = make_uint2(0u, 0u);
        const int img = img0 + n_l[i], hi = hi0 + h_l[i] * sg.stride, wi = wi0 + w_l[i] * sg.stride;
        if (img < sg.imgs && hi >= 0 && hi < sg.H && wi >= 0 && wi < sg.W)
          wd[i] = __ldg(reinterpret_cast<const uint2*>(sg.bits + (((int64_t)img * sg.H + hi) * sg.W + wi) * sg.Cw + slab * 2));
      }
    };

    uint32_t c = (uint32_t)grp;
    uint2 cur[nrow];
    if (c < c_end) load_rows(c, cur);
    while (c < c_end) {
      uint2 nxt[nrow];
      if (c + kWgStages < c_end) load_rows(c + kWgStages, nxt);
      mbar_wait(&ctl->b_empty[grp], ((c / kWgStages) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < nrow; ++i) {
        uint8_t* row = tile_b + row_off[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t byte = ((j < 4 ? cur[i].x : cur[i].y) >> (8 * (j & 3))) & 0xFFu;
          uint4 o;
          o.x = bits2_to_bf16x2(byte);
          o.y = bits2_to_bf16x2(byte >> 2);
          o.z = bits2_to_bf16x2(byte >> 4);
          o.w = bits2_to_bf16x2(byte >> 6);
          *reinterpret_cast<uint4*>(row + (((uint32_t)j ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->b_full[grp]);
#pragma unroll
      for (int i = 0; i < nrow; ++i) cur[i] = nxt[i];
      c += kWgStages;
    }
  } else {
    // ---- epilogue: TMEM -> red.global into dW[co][k] ----
    if (ntiles > 0) {
      mbar_wait(&ctl->done, 0);
      tc_fence_after_sync();
      const int co = co0 + warp * 32 + lane;
      for (int b = 0; b < g.bpc; ++b) {
        const int blk = blk0 + b;
        const int tap = blk / sg.nslab, slab = blk - tap * sg.nslab;
        const int kbase = (tap * sg.nslab + slab) * 64;
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + b * 64 + c0, v);
          tmem_ld_wait();
          if (co < g.Cout) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(g.dw + (int64_t)co * g.K + kbase + c0 + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of the spike convolution, operand roles swapped (default):  D[(blk, ci)][co], i.e. the M side of the MMA is
// a PAIR of expanded (tap, slab) spike blocks (128 = 2 x 64 input channels: always a full tile) and the N side is the
// gy tile (NB = min(Cout, 128) output channels).  k_umma_wgrad read the whole gy tile (32 KB, half of it zero fill for
// Cout = 64) from shared memory once per 64-channel spike block: 64 KB of shared-memory traffic per 256 MMA cycles, twice
// what the SM delivers.  Here the gy tile is read once per PAIR of blocks and sized to the real Cout: 40 KB (Cout = 64) /
// 48 KB (Cout >= 128) per block.  The epilogue lane is an input channel, so consecutive lanes add to consecutive dW addresses.
// The (tap, slab) blocks of a CTA are padded to an even count with an all-zero block.
// ---------------------------------------------------------------------------------------------
struct WgtArgs {
  int m_tiles;      // pixel tiles
  int splits_m;     // CTAs along the pixel tiles
  int bpc;          // (tap, slab) blocks per CTA, even
  int nblk;         // real blocks (kh * kw * nslab); blocks >= nblk are zero padding
  int Cout, K;      // dW is [Cout][K]
  float* dw;
};

template <int G_SPLIT, int NB>
__global__ void __launch_bounds__(kSpikeThreads, 1)
k_umma_wgrad_t(const __grid_constant__ CUtensorMap tm_g0, const __grid_constant__ CUtensorMap tm_g1, const WgtArgs g,
               const SpikeGeom sg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kCoBlk = NB / 64;
  constexpr int kABuf = G_SPLIT * kCoBlk * kATileBytes;   // planes x 64-co blocks x (128 pixels x 128 B)
  uint8_t* a_smem = smem;                                  // 2 buffers of gy
  uint8_t* b_smem = smem + 2 * kABuf;                      // kWgStages x 16 KB of expanded spikes
  WgCtl* ctl = reinterpret_cast<WgCtl*>(b_smem + kWgStages * kATileBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * NB;
  const int blk0 = blockIdx.z * g.bpc;
  const int per = (g.m_tiles + g.splits_m - 1) / g.splits_m;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(t_begin + per, g.m_tiles);
  const int ntiles = max(t_end - t_begin, 0);
  constexpr int kWpg = kExpWarps / kWgStages;  // expander warps per stage group

  if (warp == 4 && lane == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->a_full[b], 1); mbar_init(&ctl->a_empty[b], 1); }
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&ctl->b_full[s], kWpg); mbar_init(&ctl->b_empty[s], 1); }
    mbar_init(&ctl->done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const int tiles_hw = sg.tiles_h * sg.tiles_w;

  if (warp == 4) {
    // ---- TMA: gy tiles (kCoBlk 64-co blocks per plane), one per pixel tile ----
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ab = i & 1, ph = (i >> 1) & 1;
      mbar_wait(&ctl->a_empty[ab], ph ^ 1);
      if (lane == 0) {
        const int m_tile = t_begin + i;
        const int tn = m_tile / tiles_hw;
        const int rem = m_tile - tn * tiles_hw;
        const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
        uint8_t* dst = a_smem + ab * kABuf;
        mbar_arrive_expect_tx(&ctl->a_full[ab], (uint32_t)kABuf);
#pragma unroll
        for (int sp = 0; sp < G_SPLIT; ++sp) {
          const CUtensorMap* tm = sp == 0 ? &tm_g0 : &tm_g1;
#pragma unroll
          for (int cb = 0; cb < kCoBlk; ++cb)
            tma_load_4d(dst + (sp * kCoBlk + cb) * kATileBytes, tm, &ctl->a_full[ab], co0 + cb * 64, tw * sg.tw_b,
                        th * sg.th_b, tn * sg.tn_b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ---- MMA issuer: one 128 x NB MMA set per PAIR of spike blocks (stages 2p, 2p+1 are adjacent in shared memory) ----
    const uint32_t idesc = umma_idesc_bf16_mn(128, NB);
    uint32_t c = 0;
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ab = i & 1, ph = (i >> 1) & 1;
      mbar_wait(&ctl->a_full[ab], ph);
      for (int b = 0; b < g.bpc; b += 2, c += 2) {
        const uint32_t stage = c % kWgStages, sph = (c / kWgStages) & 1;
        mbar_wait(&ctl->b_full[stage], sph);
        mbar_wait(&ctl->b_full[stage + 1], sph);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(a_smem + ab * kABuf);
          const uint32_t b_addr = smem_u32(b_smem + stage * kATileBytes);
          uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
          for (int sp = 0; sp < G_SPLIT; ++sp) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 128 pixels = 8 K steps of 16 rows (2048 bytes each)
              const uint64_t dm = umma_desc_sw128_mn(b_addr + k * 2048, kATileBytes);                      // spikes: M
              const uint64_t dn = umma_desc_sw128_mn(a_addr + sp * kCoBlk * kATileBytes + k * 2048, kATileBytes);   // gy: N
              umma_f16(tmem_base + (b >> 1) * NB, dm, dn, idesc, acc);
              acc = 1u;
            }
          }
          umma_commit(&ctl->b_empty[stage]);
          umma_commit(&ctl->b_empty[stage + 1]);
          if (b + 2 >= g.bpc) {
            umma_commit(&ctl->a_empty[ab]);
            if (i == ntiles - 1) umma_commit(&ctl->done);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 6) {
    // ---- spike expanders (same tile the forward conv builds; stage group = stage) ----
    const int ew = warp - 6;
    const int grp = ew / kWpg, sub = ew - grp * kWpg;
    constexpr int rows_per_warp = 128 / kWpg;
    constexpr int nrow = rows_per_warp >> 5;
    int n_l[nrow], h_l[nrow], w_l[nrow];
    uint32_t row_off[nrow];
#pragma unroll
    for (int i = 0; i < nrow; ++i) {
      const int r = sub * rows_per_warp + i * 32 + lane;
      w_l[i] = r & (sg.tw_b - 1);
      h_l[i] = (r >> sg.tw_sh) & (sg.th_b - 1);
      n_l[i] = r >> (sg.tw_sh + sg.th_sh);
      row_off[i] = (uint32_t)r * 128u;
    }
    const uint32_t sw = (uint32_t)(lane & 7);
    uint8_t* tile_b = b_smem + (size_t)grp * kATileBytes;
    const uint32_t c_end = (uint32_t)ntiles * (uint32_t)g.bpc;

    auto load_rows = [&](uint32_t c, uint2 (&wd)[nrow]) {
      const uint32_t ti = c / (uint32_t)g.bpc;
      const int blk = blk0 + (int)(c - ti * g.bpc);
      const int m_tile = t_begin + (int)ti;
      const int tn = m_tile / tiles_hw;
      const int rem = m_tile - tn * tiles_hw;
      const int th = rem / sg.tiles_w, tw = rem - th * sg.tiles_w;
      const int tap = blk / sg.nslab, slab = blk - tap * sg.nslab;
      const int ky = tap / sg.kw, kx = tap - ky * sg.kw;
      const int img0 = tn * sg.tn_b;
      const int hi0 = th * sg.th_b * sg.stride - sg.pad + ky;
      const int wi0 = tw * sg.tw_b * sg.stride - sg.pad + kx;
#pragma unroll
      for (int i = 0; i < nrow; ++i) {
        wd[i] = make_uint2(0u, 0u);
        const int img = img0 + n_l[i], hi = hi0 + h_l[i] * sg.stride, wi = wi0 + w_l[i] * sg.stride;
        if (blk < g.nblk && img < sg.imgs && hi >= 0 && hi < sg.H && wi >= 0 && wi < sg.W)
          wd[i] = __ldg(reinterpret_cast<const uint2*>(sg.bits + (((int64_t)img * sg.H + hi) * sg.W + wi) * sg.Cw + slab * 2));
      }
    };

    uint32_t c = (uint32_t)grp;
    uint2 cur[nrow];
    if (c < c_end) load_rows(c, cur);
    while (c < c_end) {
      uint2 nxt[nrow];
      if (c + kWgStages < c_end) load_rows(c + kWgStages, nxt);
      mbar_wait(&ctl->b_empty[grp], ((c / kWgStages) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < nrow; ++i) {
        uint8_t* row = tile_b + row_off[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t byte = ((j < 4 ? cur[i].x : cur[i].y) >> (8 * (j & 3))) & 0xFFu;
          uint4 o;
          o.x = bits2_to_bf16x2(byte);
          o.y = bits2_to_bf16x2(byte >> 2);
          o.z = bits2_to_bf16x2(byte >> 4);
          o.w = bits2_to_bf16x2(byte >> 6);
          *reinterpret_cast<uint4*>(row + (((uint32_t)j ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->b_full[grp]);
#pragma unroll
      for (int i = 0; i < nrow; ++i) cur[i] = nxt[i];
      c += kWgStages;
    }
  } else {
    // ---- epilogue: TMEM lane = (block of the pair, input channel), column = output channel -> red.global into dW ----
    if (ntiles > 0) {
      mbar_wait(&ctl->done, 0);
      tc_fence_after_sync();
      const int half = warp >> 1;                         // lanes 0-63: first block of the pair, 64-127: second
      const int ci = (warp & 1) * 32 + lane;
      for (int p = 0; p < g.bpc / 2; ++p) {
        const int blk = blk0 + 2 * p + half;
        const int64_t kbase = (int64_t)blk * 64 + ci;     // blk = tap * nslab + slab: K index (tap * nslab + slab) * 64 + ci
        for (int c0 = 0; c0 < NB; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + p * NB + c0, v);
          tmem_ld_wait();
          if (blk < g.nblk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = co0 + c0 + j;
              if (co < g.Cout) atomicAdd(g.dw + (int64_t)co * g.K + kbase, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

std::mutex g_tm_mu;
std::map<std::tuple<const void*, uint64_t, uint64_t, uint32_t>, CUtensorMap> g_tm_cache;

}  // namespace

// 2-D bf16 row-major [rows][cols] tensor map with a {64, box_rows} box, 128-byte swizzle, zero OOB fill.
int ecsy_tensor_map_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMap* out) {
  ECSY_CHECK_ARG(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: pointer must be 16-byte aligned");
  ECSY_CHECK_ARG(cols % 8 == 0 && box_rows >= 1 && box_rows <= 256, "tensor map: cols %% 8, box rows <= 256");
  auto key = std::make_tuple(ptr, rows, cols, box_rows);
  {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    auto it = g_tm_cache.find(key);
    if (it != g_tm_cache.end()) {
      *out = it->second;
      return ECSY_OK;
    }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    ecsy_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECSY_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecsy_set_error("cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu box=%u)", (int)r,
                   (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return ECSY_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_tm_mu);
  if (g_tm_cache.size() > 4096) g_tm_cache.clear();
  g_tm_cache[key] = *out;
  return ECSY_OK;
}

// 4-D bf16 NHWC [imgs][H][W][C] tensor map with a {64, tw, th, tn} box, 128-byte swizzle, zero OOB fill.
int ecsy_tensor_map_bf16_nhwc(const void* ptr, int imgs, int H, int W, int C, int tn, int th, int tw, CUtensorMap* out) {
  ECSY_CHECK_ARG(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && C % 8 == 0, "nhwc tensor map: alignment");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    ecsy_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECSY_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)imgs};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecsy_set_error("cuTensorMapEncodeTiled(4d) failed with %d (imgs=%d H=%d W=%d C=%d box=%d,%d,%d)", (int)r, imgs, H, W,
                   C, tn, th, tw);
    return ECSY_ERR_CUDA;
  }
  return ECSY_OK;
}

// 4-D fp32 NHWC [imgs][H][W][C] tensor map with a {32 floats, tw, th, tn} box (128-byte rows), 128-byte swizzle:
// the staging slices of the TMA-store epilogue.
static int ecsy_tensor_map_f32_nhwc(const void* ptr, int imgs, int H, int W, int C, int tn, int th, int tw,
                                    CUtensorMap* out) {
  ECSY_CHECK_ARG(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && C % 32 == 0, "f32 nhwc tensor map: alignment");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    ecsy_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECSY_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)imgs};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecsy_set_error("cuTensorMapEncodeTiled(f32 4d) failed with %d (imgs=%d H=%d W=%d C=%d box=%d,%d,%d)", (int)r, imgs, H,
                   W, C, tn, th, tw);
    return ECSY_ERR_CUDA;
  }
  return ECSY_OK;
}

namespace {

constexpr int kSmemLimit = 227 * 1024;

template <int BN, int A_MODE, int A_SPLIT, int B_SPLIT, int EPI>
int launch_one(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, GemmArgs g, const SpikeGeom& sg,
               const typename EpiSel<EPI>::type& ep, int patch_bytes, cudaStream_t st) {
  constexpr bool kTS = A_MODE == kASpikesT;
  constexpr int stage_bytes = (kTS ? 0 : A_SPLIT * kATileBytes) + B_SPLIT * BN * 128;
  auto kern = k_umma_gemm<BN, A_MODE, A_SPLIT, B_SPLIT, EPI>;
  static int dyn_limit = 0;  // opt-in dynamic shared memory = 227 KB minus the kernel's static usage
  if (dyn_limit == 0) {
    cudaFuncAttributes fa;
    ECSY_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int lim = kSmemLimit - (int)fa.sharedSizeBytes;
    ECSY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    dyn_limit = lim;
  }
  (void)patch_bytes;
  constexpr int kEpiStage = 4 * kEpiWarpFloats * 4;
  const int fixed = 1024 /*align slack*/ + (int)sizeof(SharedCtl) + 128 + kEpiStage + (A_MODE == kAGather ? g.kb_total * 64 * 8 : 0);
  int stages = (dyn_limit - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (A_MODE == kASpikes) stages = stages >= 8 ? 8 : (stages >= 4 ? 4 : (stages >= 2 ? 2 : 0));
  if (kTS) {  // the A ring lives in the TMEM columns behind the two accumulators (32 columns per stage)
    const int tmem_stages = (512 - 2 * BN) / 32;
    if (stages > tmem_stages) stages = tmem_stages;
    stages &= ~1;
  }
  if (stages < 2) {
    ecsy_set_error("umma gemm: shared memory budget allows only %d stage(s) (patch %d bytes)", stages, patch_bytes);
    return ECSY_ERR_UNSUPPORTED;
  }
  g.stages = stages;
  g.wpg = A_MODE == kASpikes ? kExpWarps / stages : 1;
  const uint32_t ctl_off = (uint32_t)stages * stage_bytes;
  g.epi_off = (ctl_off + (uint32_t)sizeof(SharedCtl) + 63u) & ~63u;
  g.ktab_off = g.epi_off + (uint32_t)kEpiStage;
  const int smem = 1024 + (int)g.epi_off + kEpiStage + (A_MODE == kAGather ? g.kb_total * 64 * 8 : 0);
  int grid = g.m_tiles * g.n_tiles;
  const int sms = ecsy_num_sms();
  if (grid > sms) grid = sms;
  kern<<<grid, (A_MODE == kASpikes || A_MODE == kADw || A_MODE == kAGather || kTS) ? kSpikeThreads : 192, smem, st>>>(a0, a1, b, g, sg, ep);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

template <int A_MODE, int A_SPLIT, int B_SPLIT, int EPI>
int launch_bn(int BN, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmArgs& g,
              const SpikeGeom& sg, const typename EpiSel<EPI>::type& ep, int patch_bytes, cudaStream_t st) {
  switch (BN) {
    case 64: return launch_one<64, A_MODE, A_SPLIT, B_SPLIT, EPI>(a0, a1, b, g, sg, ep, patch_bytes, st);
    case 128: return launch_one<128, A_MODE, A_SPLIT, B_SPLIT, EPI>(a0, a1, b, g, sg, ep, patch_bytes, st);
    case 256:
      if constexpr (A_MODE != kASpikesT) return launch_one<256, A_MODE, A_SPLIT, B_SPLIT, EPI>(a0, a1, b, g, sg, ep, patch_bytes, st);
      break;
  }
  ecsy_set_error("umma gemm: unsupported BN=%d", BN);
  return ECSY_ERR_UNSUPPORTED;
}

int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

}  // namespace

// choose the 128-pixel tile box (tn, th, tw), powers of two: least padding waste, then widest rows
static bool pick_tile_box(SpikeGeom& sg, int imgs, int Ho, int Wo, int Cw, int k, int stride) {
  double best = 1e30;
  for (int tw = 1; tw <= 128; tw <<= 1)
    for (int th = 1; th * tw <= 128; th <<= 1) {
      const int tn = 128 / (tw * th);
      const int64_t cov = (int64_t)((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th * ((imgs + tn - 1) / tn) * tn;
      const int Hp = (th - 1) * stride + k, Wp = (tw - 1) * stride + k;
      const int64_t patch_bytes = (int64_t)tn * Hp * Wp * Cw * 4;
      const double score = (double)cov * (1.0 + 1e-3 * (double)patch_bytes / (128.0 * Cw * 4)) - 1e-6 * tw;
      if (score < best) {
        best = score;
        sg.tn_b = tn; sg.th_b = th; sg.tw_b = tw;
        sg.Hp = Hp; sg.Wp = Wp; sg.PP = tn * Hp * Wp;
      }
    }
  return best < 1e30;
}

int ecsy_pick_bn(int cout, int splits) {
  // wide tiles amortise the A expansion; the split-weight mode doubles the B tile, so cap at 128
  const int cap = splits == 2 ? 128 : 256;
  for (int bn = cap; bn >= 64; bn >>= 1)
    if (cout % bn == 0) return bn;
  return 0;
}

// Spike convolution: out[imgs][Ho][Wo][Cout] = conv(spikes, W) (*scale + shift) (+ residual)
namespace {
template <int BN, int B_SPLIT, int V>
int launch_ts_v(const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tres, TsArgs g, const SpikeGeom& sg,
                cudaStream_t st) {
  auto kern = k_spike_conv_ts<BN, B_SPLIT, V>;
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr = true;
  }
  constexpr int block_bytes = B_SPLIT * BN * 128;            // one weight K block
  constexpr int stage_bytes = V == 2 ? 2 * block_bytes : block_bytes;
  constexpr int tmem_stages = (512 - 2 * BN) / (V == 2 ? 64 : 32);
  const int aff_bytes = (2 * g.cout * 4 + 127) & ~127;
  const int fixed = 1024 + aff_bytes + (int)sizeof(TsCtl) + 128;
  static const bool allow_res = getenv("ECSY_TS_RESIDENT") == nullptr || getenv("ECSY_TS_RESIDENT")[0] != '0';
  g.nbuf = g.has_res ? 3 : 2;
  // all weight K blocks resident (n_tiles == 1: one weight panel for every tile of the CTA)?
  int b_bytes = 0, stages = 0;
  auto plan = [&]() {
    const int avail = kSmemLimit - fixed - 2 * g.nbuf * kTsSlice;
    g.resident = (allow_res && g.n_tiles == 1 && g.kb_total >= 4 && g.kb_total <= kTsMaxStages &&
                  g.kb_total * block_bytes <= avail && (V == 2 || g.kb_total <= tmem_stages)) ? 1 : 0;
    if (g.resident) {
      b_bytes = g.kb_total * block_bytes;
      stages = V == 2 ? tmem_stages : g.kb_total;
    } else {
      stages = avail / stage_bytes;
      if (stages > tmem_stages) stages = tmem_stages;
      if (stages > 8) stages = 8;
      b_bytes = stages * stage_bytes;
    }
  };
  plan();
  if (!g.resident && stages < (V == 2 ? 3 : 4) && g.nbuf == 3) {   // the weight ring matters more than the residual prefetch
    g.nbuf = 2;
    plan();
  }
  if (stages < (V == 2 ? 3 : 2)) {   // V == 2: an expander hops up to 3 rounds at a tile boundary; the ring must be at least that deep
    ecsy_set_error("spike_conv_ts: shared memory budget allows only %d stage(s)", stages);
    return ECSY_ERR_UNSUPPORTED;
  }
  g.stages = stages;
  g.stg_off = (uint32_t)b_bytes;                                    // multiples of 8 KB: 1024-aligned
  g.aff_off = g.stg_off + 2u * (uint32_t)g.nbuf * kTsSlice;
  g.ctl_off = g.aff_off + (uint32_t)aff_bytes;
  const int smem = 1024 + (int)g.ctl_off + (int)sizeof(TsCtl) + 64;
  int grid = g.m_tiles * g.n_tiles;
  const int sms = ecsy_num_sms();
  if (grid > sms) grid = sms;
  kern<<<grid, V == 2 ? kTsThreads2 : kTsThreads, smem, st>>>(tb, tout, tres, g, sg);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

template <int BN, int B_SPLIT>
int launch_ts(const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tres, const TsArgs& g, const SpikeGeom& sg,
              cudaStream_t st) {
  static const bool v2 = getenv("ECSY_TS_V2") == nullptr || getenv("ECSY_TS_V2")[0] != '0';
  return v2 ? launch_ts_v<BN, B_SPLIT, 2>(tb, tout, tres, g, sg, st) : launch_ts_v<BN, B_SPLIT, 1>(tb, tout, tres, g, sg, st);
}
}  // namespace

int ecsy_pick_bn_ts(int cout) {
  // TS mode: the A ring shares the 512 TMEM columns with the two accumulators -> at most 128 columns per tile
  for (int bn = 128; bn >= 64; bn >>= 1)
    if (cout % bn == 0) return bn;
  return 0;
}

int ecsy_umma_spike_conv(const uint32_t* bits, const void* w_packed, int splits, float* out, const float* scale,
                         const float* shift, const float* residual, int64_t res_imgs, int imgs, int H, int W, int Cin,
                         int Cout, int k, int stride, int pad, cudaStream_t st, int ts) {
  ECSY_CHECK_ARG(Cin % 64 == 0 && Cin >= 64, "spike_conv: Cin=%d must be a multiple of 64", Cin);
  // ts: 0 = smem operand, natural weight layout; 1 = tensor-memory operand kernel; 2 = tensor-memory operand with the
  // legacy epilogue; 3 = smem operand (wide tiles) with weights in the tensor-memory layout (pair expansion)
  const bool tsk = ts == 1 || ts == 2;
  int BN = tsk ? ecsy_pick_bn_ts(Cout) : ecsy_pick_bn(Cout, splits);
  // Two weight planes (parity precision) on a 128-column tile leave room for only TWO ring stages of two K blocks each.
  // The expanders hop up to three rounds when they cross a tile boundary (their parity of K blocks), i.e. over a whole
  // period of a two-stage ring, and an mbarrier parity wait cannot tell phase p from phase p + 2: measured as a hang on
  // resnet10's 64 -> 128 stride-2 conv at batch 2 (tests/test_gpu_baseline_cfgs.py).  64-column tiles keep >= 4 stages.
  if (tsk && splits == 2 && BN == 128) BN = 64;
  ECSY_CHECK_ARG(BN != 0, "spike_conv: Cout=%d must be a multiple of 64", Cout);
  ECSY_CHECK_ARG(splits == 1 || splits == 2, "spike_conv: splits must be 1 or 2");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0, "spike_conv: empty output");
  SpikeGeom sg{};
  if (!pick_tile_box(sg, imgs, Ho, Wo, Cin / 32, k, stride)) {
    ecsy_set_error("spike_conv: no tile shape found");
    return ECSY_ERR_ARG;
  }
  const int Cw = Cin / 32;
  sg.bits = bits; sg.imgs = imgs; sg.H = H; sg.W = W; sg.Cw = Cw; sg.Ho = Ho; sg.Wo = Wo;
  sg.kh = k; sg.kw = k; sg.stride = stride; sg.pad = pad;
  sg.tn_sh = ilog2(sg.tn_b); sg.th_sh = ilog2(sg.th_b); sg.tw_sh = ilog2(sg.tw_b);
  sg.tiles_h = (Ho + sg.th_b - 1) / sg.th_b; sg.tiles_w = (Wo + sg.tw_b - 1) / sg.tw_b;
  sg.nslab = Cin / 64;
  GemmArgs g{};
  g.m_tiles = ((imgs + sg.tn_b - 1) / sg.tn_b) * sg.tiles_h * sg.tiles_w;
  g.n_tiles = Cout / BN;
  g.kb_total = k * k * sg.nslab;
  g.M = 0;
  const int K = k * k * Cin;
  CUtensorMap tb, dummy{};
  int rc = ecsy_tensor_map_bf16(w_packed, (uint64_t)splits * Cout, (uint64_t)K, (uint32_t)BN, &tb);
  if (rc) return rc;
  EpiConv e{out, scale, shift, residual, (residual ? res_imgs : (int64_t)imgs) * Ho * Wo, Cout, 0};
  const int patch_bytes = sg.PP * Cw * 4;
  g.pair_expand = ts == 3 ? 1 : 0;
  if (tsk) {
    // TMA-store epilogue unless a T-broadcast residual would straddle the wrap inside one tile box
    const int64_t rimgs = residual ? res_imgs : imgs;
    const bool tma_epi = ts != 2 && (residual == nullptr || rimgs == imgs || rimgs % sg.tn_b == 0);
    if (tma_epi) {
      CUtensorMap tout, tres{};
      rc = ecsy_tensor_map_f32_nhwc(out, imgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &tout);
      if (rc) return rc;
      if (residual) {
        rc = ecsy_tensor_map_f32_nhwc(residual, (int)rimgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &tres);
        if (rc) return rc;
      }
      TsArgs ta{};
      ta.m_tiles = g.m_tiles; ta.n_tiles = g.n_tiles; ta.kb_total = g.kb_total;
      ta.has_res = residual ? 1 : 0; ta.res_imgs = (int)rimgs; ta.cout = Cout;
      ta.scale = scale; ta.shift = shift;
      if (BN == 64) return splits == 1 ? launch_ts<64, 1>(tb, tout, tres, ta, sg, st) : launch_ts<64, 2>(tb, tout, tres, ta, sg, st);
      return splits == 1 ? launch_ts<128, 1>(tb, tout, tres, ta, sg, st) : launch_ts<128, 2>(tb, tout, tres, ta, sg, st);
    }
    if (splits == 1) return launch_bn<kASpikesT, 1, 1, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, patch_bytes, st);
    return launch_bn<kASpikesT, 1, 2, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, patch_bytes, st);
  }
  if (splits == 1) return launch_bn<kASpikes, 1, 1, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, patch_bytes, st);
  return launch_bn<kASpikes, 1, 2, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, patch_bytes, st);
}

// 2-D fp16 row-major [rows][cols] tensor map with a {64, 128} box (128-byte rows), 128-byte swizzle.
static int ecsy_tensor_map_f16_2d(const void* ptr, uint64_t rows, uint64_t cols, CUtensorMap* out) {
  ECSY_CHECK_ARG(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && cols % 64 == 0, "f16 tensor map: alignment");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    ecsy_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECSY_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecsy_set_error("cuTensorMapEncodeTiled(f16 2d) failed with %d (rows=%llu cols=%llu)", (int)r, (unsigned long long)rows,
                   (unsigned long long)cols);
    return ECSY_ERR_CUDA;
  }
  return ECSY_OK;
}

static int ecsy_tensor_map_f32_2d(const void* ptr, uint64_t rows, uint64_t cols, CUtensorMap* out) {
  ECSY_CHECK_ARG(ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && cols % 32 == 0, "f32 tensor map: alignment");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    ecsy_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECSY_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecsy_set_error("cuTensorMapEncodeTiled(f32 2d) failed with %d (rows=%llu cols=%llu)", (int)r, (unsigned long long)rows,
                   (unsigned long long)cols);
    return ECSY_ERR_CUDA;
  }
  return ECSY_OK;
}

namespace {
template <int BN, bool F32 = false>
int launch_dense_tma(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, DtArgs g, cudaStream_t st) {
  auto kern = k_dense_tma_h<BN, F32>;
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr = true;
  }
  constexpr int stage_bytes = kATileBytes + BN * 128;
  const int fixed = 1024 + (int)sizeof(DtCtl) + 128 + 4 * kATileBytes;
  int stages = (kSmemLimit - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    ecsy_set_error("dense_tma: shared memory budget allows only %d stage(s)", stages);
    return ECSY_ERR_UNSUPPORTED;
  }
  g.stages = stages;
  g.stg_off = (uint32_t)stages * stage_bytes;     // stage sizes are multiples of 8 KB: 1024-aligned
  g.ctl_off = g.stg_off + 4u * kATileBytes;
  const int smem = 1024 + (int)g.ctl_off + (int)sizeof(DtCtl) + 64;
  int grid = g.m_tiles * g.n_tiles;
  const int sms = ecsy_num_sms();
  if (grid > sms) grid = sms;
  kern<<<grid, kDtThreads, smem, st>>>(ta, tb, to, g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
}  // namespace

namespace {
template <int BN>
int launch_ecs_gemm(const CUtensorMap& ta, const CUtensorMap& tb, EgArgs g, cudaStream_t st) {
  auto kern = k_ecs_gemm<BN>;
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr = true;
  }
  constexpr int stage_bytes = kATileBytes + BN * 128;
  constexpr int stg_bytes = 8 * 32 * kEgRow * (int)sizeof(float);
  const int fixed = 1024 + (int)sizeof(DtCtl) + 128 + stg_bytes;
  int stages = (kSmemLimit - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    ecsy_set_error("ecs_gemm: shared memory budget allows only %d stage(s)", stages);
    return ECSY_ERR_UNSUPPORTED;
  }
  g.stages = stages;
  g.stg_off = (uint32_t)stages * stage_bytes;
  g.ctl_off = g.stg_off + (uint32_t)stg_bytes;
  const int smem = 1024 + (int)g.ctl_off + (int)sizeof(DtCtl) + 64;
  int grid = g.m_tiles * g.n_tiles;
  const int sms = ecsy_num_sms();
  if (grid > sms) grid = sms;
  kern<<<grid, kDtThreads, smem, st>>>(ta, tb, g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
}  // namespace

// One ECS timestep with the point-wise spread GEMM fused in front of it (fast precision: s.half_state): a_hi [M][C] bf16 =
// dw(s_t), pw_packed [C][C] bf16; everything else as ecsy_launch_ecs_step (s.spread is not used).
int ecsy_umma_ecs_step(const void* a_hi, int64_t M, int C, const void* pw_packed, const EcsStep& s, cudaStream_t st) {
  ECSY_CHECK_ARG(a_hi && pw_packed && C % 64 == 0 && M >= 128 && s.half_state, "ecs_gemm: bad arguments");
  const int BN = C % 256 == 0 ? 256 : (C % 128 == 0 ? 128 : 64);
  CUtensorMap ta, tb;
  int rc = ecsy_tensor_map_bf16(a_hi, (uint64_t)M, (uint64_t)C, 128, &ta);
  if (rc) return rc;
  rc = ecsy_tensor_map_bf16(pw_packed, (uint64_t)C, (uint64_t)C, (uint32_t)BN, &tb);
  if (rc) return rc;
  EgArgs g{};
  g.m_tiles = (int)((M + 127) / 128);
  g.n_tiles = C / BN;
  g.kb_total = C / 64;
  g.M = M;
  g.C = C;
  g.s = s;
  switch (BN) {
    case 64: return launch_ecs_gemm<64>(ta, tb, g, st);
    case 128: return launch_ecs_gemm<128>(ta, tb, g, st);
    default: return launch_ecs_gemm<256>(ta, tb, g, st);
  }
}

// Dense GEMM on a bf16 A matrix (hi [+ lo]) : out[M][Cout] = A * W^T (*scale + shift) (+ residual)
int ecsy_umma_dense(const void* a_hi, const void* a_lo, int64_t M, int K, const void* w_packed, int splits, float* out,
                    int Cout, const float* scale, const float* shift, const float* residual, int64_t res_rows,
                    cudaStream_t st, int out_half) {
  ECSY_CHECK_ARG(K % 64 == 0 && M > 0, "dense gemm: K=%d must be a multiple of 64", K);
  const int BN = ecsy_pick_bn(Cout, splits);
  ECSY_CHECK_ARG(BN != 0, "dense gemm: Cout=%d must be a multiple of 64", Cout);
  ECSY_CHECK_ARG((splits == 2) == (a_lo != nullptr), "dense gemm: lo plane required iff splits == 2");
  CUtensorMap ta0, ta1{}, tb;
  int rc = ecsy_tensor_map_bf16(a_hi, (uint64_t)M, (uint64_t)K, 128, &ta0);
  if (rc) return rc;
  if (a_lo) {
    rc = ecsy_tensor_map_bf16(a_lo, (uint64_t)M, (uint64_t)K, 128, &ta1);
    if (rc) return rc;
  }
  rc = ecsy_tensor_map_bf16(w_packed, (uint64_t)splits * Cout, (uint64_t)K, (uint32_t)BN, &tb);
  if (rc) return rc;
  static const bool dense_tma = getenv("ECSY_DENSE_TMA") == nullptr || getenv("ECSY_DENSE_TMA")[0] != '0';
  if (dense_tma && splits == 1 && out_half && !scale && !residual && M >= 128) {
    // fast-mode ECS spread: fp16 output through the two-group TMA-store epilogue
    CUtensorMap to;
    rc = ecsy_tensor_map_f16_2d(out, (uint64_t)M, (uint64_t)Cout, &to);
    if (rc) return rc;
    DtArgs d{};
    d.m_tiles = (int)((M + 127) / 128);
    d.n_tiles = Cout / BN;
    d.kb_total = K / 64;
    switch (BN) {
      case 64: return launch_dense_tma<64>(ta0, tb, to, d, st);
      case 128: return launch_dense_tma<128>(ta0, tb, to, d, st);
      case 256: return launch_dense_tma<256>(ta0, tb, to, d, st);
    }
  }
  static const bool dense_f32 = getenv("ECSY_DENSE_TMA_F32") == nullptr || getenv("ECSY_DENSE_TMA_F32")[0] != '0';
  if (dense_tma && dense_f32 && splits == 1 && !out_half && !scale && !residual && M >= 128) {
    // fast precision, plain fp32 output (G1 = ge * Wpw of the LIF backward): the same two-group TMA-store epilogue
    CUtensorMap to;
    rc = ecsy_tensor_map_f32_2d(out, (uint64_t)M, (uint64_t)Cout, &to);
    if (rc) return rc;
    DtArgs d{};
    d.m_tiles = (int)((M + 127) / 128);
    d.n_tiles = Cout / BN;
    d.kb_total = K / 64;
    switch (BN) {
      case 64: return launch_dense_tma<64, true>(ta0, tb, to, d, st);
      case 128: return launch_dense_tma<128, true>(ta0, tb, to, d, st);
      case 256: return launch_dense_tma<256, true>(ta0, tb, to, d, st);
    }
  }
  GemmArgs g{};
  g.m_tiles = (int)((M + 127) / 128);
  g.n_tiles = Cout / BN;
  g.kb_total = K / 64;
  g.M = M;
  SpikeGeom sg{};
  EpiConv e{out, scale, shift, residual, residual ? res_rows : M, Cout, out_half};
  if (splits == 1) return launch_bn<kATma, 1, 1, kEpiConv>(BN, ta0, ta1, tb, g, sg, e, 0, st);
  return launch_bn<kATma, 2, 2, kEpiConv>(BN, ta0, ta1, tb, g, sg, e, 0, st);
}


// Stride-1 convolution over a REAL-valued bf16 NHWC tensor (hi [+ lo] planes), implicit GEMM with 4-D TMA.
int ecsy_umma_conv_bf16(const void* a_hi, const void* a_lo, const void* w_packed, int splits, float* out,
                        const float* scale, const float* shift, const float* residual, int64_t res_imgs, int imgs, int H,
                        int W, int Cin, int Cout, int k, int pad, cudaStream_t st) {
  ECSY_CHECK_ARG(Cin % 64 == 0 && Cin >= 64, "conv_bf16: Cin=%d must be a multiple of 64", Cin);
  const int BN = ecsy_pick_bn(Cout, splits);
  ECSY_CHECK_ARG(BN != 0, "conv_bf16: Cout=%d must be a multiple of 64", Cout);
  ECSY_CHECK_ARG((splits == 2) == (a_lo != nullptr), "conv_bf16: lo plane required iff splits == 2");
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0, "conv_bf16: empty output");
  SpikeGeom sg{};
  if (!pick_tile_box(sg, imgs, Ho, Wo, Cin / 32, k, 1)) {
    ecsy_set_error("conv_bf16: no tile shape found");
    return ECSY_ERR_ARG;
  }
  sg.bits = nullptr; sg.imgs = imgs; sg.H = H; sg.W = W; sg.Cw = Cin / 32; sg.Ho = Ho; sg.Wo = Wo;
  sg.kh = k; sg.kw = k; sg.stride = 1; sg.pad = pad;
  sg.tn_sh = ilog2(sg.tn_b); sg.th_sh = ilog2(sg.th_b); sg.tw_sh = ilog2(sg.tw_b);
  sg.tiles_h = (Ho + sg.th_b - 1) / sg.th_b; sg.tiles_w = (Wo + sg.tw_b - 1) / sg.tw_b;
  sg.nslab = Cin / 64;
  GemmArgs g{};
  g.m_tiles = ((imgs + sg.tn_b - 1) / sg.tn_b) * sg.tiles_h * sg.tiles_w;
  g.n_tiles = Cout / BN;
  g.kb_total = k * k * sg.nslab;
  CUtensorMap ta0, ta1{}, tb;
  int rc = ecsy_tensor_map_bf16_nhwc(a_hi, imgs, H, W, Cin, sg.tn_b, sg.th_b, sg.tw_b, &ta0);
  if (rc) return rc;
  if (a_lo) {
    rc = ecsy_tensor_map_bf16_nhwc(a_lo, imgs, H, W, Cin, sg.tn_b, sg.th_b, sg.tw_b, &ta1);
    if (rc) return rc;
  }
  rc = ecsy_tensor_map_bf16(w_packed, (uint64_t)splits * Cout, (uint64_t)k * k * Cin, (uint32_t)BN, &tb);
  if (rc) return rc;
  EpiConv e{out, scale, shift, residual, (residual ? res_imgs : (int64_t)imgs) * Ho * Wo, Cout, 0};
  if (splits == 1) return launch_bn<kATma4, 1, 1, kEpiConv>(BN, ta0, ta1, tb, g, sg, e, 0, st);
  return launch_bn<kATma4, 2, 2, kEpiConv>(BN, ta0, ta1, tb, g, sg, e, 0, st);
}

// Convolution over a real-valued fp32 NHWC tensor with a small channel count (the stem: 3 -> 64, 7x7, stride 2):
// implicit GEMM whose A tiles are gathered on the fly (no im2col matrix).  Single weight plane only.
int ecsy_umma_conv_gather(const float* x, int64_t x_imgs, const void* w_packed, float* out, const float* scale,
                          const float* shift, int imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                          cudaStream_t st) {
  const int BN = ecsy_pick_bn(Cout, 1);
  ECSY_CHECK_ARG(BN != 0, "conv_gather: Cout=%d must be a multiple of 64", Cout);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0 && k < 256, "conv_gather: geometry");
  const int K = k * k * Cin, Kpad = (K + 63) / 64 * 64;
  ECSY_CHECK_ARG(Kpad <= 1024, "conv_gather: K=%d too large", K);
  SpikeGeom sg{};
  if (!pick_tile_box(sg, imgs, Ho, Wo, 1, k, stride)) {
    ecsy_set_error("conv_gather: no tile shape found");
    return ECSY_ERR_ARG;
  }
  sg.bits = nullptr; sg.imgs = imgs; sg.H = H; sg.W = W; sg.Cw = 0; sg.Ho = Ho; sg.Wo = Wo;
  sg.kh = k; sg.kw = k; sg.stride = stride; sg.pad = pad;
  sg.tn_sh = ilog2(sg.tn_b); sg.th_sh = ilog2(sg.th_b); sg.tw_sh = ilog2(sg.tw_b);
  sg.tiles_h = (Ho + sg.th_b - 1) / sg.th_b; sg.tiles_w = (Wo + sg.tw_b - 1) / sg.tw_b;
  sg.nslab = 1;
  GemmArgs g{};
  g.m_tiles = ((imgs + sg.tn_b - 1) / sg.tn_b) * sg.tiles_h * sg.tiles_w;
  g.n_tiles = Cout / BN;
  g.kb_total = Kpad / 64;
  g.gx = x; g.gx_imgs = x_imgs; g.gcin = Cin; g.gK = K;
  CUtensorMap tb, dummy{};
  int rc = ecsy_tensor_map_bf16(w_packed, (uint64_t)Cout, (uint64_t)Kpad, (uint32_t)BN, &tb);
  if (rc) return rc;
  EpiConv e{out, scale, shift, nullptr, (int64_t)imgs * Ho * Wo, Cout, 0};
  return launch_bn<kAGather, 1, 1, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, 0, st);
}

// dW[Cout][k*k*Cin] (fp32, accumulated) for a spike convolution; gy as bf16 planes [imgs][Ho][Wo][Cout].
int ecsy_umma_spike_wgrad(const void* gy_hi, const void* gy_lo, const uint32_t* bits, float* dw, int imgs, int H, int W,
                          int Cin, int Cout, int k, int stride, int pad, cudaStream_t st) {
  ECSY_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0, "spike_wgrad: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  SpikeGeom sg{};
  if (!pick_tile_box(sg, imgs, Ho, Wo, Cin / 32, k, stride)) {
    ecsy_set_error("spike_wgrad: no tile shape found");
    return ECSY_ERR_ARG;
  }
  sg.bits = bits; sg.imgs = imgs; sg.H = H; sg.W = W; sg.Cw = Cin / 32; sg.Ho = Ho; sg.Wo = Wo;
  sg.kh = k; sg.kw = k; sg.stride = stride; sg.pad = pad;
  sg.tn_sh = ilog2(sg.tn_b); sg.th_sh = ilog2(sg.th_b); sg.tw_sh = ilog2(sg.tw_b);
  sg.tiles_h = (Ho + sg.th_b - 1) / sg.th_b; sg.tiles_w = (Wo + sg.tw_b - 1) / sg.tw_b;
  sg.nslab = Cin / 64;
  const int gsplit_t = gy_lo ? 2 : 1;
  static const bool wgrad_old = getenv("ECSY_WGRAD_V") != nullptr && getenv("ECSY_WGRAD_V")[0] == '0';
  if (!wgrad_old) {
    // role-swapped kernel: pairs of spike blocks on the M side, gy on the N side
    const int NB = Cout >= 128 ? 128 : 64;
    const int nblk = k * k * sg.nslab, nblk_even = (nblk + 1) & ~1;
    const int max_bpc = NB == 128 ? 8 : 16;               // 512 TMEM columns = (bpc / 2) * NB
    int bpc = max_bpc;
    while (nblk_even % bpc != 0) bpc -= 2;
    WgtArgs g{};
    g.m_tiles = ((imgs + sg.tn_b - 1) / sg.tn_b) * sg.tiles_h * sg.tiles_w;
    g.bpc = bpc; g.nblk = nblk; g.Cout = Cout; g.K = k * k * Cin; g.dw = dw;
    const int co_tiles = (Cout + NB - 1) / NB, groups = nblk_even / bpc;
    int sm = (2 * ecsy_num_sms()) / (co_tiles * groups);
    if (sm < 1) sm = 1;
    if (sm > g.m_tiles) sm = g.m_tiles;
    g.splits_m = sm;
    CUtensorMap t0, t1{};
    int rc = ecsy_tensor_map_bf16_nhwc(gy_hi, imgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &t0);
    if (rc) return rc;
    if (gy_lo) {
      rc = ecsy_tensor_map_bf16_nhwc(gy_lo, imgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &t1);
      if (rc) return rc;
    }
    const int smem = 1024 + 2 * gsplit_t * (NB / 64) * kATileBytes + kWgStages * kATileBytes + (int)sizeof(WgCtl) + 64;
    dim3 grid((unsigned)sm, co_tiles, groups);
#define ECSY_WGT_LAUNCH(GS, NBV)                                                                                       \
    {                                                                                                                  \
      static bool done = false;                                                                                        \
      if (!done) {                                                                                                     \
        ECSY_CUDA(cudaFuncSetAttribute(k_umma_wgrad_t<GS, NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); \
        done = true;                                                                                                   \
      }                                                                                                                \
      k_umma_wgrad_t<GS, NBV><<<grid, kSpikeThreads, smem, st>>>(t0, t1, g, sg);                                       \
    }
    if (gsplit_t == 1 && NB == 64) ECSY_WGT_LAUNCH(1, 64)
    else if (gsplit_t == 1) ECSY_WGT_LAUNCH(1, 128)
    else if (NB == 64) ECSY_WGT_LAUNCH(2, 64)
    else ECSY_WGT_LAUNCH(2, 128)
#undef ECSY_WGT_LAUNCH
    ECSY_LAUNCH_CHECK();
    return ECSY_OK;
  }
  WgArgs g{};
  g.m_tiles = ((imgs + sg.tn_b - 1) / sg.tn_b) * sg.tiles_h * sg.tiles_w;
  const int nblk = k * k * sg.nslab;
  int bpc = 8;
  while (nblk % bpc != 0) --bpc;
  g.bpc = bpc; g.Cout = Cout; g.K = k * k * Cin; g.dw = dw;
  const int co_tiles = (Cout + 127) / 128, groups = nblk / bpc;
  int sm = (2 * ecsy_num_sms()) / (co_tiles * groups);
  if (sm < 1) sm = 1;
  if (sm > g.m_tiles) sm = g.m_tiles;
  g.splits_m = sm;
  CUtensorMap t0, t1{};
  int rc = ecsy_tensor_map_bf16_nhwc(gy_hi, imgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &t0);
  if (rc) return rc;
  const int gsplit = gy_lo ? 2 : 1;
  if (gy_lo) {
    rc = ecsy_tensor_map_bf16_nhwc(gy_lo, imgs, Ho, Wo, Cout, sg.tn_b, sg.th_b, sg.tw_b, &t1);
    if (rc) return rc;
  }
  const int smem = 1024 + 2 * gsplit * 2 * kATileBytes + kWgStages * kATileBytes + (int)sizeof(WgCtl) + 64;
  dim3 grid((unsigned)sm, co_tiles, groups);
  if (gsplit == 1) {
    static bool d1 = false;
    if (!d1) { ECSY_CUDA(cudaFuncSetAttribute(k_umma_wgrad<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); d1 = true; }
    k_umma_wgrad<1><<<grid, kSpikeThreads, smem, st>>>(t0, t1, g, sg);
  } else {
    static bool d2 = false;
    if (!d2) { ECSY_CUDA(cudaFuncSetAttribute(k_umma_wgrad<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)); d2 = true; }
    k_umma_wgrad<2><<<grid, kSpikeThreads, smem, st>>>(t0, t1, g, sg);
  }
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}

// ECS spread of one timestep, fused: out[M][C] = dw3x3(spikes) * Wpw^T  (dw computed by the producer warps).
int ecsy_umma_dw_gemm(const uint32_t* bits, const float* dw_w, const float* dw_b, const void* pw_packed, int splits,
                      float* out, int out_half, int N, int H, int W, int C, cudaStream_t st) {
  ECSY_CHECK_ARG(C % 64 == 0, "dw_gemm: C=%d must be a multiple of 64", C);
  const int BN = ecsy_pick_bn(C, splits);
  const int64_t M = (int64_t)N * H * W;
  CUtensorMap tb, dummy{};
  int rc = ecsy_tensor_map_bf16(pw_packed, (uint64_t)splits * C, (uint64_t)C, (uint32_t)BN, &tb);
  if (rc) return rc;
  GemmArgs g{};
  g.m_tiles = (int)((M + 127) / 128);
  g.n_tiles = C / BN;
  g.kb_total = C / 64;
  g.M = M;
  g.dw_w = dw_w; g.dw_b = dw_b;
  SpikeGeom sg{};
  sg.bits = bits; sg.imgs = N; sg.H = H; sg.W = W; sg.Cw = C / 32;
  EpiConv e{out, nullptr, nullptr, nullptr, M, C, out_half};
  if (splits == 1) return launch_bn<kADw, 1, 1, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, 0, st);
  return launch_bn<kADw, 2, 2, kEpiConv>(BN, dummy, dummy, tb, g, sg, e, 0, st);
}
