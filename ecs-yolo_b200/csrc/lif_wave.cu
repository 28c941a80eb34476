// ECS-LIF forward (mem_update, models/common.py:236-309) for C = 64 channels as ONE kernel in which membrane
// potential and ECS trace of all T timesteps stay on the SM -- the time loop runs as a WAVEFRONT down the image.
//
// Why a wavefront.  s_t at a pixel needs f_{t-1} there, which needs the 3x3 spread of s_{t-1}: one image row of
// dependency per timestep.  A CTA owns a vertical BAND of one image (region width Wb = 64 or 32 pixels including a
// halo of T-1 columns towards interior band edges, one zero column at image borders) and walks down it in BLOCKS
// of R = 128 / Wb image rows = 128 pixels = one tensor-core M tile.  Timestep t works on block j while timestep
// t-1 works on block j+1, so only the 2(T-1) blocks "in flight" between the first and the last timestep hold state.
// Every input element x_t is read from HBM exactly once (x the horizontal halo, 160-wide images: 1.12), nothing but
// the bit-packed spikes is written: 4 B + 1/8 B per element-step against 16-24 B for the per-timestep pipeline (lif.cu).
//
// Work units.  The bands of all images form one STREAM of blocks, bands separated by one all-zero "gap" block (top /
// bottom zero padding of the 3x3 spread for free).  The stream is cut into gridDim.x equal contiguous segments; a
// CTA additionally processes ceil((T-1)/R) warm-up blocks before and T-1 cool-down blocks after its segment (their
// results are discarded), so there is no tail effect and no per-band pipeline drain.
//
// Roles (T + 1 warp groups).  Warp group t (128 threads, t = 0..T-1) performs timestep t of every block, one block
// after the other; a last warp issues the tensor-core work.  Level t, block j:
//     E     = D + bconst              D: accumulated spread convolution, read from TENSOR MEMORY
//     f     = beta * tanh(alpha * E)  (models/common.py:263-267)
//     m_t   = (m_{t-1} * decay) * (1 - s_{t-1}) + x_t + f          (reference order, :306-309; m_0 = x_0)
//     s_t   = m_t > thresh            -> 8 bytes per pixel to HBM, and as a {0,1} bf16 row into the level's smem ring
//     kappa * E and m_t go back to tensor memory (levels < T-1).
// Spread = the depth-wise 3x3 and point-wise 1x1 of `spread` folded into one 3x3 spike convolution W_eff[co][ci][tap] =
// pw[co][ci] * dw[ci][tap] (bf16, all nine 64x64 tap tiles resident in smem, 72 KB); the nine taps of a block are
// nine UMMA descriptors whose start row is shifted by ky * Wb + kx rows of the ring (rows are written with the 128-byte
// swizzle of their ABSOLUTE address, so a descriptor may start at any row).  The tensor core accumulates S(s_t) onto
// kappa * E_{t-1}, which already sits in the accumulator: e_t = alpha * (S_t + bconst + kappa * E_{t-1}).
// The taps of block j split in two groups: ky in {-1, 0} need rows of blocks j-1, j and are issued as soon as level
// t finishes block j; ky = +1 needs the first row of block j+1 and follows one block later.  Only those 12 MMAs sit
// between "level t finished block j+1" and "level t+1 may start block j" -- the critical path of the wavefront.
//
// Thread <-> data.  Tensor-memory accesses use the 16x256b shape (thread i of a warp: rows i/4 and i/4 + 8 of a
// 16-lane half, columns 8k + 2(i%4) + {0,1}): a warp's LDG.64 of x then covers 8 pixels x 32 contiguous bytes -- every
// 32-byte sector it touches is fully used (the 32x32b shape, one pixel per thread, touches 32 half-used sectors per
// load and made the first fused kernel, lif_fused.cu, LSU-bound).  The K order of W_eff is permuted on the host to the
// channel order in which a thread holds its 16 channels of a pixel, so a spike row is written with two 16-byte stores
// per thread and 32-column chunk, bank-conflict free.
//
// Tensor memory (512 columns): E of 5 blocks (64 columns each) + membrane of 3 blocks.  m_0 = x_0 is not stored: level
// 1 re-reads x_0 (an L2 hit, the same CTA read it one block earlier).  Ring per level: two block slots + one mirrored
// image row at either end (ky = -1 of the block in the first slot, ky = +1 of the block in the second), 50 KB at Wb = 64.
//
// Fast precision only (one bf16 plane of W_eff; spikes are exact in bf16, accumulation fp32); 2 <= T <= 4.
#include <stdlib.h>

#include "ecsy_common.cuh"
#include "../../include/ecsy.h"

using namespace ecsy;

namespace {

constexpr int kWvC = 64;
constexpr int kWvMaxT = 4;
constexpr int kWvESlots = 5;
constexpr int kWvMSlots = 3;
constexpr int kWvMCol0 = kWvESlots * 64;        // membrane slots start at column 320
constexpr int kWvWBytes = 9 * 64 * 128;         // nine [64 co x 64 kk] bf16 tap tiles
constexpr int kWvFullDepth = 8;                 // `full` barriers per level: a level may run up to kWvESlots blocks ahead of the next

struct WvCtl {
  uint64_t ready[kWvMaxT];                  // level t finished a block: spike rows + tensor-memory stores visible (4 warps)
  uint64_t ring_free[kWvMaxT];              // every MMA that reads the ring rows level t is about to overwrite has completed
  uint64_t full[kWvMaxT][kWvFullDepth];     // every MMA into E(block j) for level t has completed (barrier j % 8)
  uint64_t e_free[kWvESlots];               // the last level has read E of the block that held this slot
  uint64_t m_free[kWvMSlots];
  uint32_t tmem_base;
  uint32_t pad;
};

struct WvArgs {
  const float* x;
  int64_t x_tstride;
  const float* in_scale;
  const float* in_shift;
  const uint16_t* w_eff;     // [9][64][64] bf16, K-permuted (see ecsy_lif_wave_pack)
  const float* bconst;       // [64] pw * b_dw + b_pw
  uint32_t* spikes;          // [T][N][H][W][2]
  int T, N, H, W;
  int R, Wb, logWb;          // block = R image rows x Wb region columns
  int nb, hb;                // bands per image, blocks per band (+1 gap entry in the stream)
  int64_t S;                 // stream length = N * nb * (hb + 1)
  int warm, cool;
  float thresh, decay, alpha, beta, kappa;
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// 16 lanes x 32 columns: thread i <-> rows i/4 (+8), columns 8k + 2(i%4) + {0,1}; v[4k + 2r + e].
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float wv_tanh(float v) {   // absolute error ~1e-7 (see elementwise.cu: tanh_fast)
  const float y = __expf(2.f * v);
  return 1.f - __fdividef(2.f, y + 1.f);
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// Ring geometry (rows of 128 bytes): [8 pad][guard-top: Wb][slot A: 128][slot B: 128][guard-bottom: Wb][8 pad]
__host__ __device__ constexpr int wv_ring_rows(int Wb) { return 8 + Wb + 256 + Wb + 8; }

struct BlockGeom {   // one stream entry
  bool gap;
  int n, row0, rx0, ox0, ox1;
};

__device__ __forceinline__ BlockGeom wv_locate(const WvArgs& g, int64_t p) {
  BlockGeom b;
  const int per = g.hb + 1;
  b.gap = p < 0 || p >= g.S;
  int64_t band = b.gap ? 0 : p / per;
  const int r = b.gap ? 0 : (int)(p - band * per);
  if (r == g.hb) b.gap = true;
  b.n = (int)(band / g.nb);
  const int bi = (int)(band - (int64_t)b.n * g.nb);
  b.row0 = r * g.R;
  b.ox0 = (int)(((int64_t)g.W * bi) / g.nb);
  b.ox1 = (int)(((int64_t)g.W * (bi + 1)) / g.nb);
  b.rx0 = b.ox0 - (bi == 0 ? 1 : g.T - 1);
  return b;
}

struct LevelCtx {
  WvCtl* ctl;
  uint32_t tmem_base, rings_u32, c_bconst, c_scale, c_shift;
  int ring_bytes, P, lvl;
  int64_t p_begin, seg0, seg1;
};

// =============================== one level: timestep `lvl` of every block ===============================
// KIND 0: level 0 (m_0 = x_0, no tensor memory); 1: level 1 (m_0 re-read from x_0); 2: levels >= 2 (membrane from
// tensor memory).  LAST: level T-1 (nothing handed on: no ring rows, no tensor-memory stores).
template <int KIND, bool LAST>
__device__ __forceinline__ void wv_level(const WvArgs& g, const LevelCtx& cx) {
  WvCtl* ctl = cx.ctl;
  const uint32_t tmem_base = cx.tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = g.T, lvl = cx.lvl, P = cx.P;
  const int64_t p_begin = cx.p_begin, seg0 = cx.seg0, seg1 = cx.seg1;
  const bool affine = g.in_scale != nullptr;
  const int SA = 8 + g.Wb, SB = SA + 128, GT = 8, GB = SB + 128;
  constexpr bool last = LAST;
  {
    const int q = warp & 3;                   // tensor-memory lane quarter (hardware: warp id % 4)
    const int q4 = lane & 3;
    const uint32_t ring = cx.rings_u32 + (uint32_t)lvl * (uint32_t)cx.ring_bytes;          // written by levels < T-1
    const uint32_t c_bconst = cx.c_bconst + (uint32_t)q4 * 8u;
    const uint32_t c_scale = cx.c_scale + (uint32_t)q4 * 8u;
    const uint32_t c_shift = cx.c_shift + (uint32_t)q4 * 8u;
    const int64_t bits_tstride = (int64_t)g.N * g.H * g.W * 2;
    const float* xt = g.x + (int64_t)lvl * g.x_tstride;
    const int nblk = P - lvl;                 // level lvl processes local blocks 0 .. P-1-lvl

    for (int j = 0; j < nblk; ++j) {
      const int64_t p = p_begin + j;
      const BlockGeom bg = wv_locate(g, p);
      const bool seg_out = p >= seg0 && p < seg1;

      // ---- this thread's four pixels: (half h, row-select rs) -> block row a = 32q + 16h + 8rs + lane/4 ----
      int pix[4];
      uint32_t flags = 0;       // bit i: inside the image, bit 4+i: result is written to HBM
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = 32 * q + 16 * (i >> 1) + 8 * (i & 1) + (lane >> 2);
        const int ry = a >> g.logWb, rx = a & (g.Wb - 1);
        const int y = bg.row0 + ry, xx = bg.rx0 + rx;
        const bool inside = !bg.gap && y < g.H && xx >= 0 && xx < g.W;
        pix[i] = inside ? (bg.n * g.H + y) * g.W + xx : 0;
        flags |= (inside ? 1u : 0u) << i;
        flags |= ((inside && seg_out && xx >= bg.ox0 && xx < bg.ox1) ? 1u : 0u) << (4 + i);
      }

      float xa[2][16];          // input current of the chunk being computed / the next chunk (software pipeline)
      float x0a[2][16];         // level 1 only: x_0 = m_0
      auto load_chunk = [&](int u, float (&xv)[16], float (&x0v)[16]) {
        const int h = u >> 1, gq = u & 1;
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const int pi = 2 * h + rs;
          const bool ok = (flags >> pi) & 1u;
          const float* src = xt + (int64_t)pix[pi] * 64 + 32 * gq + 2 * q4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 v = ok ? ldg_stream2(src + 8 * k) : make_float2(0.f, 0.f);
            xv[4 * k + 2 * rs] = v.x;
            xv[4 * k + 2 * rs + 1] = v.y;
            if (KIND == 1) {
              float2 v0 = ok ? ldg_stream2(src - g.x_tstride + 8 * k) : make_float2(0.f, 0.f);
              x0v[4 * k + 2 * rs] = v0.x;
              x0v[4 * k + 2 * rs + 1] = v0.y;
            }
          }
        }
      };
      load_chunk(0, xa[0], x0a[0]);

      // ---- wait for this block's spread accumulator, the ring rows and the tensor-memory slots we are about to write ----
      if (KIND > 0) {
        mbar_wait(&ctl->full[lvl][j & (kWvFullDepth - 1)], (uint32_t)(j / kWvFullDepth) & 1u);
        if (KIND == 1 && !LAST && j >= kWvMSlots) mbar_wait(&ctl->m_free[j % kWvMSlots], (uint32_t)(j / kWvMSlots - 1) & 1u);
        tc_fence_after_sync();
      }
      if (!last && j >= 1) mbar_wait(&ctl->ring_free[lvl], (uint32_t)(j - 1) & 1u);

      const uint32_t e_col = (uint32_t)(j % kWvESlots) * 64u;
      const uint32_t m_col = (uint32_t)kWvMCol0 + (uint32_t)(j % kWvMSlots) * 64u;
      const uint32_t slot_row = (uint32_t)((j & 1) ? SB : SA);
      uint32_t wlo[4] = {0u, 0u, 0u, 0u}, whi[4] = {0u, 0u, 0u, 0u};   // HBM spike words (channels 0-31 / 32-63) per pixel

#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int h = u >> 1, gq = u & 1;
        if (u + 1 < 4) load_chunk(u + 1, xa[(u + 1) & 1], x0a[(u + 1) & 1]);
        float (&xv)[16] = xa[u & 1];
        float (&x0v)[16] = x0a[u & 1];
        const uint32_t t_addr = tmem_base + ((uint32_t)(32 * q + 16 * h) << 16);
        uint32_t ev[16], mv[16];
        if (KIND > 0) {
          tmem_ld_16x256b_x4(t_addr + e_col + 32u * gq, ev);
          if (KIND > 1) tmem_ld_16x256b_x4(t_addr + m_col + 32u * gq, mv);
          tmem_ld_wait();
        }
        uint32_t sbits = 0;      // bit 4k + 2rs + e of this chunk
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 32 * gq + 8 * k;     // + 2*q4 + e
          float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f), bc = make_float2(0.f, 0.f);
          if (affine) { sc = lds64f(c_scale + c * 4); sh = lds64f(c_shift + c * 4); }
          if (KIND > 0) bc = lds64f(c_bconst + c * 4);
#pragma unroll
          for (int rs = 0; rs < 2; ++rs) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int idx = 4 * k + 2 * rs + e;
              float xin = xv[idx];
              if (affine) xin = add_rn(mul_rn(xin, e ? sc.y : sc.x), e ? sh.y : sh.x);
              float m;
              if (KIND == 0) {
                m = xin;
              } else {
                float mo;
                if (KIND == 1) {
                  mo = x0v[idx];
                  if (affine) mo = add_rn(mul_rn(mo, e ? sc.y : sc.x), e ? sh.y : sh.x);
                } else {
                  mo = __uint_as_float(mv[idx]);
                }
                const float E = add_rn(__uint_as_float(ev[idx]), e ? bc.y : bc.x);
                const float f = mul_rn(g.beta, wv_tanh(mul_rn(g.alpha, E)));
                const float keep = mo > g.thresh ? 0.f : 1.f;
                m = add_rn(add_rn(mul_rn(mul_rn(mo, g.decay), keep), xin), f);
                ev[idx] = __float_as_uint(mul_rn(g.kappa, E));
              }
              mv[idx] = __float_as_uint(m);
              sbits |= (m > g.thresh ? 1u : 0u) << idx;
            }
          }
        }
        if (!last) {
          if (KIND > 0) {
            tmem_st_16x256b_x4(t_addr + e_col + 32u * gq, ev);
            tmem_st_16x256b_x4(t_addr + m_col + 32u * gq, mv);
          }
        }
        // per pixel: zero outside the image (zero padding of the spread), bf16 row chunk, HBM word
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const int pi = 2 * h + rs;
          const bool inside = (flags >> pi) & 1u;
          uint32_t w4[4];
          uint32_t word = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t b0 = inside ? (sbits >> (4 * k + 2 * rs)) & 1u : 0u;
            const uint32_t b1 = inside ? (sbits >> (4 * k + 2 * rs + 1)) & 1u : 0u;
            w4[k] = b0 * 0x3F80u + b1 * 0x3F800000u;
            word |= (b0 | (b1 << 1)) << (8 * k);
          }
          word <<= 2 * q4;
          if (gq == 0) wlo[pi] = word; else whi[pi] = word;
          if (!last) {
            const int a = 32 * q + 16 * h + 8 * rs + (lane >> 2);
            const uint32_t chunk = (uint32_t)(2 * q4 + gq);
            const uint32_t row = slot_row + (uint32_t)a;
            sts128(ring + row * 128u + ((chunk ^ (row & 7u)) << 4), w4[0], w4[1], w4[2], w4[3]);
            const int ry = a >> g.logWb, rx = a & (g.Wb - 1);
            // mirrors: first image row of a block in slot A -> guard-bottom, last image row of a block in slot B -> guard-top
            if (!(j & 1) && ry == 0) {
              const uint32_t mr = (uint32_t)(GB + rx);
              sts128(ring + mr * 128u + ((chunk ^ (mr & 7u)) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
            if ((j & 1) && ry == g.R - 1) {
              const uint32_t mr = (uint32_t)(GT + rx);
              sts128(ring + mr * 128u + ((chunk ^ (mr & 7u)) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
          }
        }
      }
      // ---- HBM spikes: combine the four threads of a pixel, lane q4 writes pixel q4 ----
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        wlo[i] |= __shfl_xor_sync(0xffffffffu, wlo[i], 1);
        wlo[i] |= __shfl_xor_sync(0xffffffffu, wlo[i], 2);
        whi[i] |= __shfl_xor_sync(0xffffffffu, whi[i], 1);
        whi[i] |= __shfl_xor_sync(0xffffffffu, whi[i], 2);
      }
      {
        uint32_t lo = wlo[0], hi = whi[0];
        int pp = pix[0];
        bool ok = (flags >> 4) & 1u;
#pragma unroll
        for (int i = 1; i < 4; ++i)
          if (q4 == i) { lo = wlo[i]; hi = whi[i]; pp = pix[i]; ok = (flags >> (4 + i)) & 1u; }
        if (ok) {
          uint2 v = make_uint2(lo, hi);
          *reinterpret_cast<uint2*>(g.spikes + (int64_t)lvl * bits_tstride + (int64_t)pp * 2) = v;
        }
      }
      // ---- hand over ----
      if (!last) {
        if (KIND > 0) tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->ready[lvl]);
      } else {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctl->e_free[j % kWvESlots]);
          if (T > 2) mbar_arrive(&ctl->m_free[j % kWvMSlots]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(32 * (4 * kWvMaxT + 1), 1)
k_lif_ecs_wave64(const WvArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int T = g.T;
  const int ring_bytes = wv_ring_rows(g.Wb) * 128;
  uint8_t* w_smem = smem;
  uint8_t* rings = smem + kWvWBytes;                                   // (T-1) rings, 1024-byte aligned
  float* s_bconst = reinterpret_cast<float*>(rings + (T - 1) * ring_bytes);
  float* s_scale = s_bconst + 64;
  float* s_shift = s_scale + 64;
  WvCtl* ctl = reinterpret_cast<WvCtl*>(s_shift + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthreads = blockDim.x;
  const bool affine = g.in_scale != nullptr;

  // ---- this CTA's segment of the stream ----
  const int64_t seg0 = (g.S * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t seg1 = (g.S * (int64_t)(blockIdx.x + 1)) / gridDim.x;
  int64_t p_begin = seg0 - g.warm;
  if (p_begin < 0) p_begin = 0;
  const int P = (int)(seg1 + g.cool - p_begin);      // local blocks 0..P-1 (positions >= S are gaps)

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < (T - 1) * ring_bytes / 16; i += nthreads)
    reinterpret_cast<uint4*>(rings)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 9 * 64 * 8; i += nthreads) {         // W_eff -> swizzled K-major tiles
    const int chunk = i & 7, row = (i >> 3) & 63, tap = i >> 9;
    const uint4 v = reinterpret_cast<const uint4*>(g.w_eff)[i];
    *reinterpret_cast<uint4*>(w_smem + tap * 8192 + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
  }
  if (threadIdx.x < 64) {
    s_bconst[threadIdx.x] = g.bconst[threadIdx.x];
    s_scale[threadIdx.x] = affine ? g.in_scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = affine ? g.in_shift[threadIdx.x] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int t = 0; t < kWvMaxT; ++t) {
      mbar_init(&ctl->ready[t], 4);
      mbar_init(&ctl->ring_free[t], 1);
      for (int k = 0; k < kWvFullDepth; ++k) mbar_init(&ctl->full[t][k], 1);
    }
    for (int k = 0; k < kWvESlots; ++k) mbar_init(&ctl->e_free[k], 4);
    for (int k = 0; k < kWvMSlots; ++k) mbar_init(&ctl->m_free[k], 4);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&ctl->tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t rings_u32 = smem_u32(rings);
  const int SA = 8 + g.Wb, SB = SA + 128;

  if (warp == 4 * T) {
    // =============================== MMA issuer (event driven) ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint64_t db0 = umma_desc_sw128(smem_u32(w_smem));
    int nj[kWvMaxT] = {0, 0, 0, 0};
    int remaining = 0;
    for (int t = 0; t + 1 < T; ++t) remaining += (P - t > 0 ? P - t : 0);
    long long t_last = clock64();
    while (remaining > 0) {
      bool progress = false;
      for (int t = 0; t + 1 < T; ++t) {
        const int j = nj[t];
        if (j > P - 1 - t) continue;
        if (!mbar_test(&ctl->ready[t], (uint32_t)j & 1u)) continue;
        const bool has_p1 = j <= P - 2 - t;              // level t+1 processes blocks 0 .. P-2-t
        if (t == 0 && has_p1 && j >= kWvESlots &&
            !mbar_test(&ctl->e_free[j % kWvESlots], (uint32_t)(j / kWvESlots - 1) & 1u))
          continue;
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t ring = rings_u32 + (uint32_t)t * (uint32_t)ring_bytes;
          if (j >= 1) {
            // taps ky = +1 of block j-1 (level t+1): rows of block j-1 and the first image row of block j
            const int jb = j - 1;
            const uint32_t d_tmem = tmem_base + (uint32_t)(jb % kWvESlots) * 64u;
            const uint32_t base_row = (uint32_t)((jb & 1) ? SB : SA);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t da = umma_desc_sw128(ring + (base_row + (uint32_t)g.Wb + (uint32_t)kx - 1u) * 128u);
              const uint64_t db = db0 + (uint64_t)((6 + kx) * 512);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, 1u);
            }
            umma_commit(&ctl->full[t + 1][jb & (kWvFullDepth - 1)]);
          }
          if (has_p1) {
            // taps ky = -1, 0 of block j (level t+1): last image row of block j-1 (or the mirrored guard) and block j
            const uint32_t d_tmem = tmem_base + (uint32_t)(j % kWvESlots) * 64u;
            const uint32_t base_row = (uint32_t)((j & 1) ? SB : SA);
#pragma unroll
            for (int tap = 0; tap < 6; ++tap) {
              const int ky = tap / 3 - 1, kx = tap % 3 - 1;
              const uint64_t da = umma_desc_sw128(ring + (uint32_t)((int)base_row + ky * g.Wb + kx) * 128u);
              const uint64_t db = db0 + (uint64_t)(tap * 512);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc, (t > 0 || tap > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&ctl->ring_free[t]);
        }
        __syncwarp();
        ++nj[t];
        --remaining;
        progress = true;
      }
      if (progress) {
        t_last = clock64();
      } else if (clock64() - t_last > 4000000000LL) {
        if (lane == 0) printf("ecsy: lif_wave issuer stalled, block %d (nj %d %d %d, P %d)\n", blockIdx.x, nj[0], nj[1], nj[2], P);
        __trap();
      }
    }
  } else if (warp < 4 * T) {
    const int lvl = warp >> 2;
    const bool last = lvl == T - 1;
    LevelCtx c;
    c.ctl = ctl; c.tmem_base = tmem_base; c.rings_u32 = rings_u32; c.ring_bytes = ring_bytes;
    c.c_bconst = smem_u32(s_bconst); c.c_scale = smem_u32(s_scale); c.c_shift = smem_u32(s_shift);
    c.p_begin = p_begin; c.seg0 = seg0; c.seg1 = seg1; c.P = P; c.lvl = lvl;
    if (lvl == 0) wv_level<0, false>(g, c);
    else if (lvl == 1) { if (last) wv_level<1, true>(g, c); else wv_level<1, false>(g, c); }
    else { if (last) wv_level<2, true>(g, c); else wv_level<2, false>(g, c); }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// Band plan for an image of width W: R rows x Wb columns per block, nb bands.  Returns 0 when no plan fits.
static int wv_plan(int T, int W, int* R, int* Wb, int* nb) {
  double best = 0.0;
  int ok = 0;
  for (int wb = 64; wb >= 32; wb >>= 1) {
    for (int n = 1; n <= 64; ++n) {
      // widest region over the bands of an equal split: output width + halo (T-1 towards a neighbour band, 1 at the image border)
      int need = 0;
      for (int b = 0; b < n; ++b) {
        const int o0 = (int)(((long long)W * b) / n), o1 = (int)(((long long)W * (b + 1)) / n);
        const int w = (o1 - o0) + (b == 0 ? 1 : T - 1) + (b == n - 1 ? 1 : T - 1);
        if (w > need) need = w;
      }
      if (need <= wb) {
        const double eff = (double)W / ((double)n * wb);
        if (eff > best + 1e-9) { best = eff; *R = 128 / wb; *Wb = wb; *nb = n; ok = 1; }
        break;
      }
    }
  }
  return ok;
}

}  // namespace

extern "C" int ecsy_lif_ecs_wave_supported(int T, int C, int H, int W) {
  int R, Wb, nb;
  return (C == kWvC && T >= 2 && T <= kWvMaxT && H >= 1 && wv_plan(T, W, &R, &Wb, &nb)) ? 1 : 0;
}

extern "C" int ecsy_lif_ecs_wave_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                                     const void* w_eff, const float* bconst, uint32_t* spikes, int T, int64_t N, int H,
                                     int W, int C, float thresh, float decay, float alpha, float beta, float kappa,
                                     void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(x && spikes && w_eff && bconst && N > 0 && H > 0 && W > 0, "lif_ecs_wave_fwd: bad arguments");
  ECSY_CHECK_ARG(ecsy_lif_ecs_wave_supported(T, C, H, W), "lif_ecs_wave_fwd: unsupported T=%d / C=%d / W=%d (C == 64, 2 <= T <= 4)",
                 T, C, W);
  ECSY_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "lif_ecs_wave_fwd: scale/shift pair");
  ECSY_CHECK_ARG(N * H * W < (1LL << 31), "lif_ecs_wave_fwd: tensor too large");
  WvArgs g{};
  g.x = x; g.x_tstride = x_tstride; g.in_scale = in_scale; g.in_shift = in_shift;
  g.w_eff = reinterpret_cast<const uint16_t*>(w_eff); g.bconst = bconst; g.spikes = spikes;
  g.T = T; g.N = (int)N; g.H = H; g.W = W;
  wv_plan(T, W, &g.R, &g.Wb, &g.nb);
  g.logWb = g.Wb == 64 ? 6 : 5;
  g.hb = (H + g.R - 1) / g.R;
  g.S = (int64_t)N * g.nb * (g.hb + 1);
  g.warm = (T - 1 + g.R - 1) / g.R;
  g.cool = T - 1;
  g.thresh = thresh; g.decay = decay; g.alpha = alpha; g.beta = beta; g.kappa = kappa;
  const int ring_bytes = wv_ring_rows(g.Wb) * 128;
  const int smem = 1024 + kWvWBytes + (T - 1) * ring_bytes + 3 * 64 * 4 + (int)sizeof(WvCtl) + 64;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ECSY_CUDA(cudaFuncSetAttribute(k_lif_ecs_wave64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  // every CTA should own at least ~8 blocks of useful work next to its warm-up / cool-down blocks
  int64_t grid = ecsy_num_sms();
  if (g.S / 8 < grid) grid = g.S / 8 > 0 ? g.S / 8 : 1;
  k_lif_ecs_wave64<<<(int)grid, 32 * (4 * T + 1), smem, st>>>(g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
