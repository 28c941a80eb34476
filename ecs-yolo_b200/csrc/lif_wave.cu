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
constexpr int kWvESlots = 8;                    // all 512 tensor-memory columns hold spread accumulators
constexpr int kWvMSlots = 6;                    // membrane slots: CTA-private global scratch (stays in L2)
constexpr int kWvMSlotFloats = 128 * 64;         // one block of membranes, stored in the writing thread's own order
constexpr int kWvWBytes = 9 * 64 * 128;         // nine [64 co x 64 kk] bf16 tap tiles
constexpr int kWvThreads = 608;                 // 16 level warps + 3 issuer warps (a CTA is allocated in units of 4 warps: 96 registers)
constexpr int kWvIssuerWarp = 16;
constexpr int kWvFullDepth = 8;                 // `full` barriers per level: a level may run up to kWvESlots blocks ahead of the next

struct WvCtl {
  uint64_t ready[kWvMaxT];                  // level t finished a block: spike rows + tensor-memory stores visible (4 warps)
  uint64_t ring_free[kWvMaxT];              // every MMA that reads the ring rows level t is about to overwrite has completed
  uint64_t full[kWvMaxT][kWvFullDepth];     // every MMA into E(block j) for level t has completed (barrier j % 8)
  uint64_t e_free[kWvESlots];               // the last level has read E of the block that held this slot
  uint64_t m_free[kWvMSlots];               // the last level has read the membranes of the block that held this slot
  uint64_t m_ready[kWvMSlots];              // a level has written its membranes of the block in this slot (T-2 phases per use)
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
  float* mscr;               // [gridDim.x][kWvMSlots][128 * 64] membrane scratch
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

// Blocking wait without a clock read in the loop (CS2R shares the XU pipe with the MUFU ops of the working warps): the
// hardware suspends the warp up to the hint and wakes it on the arrival; a pipeline bug traps instead of hanging the box.
__device__ __forceinline__ void wv_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  printf("ecsy: lif_wave barrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
  __trap();
}
__device__ __forceinline__ uint4 ld_cg_u4(const float* p) {     // L2-coherent (another warp group of this CTA wrote it)
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_cg_u4(float* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
// 16 lanes x 16 columns: v[4k + 2r + e], k = 0, 1.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// beta * tanh(alpha E) = beta - 2 beta / (2^(2 log2(e) alpha E) + 1): ex2.approx + rcp.approx, absolute error ~1e-7 (see
// elementwise.cu: tanh_fast); 2^t -> inf gives beta, 2^t -> 0 gives -beta.
__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float exp2f_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
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

#ifdef WV_PROF
#define WV_DECL long long prof_full = 0, prof_ring = 0, prof_work = 0, prof_pre = 0, pc_issue = 0, pc_tmem = 0, pc_math = 0, pc_store = 0, pc_rot = 0
#define WV_T(v) const long long v = clock64()
#define WV_ACC(a, d) a += (d)
#define WV_REPORT                                                                                                        \
  if (blockIdx.x == 1 && (threadIdx.x & 127) == 0)                                                                       \
    printf("wave prof: level %d blocks %d  per block: pre %lld  wait_full %lld  wait_ring %lld  work %lld | per chunk: issue %lld tmem %lld math %lld store %lld rotate %lld\n", lvl, nblk, \
           prof_pre / nblk, prof_full / nblk, prof_ring / nblk, prof_work / nblk, pc_issue / (8 * nblk), pc_tmem / (8 * nblk), pc_math / (8 * nblk), pc_store / (8 * nblk), pc_rot / (8 * nblk))
#else
#define WV_DECL
#define WV_T(v)
#define WV_ACC(a, d)
#define WV_REPORT
#endif

struct LevelCtx {
  WvCtl* ctl;
  uint32_t tmem_base, rings_u32, c_bconst, c_scale, c_shift;
  int ring_bytes, P, lvl;
  int p_begin, seg0, seg1;
};

// =============================== one level: timestep `lvl` of every block ===============================
// KIND 0: level 0 (m_0 = x_0, no tensor memory); 1: level 1 (m_0 re-read from x_0); 2: levels >= 2 (membrane from
// tensor memory).  LAST: level T-1 (nothing handed on: no ring rows, no tensor-memory stores).  AFF: a tdBN affine is
// pending on the input current.
//
// A level runs on ONE warp per scheduler partition, so its block time is set by instruction count and dependency
// latency, not by memory: the loop below is written for few instructions per element (folded constants, one FSET per
// spike, spike rows and HBM words assembled with byte permutes) and no spills (with ~225 KB of shared memory in use the
// L1 is ~1 KB: a spilled register is an L2 round trip).
template <int KIND, bool LAST, bool AFF>
__device__ __forceinline__ void wv_level(const WvArgs& g, const LevelCtx& cx) {
  WvCtl* ctl = cx.ctl;
  const uint32_t tmem_base = cx.tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lvl = cx.lvl, P = cx.P;
  const int p_begin = cx.p_begin, seg0 = cx.seg0, seg1 = cx.seg1;
  const int SA = 8 + g.Wb, SB = SA + 128, GT = 8, GB = SB + 128;
  const int q = warp & 3;                   // tensor-memory lane quarter (hardware: warp id % 4)
  const int q4 = lane & 3;
  const uint32_t ring = cx.rings_u32 + (uint32_t)lvl * (uint32_t)cx.ring_bytes;          // written by levels < T-1
  const uint32_t c_bconst = cx.c_bconst + (uint32_t)q4 * 8u;
  const uint32_t c_scale = cx.c_scale + (uint32_t)q4 * 8u;
  const uint32_t c_shift = cx.c_shift + (uint32_t)q4 * 8u;
  const int64_t bits_tstride = (int64_t)g.N * g.H * g.W * 2;
  const float* xt = g.x + (int64_t)lvl * g.x_tstride + 2 * q4;
  const float* x0p = g.x + 2 * q4;
  const int nblk = P - lvl;                 // level lvl processes local blocks 0 .. P-1-lvl
  // this thread's four pixels i = 2h + rs: block row a_i = 32q + 16h + 8rs + lane/4 -> (ry_i, rx_i) inside the block
  const int a0 = 32 * q + (lane >> 2);
  const float c1 = 2.f * 1.4426950408889634f * g.alpha;      // tanh(alpha E) = 1 - 2 / (2^(c1 E) + 1)
  const float nb2 = -2.f * g.beta;

  // ---- stream position -> band geometry, kept incrementally (three divisions only when a new band starts) ----
  int r_in_band, band_pix0, rx0, ox0, ox1;      // block index in the band; pixel index of (row 0, region column 0); columns
  bool band_valid;
  auto set_band = [&](int p) {
    const uint32_t per = (uint32_t)(g.hb + 1);
    band_valid = p >= 0 && p < (int)g.S;
    const uint32_t up = band_valid ? (uint32_t)p : 0u;
    const uint32_t band = up / per;
    r_in_band = (int)(up - band * per);
    const int n = (int)(band / (uint32_t)g.nb);
    const int bi = (int)(band - (uint32_t)n * (uint32_t)g.nb);
    ox0 = (int)(((uint32_t)g.W * (uint32_t)bi) / (uint32_t)g.nb);
    ox1 = (int)(((uint32_t)g.W * (uint32_t)(bi + 1)) / (uint32_t)g.nb);
    rx0 = ox0 - (bi == 0 ? 1 : g.T - 1);
    band_pix0 = n * g.H * g.W + rx0;
  };
  set_band(p_begin);
  // pixel index + flags (bit i: inside the image, bit 4+i: written to HBM) of the four pixels in the CURRENT block
  auto locate = [&](int p, int (&pix)[4]) -> uint32_t {
    const bool gap = !band_valid || r_in_band == g.hb;
    const bool seg_out = p >= seg0 && p < seg1;
    const int row0 = r_in_band * g.R;
    uint32_t flags = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = a0 + 16 * (i >> 1) + 8 * (i & 1);
      const int ry = a >> g.logWb, rx = a & (g.Wb - 1);
      const int y = row0 + ry, xx = rx0 + rx;
      const bool inside = !gap && y < g.H && xx >= 0 && xx < g.W;
      pix[i] = inside ? band_pix0 + y * g.W + rx : 0;
      flags |= (inside ? 1u : 0u) << i;
      flags |= ((inside && seg_out && xx >= ox0 && xx < ox1) ? 1u : 0u) << (4 + i);
    }
    return flags;
  };
  // 8 values of one chunk u = (h, gq, kh): pixels (h, rs = 0, 1), channels 16 (u & 3) + 8 k + 2 q4 + e (k, e = 0, 1).  A
  // pixel outside the image reads a huge negative current: it never fires (zero padding of the spread), its state stays
  // finite.  The chunk loop is NOT unrolled (four levels x eight unrolled chunks were 140 KB of hot code on one SM: a
  // quarter of all warp stalls were instruction-cache misses); buffers rotate through registers instead.
  // The kernel is bound by the load / store unit's WAVEFRONTS (one per 128-byte line an instruction touches, ~2 cycles
  // each; measured with in-kernel clocks: a level spent 900-1400 cycles per chunk just ISSUING its loads).  In the fragment
  // layout a thread owns channel pairs, so a warp's 8-byte loads touch 8 lines for 256 bytes.  Here a thread loads 16 bytes
  // (four channels of ONE of the chunk's two channel groups) and swaps one pair with its neighbour lane: 8 lines for 512
  // bytes, half the wavefronts.  Lane q4: group kq = q4 & 1, channels 4 (q4 >> 1) .. + 3 of it; it needs the pairs 2 q4,
  // 2 q4 + 1 of BOTH groups: even lanes keep (x, y) as their group-0 pair and get the group-1 pair from the odd neighbour,
  // odd lanes keep (z, w) as their group-1 pair and get the group-0 pair from the even neighbour.
  // load_x only REQUESTS the data (raw[4 rs + i] = element i of the lane's 16-byte slice of pixel rs); the pair swap
  // happens in unpack_x when the chunk is consumed -- a shuffle right after the load would wait for it.
  auto load_x = [&](int u, const int (&pix)[4], uint32_t flags, float (&raw)[8], const float* base) {
    const bool hi = u >= 4;
    const int cb = 16 * (u & 3);
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const int pp = hi ? pix[2 + rs] : pix[rs];
      const bool ok = (flags >> ((hi ? 2 : 0) + rs)) & 1u;
      // `base` carries + 2 q4; the 16-byte slice of this lane starts at 8 (q4 & 1) + 4 (q4 >> 1)
      const float* src = base - 2 * q4 + (int64_t)pp * 64 + cb + 8 * (q4 & 1) + 4 * (q4 >> 1);
      const float4 v = ok ? ldg_stream(reinterpret_cast<const float4*>(src)) : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
      raw[4 * rs] = v.x; raw[4 * rs + 1] = v.y; raw[4 * rs + 2] = v.z; raw[4 * rs + 3] = v.w;
    }
  };
  auto unpack_x = [&](const float (&raw)[8], float (&xv)[8]) {
    const int kq = q4 & 1;
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const float r0 = __shfl_xor_sync(0xffffffffu, kq ? raw[4 * rs] : raw[4 * rs + 2], 1);
      const float r1 = __shfl_xor_sync(0xffffffffu, kq ? raw[4 * rs + 1] : raw[4 * rs + 3], 1);
      xv[2 * rs] = kq ? r0 : raw[4 * rs];               // group k = 0: channels cb + 2 q4 + {0, 1}
      xv[2 * rs + 1] = kq ? r1 : raw[4 * rs + 1];
      xv[4 + 2 * rs] = kq ? raw[4 * rs + 2] : r0;       // group k = 1: channels cb + 8 + 2 q4 + {0, 1}
      xv[4 + 2 * rs + 1] = kq ? raw[4 * rs + 3] : r1;
    }
  };

  WV_DECL;
  for (int j = 0; j < nblk; ++j) {
    WV_T(tpb);
    const int p = p_begin + j;
    int pix[4];
    const uint32_t flags = locate(p, pix);
    // L2 prefetch of the NEXT block's input tile (same band: R rows further down): one 256-byte pixel row per thread
    if (r_in_band + 1 < g.hb) {
      int pp = pix[0];
      bool ok = flags & 1u;
#pragma unroll
      for (int i = 1; i < 4; ++i)
        if (q4 == i) { pp = pix[i]; ok = (flags >> i) & 1u; }
      if (ok) {
        const float* nx = xt + ((int64_t)pp + (int64_t)g.R * g.W) * 64 - 2 * q4;
        prefetch_l2(nx);
        prefetch_l2(nx + 32);
      }
    }
    float xc[8], xn[8], xn2[8];   // input current: chunk being computed / next / next-but-one (software pipeline, depth 2)
    float x0c[8], x0n[8];         // level 1 only: x_0 = m_0 (an L2 hit: level 0 read it a block earlier; depth 1)
    uint32_t mc[8], mn[8];        // levels >= 2: m_{t-1} from the scratch (L2; depth 1)
    load_x(0, pix, flags, xc, xt);                    // in flight during the waits
    load_x(1, pix, flags, xn, xt);
    if (KIND == 1) load_x(0, pix, flags, x0c, x0p);

    // ---- wait for this block's spread accumulator, the ring rows and the tensor-memory slots we are about to write ----
    WV_T(tp0);
    if (KIND > 0) {
      wv_wait(&ctl->full[lvl][j & (kWvFullDepth - 1)], (uint32_t)(j / kWvFullDepth) & 1u);
      WV_T(tp1);
      WV_ACC(prof_full, tp1 - tp0);
      if (KIND == 1 && !LAST && j >= kWvMSlots) wv_wait(&ctl->m_free[j % kWvMSlots], (uint32_t)(j / kWvMSlots - 1) & 1u);
      // membranes of this block written by level lvl-1: phase (use * (T-2) + lvl - 2) of the slot's barrier
      if (KIND > 1) wv_wait(&ctl->m_ready[j % kWvMSlots], (uint32_t)((j / kWvMSlots) * (g.T - 2) + lvl - 2) & 1u);
      tc_fence_after_sync();
    }
    WV_T(tp2);
    if (!LAST && j >= 1) wv_wait(&ctl->ring_free[lvl], (uint32_t)(j - 1) & 1u);
    WV_T(tp3);
    WV_ACC(prof_ring, tp3 - tp2);
    WV_ACC(prof_pre, tp0 - tpb);

    const uint32_t e_col = (uint32_t)(j % kWvESlots) * 64u;
    // membrane scratch of this block: [chunk u][thread of the level][8 values] -> two coalesced 16-byte accesses per chunk
    float* mslot = g.mscr + ((size_t)blockIdx.x * kWvMSlots + (size_t)(j % kWvMSlots)) * kWvMSlotFloats + (threadIdx.x & 127) * 8;
    const uint32_t slot_row = (uint32_t)((j & 1) ? SB : SA);
    uint32_t wlo[2] = {0u, 0u}, whi[2] = {0u, 0u};     // HBM spike words (channels 0-31 / 32-63) of the half's two pixels
    auto load_m = [&](int u, uint32_t (&m)[8]) {
      const uint4 m0 = ld_cg_u4(mslot + u * 1024), m1 = ld_cg_u4(mslot + u * 1024 + 4);
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
    };
    if (KIND > 1) load_m(0, mc);

#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      const int h = u >> 2, gq = (u >> 1) & 1, kh = u & 1, cb = 16 * (u & 3);     // cb: first channel of the chunk (+ 2 q4 + e)
      WV_T(c0);
      // next chunk's loads first, then this chunk's tensor-memory load: all in flight while the previous results retire
      if (u + 2 < 8) load_x(u + 2, pix, flags, xn2, xt);
      if (u + 1 < 8) {
        if (KIND == 1) load_x(u + 1, pix, flags, x0n, x0p);
        if (KIND > 1) load_m(u + 1, mn);
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * q + 16 * h) << 16) + (uint32_t)cb;
      uint32_t ev[8];
      if (KIND > 0) tmem_ld_16x256b_x2(t_addr + e_col, ev);
      // the chunk's input current arrives as raw 16-byte slices: swap pairs with the neighbour lane now (see load_x)
      {
        float t8[8];
        unpack_x(xc, t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) xc[i] = t8[i];
        if (KIND == 1) {
          unpack_x(x0c, t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) x0c[i] = t8[i];
        }
      }
      float2 scv[2], shv[2];
      if (AFF) {
#pragma unroll
        for (int k = 0; k < 2; ++k) { scv[k] = lds64f(c_scale + (cb + 8 * k) * 4); shv[k] = lds64f(c_shift + (cb + 8 * k) * 4); }
#pragma unroll
        for (int idx = 0; idx < 8; ++idx)
          xc[idx] = add_rn(mul_rn(xc[idx], (idx & 1) ? scv[idx >> 2].y : scv[idx >> 2].x), (idx & 1) ? shv[idx >> 2].y : shv[idx >> 2].x);
      }
      uint32_t mv[8];       // m_t
      float2 bcv[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      if (KIND > 0) {
        bcv[0] = lds64f(c_bconst + cb * 4);
        bcv[1] = lds64f(c_bconst + (cb + 8) * 4);
        WV_T(c1);
        tmem_ld_wait();
        WV_T(c2);
        WV_ACC(pc_issue, c1 - c0);
        WV_ACC(pc_tmem, c2 - c1);
      }
      WV_T(c2b);
      uint32_t w2[2][2];      // {0,1} bf16 pairs of the chunk's two pixels (4 channels = 8 bytes each) for the ring
      uint32_t word[2] = {0u, 0u};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          uint32_t sp[2];      // all-ones where the neuron fires
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int idx = 4 * k + 2 * rs + e;
            float m;
            if (KIND == 0) {
              m = xc[idx];
            } else {
              float mo;
              if (KIND == 1) {
                mo = x0c[idx];
                if (AFF) mo = add_rn(mul_rn(mo, e ? scv[k].y : scv[k].x), e ? shv[k].y : shv[k].x);
              } else {
                mo = __uint_as_float(mc[idx]);
              }
              const float pre = add_rn(mo > g.thresh ? 0.f : mul_rn(mo, g.decay), xc[idx]);
              const float E = add_rn(__uint_as_float(ev[idx]), e ? bcv[k].y : bcv[k].x);
#ifdef ECSY_ACCURATE_TANH
              const float y = exp2f_approx(mul_rn(E, c1));
              const float f = __fmaf_rn(rcp_approx(add_rn(y, 1.f)), nb2, g.beta);      // beta * tanh(alpha * E)
#else
              const float f = mul_rn(g.beta, tanh_approx(mul_rn(g.alpha, E)));         // one MUFU op (see elementwise.cu)
#endif
              m = add_rn(pre, f);
              ev[idx] = __float_as_uint(mul_rn(g.kappa, E));
            }
            mv[idx] = __float_as_uint(m);
            sp[e] = m > g.thresh ? 0xffffffffu : 0u;
          }
          w2[rs][k] = __byte_perm(sp[0], sp[1], 0x5410) & 0x3F803F80u;          // low half: channel e = 0, high half: e = 1
          word[rs] |= ((sp[0] & 1u) | (sp[1] & 2u)) << (8 * k);
        }
      }
      WV_T(c3);
      WV_ACC(pc_math, c3 - c2b);
      if (!LAST && KIND > 0) {
        tmem_st_16x256b_x2(t_addr + e_col, ev);
        st_cg_u4(mslot + u * 1024, make_uint4(mv[0], mv[1], mv[2], mv[3]));
        st_cg_u4(mslot + u * 1024 + 4, make_uint4(mv[4], mv[5], mv[6], mv[7]));
      }
      const uint32_t wshift = (uint32_t)(16 * kh + 2 * q4);
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        if (gq == 0) wlo[rs] |= word[rs] << wshift; else whi[rs] |= word[rs] << wshift;
        if (!LAST) {
          const int a = a0 + 16 * h + 8 * rs;
          const uint32_t chunk = (uint32_t)(2 * q4 + gq);
          const uint32_t row = slot_row + (uint32_t)a;
          sts64(ring + row * 128u + ((chunk ^ (row & 7u)) << 4) + 8u * kh, w2[rs][0], w2[rs][1]);
          const int ry = a >> g.logWb, rx = a & (g.Wb - 1);
          // mirrors: first image row of a block in slot A -> guard-bottom, last image row of a block in slot B -> guard-top
          if (!(j & 1) && ry == 0) {
            const uint32_t mr = (uint32_t)(GB + rx);
            sts64(ring + mr * 128u + ((chunk ^ (mr & 7u)) << 4) + 8u * kh, w2[rs][0], w2[rs][1]);
          }
          if ((j & 1) && ry == g.R - 1) {
            const uint32_t mr = (uint32_t)(GT + rx);
            sts64(ring + mr * 128u + ((chunk ^ (mr & 7u)) << 4) + 8u * kh, w2[rs][0], w2[rs][1]);
          }
        }
      }
      if ((u & 3) == 3) {
        // ---- HBM spikes of this half: combine the four threads of a pixel; lane q4 = 0 / 1 writes pixel rs = 0 / 1 ----
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          wlo[rs] |= __shfl_xor_sync(0xffffffffu, wlo[rs], 1);
          wlo[rs] |= __shfl_xor_sync(0xffffffffu, wlo[rs], 2);
          whi[rs] |= __shfl_xor_sync(0xffffffffu, whi[rs], 1);
          whi[rs] |= __shfl_xor_sync(0xffffffffu, whi[rs], 2);
        }
        const int rsel = q4 & 1;
        const int pA = h ? pix[2] : pix[0], pB = h ? pix[3] : pix[1];
        if (q4 < 2 && ((flags >> (4 + 2 * h + rsel)) & 1u))
          *reinterpret_cast<uint2*>(g.spikes + (int64_t)lvl * bits_tstride + (int64_t)(rsel ? pB : pA) * 2) =
              make_uint2(rsel ? wlo[1] : wlo[0], rsel ? whi[1] : whi[0]);
        wlo[0] = wlo[1] = whi[0] = whi[1] = 0u;
      }
      WV_T(c4);
      WV_ACC(pc_store, c4 - c3);
      // rotate the software pipeline
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xc[i] = xn[i];
        xn[i] = xn2[i];
        if (KIND == 1) x0c[i] = x0n[i];
        if (KIND > 1) mc[i] = mn[i];
      }
      WV_T(c5);
      WV_ACC(pc_rot, c5 - c4);
    }
    WV_T(tp4);
    WV_ACC(prof_work, tp4 - tp3);
    // ---- hand over ----
    if (!LAST) {
      if (KIND > 0) tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (KIND > 0) mbar_arrive(&ctl->m_ready[j % kWvMSlots]);     // release: this warp's membrane stores are visible
        mbar_arrive(&ctl->ready[lvl]);
      }
    } else {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctl->e_free[j % kWvESlots]);
        if (g.T > 2) mbar_arrive(&ctl->m_free[j % kWvMSlots]);
      }
    }
    // ---- next stream entry ----
    if (band_valid && r_in_band < g.hb) ++r_in_band; else set_band(p + 1);
  }
  WV_REPORT;
}

__global__ void __launch_bounds__(kWvThreads, 1)
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
    for (int k = 0; k < kWvMSlots; ++k) { mbar_init(&ctl->m_free[k], 4); mbar_init(&ctl->m_ready[k], 4); }
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

  if (warp >= kWvIssuerWarp) {
    if (warp - kWvIssuerWarp + 1 < T) {
    // =============================== MMA issuers: one warp per producing level ===============================
    // Issuer t sleeps on ready[t] (hardware-suspended wait, no polling: a polling issuer took most of its scheduler
    // partition's issue slots from the four level warps that share it) and, when level t has finished block j, issues
    //   taps ky = +1 of block j-1 for level t+1 (they needed the first image row of block j)  -> full[t+1] of block j-1
    //   taps ky = -1, 0 of block j for level t+1                                              -> ring_free[t]
    // One thread issues every MMA into a given accumulator; accumulation rounds of different levels on the same E slot are
    // separated by the level that reads and rewrites it in between (ready / full handshakes).
    const int t = warp - kWvIssuerWarp;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint64_t db0 = umma_desc_sw128(smem_u32(w_smem));
    const uint32_t ring = rings_u32 + (uint32_t)t * (uint32_t)ring_bytes;
    for (int j = 0; j <= P - 1 - t; ++j) {
      wv_wait(&ctl->ready[t], (uint32_t)j & 1u);
      const bool has_p1 = j <= P - 2 - t;              // level t+1 processes blocks 0 .. P-2-t
      // the accumulator slot of block j is recycled from block j - kWvESlots: the last level must have read it
      if (t == 0 && has_p1 && j >= kWvESlots) wv_wait(&ctl->e_free[j % kWvESlots], (uint32_t)(j / kWvESlots - 1) & 1u);
      tc_fence_after_sync();
      if (elect_one()) {
        if (j >= 1) {
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
    }
    }
  } else if (warp < 4 * T) {
    const int lvl = warp >> 2;
    const bool last = lvl == T - 1;
    LevelCtx c;
    c.ctl = ctl; c.tmem_base = tmem_base; c.rings_u32 = rings_u32; c.ring_bytes = ring_bytes;
    c.c_bconst = smem_u32(s_bconst); c.c_scale = smem_u32(s_scale); c.c_shift = smem_u32(s_shift);
    c.p_begin = (int)p_begin; c.seg0 = (int)seg0; c.seg1 = (int)seg1; c.P = P; c.lvl = lvl;
    if (affine) {
      if (lvl == 0) wv_level<0, false, true>(g, c);
      else if (lvl == 1) { if (last) wv_level<1, true, true>(g, c); else wv_level<1, false, true>(g, c); }
      else { if (last) wv_level<2, true, true>(g, c); else wv_level<2, false, true>(g, c); }
    } else {
      if (lvl == 0) wv_level<0, false, false>(g, c);
      else if (lvl == 1) { if (last) wv_level<1, true, false>(g, c); else wv_level<1, false, false>(g, c); }
      else { if (last) wv_level<2, true, false>(g, c); else wv_level<2, false, false>(g, c); }
    }
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

static int64_t wv_grid(int64_t S) {
  // every CTA should own at least ~8 blocks of useful work next to its warm-up / cool-down blocks
  int64_t grid = ecsy_num_sms();
  if (S / 8 < grid) grid = S / 8 > 0 ? S / 8 : 1;
  return grid;
}

extern "C" size_t ecsy_lif_ecs_wave_ws_bytes(int T, int64_t N, int H, int W, int C) {
  (void)T; (void)N; (void)H; (void)W; (void)C;
  return (size_t)ecsy_num_sms() * kWvMSlots * kWvMSlotFloats * sizeof(float) + 256;
}

// Measured on B200 (tools/wave_bench.py, batch 64, T = 4, fast precision): against the per-timestep pipeline the wavefront
// kernel takes 1.36 vs 1.47 ms at 64 ch @ 160 x 160, 4.8 vs 5.6 ms on the T-broadcast 320 x 320 layer, but 0.51 vs 0.42 ms at
// 80 x 80, where the bands are 32 columns wide (four rows per block: a larger share of halo and mirror work).
extern "C" int ecsy_lif_ecs_wave_prefers(int T, int C, int H, int W) {
  int R, Wb, nb;
  if (!ecsy_lif_ecs_wave_supported(T, C, H, W)) return 0;
  wv_plan(T, W, &R, &Wb, &nb);
  return Wb == 64 ? 1 : 0;
}

extern "C" int ecsy_lif_ecs_wave_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                                     const void* w_eff, const float* bconst, uint32_t* spikes, int T, int64_t N, int H,
                                     int W, int C, float thresh, float decay, float alpha, float beta, float kappa,
                                     void* ws, size_t ws_bytes, void* stream) {
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
  const int64_t grid = wv_grid(g.S);
  const size_t need = ecsy_lif_ecs_wave_ws_bytes(T, N, H, W, C);
  if (T > 2 && (ws == nullptr || ws_bytes < need)) {
    ecsy_set_error("lif_ecs_wave_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  g.mscr = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  k_lif_ecs_wave64<<<(int)grid, kWvThreads, smem, st>>>(g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
