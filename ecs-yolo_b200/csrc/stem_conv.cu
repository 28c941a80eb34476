// Stem convolution on the real-valued image (Conv_1 / Snn_Conv2d on a non-spike input, models/common.py:409-425 with
// :609-624): Cin <= 4 channels, k x k taps (k <= 8), 64 output channels, fast precision (one bf16 plane) -- the 7x7 / stride-2
// stem of cfg/resnet10|18|34.yaml.  The generic im2col-on-the-fly producer of k_umma_gemm<kAGather> spent 2.3 ms per
// batch-64 step here (3 % tensor-pipe active): a table look-up, two bounds tests and one scalar load per operand element,
// with the full global-load latency exposed once per 64-wide K block.  This kernel
//   * reads the image as NHWC4 bf16 (8 bytes per pixel, written by a one-pass converter), so that the operand row of an
//     output pixel for one kernel ROW ky is 8 consecutive pixels = 64 contiguous bytes = 32 K entries (kx < 8, 4 channels;
//     the weight rows are zero where kx >= k or ci >= Cin) -- no conversion, no table, 8-byte loads;
//   * keeps the whole packed weight [64][ceil(k/2) * 64] resident in shared memory (TMA, once per CTA);
//   * builds one 128-pixel tile (2 output rows x 64 columns) per producer group with ALL of its loads in flight at once,
//     two groups of 8 warps alternating tiles, so the latency of one tile hides behind the stores of the other;
//   * accumulates in tensor memory (tcgen05.mma, two buffers) and writes each warp's 32 pixels x 64 channels as one
//     contiguous 8 KB run through a shared-memory transpose (folded tdBN scale / shift applied on the way).
#include <algorithm>

#include "ecsy_common.cuh"
#include "../../include/ecsy.h"
#include "umma_gemm.h"

using namespace ecsy;

namespace {

constexpr int kCout = 64;
constexpr int kMaxKB = 4;                         // K blocks of 64 = two kernel rows each: k <= 8
constexpr int kProdWarps = 16;                    // two groups of 8
constexpr int kEpiWarps = 8;                      // two epilogue groups of 4
constexpr int kThreads = (kEpiWarps + 2 + kProdWarps) * 32;   // epilogue + TMA + MMA + producers
constexpr int kTile = 128 * 128;                  // bytes of a [128 rows][64 bf16] operand block
constexpr int kWBlk = kCout * 128;                // bytes of a [64 co][64 k] weight block
constexpr int kStageRow = 36;                     // floats per staged row of a 32-column chunk (padding: conflict-free float4 access)

struct StemCtl {
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2], w_full;
  uint32_t tmem_base, pad;
};

struct StemArgs {
  const uint2* xq;      // [imgs][H][W] pixels of 4 bf16
  float* out;           // [imgs][Ho][Wo][64]
  const float* scale;   // optional folded tdBN
  const float* shift;
  int imgs, H, W, Ho, Wo, k, stride, pad, KB, tiles_h, tiles_w, total_tiles;
  int dbg;   // profiling only (ECSY_STEM_DBG): 1 = no global stores, 2 = no global loads
};

__global__ void __launch_bounds__(256) k_to_nhwc4_bf16(const float* __restrict__ x, uint2* __restrict__ out, int64_t pixels,
                                                       int Cin) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixels; p += stride) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < Cin; ++c) v[c] = x[p * Cin + c];
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    out[p] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_stem_umma(const __grid_constant__ CUtensorMap tm_w, const StemArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                                        // 2 slots x kMaxKB blocks
  uint8_t* w_smem = a_smem + 2 * kMaxKB * kTile;                 // kMaxKB blocks
  float* o_smem = reinterpret_cast<float*>(w_smem + kMaxKB * kWBlk);   // 8 warps x 32 rows x kStageRow
  StemCtl* ctl = reinterpret_cast<StemCtl*>(o_smem + kEpiWarps * 32 * kStageRow);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (int)blockIdx.x < g.total_tiles ? (g.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int tiles_hw = g.tiles_h * g.tiles_w;

  if (warp == kEpiWarps && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->a_full[s], kProdWarps / 2);
      mbar_init(&ctl->a_empty[s], 1);
      mbar_init(&ctl->t_full[s], 1);
      mbar_init(&ctl->t_empty[s], 4);
    }
    mbar_init(&ctl->w_full, 1);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc<128>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == kEpiWarps) {
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(&ctl->w_full, (uint32_t)(g.KB * kWBlk));
      for (int kb = 0; kb < g.KB; ++kb) tma_load_2d(w_smem + kb * kWBlk, &tm_w, &ctl->w_full, kb * 64, 0);
    }
  } else if (warp == kEpiWarps + 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, kCout);
    if (my_tiles > 0) mbar_wait(&ctl->w_full, 0);
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t slot = it & 1, ph = (it >> 1) & 1;
      mbar_wait(&ctl->t_empty[slot], ph ^ 1);
      mbar_wait(&ctl->a_full[slot], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(a_smem + slot * kMaxKB * kTile), w_addr = smem_u32(w_smem);
        uint32_t acc = 0u;
        for (int kb = 0; kb < g.KB; ++kb) {
          const uint64_t da = umma_desc_sw128(a_addr + kb * kTile), db = umma_desc_sw128(w_addr + kb * kWBlk);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem_base + slot * kCout, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
            acc = 1u;
          }
        }
        umma_commit(&ctl->a_empty[slot]);
        umma_commit(&ctl->t_full[slot]);
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarps + 2) {
    // =============================== operand producers ===============================
    const int pw = warp - (kEpiWarps + 2), grp = pw >> 3;
    const int t = (pw & 7) * 32 + lane;
    const int row = t & 127, part = t >> 7;
    const int h_l = row >> 6, w_l = row & 63;
    const uint32_t sw = (uint32_t)row & 7u;
    uint8_t* slot_base = a_smem + grp * kMaxKB * kTile + (uint32_t)row * 128u;
    uint32_t coff[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) coff[m] = (((uint32_t)(part * 4 + m)) ^ sw) << 4;
    for (int it = grp; it < my_tiles; it += 2) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / tiles_hw;
      const int rem = tile - img * tiles_hw;
      const int th = rem / g.tiles_w, tw = rem - th * g.tiles_w;
      const int ho = th * 2 + h_l, wo = tw * 64 + w_l;
      const bool valid = ho < g.Ho && wo < g.Wo;
      const int hi0 = ho * g.stride - g.pad, wi0 = wo * g.stride - g.pad;
      // interior pixels (all 8 columns of the kernel row inside the image, even image width): the 64 bytes of a row are
      // fetched with 16-byte loads where they are aligned -- 4 or 5 load instructions and one address per row instead of
      // 8 predicated 8-byte loads
      const bool w_in = valid && wi0 >= 0 && wi0 + 8 <= g.W && (g.W & 1) == 0;
      const bool odd = (wi0 & 1) != 0;
      const uint2* img_base = g.xq + (int64_t)img * g.H * g.W;
      // two K blocks (= four kernel rows per row pair of threads) at a time: 32 registers of loads in flight
#pragma unroll
      for (int half = 0; half < kMaxKB / 2; ++half) {
        uint2 v[2][8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kb = 2 * half + q;
          const int ky = 2 * kb + part, hi = hi0 + ky;
          const bool rok = valid && kb < g.KB && ky < g.k && (unsigned)hi < (unsigned)g.H && !(g.dbg & 2);
          const uint2* src = img_base + (rok ? hi : 0) * g.W;
          if (rok && w_in) {
            const uint2* p = src + wi0;
            if (odd) {
              v[q][0] = __ldg(p);
              const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + 1));
              const uint4 b = __ldg(reinterpret_cast<const uint4*>(p + 3));
              const uint4 c = __ldg(reinterpret_cast<const uint4*>(p + 5));
              v[q][7] = __ldg(p + 7);
              v[q][1] = make_uint2(a.x, a.y); v[q][2] = make_uint2(a.z, a.w);
              v[q][3] = make_uint2(b.x, b.y); v[q][4] = make_uint2(b.z, b.w);
              v[q][5] = make_uint2(c.x, c.y); v[q][6] = make_uint2(c.z, c.w);
            } else {
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + 2 * m));
                v[q][2 * m] = make_uint2(a.x, a.y);
                v[q][2 * m + 1] = make_uint2(a.z, a.w);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int wi = wi0 + j;
              v[q][j] = (rok && (unsigned)wi < (unsigned)g.W) ? __ldg(src + wi) : make_uint2(0u, 0u);
            }
          }
        }
        if (half == 0) mbar_wait(&ctl->a_empty[grp], ((it >> 1) & 1) ^ 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kb = 2 * half + q;
          if (kb < g.KB) {
            uint8_t* dst = slot_base + kb * kTile;
#pragma unroll
            for (int m = 0; m < 4; ++m)   // pixels 2m, 2m+1 of this kernel row = K entries part*32 + 8m .. +7
              *reinterpret_cast<uint4*>(dst + coff[m]) = make_uint4(v[q][2 * m].x, v[q][2 * m].y, v[q][2 * m + 1].x, v[q][2 * m + 1].y);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->a_full[grp]);
    }
  } else {
    // =============================== epilogue: two groups of 4 warps, group ge drains accumulator buffer ge ===============
    const int ge = warp >> 2, q = warp & 3;
    float* stage = o_smem + warp * 32 * kStageRow;
    const int h_l = q >> 1, w_l0 = (q & 1) * 32;   // the warp's 32 rows are 32 consecutive pixels of one output row
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    // the folded tdBN is applied in the write-out phase, where a lane owns FIXED channels: 4 registers of scale / shift per
    // 32-column chunk instead of 16 loads per chunk in the dependent chain of the staging phase
    float4 sc[2], sf[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      sc[c] = make_float4(1.f, 1.f, 1.f, 1.f);
      sf[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g.scale != nullptr) {
        sc[c] = __ldg(reinterpret_cast<const float4*>(g.scale + c * 32 + c4));
        sf[c] = __ldg(reinterpret_cast<const float4*>(g.shift + c * 32 + c4));
      }
    }
    for (int it = ge; it < my_tiles; it += 2) {
      const uint32_t ph = (it >> 1) & 1;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / tiles_hw;
      const int rem = tile - img * tiles_hw;
      const int th = rem / g.tiles_w, tw = rem - th * g.tiles_w;
      const int ho = th * 2 + h_l, wo0 = tw * 64 + w_l0;
      float* dst = g.out + (((int64_t)img * g.Ho + ho) * g.Wo + wo0) * kCout;
      const bool row_ok = ho < g.Ho && !(g.dbg & 1);
      mbar_wait(&ctl->t_full[ge], ph);
      tc_fence_after_sync();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + ge * kCout, v0);
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + ge * kCout + 32, v1);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->t_empty[ge]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t* v = c == 0 ? v0 : v1;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(stage + lane * kStageRow + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + sub;
            float4 o = *reinterpret_cast<const float4*>(stage + r * kStageRow + c4);
            if (g.scale != nullptr) {
              o.x = add_rn(mul_rn(o.x, sc[c].x), sf[c].x); o.y = add_rn(mul_rn(o.y, sc[c].y), sf[c].y);
              o.z = add_rn(mul_rn(o.z, sc[c].z), sf[c].z); o.w = add_rn(mul_rn(o.w, sc[c].w), sf[c].w);
            }
            if (wo0 + r < g.Wo) *reinterpret_cast<float4*>(dst + r * kCout + c * 32 + c4) = o;
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after_sync();
    tmem_dealloc<128>(tmem_base);
  }
}

inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

extern "C" int ecsy_stem_conv_supported(int Cin, int Cout, int k, int splits) {
  return (Cin >= 1 && Cin <= 4 && Cout == kCout && k >= 1 && k <= 2 * kMaxKB && splits == 1) ? 1 : 0;
}

extern "C" size_t ecsy_stem_conv_ws_bytes(int64_t imgs, int H, int W) {
  return 256 + al256(static_cast<size_t>(imgs) * H * W * sizeof(uint2));
}

// x: [imgs][H][W][Cin] fp32; w_stem: bf16 [64][ceil(k/2) * 64], entry (co, kb*64 + part*32 + kx*4 + ci) = W[co][ci][2kb+part][kx]
// (zero where ky >= k, kx >= k or ci >= Cin; functional.pack_stem_weight); out: [imgs][Ho][Wo][64] fp32 = conv * scale + shift.
extern "C" int ecsy_stem_conv(const float* x, int64_t imgs, int H, int W, int Cin, const void* w_stem, float* out,
                              const float* scale, const float* shift, int Cout, int k, int stride, int pad, void* ws,
                              size_t ws_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ECSY_CHECK_ARG(x && w_stem && out && imgs > 0 && H > 0 && W > 0 && stride >= 1 && pad >= 0, "stem_conv: bad arguments");
  ECSY_CHECK_ARG(ecsy_stem_conv_supported(Cin, Cout, k, 1), "stem_conv: Cin=%d Cout=%d k=%d unsupported", Cin, Cout, k);
  ECSY_CHECK_ARG((scale == nullptr) == (shift == nullptr), "stem_conv: scale / shift pair");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ECSY_CHECK_ARG(Ho > 0 && Wo > 0, "stem_conv: geometry");
  const size_t need = ecsy_stem_conv_ws_bytes(imgs, H, W);
  if (ws == nullptr || ws_bytes < need) {
    ecsy_set_error("stem_conv: workspace %zu < %zu bytes", ws_bytes, need);
    return ECSY_ERR_WS;
  }
  uint2* xq = reinterpret_cast<uint2*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  const int64_t pixels = imgs * H * W;
  const int cgrid = (int)std::min<int64_t>((pixels + 255) / 256, (int64_t)ecsy_num_sms() * 8);
  k_to_nhwc4_bf16<<<cgrid, 256, 0, st>>>(x, xq, pixels, Cin);
  ECSY_LAUNCH_CHECK();
  StemArgs g{};
  g.xq = xq; g.out = out; g.scale = scale; g.shift = shift;
  g.imgs = (int)imgs; g.H = H; g.W = W; g.Ho = Ho; g.Wo = Wo; g.k = k; g.stride = stride; g.pad = pad;
  g.KB = (k + 1) / 2;
  g.tiles_h = (Ho + 1) / 2; g.tiles_w = (Wo + 63) / 64;
  const int64_t total = imgs * g.tiles_h * g.tiles_w;
  ECSY_CHECK_ARG(total < (1LL << 31), "stem_conv: too many tiles");
  g.total_tiles = (int)total;
  static const int dbg = getenv("ECSY_STEM_DBG") ? atoi(getenv("ECSY_STEM_DBG")) : 0;
  g.dbg = dbg;
  CUtensorMap tw;
  int rc = ecsy_tensor_map_bf16(w_stem, (uint64_t)kCout, (uint64_t)g.KB * 64, (uint32_t)kCout, &tw);
  if (rc) return rc;
  const int smem = 1024 + 2 * kMaxKB * kTile + kMaxKB * kWBlk + kEpiWarps * 32 * kStageRow * (int)sizeof(float) + (int)sizeof(StemCtl) + 64;
  static bool attr = false;
  if (!attr) {
    ECSY_CUDA(cudaFuncSetAttribute(k_stem_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int grid = (int)std::min<int64_t>(total, (int64_t)ecsy_num_sms());
  k_stem_umma<<<grid, kThreads, smem, st>>>(tw, g);
  ECSY_LAUNCH_CHECK();
  return ECSY_OK;
}
