/* ecsy.h -- C ABI of the B200-native ECS-YOLO spiking hot path (libecsy.so, sm_100a).
 *
 * Drop-in boundary for the reference's PyTorch operator calls on the time-unrolled spiking backbone
 * and detection heads (mowanggui/ECS-YOLO, models/common.py, models/yolo.py, models/yolo_snn.py).
 * Every entry point cites the reference interface it replaces.  Conventions:
 *   - plain pointers + sizes, no framework types; all pointers are DEVICE pointers;
 *   - the caller owns every buffer including workspaces (size via the *_ws_bytes queries); the library
 *     allocates nothing and keeps no state beyond cached TMA descriptors keyed by pointer/shape;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*), no hidden synchronisation;
 *   - return 0 on success, a negative code otherwise; ecsy_last_error() (thread-local) explains;
 *   - there is no CPU fallback.
 * Layouts: real activations are fp32 NHWC "[imgs][H][W][C]" with imgs = T*N (t-major); spikes are
 * bit-packed along C: uint32 "[imgs][H][W][C/32]", bit (c & 31) of word (c >> 5).
 * A source with `*_imgs` < imgs is broadcast over T (image index taken modulo `*_imgs`): the direct-coded
 * input of Model.forward (models/yolo.py:248-251) repeats the same image every timestep.
 */
#ifndef ECSY_H_
#define ECSY_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECSY_ABI_VERSION 2

const char* ecsy_last_error(void);
int ecsy_abi_version(void);
int ecsy_sm_count(void);

/* ---- layout at the model edge (reference tensors are NCHW per timestep, models/common.py:619-624) ---- */
int ecsy_nchw_to_nhwc_f32(const float* in, float* out, int64_t imgs, int C, int H, int W, void* stream);
int ecsy_nhwc_to_nchw_f32(const float* in, float* out, int64_t imgs, int C, int H, int W, void* stream);
/* ActFun.forward (models/common.py:61-64): bit = x > thresh.  unpack gives the {0,1} float tensor back. */
int ecsy_spikes_pack(const float* x_nhwc, uint32_t* bits, int64_t pixels, int C, float thresh, void* stream);
int ecsy_spikes_unpack(const uint32_t* bits, float* x_nhwc, int64_t pixels, int C, void* stream);

/* ---- weights: nn.Conv2d weight [Co][Ci][kh][kw] fp32 (Snn_Conv2d.weight, mem_update.spread[1].weight)
 *      -> [splits][Co][Kpad] bf16, k = (ky*kw+kx)*Ci + ci; plane 0 = bf16(w), plane 1 = bf16(w - plane 0). */
int ecsy_pack_conv_weight(const float* w, void* out_bf16, int Co, int Ci, int kh, int kw, int Kpad, int splits,
                          void* stream);

/* ---- mem_update.forward (models/common.py:252-283): ECS-LIF over T steps.
 * x: [T][N][H][W][C] real input current (t stride `x_tstride` elements; 0 = same tensor every step),
 * optional per-channel affine on x (a folded tdBN), dw_w: spread[0].weight as [9][C], dw_b: [C],
 * pw_packed: spread[1].weight packed by ecsy_pack_conv_weight (Kpad = C), pw_b: [C].
 * spikes: [T][N][H][W][C/32] out.  mem_save / ecs_save (optional): membranes m_t [T][..][C] and ECS traces e_t
 * [T-1][..][C], written when the backward pass re-runs the forward (recompute instead of store). */
size_t ecsy_lif_ecs_ws_bytes(int T, int64_t N, int H, int W, int C, int splits);
int ecsy_lif_ecs_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                     const float* dw_w, const float* dw_b, const void* pw_packed, const float* pw_b, int splits,
                     uint32_t* spikes, float* mem_save, float* ecs_save, int T, int64_t N, int H, int W, int C,
                     float thresh,
                     float decay, float alpha, float beta, float kappa, void* ws, size_t ws_bytes, void* stream);

/* ---- the depth-wise half of mem_update.spread on its own (models/common.py:289-294, spread[0] = Conv2d(C, C, 3,
 * padding 1, groups C) applied to the spikes of one step): bits [N][H][W][C/32] -> bf16 rows [N*H*W][C] a_hi
 * (+ residual plane a_lo = bf16(v - a_hi), or NULL), the A operand of the point-wise spread GEMM.  `version`: 0 = the
 * kernel ecsy_lif_ecs_fwd uses, 1 = per-pixel global byte loads, 2 = shared-memory staged tiles (bit-identical). */
int ecsy_spread_dw(const uint32_t* bits, const float* dw_w, const float* dw_b, void* a_hi, void* a_lo, int64_t N,
                   int H, int W, int C, int version, void* stream);

/* ---- mem_update(act=True).forward: the SiLU "analog spike" neuron of class Conv (models/common.py:362-375,
 * 252-283).  out: [T][N][H][W][C] fp32 = silu(mem_t).  inplace != 0 reproduces the reference models, where
 * initialize_weights() makes nn.SiLU in-place so that mem_old holds silu(mem) (utils/torch_utils.py:165-166). */
/* Fused forward of the same neuron for C == 64 (fast / single-plane mode): every timestep of a 22x22 pixel tile runs
 * inside one CTA, membrane and ECS trace stay in tensor memory, only x is read and only spike bits are written.
 * w_eff_ts: ecsy_pack_spike_conv_weight of W_eff[co][ci][ky][kx] = pw[co][ci] * dw[ci][ky][kx] (splits = 1);
 * bconst[co] = sum_ci pw[co][ci] * dw_b[ci] + pw_b[co].  No workspace. */
int ecsy_lif_ecs_fused_supported(int T, int C);
int ecsy_lif_ecs_fused_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                           const void* w_eff_ts, const float* bconst, uint32_t* spikes, int T, int64_t N, int H, int W,
                           int C, float thresh, float decay, float alpha, float beta, float kappa, void* stream);

/* Wavefront forward of the same neuron for C == 64, 2 <= T <= 4 (fast / single-plane precision): the DEFAULT inference
 * path for 64-channel layers.  The time loop runs as a wavefront down vertical bands of the image inside one persistent
 * kernel (csrc/lif_wave.cu): membrane and ECS trace of all T steps stay in tensor memory, x is read once, only spike
 * bits are written (reference loop: models/common.py:252-283, state `mem` / `ecs` carried across `for i in
 * range(time_window)`).  w_eff: [9][64][64] bf16, tap-major, W_eff[tap][co][kk] = pw[co][c(kk)] * dw[c(kk)][tap] with
 * the channel permutation c(kk) = 8 * ((kk % 16) / 2) + 2 * (kk / 16) + kk % 2 (functional.pack_lif_wave_weight);
 * bconst as above.  ws: ecsy_lif_ecs_wave_ws_bytes() of CTA-private membrane scratch (28 MB, written and re-read within
 * microseconds: it lives in the L2). */
int ecsy_lif_ecs_wave_supported(int T, int C, int H, int W);
int ecsy_lif_ecs_wave_prefers(int T, int C, int H, int W);   /* measured dispatch rule (csrc/lif_wave.cu) */
size_t ecsy_lif_ecs_wave_ws_bytes(int T, int64_t N, int H, int W, int C);
int ecsy_lif_ecs_wave_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                          const void* w_eff, const float* bconst, uint32_t* spikes, int T, int64_t N, int H, int W,
                          int C, float thresh, float decay, float alpha, float beta, float kappa, void* ws,
                          size_t ws_bytes, void* stream);

size_t ecsy_lif_silu_ws_bytes(int T, int64_t N, int H, int W, int C, int splits);
int ecsy_lif_silu_fwd(const float* x, int64_t x_tstride, const float* in_scale, const float* in_shift,
                      const float* dw_w, const float* dw_b, const void* pw_packed, const float* pw_b, int splits,
                      float* out, float* mem_save, float* ecs_save, int inplace, int T, int64_t N, int H, int W, int C,
                      float decay, float alpha, float beta, float kappa, void* ws, size_t ws_bytes, void* stream);
/* backward of the in-place SiLU neuron (out / mem / ecs from a re-run of ecsy_lif_silu_fwd with mem_save, ecs_save);
 * same argument meaning as ecsy_lif_ecs_bwd, workspace size from ecsy_lif_ecs_bwd_ws_bytes. */
int ecsy_lif_silu_bwd(const float* gout, const float* out, const float* mem, const float* ecs, const float* dw_w,
                      const float* dw_b, const void* pwT_packed, int splits, float* gx, float* g_dw_w, float* g_dw_b,
                      float* g_pw_w, float* g_pw_b, int T, int64_t N, int H, int W, int C, float decay, float alpha,
                      float beta, float kappa, void* ws, size_t ws_bytes, void* stream);

/* ---- Snn_Conv2d.forward on spikes (models/common.py:609-624), tcgen05 implicit GEMM.
 * out[imgs][Ho][Wo][Cout] = conv(spikes, W) * scale[c] + shift[c] (+ residual).  scale/shift: folded
 * eval-mode tdBN (models/common.py:674-679) or NULL; residual: the membrane shortcut (common.py:1216). */
int ecsy_spike_conv_fwd(const uint32_t* spikes, const void* w_packed, int splits, float* out, const float* scale,
                        const float* shift, const float* residual, int64_t res_imgs, int64_t imgs, int H, int W,
                        int Cin, int Cout, int k, int stride, int pad, void* stream);

/* Same operator with the spike operand staged in TENSOR MEMORY (tcgen05.mma with A in TMEM; Cin % 64 == 0,
 * Cout % 64 == 0 -- ask ecsy_spike_conv_ts_supported).  w_ts: [splits][Cout][k*k*Cin] bf16 from
 * ecsy_pack_spike_conv_weight (weights * 0.5 -- a spike is emitted as 2.0 -- in the expander's channel order). */
int ecsy_pack_spike_conv_weight(const float* w, void* out_bf16, int Co, int Ci, int kh, int kw, int splits,
                                void* stream);
int ecsy_spike_conv_ts_supported(int Cin, int Cout);
int ecsy_spike_conv_prefers_ts(int Cin, int Cout, int splits);   /* measured dispatch rule (wide layers keep smem A) */
/* shared-memory-operand kernel (wide 256-column tiles) on the same w_ts weights: the layers prefers_ts rejects */
int ecsy_spike_conv_pair_fwd(const uint32_t* spikes, const void* w_ts, int splits, float* out, const float* scale,
                             const float* shift, const float* residual, int64_t res_imgs, int64_t imgs, int H, int W,
                             int Cin, int Cout, int k, int stride, int pad, void* stream);
int ecsy_spike_conv_ts_fwd(const uint32_t* spikes, const void* w_ts, int splits, float* out, const float* scale,
                           const float* shift, const float* residual, int64_t res_imgs, int64_t imgs, int H, int W,
                           int Cin, int Cout, int k, int stride, int pad, void* stream);

/* ---- Snn_Conv2d.forward on REAL inputs: stem Conv_1 (common.py:409-425), Conv (common.py:362-375),
 * Detect.m (yolo.py:73), DDetect cv2/cv3[-1] (yolo_snn.py:100-107).  groups == 1 && Cout % 64 == 0 && no bias:
 * im2col + tcgen05 GEMM (needs w_packed + workspace); otherwise SIMT with w_simt = [kh][kw][Ci/g][Co] fp32.
 * bias is added as bias * bias_mul (Detect fuses Conv_7's sum of T weights into the bias). */
size_t ecsy_real_conv_ws_bytes(int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad, int groups,
                               int splits);
int ecsy_real_conv_fwd(const float* x, int64_t x_imgs, const void* w_packed, const float* w_simt, int splits,
                       const float* bias, float bias_mul, const float* scale, const float* shift, float* out,
                       int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad, int groups, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- batch_norm_2d / batch_norm_2d1 (models/common.py:668-700,753-758): per-channel mean and biased
 * variance over all rows = T*N*H*W of x[rows][C] (train mode).  Normalisation itself is applied as a
 * per-channel affine by the consumers (lif in_scale, conv scale, affine_add). */
size_t ecsy_tdbn_stats_ws_bytes(int64_t rows, int C);
int ecsy_tdbn_stats(const float* x, int64_t rows, int C, float* mean, float* var_biased, void* ws, size_t ws_bytes,
                    void* stream);
/* Train-mode tdBN after ecsy_tdbn_stats, one launch (nn.BatchNorm3d inside batch_norm_2d, models/common.py:668-700):
 * running statistics updated `updates` times with `momentum` (unbias = n / (n - 1); running_* / num_batches_tracked may be
 * NULL), rstd = rsqrt(var + eps), scale = weight * rstd, shift = bias - mean * scale -- all device [C] vectors. */
int ecsy_tdbn_finish(const float* mean, const float* var_biased, const float* weight, const float* bias, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float momentum, float unbias, float eps, int updates,
                     float* scale, float* shift, float* rstd, int C, void* stream);
/* tdBN backward coefficients, one launch: from sg = sum g and sgy = sum g*y (ecsy_colsum2) the per-channel A, B, C of
 * g_y = A g + B y + C (B and C times tfac = T / Tp for a T-broadcast y) and the weight gradient; the bias gradient is sg. */
int ecsy_tdbn_bwd_coef(const float* sg, const float* sgy, const float* mean, const float* rstd, const float* weight, float n,
                       float tfac, float* A, float* B, float* Cc, float* g_weight, int C, void* stream);


/* ---- block output `residual_function(x) + shortcut(x)` (models/common.py:1074,1216,1484) with the pending
 * tdBN affines of both operands: out = a*sa+ba (+ b*sb+bb). */
int ecsy_affine_add(const float* a, int64_t a_imgs, const float* sa, const float* ba, const float* b, int64_t b_imgs,
                    const float* sb, const float* bb, float* out, int64_t imgs, int64_t hw, int C, void* stream);

/* ---- nn.MaxPool3d((1,s,s)) (common.py:1209,1481), Sample nearest x`up` (common.py:856-868), Concat on the
 * channel axis (common.py:1764): resample `in` into channels [coff, coff+C) of out[imgs][Ho][Wo][Ctot]. */
int ecsy_resample(const float* in, int64_t in_imgs, const float* scale, const float* shift, float* out, int64_t imgs,
                  int Hi, int Wi, int C, int Ctot, int coff, int pool, int up, void* stream);

/* ---- backward of the resampling ops: nn.MaxPool3d((1,s,s)) (gradient to the first maximum of each window; the
 * pooled gradient may live in channels [gcoff, gcoff+C) of a wider tensor), and Sample / Concat
 * (sum over each s x s block of channels [coff, coff+C) of `in`; s == 1 is a plain channel slice). */
int ecsy_maxpool_bwd(const float* x, int64_t x_imgs, const float* g_pooled, float* gx, int64_t imgs, int Ho, int Wo,
                     int C, int gC, int gcoff, int s, void* stream);
int ecsy_sumpool_slice(const float* in, float* out, int64_t imgs, int Ho, int Wo, int C, int inC, int coff, int s,
                       void* stream);

/* ---- reduction over T: out = (sum_t w[t] * x[t]) / div.  Conv_7 (common.py:549-562, w = Conv3d weight) and
 * DDetect's mean over T (yolo_snn.py:115-116, w = NULL, div = T). */
int ecsy_tsum(const float* x, const float* w, float div, float* out, int T, int64_t per_t, void* stream);

/* ---- Detect.forward view/permute + eval decode (models/yolo.py:110-146).  y: [N][H][W][na*no];
 * raw: [N][na][H][W][no]; z (NULL in training): rows [row_off, row_off + na*H*W) of [N][rows_total][no]. */
int ecsy_detect_decode(const float* y, float* raw, float* z, const float* anchors, float stride_px, int N, int H,
                       int W, int na, int no, int64_t rows_total, int64_t row_off, void* stream);

/* ---- DDetect.forward concat + DFL + dist2bbox + sigmoid (models/yolo_snn.py:115-127, common.py:312-323,
 * utils/tal/anchor_generator.py:8-32).  box: [N][H][W][64], cls: [N][H][W][nc] (means over T);
 * xs: [N][64+nc][H][W]; y (NULL in training): anchors [a_off, a_off + H*W) of [N][4+nc][a_total]. */
int ecsy_ddetect_decode(const float* box, const float* cls, float* xs, float* y, float stride_px, int N, int H, int W,
                        int nc, int64_t a_total, int64_t a_off, void* stream);

/* ---- backward of `Snn_Conv2d -> batch_norm_2d` on spikes in training, one call (autograd of F.conv2d,
 * models/common.py:623, and of nn.BatchNorm3d, :674-679).  g: gradient w.r.t. the NORMALISED output [imgs][Ho][Wo][Cout],
 * y: the raw conv output, A / B / Cv [Cout]: g_y = A*g + B*y + Cv (the tdBN backward given the batch sums from
 * ecsy_colsum2; all NULL: g is already g_y).  g_y is formed once as bf16 planes and feeds the weight gradient (dw
 * [Cout][k*k*Cin], ACCUMULATED) and the input gradient (gx [imgs][H][W][Cin]); wT_packed as for ecsy_conv_dgrad. */
size_t ecsy_spike_conv_bwd_ws_bytes(int64_t imgs, int H, int W, int Cout, int k, int stride, int pad, int splits);
int ecsy_spike_conv_bwd(const float* g, const float* y, const float* A, const float* B, const float* Cv,
                        const uint32_t* spikes, const void* wT_packed, int splits, float* gx, float* dw, int64_t imgs, int H,
                        int W, int Cin, int Cout, int k, int stride, int pad, void* ws, size_t ws_bytes, void* stream);

/* ---- non_max_suppression (utils/general.py:649-741) on the decoded Detect output, all images in one call.
 * pred: [N][R][5+nc] rows (cx, cy, w, h, obj, cls...).  Candidates: obj > conf_thres, conf = obj*cls > conf_thres for
 * the best class, or for every class with multi_label; cls_ok (optional [nc] bytes) is the `classes` filter; at most
 * max_nms candidates by confidence enter the greedy scan (torchvision.ops.nms on class-offset boxes, IoU > iou_thres
 * suppresses; iou_thres is a double like torchvision's argument); the first max_det survivors are written as
 * out[n][i] = (x1, y1, x2, y2, conf, cls), i < out_count[n], in descending confidence (ties: prediction order).
 * `labels` (autolabelling, general.py:681-688) and merge-NMS (dead code in the reference) are not supported. */
size_t ecsy_nms_ws_bytes(int64_t N, int R, int nc, int multi_label);
int ecsy_nms(const float* pred, int64_t N, int R, int nc, float conf_thres, double iou_thres, int agnostic, int multi_label,
             const uint8_t* cls_ok, int max_det, int max_nms, float* out, int* out_count, void* ws, size_t ws_bytes,
             void* stream);

/* ---- optimizer.step() + ModelEMA.update() of the training loop (train.py:282-287, 576-582; utils/torch_utils.py:
 * 306-316) as one multi-tensor launch.  Device tables (caller-owned, built once per model): per tensor the address of the
 * fp32 value, of its gradient (0 = none: a buffer, or no grad this step), of its zero-initialised momentum buffer (0 =
 * none) and of its EMA copy (0 = none), its element count and parameter group; per chunk of ecsy_optim_chunk() elements
 * the tensor index and element offset.  lr / weight_decay: HOST arrays [n_groups] (n_groups <= 8).
 * SGD (dampening 0): d = g + wd*p; buf = first_step ? d : momentum*buf + d; d = nesterov ? d + momentum*buf : buf;
 * p -= lr*d.  EMA (do_ema): e = e*ema_d + ema_one_minus_d*p on the UPDATED p, for every tensor with an EMA address. */
int ecsy_optim_chunk(void);
int ecsy_sgd_ema_step(const uint64_t* val, const uint64_t* grad, const uint64_t* mom, const uint64_t* ema,
                      const int64_t* numel, const int32_t* group, int n_tensors, const int32_t* chunk_tensor,
                      const int64_t* chunk_off, int64_t n_chunks, const float* lr, const float* weight_decay, int n_groups,
                      float momentum, int nesterov, int first_step, int do_ema, float ema_d, float ema_one_minus_d,
                      void* stream);

/* ---- Gen1 event-camera input: event bins -> network input (g1-resnet/utils/give_g1_data.py:550-565 create_data,
 * g1-resnet/utils/datasets_g1T.py:518-533 cv2.resize per frame, g1-resnet/train_g1.py:298 "/ 255").
 * Events in sensor order: ex / ey pixel, ep polarity (0 / 1), eframe = sample * T + time bin.  The last event of a pixel
 * wins (255 * p over grey 127); frames [H][W] (Gen1: 240 x 304) are resized to [Ho][Wo] with OpenCV's fixed-point
 * bilinear kernel, bit-exactly, and written as float32 / 255 to out [T][N][Ho][Wo][3] (the model's NHWC input; the
 * three channels are identical, as in the reference).  oob_count: events outside the sensor / frame range (the
 * reference asserts on them; they are skipped here and counted). */
size_t ecsy_event_frames_ws_bytes(int64_t N, int T, int H, int W);
int ecsy_event_frames(const int32_t* ex, const int32_t* ey, const int32_t* ep, const int32_t* eframe, int64_t n_events,
                      int64_t N, int T, int H, int W, int Ho, int Wo, float* out, int* oob_count, void* ws,
                      size_t ws_bytes, void* stream);

/* ---- backward of mem_update.forward: surrogate-gradient BPTT (ActFun.backward, models/common.py:66-79; autograd
 * through :263-281).  gout: dL/dspikes [T][N][H][W][C]; spikes / mem / ecs: from a re-run of ecsy_lif_ecs_fwd with
 * mem_save + ecs_save; pwT_packed: ecsy_pack_conv_weight of spread[1].weight TRANSPOSED ([ci][co]).
 * gx: dL/d(input current) [T][..][C]; g_dw_w [9][C], g_dw_b [C], g_pw_w [C][C] (co-major), g_pw_b [C] are ACCUMULATED. */
size_t ecsy_lif_ecs_bwd_ws_bytes(int T, int64_t N, int H, int W, int C, int splits);
int ecsy_lif_ecs_bwd(const float* gout, const uint32_t* spikes, const float* mem, const float* ecs, const float* dw_w,
                     const float* dw_b, const void* pwT_packed, int splits, float* gx, float* g_dw_w, float* g_dw_b,
                     float* g_pw_w, float* g_pw_b, int T, int64_t N, int H, int W, int C, float thresh, float lens,
                     float decay, float alpha, float beta, float kappa, void* ws, size_t ws_bytes, void* stream);

/* ---- backward of Snn_Conv2d.forward (autograd of F.conv2d, models/common.py:623).
 * dgrad: gx[imgs][H][W][Cin] from gy[imgs][Ho][Wo][Cout]; wT_packed = ecsy_pack_conv_weight of the flipped,
 * transposed weight W'[ci][co][ky][kx] = W[co][ci][k-1-ky][k-1-kx] (strided convs: zero-inserted gradient).
 * wgrad (spike input): dw[Cout][(ky*kw+kx)*Cin+ci] += gy^T * im2col(spikes), pixel contraction on tcgen05 with
 * MN-major operands; wgrad (real input): dw[Cout][Kpad] += gy^T * im2col(x). */
size_t ecsy_conv_dgrad_ws_bytes(int64_t imgs, int H, int W, int Cout, int k, int stride, int pad, int splits);
int ecsy_conv_dgrad(const float* gy, const void* wT_packed, int splits, float* gx, int64_t imgs, int H, int W, int Cin,
                    int Cout, int k, int stride, int pad, void* ws, size_t ws_bytes, void* stream);
size_t ecsy_spike_conv_wgrad_ws_bytes(int64_t imgs, int Ho, int Wo, int Cout, int splits);
int ecsy_spike_conv_wgrad(const float* gy, const uint32_t* spikes, float* dw, int64_t imgs, int H, int W, int Cin,
                          int Cout, int k, int stride, int pad, int splits, void* ws, size_t ws_bytes, void* stream);
size_t ecsy_real_conv_wgrad_ws_bytes(int64_t imgs, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                                     int splits);
int ecsy_real_conv_wgrad(const float* gy, const float* x, int64_t x_imgs, float* dw, int64_t imgs, int H, int W, int Cin,
                         int Cout, int k, int stride, int pad, int splits, void* ws, size_t ws_bytes, void* stream);

/* ---- per-channel sums for the tdBN / folded-affine backward: sum_g[c] = sum_r g[r][c], sum_gx[c] = sum_r
 * g[r][c] * x[r mod x_rows][c] (autograd of nn.BatchNorm3d inside batch_norm_2d, models/common.py:674-679).
 * ws: at least 16*C + 256 bytes. */
int ecsy_colsum2(const float* g, const float* x, int64_t rows, int64_t x_rows, int C, float* sum_g, float* sum_gx,
                 void* ws, size_t ws_bytes, void* stream);

/* ---- weight-gradient contraction over rows (pixels): out[Ca][Cb] += alpha * sum_p P[p][ca] * Q[p][cb],
 * P/Q row-major bf16 (hi planes + optional lo residual planes).  Replaces autograd's conv2d weight backward for
 * the ECS point-wise spread (mem_update.spread[1], models/common.py:295-297) -- tcgen05 with MN-major operands. */
int ecsy_xty_bf16(const void* p_hi, const void* p_lo, const void* q_hi, const void* q_lo, int64_t rows, int Ca, int Cb,
                  float alpha, float* out, void* stream);

/* ---- Stack-A training loss, forward + gradient in one call: replaces ComputeLoss.__call__ and
 * build_targets (utils/loss.py:162-290; SIoU box term utils/metrics.py:227-307; BCEWithLogits with pos_weight for the
 * objectness / class terms, label smoothing cp / cn, per-level balance) on the path data/hyps/hyp.scratch.yaml selects
 * (fl_gamma = 0, slide_ratio = 0).  p / gp: HOST arrays of nl DEVICE pointers to the raw Detect outputs
 * [N][na][ny_l][nx_l][5 + nc] and their gradients (gp or gp[l] may be null: forward only; otherwise gp[l] is
 * OVERWRITTEN with d loss / d p[l] for an upstream gradient of 1).  targets: device [nt][6] = (image, class, cx, cy,
 * w, h) normalised; anchors: device [nl][na][2] in grid units (Detect.anchors, models/yolo.py:230); ny / nx / balance:
 * host arrays [nl].  out: device [4 + nl] = loss (already times the batch size, :231), lbox, lobj, lcls (the
 * reference's loss_items), then the objectness BCE mean of every level (what autobalance reads, :224).
 * Duplicate (image, anchor, cell) matches resolve to the last one in the reference's row order (:205).
 * fl_gamma > 0: the BCE terms are wrapped in FocalLoss (:80-106, alpha 0.25).  slide_state != null: they are wrapped in
 * SlideLoss (:38-76); slide_state is a caller-owned, zero-initialised DEVICE array [4] = (ema of the class criterion,
 * ema of the objectness criterion, has-value flags) that the call reads and updates, like the `ema` attributes of the
 * reference's two SlideLoss objects.  The two wrappers exclude each other (in the reference too). */
size_t ecsy_yolo_loss_ws_bytes(int nl, int64_t N, int na, int64_t nt, const int* ny, const int* nx);
int ecsy_yolo_loss(const float* const* p, float* const* gp, const float* targets, int64_t nt, const float* anchors,
                   int nl, int64_t N, int na, int nc, const int* ny, const int* nx, const float* balance, float box,
                   float obj, float cls, float cls_pw, float obj_pw, float cp, float cn, float anchor_t, float gr,
                   float fl_gamma, float* slide_state, float* out, void* ws, size_t ws_bytes, void* stream);

/* Stem convolution of the real-valued image, fast precision (one bf16 plane): Conv_1 / Snn_Conv2d on a non-spike input,
 * /root/reference/models/common.py:409-425 with :609-624 (F.conv2d per timestep; a T-broadcast image is convolved once).
 * x: device [imgs][H][W][Cin] fp32 (Cin <= 4); w_stem: device bf16 [64][ceil(k/2) * 64], entry (co, kb*64 + part*32 + kx*4 + ci)
 * = W[co][ci][2*kb + part][kx], zero elsewhere; out: [imgs][Ho][Wo][64] fp32 = conv * scale + shift (scale / shift optional,
 * per output channel: the folded tdBN of inference).  ecsy_stem_conv_supported: Cin <= 4, Cout == 64, k <= 8, splits == 1. */
int ecsy_stem_conv_supported(int Cin, int Cout, int k, int splits);
size_t ecsy_stem_conv_ws_bytes(int64_t imgs, int H, int W);
int ecsy_stem_conv(const float* x, int64_t imgs, int H, int W, int Cin, const void* w_stem, float* out, const float* scale,
                   const float* shift, int Cout, int k, int stride, int pad, void* ws, size_t ws_bytes, void* stream);

/* ---- Stack-B training loss, forward + gradient in one call: replaces ComputeLoss.__call__ of
 * utils/loss_tal.py:162-215 with TaskAlignedAssigner (utils/tal/assigner.py:51-179; topk 10, alpha 0.5, beta 6, CIoU
 * overlaps), BCE class term with pos_weight, the box term as the reference evaluates it (GIoU: utils/metrics2.py:282
 * keeps the requested SIoU branch unreachable) and the distribution focal loss, reg_max = 16; fl_gamma > 0 wraps the
 * class BCE in FocalLoss (utils/loss_tal.py:32-60, 116-119).
 * feats / gfeats: HOST arrays of nl DEVICE pointers to the raw DDetect training outputs [N][64 + nc][ny_l][nx_l]
 * (models/yolo_snn.py:117-119) and their gradients (gfeats or gfeats[l] may be null: forward only; otherwise
 * OVERWRITTEN, upstream gradient 1).  targets: device [nt][6] = (image, class, cx, cy, w, h) normalised; ny / nx /
 * strides: host arrays [nl]; gains: 7.5 / 0.5 / 1.5 in the reference (:210-212); topk / alpha / beta: the assigner's
 * hyper-parameters (10 / 0.5 / 6.0 unless the YOLOM / YOLOA / YOLOB environment variables say otherwise, :134-137).
 * out: device [6] = loss (times the batch size), box, cls, dfl (the reference's loss_items), number of foreground
 * anchors, target_scores.sum().  Ties the reference leaves to torch.topk / argmax resolve to the lowest index. */
size_t ecsy_tal_loss_ws_bytes(int nl, int64_t N, int64_t nt, const int* ny, const int* nx);
int ecsy_tal_loss(const float* const* feats, float* const* gfeats, const float* targets, int64_t nt, int nl, int64_t N,
                  int nc, const int* ny, const int* nx, const float* strides, float cls_pw, float gain_box,
                  float gain_cls, float gain_dfl, float fl_gamma, int topk, float alpha, float beta, float* out, void* ws,
                  size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECSY_H_ */
