"""CPU oracle for the post-decode / training-step rows of SURVEY section 8f -- TEST INFRASTRUCTURE ONLY.

Plain torch / Python restatements of
  * `non_max_suppression` (utils/general.py:649-741; its inner torchvision.ops.nms restated from the published CPU
    algorithm, torchvision/csrc/ops/cpu/nms_kernel.cpp: stable descending sort, greedy scan, IoU = inter / (a + b -
    inter) > threshold) -- torchvision is a third-party dependency of the reference, not vendored in /root/reference;
  * the SGD-Nesterov step of train.py:259-287, 570-582 (torch.optim.SGD, three parameter groups) and `ModelEMA.update`
    (utils/torch_utils.py:306-316).
Pinned by tests/test_oracle_golden.py against fixtures produced by the UNMODIFIED reference functions
(oracle/gen_golden_post.py).  Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

MAX_WH = 4096      # utils/general.py:665
MAX_NMS = 30000    # utils/general.py:666


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """utils/general.py:603-609"""
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def greedy_nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """torchvision.ops.nms on CPU, restated: indices of the kept boxes in descending score order."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True).indices
    b = boxes[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = torch.zeros(n, dtype=torch.bool)
    keep: List[int] = []
    thr = float(iou_thres)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        xx1 = torch.maximum(x1[i], x1[i + 1:])
        yy1 = torch.maximum(y1[i], y1[i + 1:])
        xx2 = torch.minimum(x2[i], x2[i + 1:])
        yy2 = torch.minimum(y2[i], y2[i + 1:])
        w = (xx2 - xx1).clamp(min=0)
        h = (yy2 - yy1).clamp(min=0)
        inter = w * h
        ovr = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= ovr.double() > thr
    return order[torch.tensor(keep, dtype=torch.long)]


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        max_det: int = 300, use_torchvision: bool = False) -> List[torch.Tensor]:
    """utils/general.py:649-741 without the autolabelling / merge-NMS / time-limit branches.  ``use_torchvision``
    calls torchvision.ops.nms for the inner step exactly like the reference (the timing leg of tests/diag/post_bench.py;
    works on CUDA tensors too); the default is the pure-torch restatement of it."""
    nc = prediction.shape[2] - 5
    xc = prediction[..., 4] > conf_thres
    multi_label &= nc > 1
    output = [torch.zeros((0, 6), device=prediction.device)] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]].clone()
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[i], x[i, j + 5, None], j[:, None].float()), 1)
        else:
            conf, j = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        elif n > MAX_NMS:
            x = x[x[:, 4].argsort(descending=True)[:MAX_NMS]]
        c = x[:, 5:6] * (0 if agnostic else MAX_WH)
        boxes, scores = x[:, :4] + c, x[:, 4]
        if use_torchvision:
            import torchvision
            i = torchvision.ops.nms(boxes, scores, iou_thres)
        else:
            i = greedy_nms(boxes, scores, iou_thres)
        if i.shape[0] > max_det:
            i = i[:max_det]
        output[xi] = x[i]
    return output


# --------------------------------------------------------------------------------------
# optimizer step + EMA (train.py:259-287, 570-582; utils/torch_utils.py:285-316)
# --------------------------------------------------------------------------------------
def sgd_nesterov_step(p: torch.Tensor, g: torch.Tensor, buf: Optional[torch.Tensor], lr: float, momentum: float,
                      weight_decay: float, nesterov: bool = True):
    """One torch.optim.SGD update of one tensor (dampening 0): returns (p_new, buf_new)."""
    d = g
    if weight_decay != 0:
        d = d + weight_decay * p
    if momentum != 0:
        buf = d.clone() if buf is None else buf * momentum + d
        d = d + momentum * buf if nesterov else buf
    return p - lr * d, buf


def ema_decay(updates: int, decay: float = 0.9999) -> float:
    """utils/torch_utils.py:299: decay * (1 - exp(-updates / 2000))"""
    return decay * (1 - math.exp(-updates / 2000))


def ema_update(ema: Dict[str, torch.Tensor], model: Dict[str, torch.Tensor], d: float) -> None:
    """utils/torch_utils.py:311-316: every floating-point state_dict entry (parameters AND buffers)."""
    for k, v in ema.items():
        if v.dtype.is_floating_point:
            v *= d
            v += (1 - d) * model[k].detach()


# --------------------------------------------------------------------------------------
# Gen1 event -> frame input path (g1-resnet/utils/give_g1_data.py:550-565 `create_data`; the loader's per-frame
# cv2.resize, g1-resnet/utils/datasets_g1T.py:518-533; `imgs.float() / 255`, g1-resnet/train_g1.py:298).
# cv2 (OpenCV 4.x, a third-party dependency that is not vendored in /root/reference) resizes uint8 images with
# INTER_LINEAR in 11-bit fixed point: restated here from its published algorithm (imgproc/src/resize.cpp:
# resizeGeneric_ / HResizeLinear / VResizeLinear<uchar>) and pinned bit-exactly on cv2 outputs.
# --------------------------------------------------------------------------------------
EV_W, EV_H = 304, 240


def events_to_frames(samples, T: int) -> torch.Tensor:
    """samples[n][t] = dict(x, y, p) in sensor order -> uint8 [N, T, 240, 304]: grey 127, an event paints 255*p, the
    LAST event of a pixel wins (numpy fancy assignment in event order).  The 3 channels are identical."""
    out = torch.full((len(samples), T, EV_H, EV_W), 127, dtype=torch.uint8)
    for n, bins in enumerate(samples):
        for t in range(T):
            b = bins[t]
            for x, y, p in zip(b["x"].tolist(), b["y"].tolist(), b["p"].tolist()):
                out[n, t, y, x] = 255 * p
    return out


def _linear_coeffs(ssize: int, dsize: int, reset: bool):
    import numpy as np
    scale = 1.0 / (float(dsize) / float(ssize))           # resize.cpp: scale_x = 1. / inv_scale_x
    idx = np.zeros(dsize, np.int64)
    a = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if reset:                                          # x only: border columns take the edge pixel with weight 1
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        idx[d] = s
        a[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        a[d, 1] = int(np.rint(f * np.float32(2048)))
    return idx, a


def resize_linear_u8(img: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    """cv2.resize(img, (dw, dh)) (INTER_LINEAR) for a uint8 [H, W] image, bit-exact."""
    import numpy as np
    src = img.numpy().astype(np.int64)
    sh, sw = src.shape
    xi, xa = _linear_coeffs(sw, dw, True)
    yi, ya = _linear_coeffs(sh, dh, False)                 # y: rows are clamped, the weights are not reset
    x1 = np.minimum(xi + 1, sw - 1)
    H = src[:, xi] * xa[:, 0][None, :] + src[:, x1] * xa[:, 1][None, :]
    S0, S1 = H[np.clip(yi, 0, sh - 1)], H[np.clip(yi + 1, 0, sh - 1)]
    b0, b1 = ya[:, 0][:, None], ya[:, 1][:, None]
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return torch.from_numpy(out.astype(np.uint8))


def event_frames(samples, T: int, out_hw: int) -> torch.Tensor:
    """The network input of the Gen1 detector: float32 [T, N, 3, S, S] = resized frames / 255."""
    fr = events_to_frames(samples, T)
    N = fr.shape[0]
    res = torch.stack([torch.stack([resize_linear_u8(fr[n, t], out_hw, out_hw) for t in range(T)]) for n in range(N)])
    x = res.float() / 255                                  # [N, T, S, S]
    return x.permute(1, 0, 2, 3).unsqueeze(2).expand(-1, -1, 3, -1, -1).contiguous()
