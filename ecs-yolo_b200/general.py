"""Post-decode step of the inference path: `non_max_suppression` with the reference's signature
(utils/general.py:649-741) on the GPU, all images of the batch in one C-ABI call (`ecsy_nms`: candidate filter +
in-CTA sort + greedy scan per image) instead of a Python loop over images around torchvision.ops.nms."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from .functional import _chk_cuda, _p, _st, _timed

MAX_NMS = 30000  # utils/general.py:666


def nms_padded(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
               classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
               max_det: int = 300) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (det [N, max_det, 6] = (x1, y1, x2, y2, conf, cls), count [N] int32); rows >= count[n] are undefined.
    No host synchronisation: usable inside a CUDA graph / a pipelined evaluation loop."""
    _chk_cuda(prediction)
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise ValueError(f"expected [N, rows, 5 + nc], got {tuple(prediction.shape)}")
    if not (0 <= conf_thres <= 1):
        raise AssertionError(f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0')
    if not (0 <= iou_thres <= 1):
        raise AssertionError(f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0')
    pred = prediction.float().contiguous()
    N, R, no = pred.shape
    nc = no - 5
    dev = pred.device
    out = torch.empty(N, max_det, 6, device=dev, dtype=torch.float32)
    cnt = torch.empty(N, device=dev, dtype=torch.int32)
    if N == 0:
        return out, cnt
    ok = None
    if classes is not None:
        ok = torch.zeros(nc, dtype=torch.uint8)
        ok[[int(c) for c in classes if 0 <= int(c) < nc]] = 1
        ok = ok.to(dev)
    L = _cabi.lib()
    ml = 1 if (multi_label and nc > 1) else 0
    ws = torch.empty(L.ecsy_nms_ws_bytes(N, R, nc, ml), device=dev, dtype=torch.uint8)
    with _timed("nms", 2):
        _cabi.check(L.ecsy_nms(_p(pred), N, R, nc, float(conf_thres), float(iou_thres), 1 if agnostic else 0, ml, _p(ok),
                               int(max_det), MAX_NMS, _p(out), _p(cnt), _p(ws), ws.numel(), _st()), "nms")
    return out, cnt


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300) -> List[torch.Tensor]:
    """Drop-in for utils/general.py:649 `non_max_suppression`: list of (n, 6) tensors [xyxy, conf, cls] per image.
    The list form needs the per-image counts on the host (one small device-to-host copy)."""
    if labels:
        raise NotImplementedError("autolabelling (`labels=`) is a training-time validation feature outside the hot path")
    out, cnt = nms_padded(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det)
    counts = cnt.tolist()
    return [out[i, :c] for i, c in enumerate(counts)]
