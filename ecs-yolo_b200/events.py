"""Gen1 event-camera input on the GPU: the reference bins events into T ternary frames on
the CPU (g1-resnet/utils/give_g1_data.py:550-565 `create_data`), resizes each with cv2 in the data loader
(g1-resnet/utils/datasets_g1T.py:518-533) and normalises in the training loop (g1-resnet/train_g1.py:298); here the raw
events of a whole batch go to the device once and one C-ABI call (`ecsy_event_frames`) produces the model input."""
from __future__ import annotations

import torch

from . import _cabi
from .functional import _chk_cuda, _p, _st, _timed

SENSOR_H, SENSOR_W = 240, 304     # Gen1 (give_g1_data.py:551-552)


def event_frames(x: torch.Tensor, y: torch.Tensor, p: torch.Tensor, frame: torch.Tensor, N: int, T: int,
                 out_hw=(320, 320), sensor_hw=(SENSOR_H, SENSOR_W), check: bool = False) -> torch.Tensor:
    """x, y, p, frame: int32 CUDA vectors over all events of the batch in sensor order (frame = sample * T + bin).
    -> float32 [T, N, 3, Ho, Wo] view over NHWC memory, ready for ``Model._forward_once``.  ``check`` raises on
    out-of-sensor events like the reference's asserts (one device-to-host read)."""
    _chk_cuda(x, y, p, frame)
    ev = [t.to(torch.int32).contiguous() for t in (x, y, p, frame)]
    n = ev[0].numel()
    if any(t.numel() != n for t in ev):
        raise ValueError("event arrays differ in length")
    dev = ev[0].device
    H, W = sensor_hw
    Ho, Wo = out_hw
    out = torch.empty(T, N, Ho, Wo, 3, device=dev, dtype=torch.float32)
    oob = torch.empty(1, device=dev, dtype=torch.int32)
    L = _cabi.lib()
    ws = torch.empty(L.ecsy_event_frames_ws_bytes(N, T, H, W), device=dev, dtype=torch.uint8)
    with _timed("event_frames", 2):
        _cabi.check(L.ecsy_event_frames(_p(ev[0]), _p(ev[1]), _p(ev[2]), _p(ev[3]), n, N, T, H, W, Ho, Wo, _p(out), _p(oob),
                                        _p(ws), ws.numel(), _st()), "event_frames")
    if check and int(oob) != 0:
        raise AssertionError(f"out of bound events: {int(oob)} outside the {W}x{H} sensor / {N * T} frames")
    return out.permute(0, 1, 4, 2, 3)
