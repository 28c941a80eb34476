"""GPU parity of the section-8f rows (post-decode NMS, fused SGD + EMA step) against the golden outputs of the
unmodified reference functions and the CPU oracle (oracle/post_oracle.py)."""
import copy
import os

import pytest
import torch

import post_oracle as P
import seeded as S
from util import ecsy

pytestmark = pytest.mark.gpu


def _load(name):
    return torch.load(os.path.join(S.GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", list(S.NMS_CASES))
def test_nms_golden(name):
    """Bit-exact: same boxes, same order, same values as utils.general.non_max_suppression."""
    E = ecsy()
    gold = _load("post_nms")[name]
    spec = S.NMS_CASES[name]
    pred = S.nms_inputs(spec)
    out = E.general.non_max_suppression(pred.cuda(), spec["conf"], spec["iou"], classes=spec.get("classes"),
                                        agnostic=spec.get("agnostic", False), multi_label=spec.get("multi_label", False),
                                        max_det=spec.get("max_det", 300))
    assert len(out) == len(gold["out"])
    for i, (a, b) in enumerate(zip(out, gold["out"])):
        assert tuple(a.shape) == tuple(b.shape), (name, i, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), (name, i)


def test_nms_full_size_properties():
    """BASELINE shape (batch 64, 6000 rows, nc = 13): size-independent properties -- descending confidence, every
    kept pair of one class overlaps <= the threshold, idempotence (NMS of the kept boxes keeps all of them), and the
    first images equal the CPU oracle."""
    E = ecsy()
    spec = dict(N=64, R=6000, nc=13, seed=777)
    pred = S.nms_inputs(spec)
    conf, iou = 0.25, 0.45
    out, cnt = E.general.nms_padded(pred.cuda(), conf, iou)
    cnt = cnt.cpu()
    out = out.cpu()
    want = P.non_max_suppression(pred[:3].clone(), conf, iou)
    for i in range(3):
        assert int(cnt[i]) == want[i].shape[0] and torch.equal(out[i, :cnt[i]], want[i])
    for i in range(spec["N"]):
        d = out[i, :cnt[i]]
        assert bool((d[1:, 4] <= d[:-1, 4]).all()) and bool((d[:, 4] > conf).all())
        boxes = d[:, :4] + d[:, 5:6] * 4096
        keep = P.greedy_nms(boxes, d[:, 4], iou)
        assert keep.numel() == d.shape[0]
    # idempotence through the kernel itself: feed the kept boxes back as predictions (obj = conf, one-hot class)
    d = out[0, :cnt[0]]
    back = torch.zeros(1, d.shape[0], 5 + spec["nc"])
    back[0, :, 0] = (d[:, 0] + d[:, 2]) / 2
    back[0, :, 1] = (d[:, 1] + d[:, 3]) / 2
    back[0, :, 2] = d[:, 2] - d[:, 0]
    back[0, :, 3] = d[:, 3] - d[:, 1]
    back[0, :, 4] = d[:, 4]
    back[0, torch.arange(d.shape[0]), 5 + d[:, 5].long()] = 1.0
    again = E.general.non_max_suppression(back.cuda(), conf * 0.5, iou)[0]
    assert again.shape[0] == d.shape[0]


def test_nms_rejects_bad_arguments():
    E = ecsy()
    with pytest.raises(AssertionError):
        E.general.non_max_suppression(torch.zeros(1, 4, 8).cuda(), conf_thres=1.5)
    with pytest.raises(RuntimeError):
        E.general.non_max_suppression(torch.zeros(1, 4, 8))          # CPU tensor: no fallback
    assert E.general.non_max_suppression(torch.zeros(2, 4, 8).cuda())[1].shape == (0, 6)
