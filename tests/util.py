import importlib
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ecsy():
    return importlib.import_module("ecs-yolo_b200")


def load_golden(name):
    import seeded as S
    return torch.load(os.path.join(S.GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def agree(a, b):
    return float((a.cpu() == b.cpu()).float().mean())


def nhwc(x):
    """[T,N,C,H,W] cpu -> cuda NHWC Act-ready view (reference-shaped, NHWC memory)."""
    return x.cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


class forced_spikes:
    """Teacher-forcing INSIDE a block (SURVEY 8c tier 1, "feed every module its recorded input"): while active, every
    mem_update.spikes() call of `root` (a) records how many positions of OUR spikes equal the oracle's recorded spikes for
    that neuron and (b) hands the ORACLE's spikes to the consumer, so the convs / tdBNs downstream are compared on identical
    inputs and a near-threshold flip cannot masquerade as an arithmetic error (or hide one).
    `ref_spikes`: {module name (state_dict prefix without the trailing dot): [T,N,C,H,W] {0,1} tensor}; `prefix` is put in
    front of root's own module names ("model.0." for a bare block, "" for a whole model)."""

    def __init__(self, E, root, ref_spikes, prefix="", force=True, gate_spikes=None):
        """gate_spikes (optional): the spikes the agreement is MEASURED against when they differ from the forced ones --
        fast precision: the oracle neuron applied to the bf16-weight conv output it actually receives (the forced spikes
        stay those of the fp32 run, so producer and consumer see the same tensors in both implementations).  The
        agreement with the forced (fp32-run) spikes is then kept in `agree_fp32`."""
        self.E, self.ref, self.force = E, ref_spikes, force
        self.gate = gate_spikes
        self.names = {id(mod): prefix + n for n, mod in root.named_modules() if isinstance(mod, E.common.mem_update)}
        self.agree, self.rate, self.agree_fp32 = {}, {}, {}

    def __enter__(self):
        E, outer = self.E, self
        self.orig = orig = E.common.mem_update.spikes

        def patched(mod, x, affine=None):
            sp = orig(mod, x, affine)
            name = outer.names.get(id(mod))
            if name is None or name not in outer.ref:
                return sp
            ref = outer.ref[name]
            got = sp.to_act().to_ref()
            refc = ref.to(got.device)
            outer.agree_fp32[name] = float((got == refc).float().mean())
            outer.agree[name] = outer.agree_fp32[name]
            if outer.gate is not None and name in outer.gate:
                outer.agree[name] = float((got == outer.gate[name].to(got.device)).float().mean())
            outer.rate[name] = float(refc.mean())
            if not outer.force:
                return sp
            F = E.functional
            Cp = sp.C
            a = F.Act.from_ref(refc)
            if Cp != a.C:       # narrow layers run zero-padded to the 64-channel granule
                a = F.Act(F.pad_channels(a.data, Cp), a.T)
            forced = F.Spikes.from_act(a)
            return F.Spikes(forced.bits, Cp, sp.Cr)
        E.common.mem_update.spikes = patched
        return self

    def __exit__(self, *exc):
        self.E.common.mem_update.spikes = self.orig
        return False


def quantize_weights_bf16(sd):
    """The fast-precision operand set: every Snn_Conv2d weight and every point-wise spread weight (spread.1) rounded to
    bf16 (the single weight plane the tensor cores see); depth-wise spread weights, biases and tdBN parameters stay fp32
    (they are fp32 in the kernels too)."""
    out = {}
    for k, v in sd.items():
        biased = (k[:-len("weight")] + "bias") in sd      # Detect / DDetect output convs: fp32 SIMT path, not rounded
        if v.dim() == 4 and k.endswith("weight") and ".bn." not in k and ("spread.1." in k or not biased):
            out[k] = v.bfloat16().float()
        else:
            out[k] = v.clone()
    return out
