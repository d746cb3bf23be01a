"""Golden vectors for the token producer (ln_post + proj + CLS/patch split) from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_tokens_golden.py

Imports modeling/clip/model.py from /root/reference (read-only; the package __init__ files, which pull in ftfy / timm, are
skipped through pre-seeded namespace packages), builds the reference's own `VisionTransformer` (one residual block, random
init under a fixed seed), runs it on seeded images and captures, at the tail this repository replaces
(clip/model.py:485-487), the INPUT of `ln_post`, the tower's output, and for a seeded cotangent the gradients autograd
returns for that input and for ln_post.weight / ln_post.bias / proj -- in fp32 (the reference's LayerNorm subclass always computes in
fp32) next to an fp64 evaluation of the same torch ops on the same inputs (pins the values; `dev32/*` = what the
reference's fp32 arithmetic loses against it).  /root/reference does not exist on the GPU box; only the .npz files travel.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from signal_b200 import synthetic as syn  # noqa: E402  (probe vectors for the fingerprints)

CASES = {
    # name: grid h x w, tower width, heads, output dim, batch, seed; full = store full gradient tensors
    # `tower`: the input of ln_post comes out of the reference's own VisionTransformer (everything stored);
    # otherwise seeded synthetic inputs (tests regenerate them with `synthetic_inputs`) go through the reference's LayerNorm
    # subclass and the tail's two statements, and only fingerprints / small vectors of the results are stored.
    "small": dict(h=16, w=8, width=64, heads=2, out=32, B=2, seed=21, tower=True),
    "vitb_vehicle": dict(h=8, w=16, width=768, heads=12, out=512, B=4, seed=22, tower=False),
}


def synthetic_inputs(c):
    """Seeded ln_post input with per-row offsets and scales (so that the statistics matter), affine parameters away from
    (1, 0), CLIP-style proj init (clip/model.py:441-445), cotangent."""
    g = torch.Generator().manual_seed(c["seed"])
    B, L1, W, D = c["B"], c["h"] * c["w"] + 1, c["width"], c["out"]
    x = torch.randn(B, L1, W, generator=g) * (0.5 + torch.rand(B, L1, 1, generator=g) * 2.0) + 0.7 * torch.randn(B, L1, 1, generator=g)
    ln_w = 1.0 + 0.3 * torch.randn(W, generator=g)
    ln_b = 0.2 * torch.randn(W, generator=g)
    proj = W ** -0.5 * torch.randn(W, D, generator=g)
    cot = torch.randn(B, L1, D, generator=g)
    return dict(x=x, ln_w=ln_w, ln_b=ln_b, proj=proj, cot=cot)


def import_reference():
    for name, path in (("modeling", "/root/reference/modeling"), ("modeling.clip", "/root/reference/modeling/clip"),
                       ("modeling.backbones", "/root/reference/modeling/backbones")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    sys.path.insert(0, "/root/reference")
    from modeling.clip.model import VisionTransformer
    return VisionTransformer


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def fingerprint(key, g):
    g = g.detach().double().reshape(-1)
    return np.array([g.norm().item()] + (syn.probe_vector(key, g.numel()) @ g).tolist(), dtype=np.float64)


def run(VT, c, dtype):
    torch.manual_seed(c["seed"])
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(PROMPT=False, ADAPTER=False))
    vt = VT(c["h"], c["w"], 16, 16, c["width"], 1, c["heads"], c["out"], cfg)
    with torch.no_grad():   # non-trivial affine parameters (the default init is weight 1, bias 0)
        g = torch.Generator().manual_seed(c["seed"] + 1)
        vt.ln_post.weight.copy_(1.0 + 0.3 * torch.randn(c["width"], generator=g))
        vt.ln_post.bias.copy_(0.2 * torch.randn(c["width"], generator=g))
    vt = vt.to(dtype)
    g = torch.Generator().manual_seed(c["seed"] + 2)
    img = torch.randn(c["B"], 3, 16 * c["h"], 16 * c["w"], generator=g).to(dtype)
    cot = torch.randn(c["B"], c["h"] * c["w"] + 1, c["out"], generator=g).to(dtype)
    grabbed = {}

    def pre_hook(_mod, args):
        x = args[0]
        x.retain_grad()
        grabbed["x"] = x

    h = vt.ln_post.register_forward_pre_hook(pre_hook)
    y = vt(img)
    h.remove()
    y.backward(cot)
    x = grabbed["x"]
    return dict(x=x.detach(), ln_w=vt.ln_post.weight.detach(), ln_b=vt.ln_post.bias.detach(), proj=vt.proj.detach(), cot=cot,
                tokens=y.detach(), dx=x.grad.detach(), d_ln_w=vt.ln_post.weight.grad.detach(), d_ln_b=vt.ln_post.bias.grad.detach(),
                d_proj=vt.proj.grad.detach(), eps=vt.ln_post.eps)


def main():
    VT = import_reference()
    for name, c in CASES.items():
        # the reference tower as it is: its LayerNorm subclass always computes in fp32 (clip/model.py:157-160), so the
        # reference itself only exists in fp32
        full = c["tower"]
        if full:
            r32 = run(VT, c, torch.float32)
            out = {"eps": np.float64(r32["eps"])}
            for k in ("x", "ln_w", "ln_b", "proj", "cot"):
                out[k] = r32[k].numpy()
        else:
            r32 = None
            out = {k: v.numpy() for k, v in synthetic_inputs(c).items()}
            out["eps"] = np.float64(1e-5)
        # fp64 pin: the torch ops the reference's tail calls (F.layer_norm, matmul), evaluated in fp64 on the stored inputs
        from modeling.clip.model import LayerNorm
        ln = torch.nn.LayerNorm(c["width"]).double()
        with torch.no_grad():
            ln.weight.copy_(torch.from_numpy(out["ln_w"]).double())
            ln.bias.copy_(torch.from_numpy(out["ln_b"]).double())
        x = torch.from_numpy(out["x"]).double().requires_grad_(True)
        proj = torch.from_numpy(out["proj"]).double().requires_grad_(True)
        xn = ln(x)
        y = xn @ proj                                   # clip/model.py:485-487
        x_cash, global_feat = y[:, 1:], y[:, 0]         # meta_arch.py:108-110
        y.backward(torch.from_numpy(out["cot"]).double())
        ref = dict(tokens=y.detach(), dx=x.grad, d_ln_w=ln.weight.grad, d_ln_b=ln.bias.grad, d_proj=proj.grad,
                   patch_mean=x_cash.detach().mean(dim=1), cls=global_feat.detach())
        # the same tail in fp32 (deviation of the reference's own fp32 arithmetic)
        ln32 = LayerNorm(c["width"])
        with torch.no_grad():
            ln32.weight.copy_(torch.from_numpy(out["ln_w"]))
            ln32.bias.copy_(torch.from_numpy(out["ln_b"]))
        x32 = torch.from_numpy(out["x"]).requires_grad_(True)
        p32 = torch.from_numpy(out["proj"]).requires_grad_(True)
        y32 = ln32(x32) @ p32
        y32.backward(torch.from_numpy(out["cot"]))
        ref32 = dict(tokens=y32.detach(), dx=x32.grad, d_ln_w=ln32.weight.grad, d_ln_b=ln32.bias.grad, d_proj=p32.grad)
        # the tail alone reproduces what the whole reference tower returned and back-propagated (bit for bit: same ops)
        for k in ("tokens", "dx", "d_ln_w", "d_ln_b", "d_proj"):
            if r32 is not None:
                assert rel(ref32[k], r32[k]) < 1e-6, (k, rel(ref32[k], r32[k]))
            assert rel(ref32[k], ref[k]) < 1e-4, (k, rel(ref32[k], ref[k]))
        for k, v in ref.items():
            if full or k in ("d_ln_w", "d_ln_b", "patch_mean", "cls"):
                out["ref/" + k] = v.numpy()
            else:
                out["fp/" + k] = fingerprint(k, v)
        for k, v in ref32.items():
            out["dev32/" + k] = np.float64(rel(v, ref[k]))
        if not full:   # inputs are regenerated from the seed by the tests
            for k in ("x", "ln_w", "ln_b", "proj", "cot"):
                del out[k]
        np.savez_compressed(os.path.join(HERE, f"tokens_{name}.npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
