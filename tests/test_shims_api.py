"""CPU: the import-path shims resolve to the B200 classes, and the drop-in classes keep the reference's constructor /
call signatures (compared with the live reference when /root/reference is present -- it is not on the GPU box)."""
import importlib
import inspect
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "signal_b200", "shims")
REF = "/root/reference"


def _fresh(names):
    for n in list(sys.modules):
        if n.split(".")[0] in names:
            del sys.modules[n]


def test_shims_resolve_to_b200_classes():
    _fresh({"layers", "modeling", "utils"})
    sys.path.insert(0, SHIMS)
    try:
        from signal_b200 import losses, modules
        sl = importlib.import_module("layers.softmax_loss")
        tl = importlib.import_module("layers.triplet_loss")
        ua = importlib.import_module("modeling.AddModule.useA")
        ub = importlib.import_module("modeling.AddModule.useB")
        assert sl.CrossEntropyLabelSmooth is losses.CrossEntropyLabelSmooth
        assert sl.LabelSmoothingCrossEntropy is losses.LabelSmoothingCrossEntropy
        assert tl.TripletLoss is losses.TripletLoss
        assert ua.Select_Interactive_Module is modules.Select_Interactive_Module
        assert ub.AlignmentM is modules.AlignmentM
        uv = importlib.import_module("utils.volume")
        um = importlib.import_module("utils.metrics")
        from signal_b200 import evaluation
        assert uv.volume_computation3 is modules.volume_computation3 and uv.volume_computation4 is modules.volume_computation4
        assert uv.volume_computation5 is modules.volume_computation5
        assert um.R1_mAP_eval is evaluation.R1_mAP_eval and um.eval_func is evaluation.eval_func
    finally:
        sys.path.remove(SHIMS)
        _fresh({"layers", "modeling", "utils"})


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the live reference is only present in the build container")
def test_signatures_match_the_reference():
    _fresh({"layers", "modeling", "utils"})
    sys.dont_write_bytecode = True      # /root/reference is read-only
    for pkg in ("layers", "modeling"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, pkg)]
        sys.modules[pkg] = m
    sys.path.insert(0, REF)
    try:
        from layers.softmax_loss import CrossEntropyLabelSmooth as RX, LabelSmoothingCrossEntropy as RL
        from layers.triplet_loss import TripletLoss as RT
        from modeling.AddModule.useA import Select_Interactive_Module as RS, TokenSelection as RTS, ModalInteractive as RMI
        from modeling.AddModule.useB import AlignmentM as RA
        from signal_b200 import losses, modules
        pairs = [(RX, losses.CrossEntropyLabelSmooth), (RL, losses.LabelSmoothingCrossEntropy), (RT, losses.TripletLoss),
                 (RS, modules.Select_Interactive_Module), (RTS, modules.TokenSelection), (RMI, modules.ModalInteractive),
                 (RA, modules.AlignmentM)]
        for ref, ours in pairs:
            assert _params(ref.__init__) == _params(ours.__init__), ref.__name__
        assert _params(RX.forward) == _params(losses.CrossEntropyLabelSmooth.forward)
        assert _params(RT.__call__) == _params(losses.TripletLoss.__call__)
        assert _params(RS.forward) == _params(modules.Select_Interactive_Module.forward)
        assert _params(RA.forward) == _params(modules.AlignmentM.forward)
        assert _params(RA.Cls_Align) == _params(modules.AlignmentM.Cls_Align)
        assert _params(RA.patch_Align) == _params(modules.AlignmentM.patch_Align)
    finally:
        sys.path.remove(REF)
        _fresh({"layers", "modeling", "utils"})


@pytest.mark.skipif(not os.path.isdir(REF), reason="the live reference is only present in the build container")
def test_reference_state_dicts_load_into_the_drop_ins():
    """Checkpoint compatibility (processor.py:313-321, make_model.py:125-130): a state_dict of the reference's
    Select_Interactive_Module / AlignmentM loads strictly into the drop-ins -- same keys, same shapes, same dtypes -- and
    the drop-ins' own state_dict loads back into the reference modules."""
    import torch
    _fresh({"layers", "modeling", "utils"})
    sys.dont_write_bytecode = True
    for pkg in ("modeling", "utils"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, pkg)]
        sys.modules[pkg] = m
    sys.path.insert(0, REF)
    try:
        from modeling.AddModule.useA import Select_Interactive_Module as RS
        from modeling.AddModule.useB import AlignmentM as RA
        from signal_b200 import modules
        for k, keep in ((80, None), (112, 0.5)):
            ref, ours = RS(512, k=k, keep_ratio=keep), modules.Select_Interactive_Module(512, k=k, keep_ratio=keep)
            sd = ref.state_dict()
            assert list(sd) == list(ours.state_dict()) and all(sd[n].shape == v.shape and sd[n].dtype == v.dtype for n, v in ours.state_dict().items())
            ours.load_state_dict(sd, strict=True)
            ref.load_state_dict(ours.state_dict(), strict=True)
            assert ours.token_selection.k1 == ref.token_selection.k1 and ours.token_selection.k2 == ref.token_selection.k2
            assert ours.token_selection.keep_ratio == ref.token_selection.keep_ratio
        for hw in ((16, 8), (8, 16)):
            ref, ours = RA(512, *hw), modules.AlignmentM(512, *hw)
            sd = ref.state_dict()
            assert list(sd) == list(ours.state_dict()) and all(sd[n].shape == v.shape for n, v in ours.state_dict().items())
            ours.load_state_dict(sd, strict=True)
            ref.load_state_dict(ours.state_dict(), strict=True)
            assert (ours.h, ours.w, ours.feat_dim) == (ref.h, ref.w, ref.feat_dim)
        # same default initialisation under the same seed (constructor order of the sub-modules is the reference's)
        torch.manual_seed(1234)
        a = RS(512, k=80).state_dict()
        torch.manual_seed(1234)
        b = modules.Select_Interactive_Module(512, k=80).state_dict()
        assert all(torch.equal(a[n], b[n]) for n in a)
    finally:
        sys.path.remove(REF)
        _fresh({"layers", "modeling", "utils"})


@pytest.mark.skipif(not os.path.isdir(REF), reason="the live reference is only present in the build container")
def test_volume_and_tower_tail_signatures_match_the_reference():
    """utils/volume.py (all three functions) and the pieces TokenProducer wraps (clip/model.py: VisionTransformer.ln_post is
    the LayerNorm subclass, .proj a [width, output_dim] Parameter; meta_arch.py returns (x[:, 1:], x[:, 0]))."""
    _fresh({"layers", "modeling", "utils"})
    sys.dont_write_bytecode = True
    for pkg, sub in (("utils", "utils"), ("modeling", "modeling"), ("modeling.clip", "modeling/clip"), ("modeling.backbones", "modeling/backbones")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[pkg] = m
    sys.path.insert(0, REF)
    try:
        import torch
        from utils import volume as RV
        from modeling.clip.model import VisionTransformer
        from signal_b200 import modules
        from signal_b200.tokens import TokenProducer
        for name in ("volume_computation3", "volume_computation4", "volume_computation5"):
            assert _params(getattr(RV, name)) == _params(getattr(modules, name)), name
        cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(PROMPT=False, ADAPTER=False))
        vt = VisionTransformer(8, 16, 16, 16, 64, 1, 2, 32, cfg)
        tp = TokenProducer(vt.ln_post, vt.proj)            # wraps the reference's own sub-module and parameter
        assert isinstance(vt.ln_post, torch.nn.LayerNorm) and tuple(vt.proj.shape) == (64, 32)
        assert list(tp.state_dict()) == [] and list(tp.parameters()) == []      # owns nothing: checkpoints unchanged
        with pytest.raises(RuntimeError):
            tp(torch.randn(2, 129, 64))                   # CPU tensors raise: no fallback
    finally:
        sys.path.remove(REF)
        _fresh({"layers", "modeling", "utils"})


def test_new_drop_ins_refuse_cpu_tensors():
    import torch
    from signal_b200 import evaluation, modules
    with pytest.raises(RuntimeError):
        modules.volume_computation4(*[torch.randn(4, 16) for _ in range(4)])
    with pytest.raises(RuntimeError):
        evaluation.euclidean_distance(torch.randn(3, 8), torch.randn(5, 8))
