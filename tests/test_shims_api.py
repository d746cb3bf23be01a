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
