"""CPU: the C-ABI library builds, loads, and exports every symbol include/signal_b200.h declares.
(No compute calls here: those need a GPU and live in the -m gpu tests.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "signal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sig_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("sig_sim_fwd", "sig_sim_bwd", "sig_align_fwd", "sig_align_bwd", "sig_das_fwd", "sig_das_bwd",
                 "sig_volume3_fwd", "sig_volume3_bwd", "sig_sim_select_from_scores", "sig_error_string", "sig_ctx_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import lib
    handle = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == _declared(), "lib.EXPORTS and the header disagree"
    assert lib.load().sig_version() == 1
    assert b"NULL" in lib.load().sig_error_string(-1)
    assert lib.load().sig_ctx_bytes(0, 128, 128, 768, 1, 0) > 0
    assert lib.load().sig_ctx_bytes(0, 128, 129, 768, 1, 0) == 0      # L > 128 is rejected
    assert lib.load().sig_ctx_bytes(1, 128, 128, 100, 0, 0) == 0      # d % 64 != 0 is rejected


def test_product_path_refuses_cpu_tensors():
    import torch
    from signal_b200 import modules as M
    sim = M.Select_Interactive_Module(64, k=8)
    x = torch.randn(2, 129, 64)
    with pytest.raises(RuntimeError):
        sim(x[:, 1:], x[:, 1:], x[:, 1:], x[:, 0], x[:, 0], x[:, 0])


def test_ctx_sizing_uses_the_dispatch_predicate():
    """sig_ctx_bytes must size the ctx for the path the entry point will take (bf16 with d > 768 falls to the
    SIMT path, which needs the larger buffer)."""
    import __graft_entry__ as entry
    entry.build()
    from signal_b200 import lib
    L_ = lib.load()
    B, L = 16, 128
    for d in (64, 512, 768):
        tc = L_.sig_ctx_bytes(lib.CTX_ALIGN, B, L, d, lib.SIG_BF16, 0)
        simt = L_.sig_ctx_bytes(lib.CTX_ALIGN, B, L, d, lib.SIG_BF16, lib.FLAG_FORCE_SIMT)
        assert 0 < tc and 0 < simt and simt == L_.sig_ctx_bytes(lib.CTX_ALIGN, B, L, d, lib.SIG_F32, 0)
    for d in (832, 1024):
        assert L_.sig_ctx_bytes(lib.CTX_ALIGN, B, L, d, lib.SIG_BF16, 0) == L_.sig_ctx_bytes(lib.CTX_ALIGN, B, L, d, lib.SIG_F32, 0)
    # L != 128 never runs on the tensor-core path either
    assert L_.sig_ctx_bytes(lib.CTX_ALIGN, B, 64, 512, lib.SIG_BF16, 0) == L_.sig_ctx_bytes(lib.CTX_ALIGN, B, 64, 512, lib.SIG_F32, 0)
