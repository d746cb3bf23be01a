"""GPU: the tcgen05/TMA GEMM core (signal_b200/csrc/tc_gemm.cu) against torch fp32 matmul on the
same bf16-rounded operands, for every operand mode the head uses."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(torch.bfloat16).cuda()


def _check(out, ref, tol=2e-3):
    err = (out.float() - ref).norm() / ref.norm()
    assert err < tol, float(err)


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("M,N,K", [(384, 512, 512), (24, 768, 768), (128, 96, 64), (1000, 200, 136)])
def test_kmajor_nt(M, N, K, bn):
    from signal_b200 import lib
    A, B = _rand((M, K), 1), _rand((N, K), 2)
    out = lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, bn=bn)
    _check(out, A.float() @ B.float().T)


def test_bias_gelu_bf16_out():
    from signal_b200 import lib
    M, N, K = 384, 1024, 512
    A, B = _rand((M, K), 3), _rand((N, K), 4)
    bias = torch.randn(N, device="cuda")
    out = lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, out_bf16=True, bias=bias, alpha=0.5, act=1)
    ref = torch.nn.functional.gelu(0.5 * (A.float() @ B.float().T) + bias)
    _check(out, ref, 6e-3)


@pytest.mark.parametrize("M,N,K,ks", [(384, 768, 768, 6), (384, 1536, 768, 4), (200, 776, 1536, 9), (12, 768, 768, 6)])
def test_split_k_with_bias_and_accumulate(M, N, K, ks):
    """split-K: partial tiles are added with fp32 atomics; the bias is added exactly once; a non-zero C is accumulated onto
    (the small M = 3B layers of SIM's MLP chain, csrc/sim_mlp_tc.inl::split_k)."""
    from signal_b200 import lib
    A, B = _rand((M, K), 31), _rand((N, K), 32)
    bias = torch.randn(N, device="cuda")
    out = lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, bias=bias, alpha=0.25, ksplit=ks)
    ref = 0.25 * (A.float() @ B.float().T) + bias
    _check(out, ref)
    base = torch.randn(M, N, device="cuda")
    out2 = lib.debug_gemm_bf16(A, 0, B, 0, M, N, K, ksplit=ks)
    _check(out2, A.float() @ B.float().T)
    # accumulate semantics: C holds values already
    C = base.clone()
    acc = _gemm_into(lib, A, B, C, M, N, K, ks)
    _check(acc, base + A.float() @ B.float().T)


def _gemm_into(lib, A, B, C, M, N, K, ks):
    """C[M,N] += A B^T through the split-K path (the seam zero-fills only the buffers it allocates itself)."""
    import ctypes as Ct
    L = lib.load()
    ga = (Ct.c_int64 * 5)(A.stride(0), 0, 0, A.shape[0], A.shape[1])
    gb = (Ct.c_int64 * 5)(B.stride(0), 0, 0, B.shape[0], B.shape[1])
    lib.check(L.sig_debug_gemm_bf16(A.data_ptr(), 0, ga, B.data_ptr(), 0, gb, C.data_ptr(), N, 0, None, M, N, K, 1.0, 0, ks, 128, 0, 0,
                                    None, None, 0, C.device.index, lib.stream_ptr(C.device)), "sig_debug_gemm_bf16")
    return C


@pytest.mark.parametrize("bn", [128, 256])
def test_token_view_as_kmajor_a(bn):
    """A = patch view x[:,1:] of a [B,129,d] token map (3-D tensor map, no copy)."""
    from signal_b200 import lib
    Bs, d, N = 5, 512, 512
    tok = _rand((Bs, 129, d), 5)
    W = _rand((N, d), 6)
    A = tok[:, 1:]
    out = lib.debug_gemm_bf16(A, 1, W, 0, Bs * 128, N, d, bn=bn)
    _check(out, A.float().reshape(-1, d) @ W.float().T)


@pytest.mark.parametrize("ksplit", [1, 4])
def test_mn_major_both_for_weight_grad(ksplit):
    """dW[m,n] = sum_k dH[k,m] X[k,n]: A = dH [K,M] row-major (MN-major), B = token view (MN-major)."""
    from signal_b200 import lib
    Bs, d = 6, 512
    tok = _rand((Bs, 129, d), 7)
    X = tok[:, 1:]
    dH = _rand((Bs * 128, d), 8)
    out = lib.debug_gemm_bf16(dH, 2, X, 3, d, d, Bs * 128, ksplit=ksplit)
    _check(out, dH.float().T @ X.float().reshape(-1, d))


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_mn_major_b_2d(bn, out_bf16):
    """C = A . W with W [K,N] row-major as an MN-major B operand."""
    from signal_b200 import lib
    M, N, K = 1024, 512, 768
    A, W = _rand((M, K), 9), _rand((K, N), 10)
    out = lib.debug_gemm_bf16(A, 0, W, 2, M, N, K, bn=bn, out_bf16=out_bf16)
    _check(out, A.float() @ W.float(), 6e-3 if out_bf16 else 2e-3)


@pytest.mark.parametrize("bn", [128, 256])
def test_persistent_many_tiles_per_cta(bn):
    """More work units than SMs: exercises the smem ring / TMEM double buffer across tiles."""
    from signal_b200 import lib
    Bs, d = 128, 768
    tok = _rand((Bs, 129, d), 11)
    W = _rand((d, d), 12)
    A = tok[:, 1:]
    bias = torch.randn(d, device="cuda")
    out = lib.debug_gemm_bf16(A, 1, W, 0, Bs * 128, d, d, bn=bn, out_bf16=True, bias=bias)
    ref = A.float().reshape(-1, d) @ W.float().T + bias
    _check(out, ref, 6e-3)


def test_token_row_epilogue_many_tiles():
    """dX-style GEMM: K-major A, MN-major B, bf16 written at token strides into a [B,129,d] map,
    plus the per-sample row vector (GAM broadcast) scaled by a device scalar."""
    from signal_b200 import lib
    Bs, d = 128, 768
    dH = _rand((Bs * 128, d), 13)
    W = _rand((d, d), 14)
    rv = torch.randn(Bs, d, device="cuda")
    sc = torch.tensor([0.37], device="cuda")
    dst = torch.full((Bs, 129, d), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.debug_gemm_bf16(dH, 0, W, 2, Bs * 128, d, d, bn=256, out=dst[:, 1:], rowvec=rv, rowvec_scale=sc)
    ref = (dH.float() @ W.float()).reshape(Bs, 128, d) + 0.37 * rv[:, None, :]
    _check(dst[:, 1:], ref, 6e-3)
    assert bool((dst[:, 0] == 7.0).all())


def test_weight_grad_many_units_splitk():
    from signal_b200 import lib
    Bs, d = 128, 768
    tok = _rand((Bs, 129, d), 15)
    X = tok[:, 1:]
    dH = _rand((Bs * 128, d), 16)
    out = lib.debug_gemm_bf16(dH, 2, X, 3, d, d, Bs * 128, ksplit=3)
    _check(out, dH.float().T @ X.float().reshape(-1, d))


def test_256x256_units_share_the_b_tile():
    """bn=512 selects 256 x 256 work units (two M tiles against one B tile, single TMEM buffer)."""
    from signal_b200 import lib
    Bs, d = 37, 768          # odd sample count: the second M tile of the last unit is out of bounds
    tok = _rand((Bs, 129, d), 21)
    W = _rand((d, d), 22)
    A = tok[:, 1:]
    bias = torch.randn(d, device="cuda")
    out = lib.debug_gemm_bf16(A, 1, W, 0, Bs * 128, d, d, bn=512, out_bf16=True, bias=bias)
    _check(out, A.float().reshape(-1, d) @ W.float().T + bias, 6e-3)
    dH = _rand((Bs * 128, d), 23)
    dst = torch.full((Bs, 129, d), 3.0, dtype=torch.bfloat16, device="cuda")
    lib.debug_gemm_bf16(dH, 0, W, 2, Bs * 128, d, d, bn=512, out=dst[:, 1:])
    _check(dst[:, 1:], (dH.float() @ W.float()).reshape(Bs, 128, d), 6e-3)
    assert bool((dst[:, 0] == 3.0).all())
    out = lib.debug_gemm_bf16(dH, 2, A, 3, d, d, Bs * 128, ksplit=5, bn=512)
    _check(out, dH.float().T @ A.float().reshape(-1, d))


def test_cta_pair_units():
    """bn=1024 selects 256 x 256 units on CTA pairs (cta_group::2): each CTA stages its 128 rows of A and half of
    the B tile.  Same three GEMM shapes as the LAM offset net; an odd sample count leaves the second CTA's rows
    of the last unit out of bounds."""
    from signal_b200 import lib
    Bs, d = 37, 768
    tok = _rand((Bs, 129, d), 31)
    W = _rand((d, d), 32)
    A = tok[:, 1:]
    bias = torch.randn(d, device="cuda")
    out = lib.debug_gemm_bf16(A, 1, W, 0, Bs * 128, d, d, bn=1024, out_bf16=True, bias=bias)       # H = X W^T + b
    _check(out, A.float().reshape(-1, d) @ W.float().T + bias, 6e-3)
    dH = _rand((Bs * 128, d), 33)
    dst = torch.full((Bs, 129, d), 3.0, dtype=torch.bfloat16, device="cuda")
    lib.debug_gemm_bf16(dH, 0, W, 2, Bs * 128, d, d, bn=1024, out=dst[:, 1:])                        # dX = dH W
    _check(dst[:, 1:], (dH.float() @ W.float()).reshape(Bs, 128, d), 6e-3)
    assert bool((dst[:, 0] == 3.0).all())
    out = lib.debug_gemm_bf16(dH, 2, A, 3, d, d, Bs * 128, ksplit=5, bn=1024)                         # dW = dH^T X
    _check(out, dH.float().T @ A.float().reshape(-1, d))
    M, N, K = 1024, 512, 768                                                                          # plain 2-D operands
    A2, B2 = _rand((M, K), 34), _rand((N, K), 35)
    out = lib.debug_gemm_bf16(A2, 0, B2, 0, M, N, K, bn=1024)
    _check(out, A2.float() @ B2.float().T)
