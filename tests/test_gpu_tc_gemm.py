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


def test_mn_major_b_2d():
    """C = A . W with W [K,N] row-major as an MN-major B operand."""
    from signal_b200 import lib
    M, N, K = 256, 512, 768
    A, W = _rand((M, K), 9), _rand((K, N), 10)
    out = lib.debug_gemm_bf16(A, 0, W, 2, M, N, K)
    _check(out, A.float() @ W.float())
