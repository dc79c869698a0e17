"""The two GEMM families (CUDA-core fp32 and tcgen05 kind::tf32) against torch fp64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 128, 4096), (32, 32, 32), (128, 64, 32), (256, 256, 64), (1000, 320, 200), (4096, 2048, 512), (300, 100, 1000), (2048, 4096, 256),
          (136, 36, 40), (640, 200, 100)]


def _mk(M, N, K, ta, tb, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
    B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    return A, B, ref


@pytest.mark.parametrize('ta', [False, True])
@pytest.mark.parametrize('tb', [False, True])
@pytest.mark.parametrize('shape', SHAPES[:5] + [(70, 3, 50), (33, 50, 20), (1, 1, 1)])
def test_simt_gemm(shape, ta, tb):
    from multimodalautoencoder_b200 import debug_gemm
    M, N, K = shape
    A, B, ref = _mk(M, N, K, ta, tb, 1)
    C = debug_gemm(A, B, ta, tb, precision='fp32')
    torch.cuda.synchronize()
    err = (C.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    assert err < 5e-6, err      # fp32 FMA accumulation over K up to 4096


@pytest.mark.parametrize('ta', [False, True])
@pytest.mark.parametrize('tb', [False, True])
@pytest.mark.parametrize('shape', SHAPES)
def test_tc_gemm(shape, ta, tb):
    """tf32 tolerance: ||C - ref||_F / ||ref||_F <= 1e-3 (10-bit mantissa operands, fp32 accumulate)."""
    from multimodalautoencoder_b200 import debug_gemm
    M, N, K = shape
    A, B, ref = _mk(M, N, K, ta, tb, 2)
    C = debug_gemm(A, B, ta, tb, precision='tf32')
    torch.cuda.synchronize()
    rel = ((C.double() - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel
    # per element: every product carries two operand roundings of 2^-11 each, the K products add up like a random walk:
    # |error| <= 6 sigma with sigma = 2^-11 * sqrt(2 K) for unit-variance operands (a misplaced tile or a dropped k-block
    # shows up as an O(sqrt(K)) error in single elements long before it moves the Frobenius norm)
    bound = 6.0 * 2.0 ** -11 * np.sqrt(2.0 * K) + 1e-6
    assert (C.double() - ref).abs().max().item() <= bound, ((C.double() - ref).abs().max().item(), bound)


def test_tc_gemm_bias_act_beta():
    from multimodalautoencoder_b200 import debug_gemm
    A, B, ref = _mk(512, 384, 128, False, False, 3)
    bias = torch.randn(384, device='cuda')
    C = debug_gemm(A, B, bias=bias, activation='softsign', precision='tf32')
    z = ref + bias.double()
    exp = z / (1 + z.abs())
    assert ((C.double() - exp).abs().max().item()) < 3e-2      # |dz| ~ 1e-3 * sqrt(K) * |a||b| at softsign slope 1
    C0 = torch.randn(512, 384, device='cuda')
    C2 = debug_gemm(A, B, precision='tf32', C_init=C0, beta=1.0)
    assert ((C2.double() - (ref + C0.double())).norm() / ref.norm()).item() < 1e-3


def test_tc_rounding_bias():
    """Reports the signed bias of the tf32 path (truncation would show as a systematic shrink)."""
    from multimodalautoencoder_b200 import debug_gemm
    g = torch.Generator(device='cuda').manual_seed(5)
    A = torch.rand((1024, 1024), device='cuda', generator=g) + 0.5     # all positive: bias does not average out
    B = torch.rand((1024, 256), device='cuda', generator=g) + 0.5
    ref = A.double() @ B.double()
    C = debug_gemm(A, B, precision='tf32')
    bias = ((C.double() - ref) / ref).mean().item()
    print('tf32 signed relative bias: %.3e' % bias)
    assert abs(bias) < 2e-3
