"""The learner's tcgen05 GEMM (csrc/gemm_tc.cu, bl_gemm_f32) against float64 matrix products: every operand orientation the
learner uses (forward: both K-major; dgrad: B transposed; wgrad: both transposed, split over K), relu on the way in, ragged sizes,
and operands far outside fp16's range (the kernel scales by powers of two taken from max|x|)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(a, b, bias=None, a_relu=False, b_relu=False, tol=2e-6):
    from boardlaw_b200.learner import gemm
    c = gemm(a, b, bias=bias, a_relu=a_relu, b_relu=b_relu)
    ad, bd = a.double(), b.double()
    if a_relu:
        ad = ad.clamp(min=0)
    if b_relu:
        bd = bd.clamp(min=0)
    ref = ad @ bd.t()
    scale = ad.abs() @ bd.abs().t()
    if bias is not None:
        ref = ref + bias.double()
        scale = scale + bias.double().abs()
    err = (c.double() - ref).abs()
    bound = tol * scale + 1e-30
    assert c.shape == ref.shape
    assert torch.isfinite(c).all()
    assert (err <= bound).all(), f'max err/bound {(err / bound).max().item():.3g}'
    return c


@pytest.mark.parametrize('M,N,K', [(1000, 256, 256), (517, 81, 256), (300, 256, 162), (128, 32, 16), (1, 1, 1), (4099, 257, 70), (260, 512, 96)])
def test_forward_orientation(M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    a = torch.randn((M, K), device='cuda', generator=g)
    b = torch.randn((N, K), device='cuda', generator=g) / K ** .5
    bias = torch.randn((N,), device='cuda', generator=g)
    _check(a, b, bias=bias)
    _check(a, b, a_relu=True)


def test_dgrad_orientation():
    g = torch.Generator(device='cuda').manual_seed(1)
    du = torch.randn((3000, 256), device='cuda', generator=g) * 1e-7          # gradients: far below fp16's smallest normal
    w = torch.randn((256, 162), device='cuda', generator=g)
    _check(du, w.t())                                                          # du @ w: B(n, k) = w[k, n]


@pytest.mark.parametrize('Nb,out,inp', [(5000, 256, 162), (32768, 256, 256), (20000, 81, 256), (777, 256, 256)])
def test_wgrad_orientation_split_k(Nb, out, inp):
    g = torch.Generator(device='cuda').manual_seed(Nb)
    du = torch.randn((Nb, out), device='cuda', generator=g) * 3e-9
    x = torch.randn((Nb, inp), device='cuda', generator=g) * 50
    _check(du.t(), x.t(), b_relu=True)                                         # du^T @ relu(x)


def test_large_magnitudes_and_zero_operand():
    g = torch.Generator(device='cuda').manual_seed(3)
    a = torch.randn((200, 64), device='cuda', generator=g) * 1e20
    b = torch.randn((96, 64), device='cuda', generator=g) * 1e-25
    _check(a, b)
    z = torch.zeros((200, 64), device='cuda')
    c = _check(z, b)
    assert (c == 0).all()


def test_into_preallocated_strided_output():
    from boardlaw_b200.learner import gemm
    g = torch.Generator(device='cuda').manual_seed(4)
    a = torch.randn((333, 100), device='cuda', generator=g)
    b = torch.randn((50, 100), device='cuda', generator=g)
    big = torch.full((333, 80), 7., device='cuda')
    gemm(a, b, out=big[:, 10:60])
    assert (big[:, :10] == 7).all() and (big[:, 60:] == 7).all()
    assert torch.allclose(big[:, 10:60].double(), a.double() @ b.double().t(), rtol=0, atol=1e-4)
