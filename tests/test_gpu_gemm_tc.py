"""tcgen05 GEMM (rg_op_linear_tc) against float64 references: bf16 tier and bf16x3 (fp32-class) tier,
all epilogues, ragged M (TMA zero-fill + row guards), every (N, K) shape of the denoiser."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
SHAPES = [(128, 128, 64), (1, 512, 512), (43, 1536, 512), (300, 512, 1536), (2752, 1024, 512), (6880, 512, 1024),
          (130, 256, 2048)]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _bf(t):
    return t.to(torch.bfloat16).double()


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_bf16_gemm_exact_on_bf16_inputs(dev, M, N, K):
    """With bf16-representable operands the tensor-core result must match fp64 to fp32 rounding."""
    from rag_gesture_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(N, generator=g)
    ref = F.linear(x.double(), w.double(), b.double())
    out = ops.linear_tc(x.to(dev), w.to(dev), b.to(dev)).cpu()
    assert rel_l2(out, ref) < 2e-6


@pytest.mark.parametrize("M,N,K", SHAPES[2:6])
def test_epilogues_and_outputs(dev, M, N, K):
    from rag_gesture_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N * 3 + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b, r = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    xd, wd, bd, rd = x.to(dev), w.to(dev), b.to(dev), r.to(dev)
    ref_bf = F.linear(_bf(x), _bf(w), b.double())            # what a bf16-operand GEMM computes
    ref = F.linear(x.double(), w.double(), b.double())
    assert rel_l2(ops.linear_tc(xd, wd, bd).cpu(), ref_bf) < 2e-6
    assert rel_l2(ops.linear_tc(xd, wd, bd).cpu(), ref) < 1e-2
    assert rel_l2(ops.linear_tc(xd, wd, bd, residual=rd).cpu(), ref_bf + r.double()) < 2e-6
    assert rel_l2(ops.linear_tc(xd, wd, bd, epilogue=_lib.OP_GELU).cpu(), F.gelu(ref_bf)) < 2e-6
    assert rel_l2(ops.linear_tc(xd, wd, bd, epilogue=_lib.OP_SILU).cpu(), F.silu(ref_bf)) < 2e-6
    assert rel_l2(ops.linear_tc(xd, wd, None).cpu(), F.linear(_bf(x), _bf(w))) < 2e-6
    # bf16x3: fp32-class accuracy on the tensor pipe
    out3, o16 = ops.linear_tc(xd, wd, bd, split=True, want_bf16=True)
    assert rel_l2(out3.cpu(), ref) < 3e-5
    hi, lo = o16[:, :N].float().cpu(), o16[:, N:].float().cpu()
    assert rel_l2(hi + lo, out3.cpu()) < 1e-5 and rel_l2(hi, out3.cpu()) < 4e-3
    _, o16b = ops.linear_tc(xd, wd, bd, want_bf16=True)
    assert torch.equal(o16b.cpu(), ops.linear_tc(xd, wd, bd).cpu().to(torch.bfloat16))
