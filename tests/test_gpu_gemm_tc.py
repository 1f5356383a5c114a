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


# ---- the persistent 2-CTA kernel (gemm2_tc.cu: cta_group::2, 256x256 pair tiles, TMEM double buffering, TMA store) ----
SHAPES2 = [(256, 256, 64),        # one tile, one K block
           (1, 512, 512),         # one row: the peer CTA's rows are entirely out of bounds
           (300, 512, 512),       # second M tile: 44 valid rows in the leader, none in the peer
           (6880, 512, 512),      # the fused level (160 clips): 54 tiles, ragged last tile
           (6880, 1536, 512),     # 162 tiles over <= 74 pairs: 2-3 tiles per pair, both accumulator stages
           (6880, 512, 2048),     # the folded cross-attention GEMM: 32 K blocks, ring wraps 8 times
           (6880, 1024, 512), (6880, 512, 1024),
           (20000, 1024, 512)]    # 316 tiles: 4-5 tiles per pair, accumulator phases flip twice


class _Forced:
    """ops with set_gemm_kernel(<any>) re-forcing the kernel under test (the tests switch between it and mode 1)."""

    def __init__(self, ops, mode, persist_tiles):
        self._ops, self._mode, self._pt = ops, mode, persist_tiles

    def set_gemm_kernel(self, mode, *a):
        if mode == 1:
            self._ops.set_gemm_kernel(1)
        else:
            self._ops.set_gemm_kernel(self._mode, 0, self._pt)

    def __getattr__(self, name):
        return getattr(self._ops, name)


@pytest.fixture(params=[(2, 1), (2, 10 ** 6), (3, 0)], ids=["persistent", "one_tile_per_pair", "pair128"])
def gemm2(request, dev):
    """The cta_group::2 kernels: the 256x256 kernel persistent with two TMEM accumulator stages (persist_tiles = 1:
    always) and one tile per pair with two CTAs per SM (persist_tiles = 1e6: never persistent), and the pair128 kernel
    (a CTA pair shares each 128-row weight tile, 128x128 epilogue)."""
    from rag_gesture_b200 import ops
    mode, pt = request.param
    f = _Forced(ops, mode, pt)
    f.set_gemm_kernel(mode)
    yield f
    ops.set_gemm_kernel(0, 0, 296)


@pytest.mark.parametrize("M,N,K", SHAPES2)
def test_gemm2_bit_identical_to_gemm1(dev, gemm2, M, N, K):
    """Both kernels accumulate the K blocks in the same order into an fp32 TMEM accumulator and run the same
    epilogue arithmetic: every output (fp32, bf16 hi, bf16 lo) must agree bit for bit, for both tiers."""
    from rag_gesture_b200 import _lib
    ops = gemm2
    g = torch.Generator().manual_seed(M + 5 * N + K)
    x, w = torch.randn(M, K, generator=g).to(dev), (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b, r = torch.randn(N, generator=g).to(dev), torch.randn(M, N, generator=g).to(dev)
    for split in (False, True):
        for kw in (dict(), dict(residual=r), dict(epilogue=_lib.OP_GELU), dict(bias_none=True)):
            bias = None if kw.pop("bias_none", False) else b
            ops.set_gemm_kernel(2)
            o2, h2 = ops.linear_tc(x, w, bias, split=split, want_bf16=True, **kw)
            ops.set_gemm_kernel(1)
            o1, h1 = ops.linear_tc(x, w, bias, split=split, want_bf16=True, **kw)
            assert torch.equal(o2, o1), (split, kw)
            assert torch.equal(h2, h1), (split, kw)
    ops.set_gemm_kernel(2)
    ref = F.linear(_bf(x.cpu()), _bf(w.cpu()), b.cpu().double())
    assert rel_l2(ops.linear_tc(x, w, b).cpu(), ref) < 2e-6


def test_gemm2_outputs_one_at_a_time(dev, gemm2):
    """fp32-only and bf16-only launches (the denoiser's qkv / ffn1 shapes) and in-place residual (C32 == R)."""
    from rag_gesture_b200 import _lib
    ops = gemm2
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    M, N, K = 700, 512, 512
    x, w = torch.randn(M, K, generator=g).to(dev), (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b, r = torch.randn(N, generator=g).to(dev), torch.randn(M, N, generator=g).to(dev)
    ref = ops.linear_tc(x, w, b, residual=r)
    h = r.clone()
    with torch.cuda.device(dev):
        _lib.check(lib.rg_op_linear_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(h), _lib.ptr(h), None, M, N, K,
                                       _lib.OP_RESIDUAL, 0, _lib.stream_ptr()))
        o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        _lib.check(lib.rg_op_linear_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), None, None, _lib.ptr(o16), M, N, K,
                                       _lib.OP_GELU, 0, _lib.stream_ptr()))
    assert torch.equal(h, ref)
    assert torch.equal(o16, ops.linear_tc(x, w, b, epilogue=_lib.OP_GELU).to(torch.bfloat16))
