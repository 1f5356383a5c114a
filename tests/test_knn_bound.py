"""The error bound behind the tensor-core kNN certificate (rag_gesture_b200/csrc/knn_tc.cu): for bf16-rounded
operands accumulated in fp32,  |approx - exact| <= c_eps * |q| * |d|  with  c_eps = (2^-7 + 2^-16 + 2^-17 +
2.5e-7 * dim) * 1.01.  The GPU kernel relies on it to prove that no row outside the re-scored candidates can
belong to the top-k; here it is checked on the CPU against adversarial operand families."""
import numpy as np
import pytest
import torch


def c_eps(dim):
    return (0.0078125 + 0.0000153 + 0.0000077 + 2.5e-7 * dim) * 1.01      # C_EPS_BF16 + accumulation term, knn_tc.cu


def _families(dim, g):
    r = lambda *s: torch.randn(*s, generator=g)
    yield "gaussian", r(64, dim), r(256, dim)
    yield "all positive (no cancellation)", r(64, dim).abs(), r(256, dim).abs()
    yield "wide dynamic range", r(64, dim) * torch.exp(3 * r(64, dim)), r(256, dim) * torch.exp(3 * r(256, dim))
    # every element sits halfway between two bf16 values: the worst case of round-to-nearest
    base = torch.exp2(torch.randint(-4, 4, (64, dim), generator=g).float())
    yield "worst-case rounding", base * (1 + 2.0 ** -8), (base[:1] * (1 + 2.0 ** -8)).expand(256, dim).clone()
    yield "near-duplicates of the query", r(64, dim), None


@pytest.mark.parametrize("dim", [64, 768, 4096])
def test_bf16_similarity_error_bound(dim):
    g = torch.Generator().manual_seed(dim)
    for name, q, d in _families(dim, g):
        if d is None:
            d = q[:1].repeat(256, 1) + 1e-3 * torch.randn(256, dim, generator=g)
        approx = (q.bfloat16().float() @ d.bfloat16().float().T)            # fp32 accumulation of bf16 operands
        exact32 = q @ d.T                                                  # what the fp32 re-score computes
        exact64 = q.double() @ d.double().T
        bound = c_eps(dim) * q.norm(dim=1)[:, None] * d.norm(dim=1)[None, :]
        err = (approx.double() - exact32.double()).abs()
        assert bool((err <= bound.double()).all()), (name, float((err / bound.double()).max()))
        # and the slack is real: the bound also covers the distance to the infinitely precise score
        assert bool(((approx.double() - exact64).abs() <= bound.double()).all()), name


def test_slot_bits_perturbation_is_covered():
    """Candidate scores carry their list slot in the 4 low mantissa bits: a relative change below 2^-19."""
    x = torch.randn(10000, generator=torch.Generator().manual_seed(1)) * 3
    bits = x.view(torch.int32)
    for slot in (0, 7, 15):
        y = ((bits & ~15) | slot).view(torch.float32)
        assert float(((y - x).abs() / x.abs()).max()) < 2.0 ** -19
