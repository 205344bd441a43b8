"""CPU check of the error model behind the rounding-error certificate (csrc/ivf.cu rerank_certify_kernel,
csrc/kmeans_tc.cuh): the candidate keys come from split-precision TF32 tensor-core products,
    x.q ~ sum_i  x_hi q_hi + x_lo q_hi + x_hi q_lo        (hi = fp32 with the low 13 mantissa bits cleared, lo = x - hi,
                                                           every operand truncated to tf32 again when it is consumed)
and the certificate allows  3.003 * 2^-20  per unit of sum |x_i||q_i|  for the split (dropped lo.lo, truncation of the
lo parts) plus  (n + 8) * 2^-22  for the accumulation.  This file emulates the operand handling bit by bit in numpy
(truncation = clearing mantissa bits, the pessimistic round-toward-zero fp32 accumulator) and checks both allowances on
random, same-sign (worst case for cancellation-free growth) and badly scaled inputs.  The device-side counterpart is
`max_candidate_error` in the search stats, asserted in the GPU parity tests."""
import numpy as np
import pytest


def tf32(x):
    """what the tensor core sees of an fp32 operand: the low 13 mantissa bits dropped"""
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32(x)
    lo = (np.asarray(x, np.float32) - hi).astype(np.float32)  # exact in fp32
    return hi, lo


def add_rz(acc, term):
    """fp32 add rounded toward zero (a pessimistic model of the tensor core's accumulator)"""
    exact = np.float64(acc) + np.float64(term)
    y = np.float32(exact)
    if abs(np.float64(y)) > abs(exact):
        y = np.nextafter(y, np.float32(0))
    return y


def split_products(x, q):
    """the three product streams of the kernels (hi.hi, lo.hi, hi.lo); a product of two tf32 operands is exact in fp32
    (11 x 11 significant bits)"""
    xh, xl = split(x)
    qh, ql = split(q)
    f = np.float64
    return [xh.astype(f) * qh.astype(f), tf32(xl).astype(f) * qh.astype(f), xh.astype(f) * tf32(ql).astype(f)]


def split_dot_terms(x, q):
    return sum(split_products(x, q))


CASES = []
for seed in range(6):
    CASES.append(("normal", seed))
CASES += [("same_sign", 0), ("same_sign", 1), ("scaled", 0), ("scaled", 1), ("tiny_lo", 0)]


def make(kind, seed, n):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        x, q = rng.standard_normal(n), rng.standard_normal(n)
    elif kind == "same_sign":
        x, q = np.abs(rng.standard_normal(n)), np.abs(rng.standard_normal(n))
    elif kind == "scaled":
        x = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
        q = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
    else:  # mantissas whose low 13 bits are all ones: the largest lo parts
        x = (np.float32(1.0) + rng.integers(0, 1 << 10, n).astype(np.float32) * np.float32(2.0 ** -10)
             + np.float32((2 ** 13 - 1) * 2.0 ** -23)).astype(np.float32)
        q = x[::-1].copy()
    x, q = x.astype(np.float32), q.astype(np.float32)
    if kind in ("normal", "same_sign"):
        x /= np.float32(np.linalg.norm(x.astype(np.float64)))
        q /= np.float32(np.linalg.norm(q.astype(np.float64)))
    return x, q


@pytest.mark.parametrize("kind,seed", CASES)
@pytest.mark.parametrize("n", [128, 300, 768])
def test_split_precision_allowance(kind, seed, n):
    x, q = make(kind, seed, n)
    exact = float(np.dot(x.astype(np.float64), q.astype(np.float64)))
    absdot = float(np.dot(np.abs(x).astype(np.float64), np.abs(q).astype(np.float64)))
    terms = split_dot_terms(x, q)
    # split error alone (exact accumulation of the three products)
    assert abs(float(terms.sum()) - exact) <= 3.0 * 2.0 ** -20 * absdot
    # plus the accumulator: one tcgen05.mma folds K = 8 products (summed here without error, the hardware's internal
    # sum is wider than fp32) into the fp32 accumulator; modelled with the pessimistic round-toward-zero add, in the
    # kernels' issue order (per K step: hi.hi, hi.lo, lo.hi)
    prods = split_products(x, q)
    acc = np.float32(0)
    for k0 in range(0, n, 8):
        for stream in (prods[0], prods[2], prods[1]):
            acc = add_rz(acc, stream[k0:k0 + 8].sum())
    allowance = (3.003 * 2.0 ** -20 + (n + 8) * 2.0 ** -22) * absdot
    assert abs(float(acc) - exact) <= allowance
    # and the certificate's form of it: per unit of (||x||^2 + ||q||^2), with the factor 2 of the key -2 x.q folded in
    nx, nq = float(np.dot(x.astype(np.float64), x.astype(np.float64))), float(np.dot(q.astype(np.float64), q.astype(np.float64)))
    assert 2 * abs(float(acc) - exact) <= (3.003 / 1048576.0 + (n + 8) * 2.384185791015625e-07) * (nx + nq)


@pytest.mark.parametrize("n", [128, 300, 768])
def test_reference_order_distance_is_within_its_own_bound(n, vo):
    """d_ref (left-to-right fp32, no FMA: the oracle) vs the exact value: |d_ref - d_true| <= 1.01 (2n + 8) u (||x||^2 +
    ||q||^2) and d_ref >= d_true (1 - (n + 3) u), the two facts the certificate uses about the reference itself"""
    u = 2.0 ** -24
    rng = np.random.default_rng(n)
    for _ in range(50):
        x = rng.standard_normal(n).astype(np.float32)
        q = rng.standard_normal(n).astype(np.float32)
        x /= np.float32(np.linalg.norm(x))
        q /= np.float32(np.linalg.norm(q))
        d_ref = float(vo.l2sq(x, q))
        diff = x.astype(np.float64) - q.astype(np.float64)
        d_true = float(np.dot(diff, diff))
        s = float(np.dot(x.astype(np.float64), x.astype(np.float64)) + np.dot(q.astype(np.float64), q.astype(np.float64)))
        assert abs(d_ref - d_true) <= 1.01 * (2 * n + 8) * u * s
        assert d_ref >= d_true * (1 - (n + 3) * u)
