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


# ---------------------------------------------------------------------------------------------------------------------
# fp16 candidate copy of the inverted lists (csrc/ivf_tc.cuh PREC 2, rerank_certify_kernel tf32_pass == 4) and fp16
# k-means assign (csrc/kmeans_tc.cuh F16).  The copy holds x~ = fp16(x * 2^e); the query is split into
# q_hi = fp16(q), q_lo = fp16((q - q_hi) * 2^11).  The certificate allows, for the key ||x||^2 - 2 x~.q~,
#     2.002 * (sqrt(max||x - x~||^2 * ||q||^2) + (||x||_max + ||x - x~||_max) * ||q - q~||)  +  (n + 8) 2^-22 (||x||^2_max + ||q||^2)
# (Cauchy-Schwarz; fp16 x fp16 products are exact in the fp32 accumulator).  Emulated here in numpy with the pessimistic
# round-toward-zero accumulator; the scale 2^e is chosen exactly like csrc/ivf.cu ivf_ensure_h16.
def h16_scale(nxmax):
    import math

    ex = math.frexp(math.sqrt(float(nxmax)) * 1.0001)[1] if nxmax > 0 else 0
    return np.float32(2.0 ** max(-100, min(100, 14 - ex)))


def h16_copy(x, scale):
    xs = (np.asarray(x, np.float32) * scale).astype(np.float16)
    assert np.all(np.isfinite(xs)), "the scale must keep every element inside fp16's range"
    return xs, (xs.astype(np.float32) / scale).astype(np.float32)  # what the copy holds, and the value it represents


def h16_query(q):
    q = np.asarray(q, np.float32)
    hi = q.astype(np.float16)
    lo = ((q - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    rep = (hi.astype(np.float32) + lo.astype(np.float32) / np.float32(2048.0)).astype(np.float32)
    return hi, lo, rep


@pytest.mark.parametrize("kind,seed", CASES + [("unit", 0), ("unit", 1), ("huge", 0)])
@pytest.mark.parametrize("n", [96, 300, 768])
def test_fp16_candidate_copy_error_is_inside_the_certificates_allowance(kind, seed, n):
    rng = np.random.default_rng(100 + seed)
    if kind == "unit":
        rows = rng.standard_normal((40, n)).astype(np.float32)
        rows /= np.linalg.norm(rows, axis=1, keepdims=True).astype(np.float32)
        q = rows[3] + np.float32(0.05) * rng.standard_normal(n).astype(np.float32)
        q = (q / np.linalg.norm(q)).astype(np.float32)
    elif kind == "huge":
        rows = (rng.standard_normal((40, n)) * 3.0e4).astype(np.float32)
        rows[::3] *= np.float32(1e-7)
        q = (rng.standard_normal(n) * 2.0e3).astype(np.float32)
    else:
        rows = np.stack([make(kind, seed * 50 + r, n)[0] for r in range(40)]).astype(np.float32)
        q = make(kind, seed, n)[1].astype(np.float32)
    nx = (rows.astype(np.float64) ** 2).sum(1)
    scale = h16_scale(np.float32(nx.max()))
    xs16, xrep = h16_copy(rows, scale)
    qh, ql, qrep = h16_query(q)
    if not (np.all(np.isfinite(qh)) and np.all(np.isfinite(ql))):
        pytest.skip("the query does not fit fp16: the device flags such queries (qres2 = inf) and redoes them exactly")
    xlo2max = ((rows.astype(np.float64) - xrep.astype(np.float64)) ** 2).sum(1).max()
    qres2 = ((q.astype(np.float64) - qrep.astype(np.float64)) ** 2).sum()
    nq2 = (q.astype(np.float64) ** 2).sum()
    u = 2.0 ** -24
    E = (1.01 * (2 * n + 8) * u * (nx.max() + nq2)
         + 2.002 * (np.sqrt(xlo2max * nq2) + (np.sqrt(nx.max()) + np.sqrt(xlo2max)) * np.sqrt(qres2))
         + (n + 8) * 2.0 ** -22 * (nx.max() + nq2))
    worst = 0.0
    for r in range(rows.shape[0]):
        # the two accumulator blocks of the kernel: x~ . q_hi and x~ . (q_lo 2^11), fp32 accumulators rounded toward
        # zero, exact fp16 x fp16 products
        a_hh = np.float32(0)
        a_hl = np.float32(0)
        ph = xs16[r].astype(np.float64) * qh.astype(np.float64)
        pl = xs16[r].astype(np.float64) * ql.astype(np.float64)
        for i in range(n):
            a_hh = add_rz(a_hh, ph[i])
            a_hl = add_rz(a_hl, pl[i])
        dot = np.float32(np.float32(a_hl) * np.float32(1.0 / 2048.0) + np.float32(a_hh))      # epilogue: fma(hl, 2^-11, hh)
        key = np.float32(np.float32(nx[r]) + np.float32(-2.0 / float(scale)) * dot)            # fma(key_scale, dot, ||x||^2)
        d_true = ((rows[r].astype(np.float64) - q.astype(np.float64)) ** 2).sum()
        worst = max(worst, abs(float(key) + nq2 - d_true))
    assert worst <= E, (worst, E)
    # and the allowance is not vacuous: on unit rows and queries it is ~1e-3 at 768 dimensions (Cauchy-Schwarz term 5e-4,
    # accumulation allowance 4e-4, fp32 terms 2e-4), below the neighbour gaps it must split on the bench data (>= 1.7e-3)
    if kind == "unit":
        assert E < 1.5e-3


def test_fp16_scale_never_overflows_and_subnormal_elements_are_covered_by_the_assign_allowance():
    """tc_assign1_kernel<., F16>: both operands are fp16(v * 2^e) with e from max ||row||^2; the certificate's operand
    term is 2^-10 (1 + 2^-12) + 2^-30 per unit of (||x||^2 + ||c||^2_max): elements whose scaled value is subnormal in
    fp16 carry an ABSOLUTE error of 2^-25 / scale, which the 2^-30 term must cover whatever the row looks like"""
    rng = np.random.default_rng(7)
    n = 128
    for trial in range(20):
        rows = rng.standard_normal((64, n)).astype(np.float32) * np.float32(10.0 ** rng.integers(-3, 4))
        rows[:, ::5] *= np.float32(1e-9)  # far below the normal range after scaling
        cents = rows[rng.integers(0, 64, 8)] * np.float32(0.7)
        nx = (rows.astype(np.float64) ** 2).sum(1)
        nc = (cents.astype(np.float64) ** 2).sum(1)
        scale = h16_scale(np.float32(nx.max()))
        _, xr = h16_copy(rows, scale)
        _, cr = h16_copy(cents, scale)
        exact = rows.astype(np.float64) @ cents.astype(np.float64).T
        approx = xr.astype(np.float64) @ cr.astype(np.float64).T
        allow = (1.001 / 1024.0 + 2.0 ** -30) * (nx[:, None] + nc.max()) / 2.0  # per dot product; the key doubles it
        assert np.all(np.abs(exact - approx) <= allow)
