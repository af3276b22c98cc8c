"""Device arithmetic primitives vs the host build of the same header (csrc/tpgx_math.h), bit for bit.

The reference's `/`, sqrt ([REF] mnist/src/main.cpp:50,63) are single correctly rounded IEEE operations; the
device has no FP64 divide / sqrt instruction and runs nvcc's fast-path sequences branch-free (div_seq,
sqrt_seq) with its own handling of special operands.  These tests are what makes that safe: exhaustive over
the Sobel operand domain, dense random + edge cases elsewhere."""
import numpy as np
import pytest

import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    return tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)


def assert_same_bits(dev, ref, what, nan_payload_free=True):
    dn, rn = np.isnan(dev), np.isnan(ref)
    assert np.array_equal(dn, rn), "%s: NaN pattern differs at %s" % (what, np.flatnonzero(dn != rn)[:5])
    bad = np.flatnonzero(dev.view(np.uint64)[~dn] != ref.view(np.uint64)[~dn])
    assert bad.size == 0, "%s: %d mismatches, first at %s" % (what, bad.size, bad[:5])


def edge_values():
    tiny = np.array([5e-324, 1e-320, 2.2250738585072014e-308, 2.225073858507201e-308, 1e-300, 2.0 ** -967, 2.0 ** -968, 2.0 ** -960])
    big = np.array([1.7976931348623157e308, 8.98846567431158e307, 2.0 ** 1000, 2.0 ** 960, 1e300])
    mid = np.array([0.0, 1.0, 2.0, 3.0, 10.0, 0.1, 1.0 / 3.0, 255.0, 0.4375, 2.4375, 700.0, 709.78, 745.2, 1.0 + 2.0 ** -52, 1.0 - 2.0 ** -53])
    v = np.concatenate([tiny, big, mid, [np.inf, np.nan]])
    return np.concatenate([v, -v])


def test_division_all_operand_classes(ev, oracle_mod):
    """a / b correctly rounded for zeros, infinities, NaN, subnormals, extreme exponents and ordinary values —
    scalar and 4-wide sequences."""
    e = edge_values()
    a, b = [m.ravel() for m in np.meshgrid(e, e, indexing="ij")]
    rng = np.random.default_rng(11)
    n = 1 << 20
    ra = rng.standard_normal(n) * np.exp(rng.uniform(-700, 700, n))
    rb = rng.standard_normal(n) * np.exp(rng.uniform(-700, 700, n))
    px = rng.integers(0, 256, n).astype(float) * (rng.random(n) < 0.5)   # MNIST-like: half zeros
    a = np.concatenate([a, ra, px, ra, rng.integers(-1000, 1000, n).astype(float)])
    b = np.concatenate([b, rb, rng.permutation(px), px, rng.integers(-1000, 1000, n).astype(float)])
    ref = oracle_mod.math_probe(A.PROBE_DIV, a, b)
    with np.errstate(all="ignore"):
        assert_same_bits(ref, a / b, "oracle division is the hardware division")
    for fn in (A.PROBE_DIV, A.PROBE_DIV_K):
        assert_same_bits(ev.math_probe(fn, a, b), ref, "division fn %d" % fn)


def test_sobel_division_exhaustive(ev, oracle_mod):
    """(double)gy / (double)gx for EVERY Sobel gradient pair in [-1020, 1020]^2 (4.2 M quotients): proves the
    unchecked fast-path sequence on that domain."""
    g = np.arange(-1020, 1021, dtype=np.float64)
    gy, gx = [m.ravel() for m in np.meshgrid(g, g, indexing="ij")]
    assert_same_bits(ev.math_probe(A.PROBE_DIV_SMALL_INT, gy, gx), oracle_mod.math_probe(A.PROBE_DIV_SMALL_INT, gy, gx), "sobel division")


def test_sobel_sqrt_exhaustive(ev):
    """sqrt(gx*gx + gy*gy) for EVERY reachable radicand 0 .. 2*1020^2."""
    v = np.arange(0, 2 * 1020 * 1020 + 1, dtype=np.float64)
    assert_same_bits(ev.math_probe(A.PROBE_SQRT_NONNEG, v), np.sqrt(v), "sobel sqrt")


@pytest.mark.parametrize("fn,lo,hi", [
    (A.PROBE_EXP, -760.0, 720.0), (A.PROBE_EXP_K, -760.0, 720.0), (A.PROBE_EXP, -2.0, 2.0), (A.PROBE_LOG, 0.0, 300.0),
    (A.PROBE_LOG_K, 0.0, 300.0), (A.PROBE_LOG, 0.5, 2.0), (A.PROBE_ATAN, -5.0, 5.0), (A.PROBE_SIN, -50.0, 50.0),
    (A.PROBE_COS, -50.0, 50.0), (A.PROBE_TAN, -50.0, 50.0),
])
def test_shared_libm_device_equals_host(ev, oracle_mod, fn, lo, hi):
    rng = np.random.default_rng(fn * 7 + 1)
    n = 1 << 19
    x = np.concatenate([rng.uniform(lo, hi, n), edge_values(), rng.standard_normal(n) * np.exp(rng.uniform(-740, 709, n)),
                        rng.integers(0, 256, n).astype(float) * (rng.random(n) < 0.5)])
    assert_same_bits(ev.math_probe(fn, x), oracle_mod.math_probe(fn, x), "libm fn %d" % fn)


def test_division_by_ten_all_operand_classes(ev, oracle_mod):
    """a / 10.0 through the 3-operation constant-divisor sequence: dense random mantissas over the whole exponent
    range, quotients straddling every power of two (the one place a rounding tie can occur), multiples of 5 and 10,
    subnormal quotients, and the specials."""
    rng = np.random.default_rng(5)
    n = 1 << 21
    mant = rng.integers(0, 1 << 52, n, dtype=np.uint64)
    expo = rng.integers(1, 2047, n, dtype=np.uint64)
    sign = rng.integers(0, 2, n, dtype=np.uint64)
    dense = ((sign << np.uint64(63)) | (expo << np.uint64(52)) | mant).view(np.float64)
    k = np.arange(-1070, 1020)
    near = []
    for d in range(-40, 41):     # x = 10 * (2^k + d ulp): quotient next to a binade boundary, incl. 54-bit midpoints
        q = np.ldexp(1.0, k) * (1.0 + d * 2.0 ** -53)
        near.append(q * 10.0); near.append(np.nextafter(q * 10.0, np.inf)); near.append(np.nextafter(q * 10.0, -np.inf))
    ints = np.arange(-200000, 200000, dtype=np.float64)
    x = np.concatenate([dense, np.concatenate(near), ints, ints * 5.0, ints * 0.1, edge_values(), rng.integers(0, 256, n).astype(float) * rng.integers(-10, 11, n)])
    with np.errstate(all="ignore"):
        ref = x / 10.0
    assert_same_bits(oracle_mod.math_probe(A.PROBE_DIV10, x), ref, "oracle / 10")
    for fn in (A.PROBE_DIV10, A.PROBE_DIV10_K):
        assert_same_bits(ev.math_probe(fn, x), ref, "division by ten fn %d" % fn)
