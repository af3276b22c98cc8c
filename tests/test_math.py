"""Shared deterministic libm (gegelati-apps_b200/csrc/tpgx_math.h, host build through the oracle):
accuracy against mpmath, agreement with glibc on special values, and the reported mismatch rate
between the "shared" flavour (what the GPU runs) and glibc (what the real apps compute)."""
import math

import mpmath
import numpy as np
import pytest

from tpgx import abi as A

FN = {"exp": A.OP_EXP, "log": A.OP_LOG, "cos": A.OP_COS, "sin": A.OP_SIN, "tan": A.OP_TAN}


def atan_via_sobel_dir(oracle_mod, x, flavour):
    """atan is only reachable through sobelDir: a window with gx = 1, gy = v gives atan(v)."""
    out = np.empty(len(x))
    for i, v in enumerate(x):
        out[i] = oracle_mod.apply_op(A.OP_SOBEL_DIR, win=[0, 0, 0, 0, 0, 0.5, 0, float(v) / 2, 0], math=flavour)
    return out


def max_ulp(fname, x, y):
    mpmath.mp.prec = 200
    f = getattr(mpmath, fname)
    worst = 0.0
    for xi, yi in zip(x, y):
        if not math.isfinite(yi):
            continue
        t = f(mpmath.mpf(float(xi)))
        if t == 0:
            continue
        worst = max(worst, float(abs(mpmath.mpf(float(yi)) - t) / mpmath.mpf(math.ulp(float(t)))))
    return worst


@pytest.mark.parametrize("name,lo,hi,bound", [
    ("exp", -20, 20, 1.0), ("exp", -745, 709, 1.0), ("log", 1e-300, 10, 1.0), ("log", 0.9, 1.1, 1.0),
    ("sin", -10, 10, 1.0), ("sin", -1e6, 1e6, 1.0), ("cos", -10, 10, 1.0), ("cos", -1e6, 1e6, 1.0),
    ("tan", -10, 10, 2.0), ("tan", -1e6, 1e6, 2.0),
])
def test_accuracy_vs_mpmath(oracle_mod, name, lo, hi, bound):
    x = np.random.default_rng(1).uniform(lo, hi, 3000)
    y = oracle_mod.math_array(FN[name], x, math=oracle_mod.MATH_SHARED)
    assert max_ulp(name, x, y) < bound


def test_trig_huge_arguments(oracle_mod):
    """Payne-Hanek reduction stays accurate up to the largest doubles."""
    x = np.exp(np.random.default_rng(2).uniform(0, 709, 2000)) * np.random.default_rng(3).choice([-1, 1], 2000)
    for name in ("sin", "cos"):
        assert max_ulp(name, x, oracle_mod.math_array(FN[name], x)) < 1.0


def test_atan_accuracy(oracle_mod):
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-3, 3, 1500), np.exp(rng.uniform(-40, 60, 1500)) * rng.choice([-1, 1], 1500)])
    assert max_ulp("atan", x, atan_via_sobel_dir(oracle_mod, x, 1)) < 1.0


def test_special_values_match_glibc(oracle_mod):
    inf, nan = math.inf, math.nan
    cases = {
        "exp": [inf, -inf, nan, 0.0, -0.0, 709.782712893384, 709.79, 710.0, 1e308, -745.13, -745.2, -746.0, -1e308, 5e-324, 1.0],
        "log": [0.0, -0.0, -1.0, -inf, inf, nan, 5e-324, 2.2250738585072014e-308, 1.0, 1.7976931348623157e308],
        "sin": [0.0, -0.0, inf, -inf, nan, 5e-324], "cos": [0.0, inf, nan], "tan": [0.0, -0.0, inf, nan],
    }
    for name, xs in cases.items():
        x = np.array(xs)
        a = oracle_mod.math_array(FN[name], x, math=0)
        b = oracle_mod.math_array(FN[name], x, math=1)
        for xi, ai, bi in zip(xs, a, b):
            if math.isnan(ai):
                assert math.isnan(bi), (name, xi)
            elif math.isinf(ai) or ai == 0.0:
                assert ai == bi and math.copysign(1, ai) == math.copysign(1, bi), (name, xi, ai, bi)
            else:
                assert abs(ai - bi) <= 2 * math.ulp(ai), (name, xi, ai, bi)
    xs = [0.0, -0.0, inf, -inf, 1e-300, -1e-300, 1e300, 0.4375, 0.6875, 1.1875, 2.4375, 5e-324]
    a, b = atan_via_sobel_dir(oracle_mod, xs, 0), atan_via_sobel_dir(oracle_mod, xs, 1)
    for xi, ai, bi in zip(xs, a, b):
        assert abs(ai - bi) <= math.ulp(ai) and math.copysign(1, ai) == math.copysign(1, bi), (xi, ai, bi)
    assert math.isnan(atan_via_sobel_dir(oracle_mod, [nan], 1)[0])


def test_shared_vs_glibc_mismatch_rate_is_last_ulp_only(oracle_mod):
    """Reported, not hidden: the shared libm differs from glibc in a few percent of arguments, always by one ulp."""
    x = np.random.default_rng(5).uniform(-30, 30, 20000)
    for name in ("exp", "sin", "cos"):
        a = oracle_mod.math_array(FN[name], x, math=0)
        b = oracle_mod.math_array(FN[name], x, math=1)
        diff = a.view(np.int64) - b.view(np.int64)
        assert np.abs(diff).max() <= 1
        assert (diff != 0).mean() < 0.10


def test_rng_restatement_matches_libstdcxx(oracle_mod):
    """tpgx_rng.h (device draw transforms) == std::uniform_{int,real}_distribution over std::mt19937_64."""
    lib = oracle_mod.load()
    for seed in (0, 1, 2 ^ 1, 123456789):
        for lo, hi in ((0, 3), (1, 3), (0, 8), (1, 1), (0, 59999), (0, 2 ** 64 - 1), (5, 2 ** 63)):
            a = np.zeros(200, dtype=np.uint64); b = np.zeros(200, dtype=np.uint64)
            lib.orc_rng_u64(seed, 200, lo, hi, a.ctypes.data)
            lib.orc_rng_restated_u64(seed, 200, lo, hi, b.ctypes.data)
            assert np.array_equal(a, b), (seed, lo, hi)
        for lo, hi in ((-math.pi, math.pi), (-1.0, 1.0), (0.0, 1.0)):
            a = np.zeros(200); b = np.zeros(200)
            lib.orc_rng_f64(seed, 200, lo, hi, a.ctypes.data)
            lib.orc_rng_restated_f64(seed, 200, lo, hi, b.ctypes.data)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (seed, lo, hi)


def test_mnist_actions_and_scores_do_not_depend_on_the_libm_flavour(oracle_mod):
    """What the one-ulp differences between the shared libm (the GPU's) and glibc (the real apps') do to an MNIST-shaped
    evaluateAllRoots: a small share of the bids differs in the last bits (more where a later subtraction cancels), the
    non-finite pattern, every chosen action and every root score are the same.  (3 000 samples x 200 roots measured
    once for DESIGN.md: 1.9 % of the finite bids, 0 of 600 000 actions, 0 of 200 scores.)"""
    import tpgx
    g = tpgx.synthetic_graph("mnist", roots=60, edges=5, lines=20, depth=2, mix="uniform", share_prob=0.1, seed=42)
    img, lab = tpgx.synthetic_images(600, seed=3)
    want = ("score", "bid", "action", "confusion")
    res = [oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5, math=m)
           .evaluate_classification(img, lab, want=want, nb_threads=4) for m in (0, 1)]
    a, b = res
    assert np.array_equal(np.isnan(a["bid"]), np.isnan(b["bid"])) and np.array_equal(np.isinf(a["bid"]), np.isinf(b["bid"]))
    fin = np.isfinite(a["bid"])
    differ = (a["bid"].view(np.int64) != b["bid"].view(np.int64))[fin].mean()
    assert 0.0 < differ < 0.05, differ
    assert np.array_equal(a["action"], b["action"])
    assert np.array_equal(a["confusion"], b["confusion"])
    assert np.array_equal(a["score"].view(np.uint64), b["score"].view(np.uint64))
