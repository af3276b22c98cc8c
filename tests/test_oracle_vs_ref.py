"""Pins the CPU oracle against the reference's OWN code for the parts of the hot path that are in
the reference tree: the five apps' instruction lambdas and the four sequential
LearningEnvironments.

Two layers:
  * golden vectors committed under tests/golden/ (ref_instructions.npz, ref_env_traces.npz), produced
    by tests/golden/make_fixtures.py from oracle/_ref/libtpgref.so = the reference sources compiled
    unmodified (oracle/build_ref.sh) — these run everywhere, including the GPU box;
  * live comparisons against libtpgref.so on more random cases when the library is present
    (build container; it also travels to the GPU box with the snapshot).
The oracle is run in its "std" math flavour here (glibc libm = what the reference apps compute);
bit-exact equality is required (NaN == NaN)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import tpgx
from tpgx import abi as A
from oracle import refbridge as RB

KIND_OF_TYPE = {"D": RB.KIND_DOUBLE, "I": RB.KIND_INT, "C": RB.KIND_CONSTANT, "W": RB.KIND_WINDOW}
# operand signature of every device opcode (include/tpgx.h tpgx_opcode comments)
OP_SIG = {A.OP_SUB: "DD", A.OP_ADD: "DD", A.OP_MUL: "DD", A.OP_DIV: "DD", A.OP_MAX: "DD", A.OP_EXP: "D", A.OP_LOG: "D",
          A.OP_COS: "D", A.OP_SIN: "D", A.OP_TAN: "D", A.OP_PI: "D", A.OP_MULC: "DC", A.OP_FMODG: "DD", A.OP_SUB_II: "II",
          A.OP_CAST_I: "I", A.OP_EQ0: "D", A.OP_EQM1: "D", A.OP_EQ1: "D", A.OP_GE15: "D", A.OP_COND: "DD",
          A.OP_SOBEL_MAGN: "W", A.OP_SOBEL_DIR: "W"}


def same_bits(x, y):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    nan = np.isnan(x) & np.isnan(y)
    return np.array_equal(x.view(np.uint64)[~nan], y.view(np.uint64)[~nan]) and np.array_equal(np.isnan(x), np.isnan(y))


def oracle_op(oracle_mod, opcode, sig, a, b, ia, ib, win):
    if sig == "W":
        return np.array([oracle_mod.apply_op(opcode, win=w, math=oracle_mod.MATH_STD) for w in win])
    return np.array([oracle_mod.apply_op(opcode, a=x, b=y, ia=int(p), ib=int(q), c=y, math=oracle_mod.MATH_STD)
                     for x, y, p, q in zip(a, b, ia, ib)])


@pytest.fixture(scope="module")
def golden_ins():
    return np.load(os.path.join(GOLDEN, "ref_instructions.npz"))


@pytest.mark.parametrize("app", sorted(RB.APPS))
def test_instruction_lambdas_match_golden(oracle_mod, golden_ins, app):
    """Every (app, instruction index): operand signature and result on the fixed operand grid equal what the
    reference lambda ([REF] Appendix A files) returned."""
    g = golden_ins
    for idx, opcode in enumerate(tpgx.INSTRUCTION_SETS[app]):
        sig = OP_SIG[opcode]
        assert [KIND_OF_TYPE[c] for c in sig] == list(g["%s_%d_kinds" % (app, idx)]), (app, idx)
        got = oracle_op(oracle_mod, opcode, sig, g["a"], g["b"], g["ia"], g["ib"], g["win"])
        assert same_bits(got, g["%s_%d" % (app, idx)]), (app, idx)
    assert "%s_%d" % (app, len(tpgx.INSTRUCTION_SETS[app])) not in g.files  # same number of instructions


def oracle_trace(oracle_mod, kind, spr, seed, mode, actions, max_steps, n_obs=12):
    env = oracle_mod.Env(kind, spr, RB.PENDULUM_ACTIONS, math=oracle_mod.MATH_STD)
    env.reset(seed, mode)

    def row():
        s = env.state(16)
        if kind == A.ENV_STICKGAME:
            obs = [1, 2, 3, 4, s[0]]          # getDataSources(): hints first, remainingSticks second
        elif kind == A.ENV_TICTACTOE:
            obs = list(s[:9])
        else:
            obs = list(s[:2])
        o = np.zeros(n_obs); o[:len(obs)] = obs
        adv = kind in (A.ENV_STICKGAME, A.ENV_TICTACTOE)
        return np.concatenate([o, [float(env.is_terminal()), env.score(0), env.score(1) if adv else env.score(0)]])
    rows = [row()]
    for t in range(max_steps):
        if env.is_terminal():
            break
        env.do_action(actions[t])
        rows.append(row())
    return np.array(rows)


def test_environment_traces_match_golden(oracle_mod):
    """Seeded random-action episodes of gridworld / stickgame (random + self-play) / tic-tac-toe (adversarial
    + random opponent) / pendulum: observation, isTerminal, getScore(s) after every doAction are bit-identical
    to the reference classes ([REF] Appendix B files)."""
    g = np.load(os.path.join(GOLDEN, "ref_env_traces.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_meta")})
    assert len(names) >= 60
    for n in names:
        kind, spr, seed, mode = [int(v) for v in g[n + "_meta"]]
        actions = g[n + "_actions"]
        ref = g[n + "_trace"]
        got = oracle_trace(oracle_mod, kind, spr, seed, mode, actions, len(actions))
        assert got.shape == ref.shape, n
        assert same_bits(got, ref), n


needs_ref = pytest.mark.skipif(not RB.available(), reason="oracle/_ref/libtpgref.so not built (needs /root/reference)")


@needs_ref
def test_live_instruction_lambdas_random(oracle_mod):
    rng = np.random.default_rng(123)
    n = 300
    a = np.concatenate([rng.normal(0, 1e3, n), rng.integers(0, 256, n).astype(float), np.exp(rng.uniform(-700, 700, n)) * rng.choice([-1, 1], n)])
    b = rng.permutation(a)
    ia, ib = rng.integers(-50, 50, len(a)).astype(np.int32), rng.integers(-50, 50, len(a)).astype(np.int32)
    win = rng.integers(0, 256, (len(a), 9)).astype(float)
    for app in RB.APPS:
        sigs = RB.set_signature(app)
        assert len(sigs) == len(tpgx.INSTRUCTION_SETS[app])
        for idx, opcode in enumerate(tpgx.INSTRUCTION_SETS[app]):
            sig = OP_SIG[opcode]
            if sig == "W":
                ref = np.array([RB.call(app, idx, win=w) for w in win])
            else:
                ref = np.array([RB.call(app, idx, d=(x, y), i=(p, q), c=(x, y)) for x, y, p, q in zip(a, b, ia, ib)])
            assert same_bits(oracle_op(oracle_mod, opcode, sig, a, b, ia, ib, win), ref), (app, idx)


@needs_ref
def test_live_environment_traces_random(oracle_mod):
    rng = np.random.default_rng(77)
    for name, kind, spr, nact, steps in RB.ENV_CASES:
        for ep in range(25):
            seed, mode = int(rng.integers(0, 2**62)), int(rng.integers(0, 3))
            actions = rng.integers(0, nact, steps).astype(np.float64)
            ref = RB.run_trace(RB.Env(kind, spr), seed, mode, actions, steps)
            got = oracle_trace(oracle_mod, kind, spr, seed, mode, actions, steps)
            assert got.shape == ref.shape and same_bits(got, ref), (name, ep)


@needs_ref
def test_live_pendulum_kat_from_explicit_state(oracle_mod):
    """8 steps from (angle 1.0, velocity 0.0) with actions 14,13,...; hex doubles after
    step 1 are exact on any IEEE platform (only fmod/sin of glibc involved through the reference code itself)."""
    ref = RB.Env(RB.ENV_PENDULUM)
    ref.reset(0, 0)
    ref.set_pendulum(1.0, 0.0)
    orc = oracle_mod.Env(A.ENV_PENDULUM, 0, RB.PENDULUM_ACTIONS, math=oracle_mod.MATH_STD)
    orc.reset(0, 0)
    orc.set_state([1.0, 0.0])
    for a in [14, 13, 13, 13, 13, 13, 13, 13]:
        ref.do_action(a); orc.do_action(a)
        assert same_bits(ref.observe(2), orc.state(2))
    assert ref.observe(2)[0] == float.fromhex("0x1.b7c239d529b05p+0") and ref.observe(2)[1] == float.fromhex("0x1.b42e2d953e981p+1")


@needs_ref
def test_live_out_of_range_action_throws_like_reference(oracle_mod):
    """[REF] pendulum/src/Learn/pendulum.cpp:66-68 throws on an out-of-range action id."""
    ref = RB.Env(RB.ENV_PENDULUM); ref.reset(1, 0)
    orc = oracle_mod.Env(A.ENV_PENDULUM, 0, RB.PENDULUM_ACTIONS, math=oracle_mod.MATH_STD); orc.reset(1, 0)
    assert ref.do_action(15) == -1 and orc.do_action(15) == -1
    assert ref.do_action(-1) == -1 and orc.do_action(-1) == -1
