"""Known-answer tests that pin the oracle against the reference's three golden graphs
(files copied verbatim from /root/reference/*/src/CodeGen/*_out_best.dot by
tests/golden/make_fixtures.py) under the in-tree environment semantics."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
import tpgx
from tpgx import abi as A


def load(app, dot, oracle_mod, tie=A.TIE_LAST_MAX, math_flavour=1):
    g = tpgx.import_dot(os.path.join(GOLDEN, dot), tpgx.APP_CONSTANTS[app])
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app],
                          tie_break=tie, math=math_flavour)
    return g, o


def test_dot_import_shapes():
    g = tpgx.import_dot(os.path.join(GOLDEN, "TicTacToe_out_best.dot"), 0)
    assert (g.nb_teams, g.nb_programs, g.nb_edges, g.nb_roots) == (4, 17, 19, 1)
    g = tpgx.import_dot(os.path.join(GOLDEN, "StickGame_out_best.dot"), 0)
    assert (g.nb_teams, g.nb_programs, g.nb_edges, g.nb_roots) == (2, 9, 9, 1)
    g = tpgx.import_dot(os.path.join(GOLDEN, "Pendulum_out_best.dot"), 5)
    assert (g.nb_teams, g.nb_programs, g.nb_edges, g.nb_roots) == (1, 4, 5, 1)  # P2514 used by two edges
    assert g.constants.shape == (4, 5) and g.constants[2, 2] == -3.105208


def test_stickgame_bids_as_functions_of_sticks(oracle_mod):
    """Appendix C.2: decoded bids of the 9 programs as functions of s = remaining sticks."""
    g, o = load("stickgame", "StickGame_out_best.dot", oracle_mod)
    name = {p: i for i, p in enumerate(g.names["programs"])}
    s = np.arange(1, 22)
    hints = np.tile(np.array([1, 2, 3, 4], dtype=np.int32), (len(s), 1))
    bid = o.execute_programs(len(s), [hints, s.astype(np.int32).reshape(-1, 1)])
    expect = {1448: 10 + 0 * s, 1449: 3 + 0 * s, 1450: 6 + 0 * s, 1451: 10 + 0 * s, 1452: s - 1,
              1444: 4 - s, 1445: s - 2, 1446: 0 * s, 1447: 0 * s}
    for p, e in expect.items():
        assert np.array_equal(bid[name[p]], e.astype(float)), "P%d" % p


@pytest.mark.parametrize("tie,moves", [
    (A.TIE_LAST_MAX, [(21, 2), (18, 2), (15, 2), (12, 2), (9, 1), (7, 1), (5, 1), (3, 1), (1, 0)]),
    (A.TIE_FIRST_MAX, [(21, 2), (18, 2), (15, 2), (12, 2), (9, 0), (8, 0), (7, 0), (6, 0), (5, 0), (4, 0), (3, 0), (2, 0), (1, 0)]),
])
def test_stickgame_selfplay_kat(oracle_mod, tie, moves):
    """Self-play as [REF] stickgame/src/CodeGen/mainTPGInference.cpp:57-75 (second player not random):
    both tie-break rules, which first differ at 9 sticks."""
    g, o = load("stickgame", "StickGame_out_best.dot", oracle_mod, tie=tie)
    out = o.evaluate_episodes(A.ENV_STICKGAME, seeds=[0], job_roots=[[0, 0]], roots_per_job=2, max_actions=64,
                              second_player_random=0, trace_steps=16)
    trace = [int(a) for a in out["trace"][0, 0] if a >= 0]
    assert trace == [a for _, a in moves]
    sticks = 21
    for s, a in moves:
        assert sticks == s
        sticks -= a + 1
    assert sticks == 0


def test_tictactoe_empty_board_kat(oracle_mod):
    """Appendix C.4: on the empty board the path is T130 -> T120 -> T113 -> action 8 with the listed bids."""
    for tie in (A.TIE_LAST_MAX, A.TIE_FIRST_MAX):
        g, o = load("tictactoe", "TicTacToe_out_best.dot", oracle_mod, tie=tie)
        name = {p: i for i, p in enumerate(g.names["programs"])}
        board = -np.ones((1, 9))
        bid = o.execute_programs(1, [board])[:, 0]
        dblmin = 2.2250738585072014e-308
        t130 = [bid[name[p]] for p in (960, 961, 962, 963, 964)]
        t120 = [bid[name[p]] for p in (955, 956, 957, 948, 958, 959)]
        t113 = [bid[name[p]] for p in (951, 950, 952, 953, 954)]
        assert t130 == [1, -2, -1, -2, -1]
        assert t120 == [-2, 10, -11, dblmin, -10, -2]
        assert t113[:4] == [-1, 0, -2, 10] and t113[4] == 0.0 and math.copysign(1.0, t113[4]) == -1.0
        out = o.evaluate_episodes(A.ENV_TICTACTOE, seeds=[0], job_roots=[[0, 0]], roots_per_job=2, max_actions=1, trace_steps=1)
        assert out["trace"][0, 0, 0] == 8
        assert out["counters"]["team_visits"] == 3


def test_pendulum_trace_kat(oracle_mod):
    """Appendix C.3: 8 steps from (angle 1.0, velocity 0.0) with glibc libm: actions 14,13x7; step-1 bids;
    state after step 1 and step 8 as hex doubles."""
    g, o = load("pendulum", "Pendulum_out_best.dot", oracle_mod, math_flavour=0)
    bid = o.execute_programs(1, [np.array([[1.0, 0.0]])])[:, 0]
    # program order in file: P2512, P2513, P2514, P2515
    assert bid[0] == -math.inf and bid[1] == 1.0
    assert bid[2] == 0.0 and math.copysign(1, bid[2]) == -1.0
    assert bid[3] == 0.9304874999999999
    acts = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
    env = oracle_mod.Env(A.ENV_PENDULUM, pendulum_actions=acts, math=0)
    env.reset(2, 0)
    env.set_state([1.0, 0.0])
    actions = []
    for step in range(8):
        st = env.state(4)
        b = o.execute_programs(1, [st[:2].reshape(1, 2)])[:, 0]
        edges = [(b[0], 0), (b[1], 14), (b[2], 6), (b[2], 6), (b[3], 13)]
        best = edges[0]
        for e in edges[1:]:
            if e[0] >= best[0]:
                best = e
        actions.append(best[1])
        prev_total = env.state(4)[2]
        env.do_action(best[1])
        st = env.state(4)
        if step == 0:
            assert (st[2] - prev_total).hex() == "-0x1.0a3d70a000000p+0"
            assert st[0].hex() == "0x1.0415aac7b05c1p+0" and st[1].hex() == "0x1.46c55e671cc3dp-2"
    assert actions == [14] + [13] * 7
    st = env.state(4)
    assert st[0].hex() == "0x1.b7c239d529b05p+0" and st[1].hex() == "0x1.b42e2d953e981p+1"


def test_pendulum_shared_libm_within_tolerance(oracle_mod):
    """north_star: pendulum rewards within 1e-9 relative between libm flavours."""
    acts = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
    g, o_std = load("pendulum", "Pendulum_out_best.dot", oracle_mod, math_flavour=0)
    _, o_sh = load("pendulum", "Pendulum_out_best.dot", oracle_mod, math_flavour=1)
    kw = dict(seeds=np.arange(5, dtype=np.uint64), job_roots=[[0]], max_actions=1000, pendulum_actions=acts)
    a = o_std.evaluate_episodes(A.ENV_PENDULUM, **kw)
    b = o_sh.evaluate_episodes(A.ENV_PENDULUM, **kw)
    assert np.array_equal(a["episode_actions"], b["episode_actions"])
    assert np.allclose(a["episode_score"], b["episode_score"], rtol=1e-9, atol=0)
