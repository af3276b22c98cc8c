"""GPU parity of the sequential environments (episode kernels) vs the oracle: per-episode scores,
action counts, action traces, final states and counters."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu
PEND_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]  # [REF] pendulum/src/Learn/main.cpp:33


def pair(oracle_mod, app, g, tie=A.TIE_LAST_MAX):
    nc = tpgx.APP_CONSTANTS[app]
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc, tie_break=tie)
    ev.set_graph(g)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc, tie_break=tie)
    return ev, o


def check(dev, ref, exact=True, rtol=0.0):
    assert np.array_equal(dev["episode_actions"], ref["episode_actions"])
    if "trace" in ref:
        assert np.array_equal(dev["trace"], ref["trace"])
    for k in ("score", "episode_score", "final_state"):
        if exact:
            assert np.array_equal(dev[k].view(np.uint64), ref[k].view(np.uint64)), k
        else:
            assert np.allclose(dev[k], ref[k], rtol=rtol, atol=0), k
    for k in ("evaluations", "team_visits", "program_executions", "lines_executed"):
        assert dev["counters"][k] == ref["counters"][k], k


@pytest.mark.parametrize("tie", [A.TIE_LAST_MAX, A.TIE_FIRST_MAX])
def test_gridworld(oracle_mod, tie):
    g = tpgx.synthetic_graph("gridworld", roots=64, edges=6, lines=10, depth=3, share_prob=0.1, seed=3)
    ev, o = pair(oracle_mod, "gridworld", g, tie)
    kw = dict(seeds=[0], job_roots=np.arange(64), max_actions=100, trace_steps=100)
    check(ev.evaluate_episodes(A.ENV_GRIDWORLD, **kw), o.evaluate_episodes(A.ENV_GRIDWORLD, **kw))


@pytest.mark.parametrize("second_random", [1, 0])
def test_stickgame(oracle_mod, second_random):
    """stickgame/params.json: 100 iterations x <= 11 actions; the random opponent consumes 1-2 RNG draws per move."""
    g = tpgx.synthetic_graph("stickgame", roots=50, edges=5, lines=8, depth=2, seed=5)
    ev, o = pair(oracle_mod, "stickgame", g)
    seeds = (np.arange(100, dtype=np.uint64) ^ np.uint64(7))
    kw = dict(seeds=seeds, job_roots=np.arange(50), max_actions=11, second_player_random=second_random, trace_steps=11)
    check(ev.evaluate_episodes(A.ENV_STICKGAME, **kw), o.evaluate_episodes(A.ENV_STICKGAME, **kw))


def test_stickgame_golden_graph_selfplay(oracle_mod):
    g = tpgx.import_dot(os.path.join(GOLDEN, "StickGame_out_best.dot"), 0)
    for tie, n in ((A.TIE_LAST_MAX, 9), (A.TIE_FIRST_MAX, 13)):
        ev, o = pair(oracle_mod, "stickgame", g, tie)
        kw = dict(seeds=[0], job_roots=[[0, 0]], roots_per_job=2, max_actions=64, second_player_random=0, trace_steps=16)
        dev = ev.evaluate_episodes(A.ENV_STICKGAME, **kw)
        check(dev, o.evaluate_episodes(A.ENV_STICKGAME, **kw))
        assert dev["episode_actions"][0, 0] == n


@pytest.mark.parametrize("second_random,k", [(0, 2), (1, 1)])
def test_tictactoe(oracle_mod, second_random, k):
    """Adversarial jobs (2 roots alternate, board inverted between moves) and the random-opponent variant."""
    g = tpgx.synthetic_graph("tictactoe", roots=40, edges=8, lines=12, depth=2, seed=9)
    ev, o = pair(oracle_mod, "tictactoe", g)
    rng = np.random.default_rng(0)
    jobs = rng.integers(0, 40, (60, k))
    kw = dict(seeds=np.arange(3, dtype=np.uint64) + 100, job_roots=jobs, roots_per_job=k, max_actions=9 if k == 2 else 5,
              second_player_random=second_random, trace_steps=9, mode=A.MODE_VALIDATION)
    check(ev.evaluate_episodes(A.ENV_TICTACTOE, **kw), o.evaluate_episodes(A.ENV_TICTACTOE, **kw))


def test_pendulum_bit_exact_vs_shared_libm_oracle(oracle_mod):
    """Same libm sequence on both sides => identical bits, 5 iterations x 1000 steps (pendulum/params.json)."""
    g = tpgx.synthetic_graph("pendulum", roots=48, edges=5, lines=10, depth=2, seed=13)
    ev, o = pair(oracle_mod, "pendulum", g)
    kw = dict(seeds=np.arange(5, dtype=np.uint64), job_roots=np.arange(48), max_actions=1000, pendulum_actions=PEND_ACTIONS, trace_steps=64)
    check(ev.evaluate_episodes(A.ENV_PENDULUM, **kw), o.evaluate_episodes(A.ENV_PENDULUM, **kw))


def test_pendulum_golden_graph_vs_glibc_oracle(oracle_mod):
    """north_star tolerance: per-root rewards within 1e-9 relative of the CPU reference's libm."""
    g = tpgx.import_dot(os.path.join(GOLDEN, "Pendulum_out_best.dot"), 5)
    nc = 5
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["pendulum"], tpgx.APP_SOURCES["pendulum"], nb_constants=nc)
    ev.set_graph(g)
    o_std = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["pendulum"], tpgx.APP_SOURCES["pendulum"], g, nb_constants=nc, math=0)
    kw = dict(seeds=np.arange(5, dtype=np.uint64), job_roots=[[0]], max_actions=1000, pendulum_actions=PEND_ACTIONS)
    dev, ref = ev.evaluate_episodes(A.ENV_PENDULUM, **kw), o_std.evaluate_episodes(A.ENV_PENDULUM, **kw)
    assert np.array_equal(dev["episode_actions"], ref["episode_actions"])
    assert np.allclose(dev["episode_score"], ref["episode_score"], rtol=1e-9, atol=0)


@pytest.mark.parametrize("env_var,value", [("TPGX_EP_CTA_WARPS", "0"), ("TPGX_EP_CTA_WARPS", "2"), ("TPGX_EP_CTA_WARPS", "16"),
                                           ("TPGX_EP_GROUPS", "4"), ("TPGX_EP_GROUPS", "2")])
@pytest.mark.parametrize("app", ["pendulum", "stickgame", "tictactoe"])
def test_episode_kernel_mappings_agree(oracle_mod, monkeypatch, app, env_var, value):
    """The plain pass runs one CTA per job (warp = edge, lane = iteration) with as many edge warps as the largest team has
    edges; fewer warps (teams evaluated in several rounds), more warps than edges, and the lane-group kernel
    (TPGX_EP_CTA_WARPS=0: 1, 2 or 4 episodes per warp) must all give the oracle's bits."""
    monkeypatch.setenv(env_var, value)
    if env_var == "TPGX_EP_GROUPS":
        monkeypatch.setenv("TPGX_EP_CTA_WARPS", "0")
    edges = {"pendulum": 5, "stickgame": 5, "tictactoe": 7}[app]
    g = tpgx.synthetic_graph(app, roots=24, edges=edges, lines=10, depth=3, share_prob=0.1, seed=21)
    ev, o = pair(oracle_mod, app, g)
    if app == "pendulum":
        kw = dict(seeds=np.arange(5, dtype=np.uint64) + 40, job_roots=np.arange(24), max_actions=400, pendulum_actions=PEND_ACTIONS, trace_steps=64)
        dev = ev.evaluate_episodes(A.ENV_PENDULUM, **kw)
        check(dev, o.evaluate_episodes(A.ENV_PENDULUM, **kw))
    elif app == "stickgame":
        kw = dict(seeds=np.arange(37, dtype=np.uint64) + 3, job_roots=np.arange(24), max_actions=11, second_player_random=1, trace_steps=11)
        check(ev.evaluate_episodes(A.ENV_STICKGAME, **kw), o.evaluate_episodes(A.ENV_STICKGAME, **kw))
    else:
        jobs = np.random.default_rng(2).integers(0, 24, (30, 2))
        kw = dict(seeds=np.arange(3, dtype=np.uint64) + 100, job_roots=jobs, roots_per_job=2, max_actions=9, second_player_random=0, trace_steps=9)
        check(ev.evaluate_episodes(A.ENV_TICTACTOE, **kw), o.evaluate_episodes(A.ENV_TICTACTOE, **kw))


def test_pendulum_more_iterations_than_lanes(oracle_mod):
    """40 iterations per job: two CTAs per job (32 + 8 lanes), 76.8 KB of reward histories in dynamic shared memory, and
    enough steps for the 300-entry history to wrap and for isTerminal's early exit to be exercised."""
    g = tpgx.synthetic_graph("pendulum", roots=6, edges=5, lines=10, depth=2, seed=17)
    ev, o = pair(oracle_mod, "pendulum", g)
    kw = dict(seeds=np.arange(40, dtype=np.uint64) + 7, job_roots=np.arange(6), max_actions=700, pendulum_actions=PEND_ACTIONS, trace_steps=32)
    check(ev.evaluate_episodes(A.ENV_PENDULUM, **kw), o.evaluate_episodes(A.ENV_PENDULUM, **kw))


def test_wide_teams_take_several_rounds(oracle_mod):
    """tic-tac-toe/params.json allows 60 outgoing edges: teams wider than the 16 edge warps of a CTA are evaluated in
    several rounds, the admissible-edge order (ties, archive ranks) must survive that."""
    g = tpgx.synthetic_graph("tictactoe", roots=12, edges=45, lines=6, depth=2, seed=23)
    ev, o = pair(oracle_mod, "tictactoe", g)
    jobs = np.random.default_rng(3).integers(0, 12, (20, 2))
    kw = dict(seeds=np.arange(4, dtype=np.uint64) + 11, job_roots=jobs, roots_per_job=2, max_actions=9, second_player_random=0, trace_steps=9)
    check(ev.evaluate_episodes(A.ENV_TICTACTOE, **kw), o.evaluate_episodes(A.ENV_TICTACTOE, **kw))


def test_caller_supplied_rng_seeds_keep_the_library_independent_of_data_hash(oracle_mod):
    """tpgx_episode_request.rng_seeds: the value the environment's reset() hands to rng.setSeed(), computed by the caller
    with the Data::Hash of the GEGELATI it links.  With the identity hash (libstdc++) it must reproduce the default path;
    with any other value the draws follow that value, on the device and in the oracle alike."""
    g = tpgx.synthetic_graph("pendulum", roots=12, edges=4, lines=10, depth=2, seed=8)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["pendulum"], tpgx.APP_SOURCES["pendulum"], nb_constants=5)
    ev.set_graph(g)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["pendulum"], tpgx.APP_SOURCES["pendulum"], g, nb_constants=5)
    pa = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
    kw = dict(seeds=[11, 12, 13], job_roots=np.arange(12), max_actions=120, pendulum_actions=pa, mode=A.MODE_VALIDATION)
    base = ev.evaluate_episodes(A.ENV_PENDULUM, **kw)
    same = ev.evaluate_episodes(A.ENV_PENDULUM, rng_seeds=[11 ^ 1, 12 ^ 1, 13 ^ 1], **kw)
    assert np.array_equal(base["score"].view(np.uint64), same["score"].view(np.uint64))
    other = ev.evaluate_episodes(A.ENV_PENDULUM, rng_seeds=[0x9E3779B97F4A7C15, 5, 77], **kw)
    ref = o.evaluate_episodes(A.ENV_PENDULUM, rng_seeds=[0x9E3779B97F4A7C15, 5, 77], **kw)
    assert not np.array_equal(other["score"], base["score"])
    assert np.array_equal(other["score"].view(np.uint64), ref["score"].view(np.uint64))
    assert np.array_equal(other["episode_actions"], ref["episode_actions"])


def test_action_ids_outside_the_environment_are_reported(oracle_mod):
    """[REF] TicTacToe.cpp:41,46 (board handler access) and stickGameAdversarial.cpp:61 (LearningEnvironment::doAction) throw
    for an action id past the environment's actions; the device reports GRAPH_INVALID instead of touching memory."""
    for app, env in (("tictactoe", A.ENV_TICTACTOE), ("stickgame", A.ENV_STICKGAME)):
        g = tpgx.synthetic_graph(app, roots=4, edges=3, lines=6, depth=1, seed=2)
        g.edge_destination[:] = -1 - 20          # every edge leads to action 20
        ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=0)
        ev.set_graph(g)
        with pytest.raises(tpgx.TpgxError) as e:
            ev.evaluate_episodes(env, seeds=[1], job_roots=np.arange(4), max_actions=5, second_player_random=1)
        assert e.value.status == A.ERR_GRAPH_INVALID
