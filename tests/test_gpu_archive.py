"""Archive reproduction ([UP] Archive::addRecording as a side effect of TRAINING evaluation): the
records libtpgx returns equal the oracle's — same programs, same observations, same bids, same order — for several
probabilities, archive sizes, graph shapes (incl. cycles, where the visited-team rule changes which programs run) and the
app-faithful index draw."""
import numpy as np
import pytest

import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu


def pair(oracle_mod, g, tie=A.TIE_LAST_MAX):
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, tie_break=tie)
    ev.set_graph(g)
    return ev, oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5, tie_break=tie)


def same(dev, ref):
    assert len(dev["program"]) == len(ref["program"])
    assert np.array_equal(dev["program"], ref["program"]) and np.array_equal(dev["sample"], ref["sample"])
    assert np.array_equal(dev["bid"].view(np.uint64), ref["bid"].view(np.uint64))


@pytest.mark.parametrize("prob,size", [(0.01, 500), (0.05, 50), (1.0, 37), (0.3, 2000), (0.0, 10)])
@pytest.mark.parametrize("depth", [1, 3])
def test_archive_records_match_the_oracle(oracle_mod, prob, size, depth):
    g = tpgx.synthetic_graph("mnist", roots=45, edges=5, lines=10, depth=depth, share_prob=0.1, seed=60 + depth)
    img, lab = tpgx.synthetic_images(260, seed=14)
    ev, o = pair(oracle_mod, g)
    ev.set_observations(img, lab)
    seeds = np.random.default_rng(9).integers(0, 2 ** 63, g.nb_roots).astype(np.uint64)
    same(ev.archive_classification(seeds, prob, size), o.archive_classification(img, seeds, prob, size))


def test_archive_with_app_faithful_draw_and_cycles(oracle_mod):
    g = tpgx.synthetic_graph("mnist", roots=20, edges=5, lines=9, depth=3, seed=5)
    ed, teo = g.edge_destination.copy(), g.team_edge_offset
    for t in range(20, g.nb_teams):
        ed[teo[t] + 1] = t % 20      # back edge to a root: skipped when that root is on the path
    g.edge_destination = ed
    img, lab = tpgx.synthetic_images(1500, seed=3)
    ev, o = pair(oracle_mod, g, tie=A.TIE_FIRST_MAX)
    ev.set_observations(img, lab)
    idx = tpgx.mnist_episode_indices(seed=12, mode=A.MODE_TRAINING, dataset_size=1500, nb_actions=500)
    seeds = np.arange(20, dtype=np.uint64) * 7919 + 1
    dev = ev.archive_classification(seeds, 0.02, 120, sample_index=idx)
    ref = o.archive_classification(img, seeds, 0.02, 120, sample_index=idx)
    same(dev, ref)
    assert len(dev["program"]) == 120 and set(dev["sample"]).issubset(set(idx.tolist()))
    # the archive call leaves the evaluator usable for a normal evaluation afterwards
    out = ev.evaluate_classification(sample_index=idx, want=("score", "counters"))
    ref2 = o.evaluate_classification(img, lab, sample_index=idx, want=("score", "counters"))
    assert np.array_equal(out["score"].view(np.uint64), ref2["score"].view(np.uint64))


# ---- sequential environments: tpgx_archive_episodes -------------------------------------------------------------------
PEND_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]  # [REF] pendulum/src/Learn/main.cpp:33


def pair_app(oracle_mod, app, g, tie=A.TIE_LAST_MAX):
    nc = tpgx.APP_CONSTANTS[app]
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc, tie_break=tie)
    ev.set_graph(g)
    return ev, oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc, tie_break=tie)


def same_episodes(dev, ref):
    assert len(dev["program"]) == len(ref["program"])
    for k in ("program", "job", "iteration", "step"):
        assert np.array_equal(dev[k], ref[k]), k
    assert np.array_equal(dev["bid"].view(np.uint64), ref["bid"].view(np.uint64))
    if "observation" in ref:
        assert np.array_equal(dev["observation"].view(np.uint64), ref["observation"].view(np.uint64))


@pytest.mark.parametrize("prob,size", [(0.01, 2000), (0.3, 64), (1.0, 41), (0.0, 8)])
def test_episode_archive_gridworld(oracle_mod, prob, size):
    """gridworld/params.json: 1 iteration x <= 100 actions; archive 2000 @ 0.01."""
    g = tpgx.synthetic_graph("gridworld", roots=64, edges=6, lines=10, depth=3, share_prob=0.1, seed=3)
    ev, o = pair_app(oracle_mod, "gridworld", g)
    aseeds = np.random.default_rng(4).integers(0, 2 ** 63, 64).astype(np.uint64)
    kw = dict(seeds=[0], job_roots=np.arange(64), max_actions=100, archive_seeds=aseeds, probability=prob, archive_size=size, obs_len=2)
    dev = ev.archive_episodes(A.ENV_GRIDWORLD, **kw)
    same_episodes(dev, o.archive_episodes(A.ENV_GRIDWORLD, **kw))
    # the archive is a side effect of the evaluation: the scores of the same call equal a plain evaluation's
    plain = ev.evaluate_episodes(A.ENV_GRIDWORLD, seeds=[0], job_roots=np.arange(64), max_actions=100)
    assert np.array_equal(dev["score"].view(np.uint64), plain["score"].view(np.uint64))


def test_episode_archive_stickgame_random_opponent(oracle_mod):
    """100 iterations per job (stickgame/params.json): the job's draw stream runs across its iterations; int data sources."""
    g = tpgx.synthetic_graph("stickgame", roots=30, edges=5, lines=8, depth=2, seed=5)
    ev, o = pair_app(oracle_mod, "stickgame", g)
    seeds = (np.arange(100, dtype=np.uint64) ^ np.uint64(7))
    kw = dict(seeds=seeds, job_roots=np.arange(30), max_actions=11, second_player_random=1, archive_seeds=np.arange(30, dtype=np.uint64) + 5,
              probability=0.01, archive_size=500, obs_len=5)
    dev = ev.archive_episodes(A.ENV_STICKGAME, **kw)
    same_episodes(dev, o.archive_episodes(A.ENV_STICKGAME, **kw))
    assert len(dev["program"]) > 100 and dev["iteration"].max() > 50


def test_episode_archive_tictactoe_adversarial_jobs(oracle_mod):
    """Both roots of an adversarial job feed the job's single archive (tic-tac-toe/params.json: 0.05, 50)."""
    g = tpgx.synthetic_graph("tictactoe", roots=40, edges=8, lines=12, depth=2, seed=9)
    ev, o = pair_app(oracle_mod, "tictactoe", g)
    jobs = np.random.default_rng(0).integers(0, 40, (60, 2))
    kw = dict(seeds=np.arange(10, dtype=np.uint64) + 100, job_roots=jobs, roots_per_job=2, max_actions=9, second_player_random=0,
              archive_seeds=np.arange(60, dtype=np.uint64) * 3 + 1, probability=0.05, archive_size=50, obs_len=9)
    same_episodes(ev.archive_episodes(A.ENV_TICTACTOE, **kw), o.archive_episodes(A.ENV_TICTACTOE, **kw))


def test_episode_archive_pendulum(oracle_mod):
    """5 iterations x 1000 steps (pendulum/params.json), archive 500 @ 0.01: only the tail of the last jobs survives."""
    g = tpgx.synthetic_graph("pendulum", roots=24, edges=5, lines=10, depth=2, seed=13)
    ev, o = pair_app(oracle_mod, "pendulum", g)
    kw = dict(seeds=np.arange(5, dtype=np.uint64), job_roots=np.arange(24), max_actions=1000, pendulum_actions=PEND_ACTIONS,
              archive_seeds=np.arange(24, dtype=np.uint64) + 77, probability=0.01, archive_size=500, obs_len=2)
    dev = ev.archive_episodes(A.ENV_PENDULUM, **kw)
    same_episodes(dev, o.archive_episodes(A.ENV_PENDULUM, **kw))
    assert len(dev["program"]) == 500


@pytest.mark.parametrize("env_var,value", [("TPGX_EP_CTA_WARPS", "0"), ("TPGX_EP_CTA_WARPS", "3"), ("TPGX_EP_GROUPS", "4")])
def test_episode_archive_kernel_mappings_agree(oracle_mod, monkeypatch, env_var, value):
    """The resolution pass finds a recorded execution by its rank among the team's admissible edges: across the edge
    warps of a CTA (several rounds with 3 warps for 6 edges), or across the lanes of a lane group."""
    monkeypatch.setenv(env_var, value)
    if env_var == "TPGX_EP_GROUPS":
        monkeypatch.setenv("TPGX_EP_CTA_WARPS", "0")
    g = tpgx.synthetic_graph("gridworld", roots=40, edges=6, lines=10, depth=3, share_prob=0.1, seed=31)
    ev, o = pair_app(oracle_mod, "gridworld", g)
    aseeds = np.random.default_rng(5).integers(0, 2 ** 63, 40).astype(np.uint64)
    kw = dict(seeds=[3, 4, 5], job_roots=np.arange(40), max_actions=100, archive_seeds=aseeds, probability=0.05, archive_size=300, obs_len=2)
    same_episodes(ev.archive_episodes(A.ENV_GRIDWORLD, **kw), o.archive_episodes(A.ENV_GRIDWORLD, **kw))


def test_episode_archive_wide_teams(oracle_mod):
    """45 edges per team: three rounds of 16 edge warps per team evaluation; ranks continue across the rounds."""
    g = tpgx.synthetic_graph("tictactoe", roots=12, edges=45, lines=6, depth=2, seed=23)
    ev, o = pair_app(oracle_mod, "tictactoe", g)
    jobs = np.random.default_rng(3).integers(0, 12, (20, 2))
    kw = dict(seeds=np.arange(4, dtype=np.uint64) + 11, job_roots=jobs, roots_per_job=2, max_actions=9, second_player_random=0,
              archive_seeds=np.arange(20, dtype=np.uint64) + 9, probability=0.05, archive_size=200, obs_len=9)
    same_episodes(ev.archive_episodes(A.ENV_TICTACTOE, **kw), o.archive_episodes(A.ENV_TICTACTOE, **kw))
