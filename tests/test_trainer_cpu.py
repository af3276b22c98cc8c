"""Host generation loop (include/tpgx.h tpgx_trainer_*) without a GPU: graph invariants of
initRandomTPG / populateTPG / decimateWorstRoots, determinism, params.json loading, and — with the CPU oracle plugged into
the backend callbacks (test infrastructure) — that training a gridworld policy actually learns."""
import os

import numpy as np
import pytest

import tpgx
from tpgx import abi as A
from trainer_util import SMALL, oracle_backend

REF_PARAMS = "/root/reference/gridworld/params.json"
# gridworld/params.json values (the file itself only exists in the build container)
GRID = dict(nb_actions=4, nb_roots=500, max_init_outgoing_edges=4, max_outgoing_edges=10, p_edge_deletion=0.7, p_edge_addition=0.7,
            p_program_mutation=0.2, p_edge_destination_change=0.1, p_edge_destination_is_action=0.5, force_program_behavior_change=1,
            max_program_size=20, p_delete=0.5, p_add=0.5, p_mutate=1.0, p_swap=1.0, p_constant_mutation=0.5, min_const_value=-10,
            max_const_value=10, ratio_deleted_roots=0.5, archive_size=2000, archiving_probability=0.01)


def make(app="gridworld", seed=1, **over):
    p = dict(GRID, nb_actions=tpgx.APP_ACTIONS[app], **over)
    return tpgx.Trainer(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], p, nb_constants=tpgx.APP_CONSTANTS[app], seed=seed), p


@pytest.mark.skipif(not os.path.exists(REF_PARAMS), reason="reference tree not present")
def test_params_json_of_the_app_loads_with_its_comments():
    p = tpgx.load_params_json(REF_PARAMS, nb_actions=4)
    for k, v in GRID.items():
        assert p[k] == v, k
    assert p["max_actions"] == 100 and p["nb_iterations"] == 1


def check_invariants(g, p, app):
    teo, ed, ep = g.team_edge_offset, g.edge_destination, g.edge_program
    n_edges = np.diff(teo)
    assert n_edges.min() >= 1 and n_edges.max() <= p["max_outgoing_edges"]
    incoming = np.zeros(g.nb_teams, dtype=int)
    for t in range(g.nb_teams):
        dst = ed[teo[t]:teo[t + 1]]
        assert (dst < 0).any(), "every team keeps an edge to an action"
        assert dst.min() >= -p["nb_actions"]
        assert (dst[dst >= 0] < t).all(), "edges only lead to older teams: the graph stays acyclic"
        assert len(set(ep[teo[t]:teo[t + 1]].tolist())) == n_edges[t], "a team never holds the same program twice"
        np.add.at(incoming, dst[dst >= 0], 1)
    assert np.array_equal(np.flatnonzero(incoming == 0), g.root_team)
    assert set(ep.tolist()) == set(range(g.nb_programs)), "no orphan program in the view"
    assert np.diff(g.program_line_offset).min() >= 1 and np.diff(g.program_line_offset).max() <= p["max_program_size"]
    # every line is executable: the flattener (what tpgx_set_graph runs) accepts the graph
    tpgx.flatten_programs(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app])


@pytest.mark.parametrize("app", ["gridworld", "stickgame", "tictactoe", "pendulum"])
def test_populate_and_decimate_keep_the_graph_valid(app):
    tr, p = make(app, seed=3, nb_roots=120, force_program_behavior_change=0)
    g = tr.graph()
    assert g.nb_roots == g.nb_teams == p["nb_actions"]
    rng = np.random.default_rng(0)
    for gen in range(6):
        s = tr.populate()
        g = tr.graph()
        assert g.nb_roots == 120 == s["nb_roots"]
        check_invariants(g, p, app)
        d = tr.decimate(rng.normal(size=g.nb_roots))
        assert d["nb_removed_roots"] == 60
        check_invariants(tr.graph(), p, app)


def test_decimate_removes_the_lowest_scores_and_keeps_root_order_on_ties():
    tr, p = make(seed=5, nb_roots=40, force_program_behavior_change=0)
    tr.populate()
    before = tr.graph()
    scores = np.arange(40, dtype=np.float64)[::-1].copy()       # root 39 is the worst
    scores[:10] = 100.0                                         # ten equal best
    tr.decimate(scores)
    after = tr.graph()
    # survivors are the 20 best roots, in their original order; compare their first programs' lines
    def first_lines(g, r):
        t = g.root_team[r]; pr = g.edge_program[g.team_edge_offset[t]]
        return g.lines[g.program_line_offset[pr]:g.program_line_offset[pr + 1]].tobytes()
    assert after.nb_roots >= 20
    kept = [first_lines(before, r) for r in range(20)]
    got = [first_lines(after, r) for r in range(after.nb_roots)]
    it = iter(got)
    assert all(any(k == x for x in it) for k in kept), "the 20 best roots survive, order preserved"


def test_same_seed_same_evolution():
    a, _ = make(seed=11, nb_roots=80, force_program_behavior_change=0)
    b, _ = make(seed=11, nb_roots=80, force_program_behavior_change=0)
    c, _ = make(seed=12, nb_roots=80, force_program_behavior_change=0)
    rng = np.random.default_rng(1)
    for _ in range(3):
        for t in (a, b, c):
            t.populate()
        sc = rng.normal(size=80)
        assert a.fingerprint() == b.fingerprint() != c.fingerprint()
        for t in (a, b, c):
            t.decimate(sc)
    assert a.fingerprint() == b.fingerprint()


def test_gridworld_training_learns_with_the_oracle_backend(oracle_mod):
    """[REF] gridworld/src/main.cpp:57-65 with params.json scaled to 60 roots: the best policy reaches the exit
    (score > 0: +100 on the exit tile, -1 per step) within a few generations; archive and behaviour test active."""
    tr, p = make(seed=5, **SMALL)
    log = []
    be = oracle_backend(oracle_mod, "gridworld", nb_iterations=1, max_actions=100, log=log)
    stats = [tr.train_generation(be, gen) for gen in range(10)]
    best = [s["best_score"] for s in stats]
    assert best[0] <= best[-1] and max(best) > 0.0, best
    assert all(np.array_equal(np.sort(sc)[::-1][0], s["best_score"]) for (_, _, sc), s in zip(log, stats))
    assert stats[-1]["archive_records"] > 0 and sum(s["behaviour_retries"] for s in stats) >= 0
    assert all(s["nb_roots"] == 60 and s["nb_removed_roots"] == 30 for s in stats)


def test_batched_behaviour_test_equals_one_mutant_per_call(oracle_mod, monkeypatch):
    """Every new program mutates with an engine of its own, so trying several successive mutants of a program in one
    backend call (the default: 4, then the remaining 12) must give the evolution of one mutant per call."""
    prints = []
    for attempts in (None, "1", "3"):
        if attempts is None:
            monkeypatch.delenv("TPGX_TRAIN_ATTEMPTS_PER_CALL", raising=False)
        else:
            monkeypatch.setenv("TPGX_TRAIN_ATTEMPTS_PER_CALL", attempts)
        tr, _ = make(seed=9, **SMALL)
        be = oracle_backend(oracle_mod, "gridworld", nb_iterations=1, max_actions=100)
        stats = [tr.train_generation(be, gen) for gen in range(6)]
        prints.append((tr.fingerprint(), [s["behaviour_retries"] for s in stats], [s["best_score"] for s in stats]))
    assert sum(prints[0][1]) > 0, "the behaviour test never had to retry: the case does not exercise it"
    assert prints[0] == prints[1] == prints[2]


def test_evolution_does_not_depend_on_the_host_thread_count(oracle_mod, monkeypatch):
    """The new programs of a generation mutate in parallel on the host pool, each with an engine of its own seeded in order
    from the agent's: the evolution (graph fingerprint, retries, scores) must be the same with 1, 3 and 8 host threads.
    120 roots so that a generation has more new programs than one pool block (16)."""
    prints = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("TPGX_HOST_THREADS", threads)
        tr, _ = make(seed=3, nb_roots=120, archive_size=400, archiving_probability=0.05)
        be = oracle_backend(oracle_mod, "gridworld", nb_iterations=1, max_actions=100)
        stats = [tr.train_generation(be, gen) for gen in range(5)]
        assert max(s["nb_new_programs"] for s in stats) > 32
        prints.append((tr.fingerprint(), [s["behaviour_retries"] for s in stats], [s["best_score"] for s in stats]))
    assert prints[0] == prints[1] == prints[2]


def test_evaluate_job_skips_roots_at_the_evaluation_cap_and_merges_previous_results(oracle_mod):
    """[UP] LearningAgent::evaluateJob ([REF] */params.json maxNbEvaluationPerPolicy: gridworld / pendulum 10, mnist 2000):
    a root whose stored result counts maxNbEvaluationPerPolicy evaluations is NOT run again in TRAINING mode, any other
    root's new result is combined with its stored one (evaluation-count-weighted mean, upstream's operation order).  With
    ratioDeletedRoots = 0 the roots never change, so the stored results can be recomputed from the backend's log."""
    tr, p = make("stickgame", seed=4, nb_roots=24, archive_size=0, ratio_deleted_roots=0.0, force_program_behavior_change=0,
                 max_nb_evaluation_per_policy=6, nb_iterations_per_policy_evaluation=2)
    log = []
    be = oracle_backend(oracle_mod, "stickgame", nb_iterations=2, max_actions=11, second_player_random=1, log=log)
    want, n = None, 0
    for gen in range(5):
        before = len(log)
        st = tr.train_generation(be, gen)
        score, nb = tr.root_results()
        if gen < 3:                                     # 2 evaluations per generation: the cap of 6 is reached after three
            assert st["nb_evaluated_roots"] == 24 and st["nb_skipped_roots"] == 0 and len(log) == before + 1
            mode, jobs, new = log[-1]
            assert mode == A.MODE_TRAINING and jobs == 24
            if want is None:
                want, n = new.copy(), 2
            else:                                       # EvaluationResult::operator+=, `new` on the left
                want = (new * 2.0 + want * float(n)) / (2.0 + float(n))
                n += 2
        else:                                           # every root is at the cap: nothing is launched, results stand
            assert st["nb_evaluated_roots"] == 0 and st["nb_skipped_roots"] == 24 and len(log) == before
        assert np.array_equal(score.view(np.uint64), want.view(np.uint64)) and (nb == n).all()
        assert st["best_score"] == want.max()
    assert len({tuple(s) for (_, _, s) in log}) > 1, "the generations' raw scores must differ for the merge to be visible"


def test_new_roots_are_evaluated_while_capped_ones_are_skipped(oracle_mod):
    """With decimation on, the surviving roots keep their evaluation count and the fresh ones start at zero: the backend
    sees a root list holding only the roots below the cap."""
    tr, p = make(seed=5, max_nb_evaluation_per_policy=2, nb_iterations_per_policy_evaluation=1, **SMALL)
    log = []
    be = oracle_backend(oracle_mod, "gridworld", nb_iterations=1, max_actions=100, log=log)
    stats = [tr.train_generation(be, gen) for gen in range(12)]
    assert stats[0]["nb_skipped_roots"] == 0 and stats[1]["nb_skipped_roots"] == 0
    # a survivor stays a root only until a new team points at it, so few roots ever reach the cap in a small population
    assert sum(s["nb_skipped_roots"] for s in stats[2:]) > 0, [s["nb_skipped_roots"] for s in stats]
    for s, (mode, jobs, sc) in zip(stats, log):
        assert mode == A.MODE_TRAINING and jobs == s["nb_evaluated_roots"] == 60 - s["nb_skipped_roots"] and len(sc) == jobs
    score, nb = tr.root_results()
    assert nb.max() <= 2 and nb.min() >= 1 and not np.isnan(score).any()


def test_validation_pass_runs_after_the_decimation_on_the_survivors(oracle_mod):
    """[UP] trainOneGeneration with doValidation ([REF] mnist/params.json:11, pendulum/params.json:14): a second
    evaluateAllRoots in VALIDATION mode on the roots left by the decimation; its results are reported, not stored."""
    tr, p = make("pendulum", seed=6, do_validation=1, nb_iterations_per_policy_evaluation=1, **SMALL)
    log = []
    pa = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
    be = oracle_backend(oracle_mod, "pendulum", nb_iterations=1, max_actions=60, pendulum_actions=pa, log=log)
    for gen in range(3):
        st = tr.train_generation(be, gen)
        (m0, j0, s0), (m1, j1, s1) = log[-2:]
        # removing 30 roots can uncover teams only they pointed at: the survivors are the graph's roots after the decimation
        assert (m0, j0) == (A.MODE_TRAINING, 60) and m1 == A.MODE_VALIDATION and j1 == tr.graph().nb_roots >= 30
        seq = 0.0
        for v in s1:
            seq += v                                          # the trainer sums sequentially (numpy's mean sums pairwise)
        assert st["best_validation_score"] == s1.max() and st["mean_validation_score"] == seq / len(s1)
        assert not np.array_equal(np.sort(s1)[-20:], np.sort(s0)[-20:]), "validation draws its own initial states (mode enters the seed)"
    off, p2 = make("pendulum", seed=6, do_validation=0, nb_iterations_per_policy_evaluation=1, **SMALL)
    be2 = oracle_backend(oracle_mod, "pendulum", nb_iterations=1, max_actions=60, pendulum_actions=pa)
    for gen in range(3):
        st2 = off.train_generation(be2, gen)
    assert np.isnan(st2["best_validation_score"]) and off.fingerprint() == tr.fingerprint(), "validation must not change the evolution"


def test_best_root_and_keep_best_policy(oracle_mod):
    """[UP] getBestRoot / keepBestPolicy as the apps call them after training ([REF] gridworld/src/main.cpp:68,79)."""
    tr, p = make(seed=5, **SMALL)
    be = oracle_backend(oracle_mod, "gridworld", nb_iterations=1, max_actions=100)
    for gen in range(6):
        st = tr.train_generation(be, gen)
    idx, sc = tr.best_root()
    score, _ = tr.root_results()
    assert sc == score[idx] == np.nanmax(score)
    before = tr.graph()
    tr.keep_best_policy()
    after = tr.graph()
    assert after.nb_roots == 1 and after.nb_teams <= before.nb_teams
    check_invariants(after, p, "gridworld")
    assert tr.best_root()[0] == 0
