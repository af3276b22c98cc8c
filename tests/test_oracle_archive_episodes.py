"""Oracle-only checks of the episodic archive restatement ([UP] Archive::addRecording):
probability 1 records every program execution in execution order, so the record count, the per-step grouping and the
recorded observations are tied to the evaluation's own counters and traces."""
import numpy as np

import tpgx
from tpgx import abi as A


def test_probability_one_records_every_execution_in_order(oracle_mod):
    g = tpgx.synthetic_graph("gridworld", roots=12, edges=4, lines=8, depth=2, seed=2)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["gridworld"], tpgx.APP_SOURCES["gridworld"], g, nb_constants=5)
    ev = o.evaluate_episodes(A.ENV_GRIDWORLD, seeds=[0], job_roots=np.arange(12), max_actions=100)
    total = ev["counters"]["program_executions"]
    ar = o.archive_episodes(A.ENV_GRIDWORLD, seeds=[0], job_roots=np.arange(12), max_actions=100, archive_seeds=np.zeros(12, np.uint64),
                            probability=1.0, archive_size=total + 10, obs_len=2)
    assert len(ar["program"]) == total
    # job-major, then chronological
    key = ar["job"].astype(np.int64) * 1000 + ar["step"]
    assert np.all(np.diff(key) >= 0)
    # every job recorded exactly its episode's steps
    for j in range(12):
        steps = ar["step"][ar["job"] == j]
        assert steps.min() == 0 and steps.max() == ev["episode_actions"][j, 0] - 1
    # gridworld observations are the integer cell coordinates, start cell (0, 0)
    first = ar["observation"][ar["step"] == 0]
    assert np.all(first == 0.0)
    assert np.all(ar["observation"] == np.round(ar["observation"]))


def test_ring_keeps_the_last_records_and_probability_zero_none(oracle_mod):
    g = tpgx.synthetic_graph("stickgame", roots=6, edges=5, lines=8, depth=2, seed=5)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["stickgame"], tpgx.APP_SOURCES["stickgame"], g, nb_constants=0)
    kw = dict(seeds=np.arange(20, dtype=np.uint64), job_roots=np.arange(6), max_actions=11, second_player_random=1,
              archive_seeds=np.arange(6, dtype=np.uint64) + 1, obs_len=5)
    full = o.archive_episodes(A.ENV_STICKGAME, probability=0.2, archive_size=100000, **kw)
    ring = o.archive_episodes(A.ENV_STICKGAME, probability=0.2, archive_size=25, **kw)
    assert len(full["program"]) > 60 and len(ring["program"]) == 25
    # per-job rings of 25, then the merged ring of 25: the last 25 of the last job(s) of the full list
    per_job = np.concatenate([np.flatnonzero(full["job"] == j)[-25:] for j in range(6)])
    assert np.array_equal(ring["program"], full["program"][per_job][-25:])
    assert np.array_equal(ring["bid"].view(np.uint64), full["bid"][per_job][-25:].view(np.uint64))
    assert len(o.archive_episodes(A.ENV_STICKGAME, probability=0.0, archive_size=25, **kw)["program"]) == 0
