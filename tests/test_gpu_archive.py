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
