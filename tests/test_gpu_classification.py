"""GPU parity, MNIST-shaped classification path: libtpgx (through the C ABI) vs the CPU oracle on the
same seeded inputs — bids, actions, confusion tables, F1 scores and work counters, all bit-exact."""
import numpy as np
import pytest

import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu


def make(app="mnist", tie=A.TIE_LAST_MAX, **kw):
    g = tpgx.synthetic_graph(app, **kw)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=tpgx.APP_CONSTANTS[app], tie_break=tie)
    ev.set_graph(g)
    return g, ev


def oracle_for(oracle_mod, g, app="mnist", tie=A.TIE_LAST_MAX):
    return oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app], tie_break=tie)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def compare(dev, ref, keys=("score", "class_f1", "confusion", "action", "bid")):
    for k in keys:
        if k in dev and k in ref:
            if dev[k].dtype == np.float64:
                # the device leaves NaN in bid rows of programs no evaluated root can reach (never interpreted);
                # every interpreted bid is NaN-free because NaN -> -inf is applied like evaluateEdge
                done = ~np.isnan(dev[k])
                assert done.any()
                assert np.array_equal(bits(dev[k])[done], bits(ref[k])[done]), k
            else:
                assert np.array_equal(dev[k], ref[k]), k
    for k in ("evaluations", "team_visits", "program_executions", "lines_executed"):
        assert dev["counters"][k] == ref["counters"][k], k


@pytest.mark.parametrize("mix,depth,tie", [("uniform", 1, A.TIE_LAST_MAX), ("uniform", 3, A.TIE_LAST_MAX),
                                           ("scalar", 2, A.TIE_FIRST_MAX), ("arith", 1, A.TIE_FIRST_MAX)])
def test_bids_actions_scores_bit_exact(oracle_mod, mix, depth, tie):
    g, ev = make(roots=24, edges=5, lines=20, depth=depth, mix=mix, share_prob=0.1, seed=7 + depth, tie=tie, intron_lines=3)
    img, lab = tpgx.synthetic_images(300, seed=5)
    ev.set_observations(img, lab)
    want = ("score", "class_f1", "confusion", "action", "bid", "counters")
    dev = ev.evaluate_classification(want=want)
    ref = oracle_for(oracle_mod, g, tie=tie).evaluate_classification(img, lab, want=want)
    compare(dev, ref)


def test_ragged_and_empty_programs(oracle_mod):
    """Programs of different lengths, including empty ones (bid 0.0, Appendix F U5), and a sample count that is
    not a multiple of the tile."""
    g, _ = make(roots=16, edges=4, lines=6, depth=2, seed=3, line_jitter=6, intron_lines=2)
    # programs 0 and 5 lose all their lines, program 9 keeps only introns (a line writing a register nobody reads)
    off = g.program_line_offset
    keep = np.ones(len(g.lines), dtype=bool)
    for p_ in (0, 5):
        keep[off[p_]:off[p_ + 1]] = False
    keep[off[9] + 1:off[10]] = False
    g.lines["destination"][off[9]] = 3
    counts = np.add.reduceat(keep.astype(np.int64), off[:-1])
    g = tpgx.TPGraph(lines=g.lines[keep], program_line_offset=np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), constants=g.constants,
                     team_edge_offset=g.team_edge_offset, edge_program=g.edge_program, edge_destination=g.edge_destination, root_team=g.root_team)
    assert (np.diff(g.program_line_offset) == 0).sum() == 2
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    img, lab = tpgx.synthetic_images(77, seed=9)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "bid", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_sample_index_list_app_faithful_draw(oracle_mod):
    """The app-faithful episode: 500 images drawn (with repetition) from the uploaded set
    ([REF] mnist/src/mnist.cpp:16-19, maxNbActionsPerEval = 500)."""
    g, ev = make(roots=20, edges=5, lines=12, depth=2, seed=11)
    img, lab = tpgx.synthetic_images(2000, seed=2)
    ev.set_observations(img, lab)
    idx = tpgx.mnist_episode_indices(seed=7 ^ 0, mode=A.MODE_TRAINING, dataset_size=2000, nb_actions=500)   # Hash(gen) ^ Hash(it), gen 7
    assert len(np.unique(idx)) < 500 <= len(idx)
    want = ("score", "class_f1", "confusion", "action", "counters")
    dev = ev.evaluate_classification(sample_index=idx, want=want)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, sample_index=idx, want=want)
    compare(dev, ref)
    # and again over the full set afterwards (tile cache must be rebuilt)
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_special_values_nan_inf_negzero(oracle_mod):
    """All-zero and all-255 images drive log(0), x/0, 0/0, atan(NaN), exp overflow: bids still bit-exact."""
    g, ev = make(roots=32, edges=5, lines=20, depth=1, seed=21)
    img = np.zeros((96, 784), dtype=np.uint8)
    img[32:64] = 255
    img[64:] = np.random.default_rng(1).integers(0, 2, (32, 784)) * 255
    lab = (np.arange(96) % 10).astype(np.uint8)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "bid", "counters")
    dev = ev.evaluate_classification(want=want)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want)
    assert np.isinf(ref["bid"]).any()
    compare(dev, ref)


def test_root_shards_concatenate_to_full_result(oracle_mod):
    """Roots shard across GPUs: evaluating [0,a) and [a,R) separately equals the full evaluation."""
    g, ev = make(roots=30, edges=5, lines=10, depth=3, share_prob=0.2, seed=5)
    img, lab = tpgx.synthetic_images(200, seed=4)
    ev.set_observations(img, lab)
    full = ev.evaluate_classification(want=("score", "class_f1", "confusion"))
    a = ev.evaluate_classification(root_begin=0, root_end=13, want=("score", "class_f1", "confusion"))
    b = ev.evaluate_classification(root_begin=13, root_end=30, want=("score", "class_f1", "confusion"))
    for k in ("score", "class_f1", "confusion"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), full[k])


def test_multi_pass_bid_budget_matches_single_pass(oracle_mod, monkeypatch):
    """A tiny bid-matrix budget forces several sample passes; results are unchanged."""
    img, lab = tpgx.synthetic_images(1000, seed=8)
    g = tpgx.synthetic_graph("mnist", roots=12, edges=5, lines=8, depth=2, seed=2)
    outs = []
    for budget in ("0", None):
        if budget is not None:
            monkeypatch.setenv("TPGX_BID_BUDGET_MB", budget)
        else:
            monkeypatch.delenv("TPGX_BID_BUDGET_MB", raising=False)
        ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
        ev.set_graph(g)
        ev.set_observations(img, lab)
        outs.append(ev.evaluate_classification(want=("score", "confusion", "action", "counters")))
    for k in ("score", "confusion", "action"):
        assert np.array_equal(outs[0][k], outs[1][k])
    assert outs[0]["counters"] == outs[1]["counters"]


def test_graph_errors_are_reported(oracle_mod):
    """A team whose only edges lead back to visited teams has no admissible edge: the reference throws,
    the device path reports GRAPH_INVALID (no silent fallback)."""
    g = tpgx.synthetic_graph("mnist", roots=2, edges=2, lines=4, depth=1, seed=1)
    g.edge_destination[:] = [1, 1, 0, 0]  # T0 -> T1 -> T0 only
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    img, lab = tpgx.synthetic_images(40, seed=1)
    ev.set_observations(img, lab)
    with pytest.raises(tpgx.TpgxError) as e:
        ev.evaluate_classification()
    assert e.value.status == A.ERR_GRAPH_INVALID


def test_variants_agree(oracle_mod, monkeypatch):
    """Every K1 tile-shape variant produces the same bits."""
    img, lab = tpgx.synthetic_images(333, seed=3)
    g = tpgx.synthetic_graph("mnist", roots=10, edges=5, lines=15, depth=2, seed=9)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=("score", "bid", "action", "confusion", "counters"))
    for v in range(9):
        monkeypatch.setenv("TPGX_K1_VARIANT", str(v))
        ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
        ev.set_graph(g)
        ev.set_observations(img, lab)
        compare(ev.evaluate_classification(want=("score", "bid", "action", "confusion", "counters")), ref)


@pytest.mark.parametrize("lines", [28, 29, 57, 90])
def test_long_programs_cross_the_staging_buffer(oracle_mod, lines):
    """K1 stages 28 lines of a program at a time in shared memory (TPGX_K1_CAPL); programs at, just past and
    several times past that size (tic-tac-toe allows 40 lines: [REF] tic-tac-toe/params.json:30) must give the
    same bits, registers carried across the reloads."""
    g, ev = make(roots=6, edges=3, lines=lines, depth=1, seed=100 + lines, intron_lines=1)
    img, lab = tpgx.synthetic_images(130, seed=2)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "bid", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_uniform_mix_medium(oracle_mod):
    """A few thousand lines of the calibrated uniform mix (every instruction ~10 %, both operand kinds, the
    memoised exp / ln-of-pixel table and the Sobel planes) over several tiles."""
    g, ev = make(roots=60, edges=5, lines=20, depth=2, share_prob=0.1, seed=77)
    ins = np.bincount(g.lines["instruction"], minlength=10) / len(g.lines)
    assert ins.min() > 0.06 and ins.max() < 0.14
    img, lab = tpgx.synthetic_images(500, seed=21)
    ev.set_observations(img, lab)
    want = ("score", "class_f1", "confusion", "action", "bid", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_image_borders_and_extreme_pixels(oracle_mod):
    """All-0 and all-255 images and single bright pixels in every corner: window positions on the image border,
    gx = gy = 0 (sobelDir = atan(0/0) = NaN -> bid -inf), the largest gradients (+-1020) and log(0)."""
    g, ev = make(roots=12, edges=5, lines=12, depth=1, seed=31)
    img = np.zeros((64, 784), dtype=np.uint8)
    img[1] = 255
    for i, px in enumerate((0, 27, 756, 783, 28, 55, 392)):
        img[2 + i, px] = 255
    img[10:20] = np.random.default_rng(1).integers(0, 2, (10, 784)).astype(np.uint8) * 255   # hard 0 / 255 edges
    img[20:] = tpgx.synthetic_images(44, seed=4)[0]
    lab = (np.arange(64) % 10).astype(np.uint8)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "bid", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_full_size_properties():
    """BASELINE C2 size (2000 roots x 60 000 samples): too large for the oracle, checked through size-independent
    properties — every confusion table sums to the sample count, label marginals are the same for all roots,
    scores lie in [0, 1], counters equal the closed form for a depth-1 graph, and the evaluation is
    deterministic and shard-invariant (root halves concatenate to the full result)."""
    R, E, L, S = 2000, 5, 20, 60000
    g, ev = make(roots=R, edges=E, lines=L, depth=1, seed=42)
    img, lab = tpgx.synthetic_images(S)
    ev.set_observations(img, lab)
    out = ev.evaluate_classification(want=("score", "class_f1", "confusion", "counters"))
    conf = out["confusion"]
    assert conf.shape == (R, 10, 10) and (conf.sum(axis=(1, 2)) == S).all()
    assert (conf.sum(axis=2) == np.bincount(lab, minlength=10)[None, :]).all()
    assert ((out["score"] >= 0) & (out["score"] <= 1)).all() and np.isfinite(out["class_f1"]).all()
    seq = np.zeros(R)                                   # [UP] ClassificationLearningAgent: sequential sum over the classes, then / nbClasses
    for c_ in range(10):
        seq = seq + out["class_f1"][:, c_]
    assert np.array_equal(bits(out["score"]), bits(seq / 10.0))
    c = out["counters"]
    assert (c["evaluations"], c["team_visits"], c["program_executions"], c["lines_executed"]) == (R * S, R * S, R * E * S, R * E * L * S)
    again = ev.evaluate_classification(want=("score", "confusion"))
    assert np.array_equal(bits(again["score"]), bits(out["score"])) and np.array_equal(again["confusion"], conf)
    lo = ev.evaluate_classification(root_begin=0, root_end=R // 2, want=("score",))["score"]
    hi = ev.evaluate_classification(root_begin=R // 2, root_end=R, want=("score",))["score"]
    assert np.array_equal(bits(np.concatenate([lo, hi])), bits(out["score"]))


@pytest.mark.parametrize("depth,share,tie", [(1, 0.0, A.TIE_LAST_MAX), (3, 0.05, A.TIE_LAST_MAX), (3, 0.05, A.TIE_FIRST_MAX), (4, 0.0, A.TIE_FIRST_MAX)])
def test_team_mode_fused_evaluate_team(oracle_mod, depth, share, tie, monkeypatch):
    """Without a bid output the library fuses [UP] evaluateTeam into the interpreter (team mode): actions, confusion
    tables, scores and reference-equivalent counters must equal the oracle's AND the unfused bid-matrix path's."""
    g, ev = make(roots=40, edges=5, lines=14, depth=depth, share_prob=share, seed=50 + depth, tie=tie, intron_lines=2)
    img, lab = tpgx.synthetic_images(333, seed=6)
    ev.set_observations(img, lab)
    want = ("score", "class_f1", "confusion", "action", "counters")
    dev = ev.evaluate_classification(want=want)
    ref = oracle_for(oracle_mod, g, tie=tie).evaluate_classification(img, lab, want=want)
    compare(dev, ref)
    monkeypatch.setenv("TPGX_TEAM_MODE", "0")
    g2, ev2 = make(roots=40, edges=5, lines=14, depth=depth, share_prob=share, seed=50 + depth, tie=tie, intron_lines=2)
    ev2.set_observations(img, lab)
    unfused = ev2.evaluate_classification(want=want)
    compare(dev, unfused)
    assert unfused["counters"]["device_lines"] <= dev["counters"]["device_lines"]   # shared programs run once per team when fused


def test_cyclic_graph_uses_visited_team_rule(oracle_mod):
    """Edges back to already visited teams are skipped BEFORE evaluation ([UP] evaluateTeam, Appendix F U3).  A graph
    with cycles (every inner team also points back to its root and to itself) cannot be fused; results must still
    match the oracle."""
    g = tpgx.synthetic_graph("mnist", roots=12, edges=5, lines=10, depth=3, seed=91)
    ed = g.edge_destination.copy()
    teo = g.team_edge_offset
    rng = np.random.default_rng(3)
    for t in range(12, g.nb_teams):                       # inner teams: edge 1 -> a root, edge 2 -> itself (edge 0 stays an action)
        ed[teo[t] + 1] = int(rng.integers(0, 12))
        ed[teo[t] + 2] = t
    g.edge_destination = ed
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    img, lab = tpgx.synthetic_images(200, seed=12)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


@pytest.mark.parametrize("ops_from", ["pendulum", "tictactoe"])
def test_extended_instruction_build_of_k1(oracle_mod, ops_from):
    """K1 has two builds: the MNIST instruction set only and the extended one (cos sin tan pi fmod-guard, the equality
    tests, cond).  An image-classification agent is free to use any double-typed instruction, so the other apps'
    sets are run over the MNIST data source too."""
    opcodes = tpgx.INSTRUCTION_SETS[ops_from]
    nc = tpgx.APP_CONSTANTS[ops_from]
    rng = np.random.default_rng(17)
    P, R, E = 60, 12, 5
    lines, plo, consts = tpgx.random_programs(rng, P, 10, opcodes, tpgx.APP_SOURCES["mnist"], 8, nc)
    g = tpgx.TPGraph(lines=lines, program_line_offset=plo, constants=consts, team_edge_offset=(np.arange(R + 1) * E).astype(np.int32),
                     edge_program=np.arange(P, dtype=np.int32), edge_destination=(-1 - rng.integers(0, 10, P)).astype(np.int32),
                     root_team=np.arange(R, dtype=np.int32))
    img, lab = tpgx.synthetic_images(150, seed=33)
    for want in (("score", "confusion", "action", "bid", "counters"), ("score", "confusion", "action", "counters")):   # bid mode, team mode
        ev = tpgx.Evaluator(opcodes, tpgx.APP_SOURCES["mnist"], nb_constants=nc)
        ev.set_graph(g)
        ev.set_observations(img, lab)
        dev = ev.evaluate_classification(want=want)
        ref = oracle_mod.Oracle(opcodes, tpgx.APP_SOURCES["mnist"], g, nb_constants=nc).evaluate_classification(img, lab, want=want)
        compare(dev, ref)


def test_pinned_observation_upload_equals_blocking_upload(oracle_mod):
    """tpgx_set_observations_pinned only enqueues the copies (page-locked buffers); results are unchanged, and a
    pageable buffer is refused."""
    import torch
    g, ev = make(roots=10, edges=4, lines=10, depth=2, seed=4)
    img, lab = tpgx.synthetic_images(300, seed=8)
    pin_img, pin_lab = torch.from_numpy(img).pin_memory().numpy(), torch.from_numpy(lab).pin_memory().numpy()
    ev.set_observations_pinned(pin_img, pin_lab)
    want = ("score", "confusion", "action", "counters")
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))
    with pytest.raises(tpgx.TpgxError) as e:
        ev.set_observations_pinned(img.copy(), lab.copy())
    assert e.value.status == A.ERR_BAD_ARG


def test_pinned_uploads_interleave_with_graph_uploads_and_other_observations(oracle_mod):
    """The pinned upload runs on its own stream: every consumer of the observation buffers (identity tiles, an index-list
    re-tiling, a later blocking upload, the archive pass) is ordered after it, and a second pinned upload waits for the
    evaluation still reading the first."""
    import torch
    g1, ev = make(roots=12, edges=4, lines=10, depth=2, seed=21)
    g2 = tpgx.synthetic_graph("mnist", roots=9, edges=5, lines=12, depth=1, seed=22)
    want = ("score", "confusion", "action", "counters")
    sets = [tpgx.synthetic_images(n, seed=s) for n, s in ((700, 31), (333, 32), (512, 33))]
    pins = [(torch.from_numpy(i).pin_memory().numpy(), torch.from_numpy(l).pin_memory().numpy()) for i, l in sets]
    for rep in range(3):
        for k, ((img, lab), (pi, pl)) in enumerate(zip(sets, pins)):
            ev.set_observations_pinned(pi, pl)          # enqueue only
            g = g1 if (rep + k) % 2 == 0 else g2
            ev.set_graph(g)                             # graph uploads while the observations are still in flight
            o = oracle_for(oracle_mod, g)
            if k == 1:                                  # index list: re-tiling on the main stream reads the raw upload
                idx = np.random.default_rng(rep).integers(0, len(img), 200).astype(np.int32)
                compare(ev.evaluate_classification(sample_index=idx, want=want), o.evaluate_classification(img, lab, sample_index=idx, want=want))
            else:
                compare(ev.evaluate_classification(want=want), o.evaluate_classification(img, lab, want=want))
    # pinned, then a blocking upload of other data, then evaluate: the blocking upload must win
    ev.set_observations_pinned(*pins[0])
    ev.set_observations(*sets[2])
    ev.set_graph(g1)
    compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g1).evaluate_classification(*sets[2], want=want))
    # pinned, then the archive pass
    ev.set_observations_pinned(*pins[1])
    seeds = np.arange(g1.nb_roots, dtype=np.uint64) + 3
    dev = ev.archive_classification(seeds, 0.05, 64)
    ref = oracle_for(oracle_mod, g1).archive_classification(sets[1][0], seeds, 0.05, 64)
    assert np.array_equal(dev["program"], ref["program"]) and np.array_equal(dev["sample"], ref["sample"])
    assert np.array_equal(dev["bid"].view(np.uint64), ref["bid"].view(np.uint64))


# ---- oracle-checked slices of the benchmarked configurations (BASELINE C2: 2000 roots, C4: 8192 roots, 60 000 samples) ----
import os

_THREADS = max(1, os.cpu_count() or 1)


def _headline(roots):
    g, ev = make(roots=roots, edges=5, lines=20, depth=1, seed=42)
    img, lab = tpgx.synthetic_images(60000)
    ev.set_observations(img, lab)
    return g, ev, img, lab


def test_c2_root_slice_all_samples_vs_oracle(oracle_mod):
    """C2 as bench.py runs it (G(2000, 5, 20, 1, uniform, 42), 60 000 samples): 64 roots of the shard through
    root_begin / root_end on ALL samples — actions, confusion tables, score bits and counters against the oracle."""
    g, ev, img, lab = _headline(2000)
    want = ("score", "class_f1", "confusion", "action", "counters")
    for rb in (0, 1936):
        dev = ev.evaluate_classification(root_begin=rb, root_end=rb + 64, want=want)
        ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, root_begin=rb, root_end=rb + 64, want=want, nb_threads=_THREADS)
        compare(dev, ref)


@pytest.mark.parametrize("roots", [2000, 8192])
def test_headline_all_roots_sample_slice_vs_oracle(oracle_mod, roots):
    """C2 / C4: ALL roots of the graph on 2048 of the 60 000 samples (an index list spread over the whole set)."""
    g, ev, img, lab = _headline(roots)
    idx = (np.arange(2048, dtype=np.int64) * 29 + 7).astype(np.int32) % 60000
    want = ("score", "class_f1", "confusion", "action", "counters")
    dev = ev.evaluate_classification(sample_index=idx, want=want)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, sample_index=idx, want=want, nb_threads=_THREADS)
    compare(dev, ref)
    # and the full evaluation that follows (what bench.py times) is consistent with it on those samples' share:
    full = ev.evaluate_classification(want=("score", "confusion"))
    assert (full["confusion"].sum(axis=(1, 2)) == 60000).all()


def test_k1_v2_equals_k1_v1_and_oracle(oracle_mod, monkeypatch):
    """The threaded interpreter (K1 v2, the default for MNIST-set shards) and K1 v1 give the same bids, in bid and in team
    mode, on programs with introns, shared programs, ragged lengths and several sample tiles."""
    g = tpgx.synthetic_graph("mnist", roots=80, edges=5, lines=20, depth=2, share_prob=0.1, seed=7, intron_lines=4, line_jitter=6)
    img, lab = tpgx.synthetic_images(700, seed=7)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=("score", "confusion", "action", "bid", "counters"))
    for v2 in ("1", "0"):
        monkeypatch.setenv("TPGX_K1_V2", v2)
        ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
        ev.set_graph(g)
        ev.set_observations(img, lab)
        compare(ev.evaluate_classification(want=("score", "confusion", "action", "bid", "counters")), ref)
        compare(ev.evaluate_classification(want=("score", "confusion", "action", "counters")), ref)


def _wide_program(nb_live):
    """r1..r<nb_live> each hold a pixel sum that stays live until a final chain of additions into r0."""
    ln = np.zeros(2 * nb_live, dtype=tpgx.LINE_DTYPE)
    ADD, IMG = 1, 2          # mnist set index of add; source 2 = the image (0 registers, 1 constants)
    for k in range(nb_live):
        ln[k] = (ADD, (k + 1) % 8, (IMG, IMG), (10 * k + 3, 10 * k + 40))
    ln[nb_live] = (ADD, 0, (0, 0), (1 % 8, 2 % 8))
    for k in range(2, nb_live):
        ln[nb_live + k - 1] = (ADD, 0, (0, 0), (0, (k + 1) % 8))
    return ln[:2 * nb_live - 1]


@pytest.mark.parametrize("nb_live", [3, 5, 6, 7])
def test_programs_needing_more_slots_than_k1_v2_has_run_on_v1(oracle_mod, nb_live):
    """K1 v2 keeps at most 5 values of a program in shared memory besides the accumulator; a program that holds more run
    on K1 v1 — same bits either way."""
    ln = _wide_program(nb_live)
    enc = tpgx.k1v2_encode_program(tpgx.flatten_programs(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], tpgx.TPGraph(
        lines=ln, program_line_offset=np.array([0, len(ln)], dtype=np.int64), constants=np.zeros((1, 5)), team_edge_offset=np.array([0, 1], dtype=np.int32),
        edge_program=np.zeros(1, dtype=np.int32), edge_destination=np.full(1, -1, dtype=np.int32), root_team=np.zeros(1, dtype=np.int32)), nb_constants=5)["code"])
    assert enc["slots"] == nb_live
    base = tpgx.synthetic_graph("mnist", roots=6, edges=3, lines=8, depth=1, seed=3)
    lines = np.concatenate([base.lines, ln])
    plo = np.concatenate([base.program_line_offset, [len(lines)]]).astype(np.int64)
    consts = np.concatenate([base.constants, np.zeros((1, 5))])
    ep = base.edge_program.copy()
    ep[4] = base.nb_programs                     # one edge of root 1 runs the wide program
    g = tpgx.TPGraph(lines=lines, program_line_offset=plo, constants=consts, team_edge_offset=base.team_edge_offset, edge_program=ep,
                     edge_destination=base.edge_destination, root_team=base.root_team)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    img, lab = tpgx.synthetic_images(300, seed=5)
    ev.set_observations(img, lab)
    for want in (("score", "confusion", "action", "bid", "counters"), ("score", "confusion", "action", "counters")):
        compare(ev.evaluate_classification(want=want), oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want))


def test_mnist_app_shape_against_the_glibc_flavour_oracle(oracle_mod, capsys):
    """The mnist app's evaluation shape ([REF] mnist/params.json: 500 roots, 500 images drawn per generation, [REF]
    mnist/src/mnist.cpp:16-19) against the oracle computing exp / ln / atan with GLIBC (what the real app's CPU agent runs):
    the single-source libm of the GPU differs from glibc in the last ulps of some bids, so bids are compared as a histogram of
    ulp distances, while the chosen actions, confusion tables and root scores — everything selection depends on — must be
    EQUAL.  (Against the `shared` flavour every bid is bit-exact: the tests above.)"""
    g, ev = make(roots=500, edges=5, lines=20, depth=3, share_prob=0.1, seed=2024)
    img, lab = tpgx.synthetic_images(10000, seed=77)
    ev.set_observations(img, lab)
    idx = tpgx.mnist_episode_indices(seed=5 ^ 0, mode=A.MODE_TRAINING, dataset_size=10000, nb_actions=500)
    want = ("score", "class_f1", "confusion", "action", "bid", "counters")
    dev = ev.evaluate_classification(sample_index=idx, want=want)
    std = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5,
                            math=oracle_mod.MATH_STD).evaluate_classification(img, lab, sample_index=idx, want=want, nb_threads=_THREADS)
    assert np.array_equal(dev["action"], std["action"])
    assert np.array_equal(dev["confusion"], std["confusion"])
    assert np.array_equal(bits(dev["score"]), bits(std["score"])) and np.array_equal(bits(dev["class_f1"]), bits(std["class_f1"]))
    a, b = dev["bid"], std["bid"]
    done = ~np.isnan(a)                                       # rows of programs no root reaches stay NaN on the device
    assert np.array_equal(np.isinf(a[done]), np.isinf(b[done])) and not np.isnan(b[done]).any()
    fin = done & np.isfinite(a) & np.isfinite(b)
    ulp = np.abs(a.view(np.int64)[fin] - b.view(np.int64)[fin])
    hist = np.bincount(np.minimum(ulp, 5).astype(np.int64), minlength=6)
    with capsys.disabled():
        print("\n[glibc-flavour parity] finite bids %d: ulp distance 0/1/2/3/4/>=5 = %s (%.2f %% differ); actions, confusion, scores equal"
              % (fin.sum(), hist.tolist(), 100.0 * (ulp > 0).mean()))
    assert (ulp > 0).mean() < 0.05


def test_labels_outside_the_class_range_are_reported():
    """[UP] ClassificationLearningEnvironment::doAction indexes classificationTable.at(currentClass): a label past nbClasses
    throws in the reference; here it is BAD_ARG, not a silent write past the shared confusion table."""
    g, ev = make(roots=4, edges=3, lines=6, depth=1, seed=1)
    img, lab = tpgx.synthetic_images(64, seed=1)
    lab = lab.copy()
    lab[17] = 200
    ev.set_observations(img, lab)
    for want in (("score",), ("score", "bid")):
        with pytest.raises(tpgx.TpgxError) as e:
            ev.evaluate_classification(want=want)
        assert e.value.status == A.ERR_BAD_ARG and "label" in str(e.value)


@pytest.mark.parametrize("ctas,group", [("1", "1"), ("3", "5"), ("148", "2"), ("37", "64")])
def test_k1_v2_work_distribution_does_not_change_results(oracle_mod, monkeypatch, ctas, group):
    """K1 v2's warps pull (tile, unit group) items from one global counter, stage code through per-warp cp.async double
    buffers and share nothing else: any grid size / item size must give the oracle's bits, run after run (compute-sanitizer
    is not available on this pool — profiles/r2_sanitizer.md — so this and the encoder's bound checks stand in for it)."""
    monkeypatch.setenv("TPGX_K1_CTAS", ctas)
    monkeypatch.setenv("TPGX_K1_GROUP", group)
    g, ev = make(roots=70, edges=5, lines=20, depth=3, share_prob=0.1, seed=61, intron_lines=3, line_jitter=8)
    img, lab = tpgx.synthetic_images(1100, seed=6)          # 9 tiles, the last one ragged
    ev.set_observations(img, lab)
    want_b, want_t = ("score", "confusion", "action", "bid", "counters"), ("score", "confusion", "action", "counters")
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want_b)
    for _ in range(3):
        compare(ev.evaluate_classification(want=want_b), ref)
        compare(ev.evaluate_classification(want=want_t), ref)


def test_k1_v2_heavy_handlers_on_wide_operand_ranges(oracle_mod):
    """The division, exp, ln and a*c/10 sequences of K1 v2 are hand-written PTX; programs whose constants span the whole
    double range (2^-1074 .. 2^1023, both signs, zeros, infinities, NaN) push subnormal operands and results, overflow, the
    slivers next to exp's thresholds and every special-operand class through them — including the exits to the generic
    routines — and every bid must carry the oracle's bits.  Shapes: r = (px*c0/10) / (px*c1/10), exp / ln of px*c/10,
    chains of a*c/10, and the reversed-operand forms."""
    rng = np.random.default_rng(2025)
    P = 1500
    MULC, DIV, EXP, LN, SUB, ADD, MUL, MAX = 7, 3, 5, 6, 0, 1, 2, 4     # indices in the mnist instruction set
    REG, CST, IMG = 0, 1, 2
    lines, off, consts = [], [0], np.zeros((P, 5))
    for p in range(P):
        e = rng.uniform(-1074, 1023, 5)
        c = np.ldexp(rng.uniform(1.0, 2.0, 5), e.astype(int)) * rng.choice([-1.0, 1.0], 5)
        special = rng.random(5) < 0.08
        c[special] = rng.choice([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 10.0, 700.5, -745.2, 709.9], special.sum())
        if p % 3 == 0:   # exp's thresholds, in reach of px * c / 10 with px in 1 .. 255
            c[0] = rng.choice([-1.0, 1.0]) * rng.uniform(27.0, 7500.0)
        consts[p] = c
        px = lambda: int(rng.integers(0, 784))
        shape = p % 6
        prog = [(MULC, 1, (IMG, CST), (px(), 0)), (MULC, 2, (IMG, CST), (px(), 1))]
        if shape == 0: prog += [(DIV, 0, (REG, REG), (1, 2))]                                              # DIV_RA / DIV_RR forms
        elif shape == 1: prog += [(DIV, 3, (REG, IMG), (2, px())), (DIV, 0, (IMG, REG), (px(), 3))]        # DIV_AP, DIV_PA
        elif shape == 2: prog += [(EXP, 3, (REG, REG), (1, 0)), (LN, 4, (REG, REG), (2, 0)), (SUB, 0, (REG, REG), (3, 4))]
        elif shape == 3: prog += [(LN, 3, (REG, REG), (2, 0)), (EXP, 0, (REG, REG), (3, 0))]               # exp(ln(x)): accumulator forms
        elif shape == 4: prog += [(MULC, 3, (REG, CST), (2, 2)), (MULC, 3, (REG, CST), (3, 3)), (DIV, 0, (REG, REG), (3, 1))]
        else: prog += [(MUL, 3, (REG, REG), (1, 2)), (MAX, 4, (REG, REG), (3, 1)), (DIV, 5, (REG, REG), (4, 4)), (ADD, 0, (REG, REG), (5, 2))]
        lines += prog
        off.append(len(lines))
    ln = np.zeros(len(lines), dtype=tpgx.LINE_DTYPE)
    for i, (ins, dst, src, loc) in enumerate(lines):
        ln[i] = (ins, dst, src, loc)
    R, E = P // 5, 5
    g = tpgx.TPGraph(lines=ln, program_line_offset=np.array(off, dtype=np.int64), constants=consts, team_edge_offset=(np.arange(R + 1) * E).astype(np.int32),
                     edge_program=np.arange(P, dtype=np.int32), edge_destination=(-1 - rng.integers(0, 10, P)).astype(np.int32),
                     root_team=np.arange(R, dtype=np.int32))
    img = rng.integers(0, 256, (256, 784)).astype(np.uint8)
    img[:, ::7] = rng.choice([0, 1, 2, 10, 100, 255], (256, 112)).astype(np.uint8)
    img[:32] = 0
    lab = (np.arange(256) % 10).astype(np.uint8)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    ev.set_observations(img, lab)
    want = ("score", "confusion", "action", "bid", "counters")
    dev = ev.evaluate_classification(want=want)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, want=want)
    b = ref["bid"]
    assert np.isinf(b).any() and (np.abs(b[np.isfinite(b)]) < 2.3e-308).any() and (np.abs(b[np.isfinite(b)]) > 1e300).any(), "the case must reach subnormal and huge bids"
    compare(dev, ref)
    compare(ev.evaluate_classification(want=("score", "confusion", "action", "counters")), ref)      # team mode


@pytest.mark.parametrize("slices", ["1", "2", "3", "4"])
def test_pipelined_generation_call_equals_the_two_call_form(oracle_mod, monkeypatch, slices):
    """tpgx_evaluate_graph_classification_allgather (graph + evaluation in one call, host work of slice k behind the kernels
    of slice k - 1, one context per slice sharing the observation tiles) gives the records of tpgx_set_graph +
    tpgx_evaluate_classification bit for bit — full sample set and an app-style index draw — and the oracle's on a slice."""
    monkeypatch.setenv("TPGX_PIPELINE", slices)
    g, ev = make(roots=1100, edges=4, lines=12, depth=2, share_prob=0.1, seed=31, intron_lines=3)
    img, lab = tpgx.synthetic_images(700, seed=9)
    ev.set_observations(img, lab)
    two = ev.evaluate_classification(want=("score", "class_f1"))
    one = ev.evaluate_graph_allgather(g)
    assert np.array_equal(bits(one["score"]), bits(two["score"])) and np.array_equal(bits(one["class_f1"]), bits(two["class_f1"]))
    idx = tpgx.mnist_episode_indices(3000, A.MODE_TRAINING, len(img), 300)
    ev.set_graph(g)
    two = ev.evaluate_classification(sample_index=idx, want=("score", "class_f1"))
    one = ev.evaluate_graph_allgather(g, sample_index=idx)
    assert np.array_equal(bits(one["score"]), bits(two["score"])) and np.array_equal(bits(one["class_f1"]), bits(two["class_f1"]))
    # the identity tiles are back for the next full-set call, and the context still works after the pipelined calls
    again = ev.evaluate_graph_allgather(g)
    ev.set_graph(g)
    ref = oracle_for(oracle_mod, g).evaluate_classification(img, lab, root_begin=500, root_end=540, want=("score",))
    assert np.array_equal(bits(again["score"][500:540]), bits(ref["score"]))
    assert np.array_equal(bits(ev.evaluate_classification(want=("score",))["score"]), bits(again["score"]))
