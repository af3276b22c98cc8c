"""Host logic without a GPU: the flattener of libtpgx (tpgx_flatten_programs = what tpgx_set_graph runs before any
upload) against the oracle's restatement of [UP] Program::identifyIntrons, on random programs of every app, plus the
error behaviour of graph validation."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import tpgx
from tpgx import abi as A


def oracle_for(oracle_mod, app, g):
    return oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app])


@pytest.mark.parametrize("app", ["mnist", "pendulum", "gridworld", "stickgame", "tictactoe"])
def test_intron_marking_matches_the_oracle(oracle_mod, app):
    g = tpgx.synthetic_graph(app, roots=30, edges=4, lines=9, depth=2, share_prob=0.1, seed=5, intron_lines=6, line_jitter=4)
    out = tpgx.flatten_programs(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app])
    o = oracle_for(oracle_mod, app, g)
    assert np.array_equal(out["intron"], o.introns())
    assert np.array_equal(out["effective_lines"], o.effective_lines())
    assert len(out["code"]) == int(out["effective_lines"].sum()) == int((out["intron"] == 0).sum())
    assert out["intron"].any() and (out["intron"] == 0).any()


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31), lines=st.integers(1, 45), regs=st.integers(1, 8))
def test_fully_random_lines_property(oracle_mod, seed, lines, regs):
    """Completely random lines (any instruction, destination and admissible source, most of them introns): the
    flattener's liveness pass and the oracle's agree line by line; effective lines are exactly the code words."""
    rng = np.random.default_rng(seed)
    app = ["mnist", "stickgame", "tictactoe", "pendulum"][seed % 4]
    ops, sources, nc = tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], tpgx.APP_CONSTANTS[app]
    P = 6
    n_src = (2 if nc else 1) + len(sources)
    ln = np.zeros(P * lines, dtype=tpgx.LINE_DTYPE)
    ln["instruction"] = rng.integers(0, len(ops), P * lines)
    ln["destination"] = rng.integers(0, regs, P * lines)
    ln["loc"] = rng.integers(0, 1000, (P * lines, 2))
    sig = [A.OP_SIGNATURE[o] for o in ops]
    for i in range(P * lines):
        for k, t in enumerate(sig[ln["instruction"][i]]):
            adm = [s for s in range(n_src) if tpgx.address_space(sources, regs, nc, s, t) > 0]
            ln["src"][i, k] = adm[rng.integers(0, len(adm))]
    g = tpgx.TPGraph(lines=ln, program_line_offset=(np.arange(P + 1) * lines).astype(np.int64), constants=np.zeros((P, nc)),
                     team_edge_offset=np.array([0, P], dtype=np.int32), edge_program=np.arange(P, dtype=np.int32),
                     edge_destination=np.full(P, -1, dtype=np.int32), root_team=np.zeros(1, dtype=np.int32))
    out = tpgx.flatten_programs(ops, sources, g, nb_registers=regs, nb_constants=nc)
    o = oracle_mod.Oracle(ops, sources, g, nb_registers=regs, nb_constants=nc)
    assert np.array_equal(out["intron"], o.introns())
    assert len(out["code"]) == int((out["intron"] == 0).sum())


def test_validation_errors_mirror_the_reference():
    g = tpgx.synthetic_graph("stickgame", roots=3, edges=2, lines=4, seed=1)
    ops, src = tpgx.INSTRUCTION_SETS["stickgame"], tpgx.APP_SOURCES["stickgame"]
    bad = tpgx.TPGraph(**{**g.__dict__, "lines": g.lines.copy()})
    bad.lines["instruction"][0] = 99
    with pytest.raises(tpgx.TpgxError) as e:
        tpgx.flatten_programs(ops, src, bad)
    assert e.value.status == A.ERR_BAD_ARG and "instruction index" in str(e.value)
    bad = tpgx.TPGraph(**{**g.__dict__, "lines": g.lines.copy()})
    i = int(np.nonzero(bad.lines["instruction"] == 1)[0][0]) if (bad.lines["instruction"] == 1).any() else 0
    bad.lines["instruction"][i] = 1          # minus(int, int) ...
    bad.lines["src"][i] = (0, 0)             # ... reading the double-typed registers: the reference engine throws
    with pytest.raises(tpgx.TpgxError) as e:
        tpgx.flatten_programs(ops, src, bad)
    assert e.value.status == A.ERR_GRAPH_INVALID
    empty_team = tpgx.TPGraph(**{**g.__dict__, "team_edge_offset": np.array([0, 2, 2, 6], dtype=np.int32)})
    with pytest.raises(tpgx.TpgxError) as e:
        tpgx.flatten_programs(ops, src, empty_team)
    assert e.value.status == A.ERR_GRAPH_INVALID and "no outgoing edge" in str(e.value)


@pytest.mark.parametrize("seed,rb,re", [(3, 0, 17), (4, 40, 41), (5, 13, 60), (6, 0, 60)])
def test_slice_graph_is_the_renumbered_reachable_subgraph(seed, rb, re):
    """tpgx_slice_graph — what every pipeline slice of tpgx_evaluate_graph_classification_allgather is flattened, planned and
    uploaded as — against the Python restatement TPGraph.subgraph: teams breadth-first from the roots in root order, programs
    in order of first appearance, edge order within a team kept, shared programs and shared teams appear once."""
    g = tpgx.synthetic_graph("mnist", roots=60, edges=4, lines=8, depth=3, share_prob=0.3, seed=seed, intron_lines=2)
    s = tpgx.slice_graph(g, rb, re)
    ref = g.subgraph(rb, re)
    assert np.array_equal(s["team_edge_offset"], ref.team_edge_offset)
    assert np.array_equal(s["edge_program"], ref.edge_program)
    assert np.array_equal(s["edge_destination"], ref.edge_destination)
    assert np.array_equal(s["root_team"], ref.root_team)
    # the id maps: new -> old, injective; every edge of every kept team maps back to the caller's edge, in order
    assert len(set(s["team_ids"].tolist())) == len(s["team_ids"]) and len(set(s["program_ids"].tolist())) == len(s["program_ids"])
    assert np.array_equal(s["team_ids"][s["root_team"]], np.asarray(g.root_team[rb:re]))
    for new_t, old_t in enumerate(s["team_ids"]):
        old = slice(g.team_edge_offset[old_t], g.team_edge_offset[old_t + 1])
        new = slice(s["team_edge_offset"][new_t], s["team_edge_offset"][new_t + 1])
        assert np.array_equal(s["program_ids"][s["edge_program"][new]], g.edge_program[old])
        d_old, d_new = g.edge_destination[old], s["edge_destination"][new]
        assert np.array_equal(np.where(d_new >= 0, s["team_ids"][np.maximum(d_new, 0)], d_new), d_old)
    # the programs' lines are the caller's (the library reads them in place through program_ids)
    assert np.array_equal(ref.program_line_offset[1:] - ref.program_line_offset[:-1],
                          (g.program_line_offset[1:] - g.program_line_offset[:-1])[s["program_ids"]])


def test_slice_graph_rejects_bad_ranges_and_indices():
    g = tpgx.synthetic_graph("mnist", roots=10, edges=3, lines=4, depth=2, seed=1)
    for rb, re in [(-1, 3), (3, 3), (5, 11)]:
        with pytest.raises(tpgx.TpgxError):
            tpgx.slice_graph(g, rb, re)
    bad = tpgx.TPGraph(lines=g.lines, program_line_offset=g.program_line_offset, constants=g.constants, team_edge_offset=g.team_edge_offset,
                       edge_program=np.where(np.arange(g.nb_edges) == 0, g.nb_programs, g.edge_program).astype(np.int32),
                       edge_destination=g.edge_destination, root_team=g.root_team)
    with pytest.raises(tpgx.TpgxError):
        tpgx.slice_graph(bad, 0, 10)
