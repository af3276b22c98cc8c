"""Dot exporter / importer ([UP] File::TPGGraphDotExporter / Importer layout): the reference's
checkpoint format ([REF] mnist/src/main.cpp:140-143)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import tpgx


def same_graph(a, b):
    assert (a.nb_teams, a.nb_roots, a.nb_edges) == (b.nb_teams, b.nb_roots, b.nb_edges)
    assert np.array_equal(a.team_edge_offset, b.team_edge_offset)
    assert np.array_equal(a.edge_destination, b.edge_destination)
    assert np.array_equal(a.root_team, b.root_team)
    for e in range(a.nb_edges):     # program ids may be renumbered; compare what each edge carries
        pa, pb = a.edge_program[e], b.edge_program[e]
        la = a.lines[a.program_line_offset[pa]:a.program_line_offset[pa + 1]]
        lb = b.lines[b.program_line_offset[pb]:b.program_line_offset[pb + 1]]
        assert np.array_equal(la, lb), e
        assert np.allclose(a.constants[pa], b.constants[pb], rtol=0, atol=5e-7), e


@pytest.mark.parametrize("app,depth,share", [("mnist", 3, 0.2), ("stickgame", 2, 0.0), ("tictactoe", 1, 0.3), ("pendulum", 2, 0.1)])
def test_export_then_import_is_identity(tmp_path, app, depth, share):
    g = tpgx.synthetic_graph(app, roots=9, edges=4, lines=7, depth=depth, share_prob=share, seed=13, intron_lines=2)
    path = os.path.join(str(tmp_path), "out_0001.dot")
    tpgx.export_dot(g, path)
    same_graph(g, tpgx.import_dot(path, tpgx.APP_CONSTANTS[app]))


@pytest.mark.parametrize("app,dot", [("pendulum", "Pendulum_out_best.dot"), ("stickgame", "StickGame_out_best.dot"), ("tictactoe", "TicTacToe_out_best.dot")])
def test_golden_files_survive_a_round_trip(tmp_path, app, dot):
    g = tpgx.import_dot(os.path.join(GOLDEN, dot), tpgx.APP_CONSTANTS[app])
    text = tpgx.export_dot(g)
    assert text.startswith("// File exported with GEGELATI v2.0.0")
    same_graph(g, tpgx.import_dot(text, tpgx.APP_CONSTANTS[app]))
