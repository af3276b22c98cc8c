"""The C++ reference-side binding (include/tpgx_gegelati.hpp) end to end: a program that reads like the apps' own
main.cpp — the REFERENCE's fillInstructionSet and LearningEnvironment classes, compiled unmodified, a golden graph
imported with the dot importer, tpgx::LearningAgent::evaluateAllRoots — runs the hot path on the GPU; its per-root
scores must equal the oracle's.  The binary is built by oracle/build_ref.sh in the build container (it contains
reference code, so it lives in oracle/_ref/ and travels to the GPU box with the snapshot)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "oracle", "_ref", "adapter_driver")
PEND_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]


def run_driver(app, dot, generation, mode, iterations, max_actions, second_random=0):
    out = subprocess.run([DRIVER, app, dot, str(generation), str(mode), str(iterations), str(max_actions), str(second_random)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.strip().splitlines()]
    return np.array([float.fromhex(r[0]) for r in rows]), [int(r[1]) for r in rows]


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/adapter_driver not built (needs /root/reference at build time)")
@pytest.mark.parametrize("app,dot,iters,maxa,spr,exact", [
    ("stickgame", "StickGame_out_best.dot", 100, 11, 1, True),       # [REF] stickgame/params.json:15,97; random second player
    ("tictactoe", "TicTacToe_out_best.dot", 10, 5, 1, True),         # [REF] tic-tac-toe/params.json:15,97
    ("pendulum", "Pendulum_out_best.dot", 5, 1000, 0, True),         # [REF] pendulum/params.json:18,142 (vs the shared-libm oracle)
])
def test_reference_sources_drive_the_gpu_evaluator(oracle_mod, app, dot, iters, maxa, spr, exact):
    path = os.path.join(GOLDEN, dot)
    generation, mode = 3, A.MODE_TRAINING
    scores, nb_eval = run_driver(app, path, generation, mode, iters, maxa, spr)
    g = tpgx.import_dot(path, tpgx.APP_CONSTANTS[app])
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=tpgx.APP_CONSTANTS[app])
    ref = o.evaluate_episodes(tpgx.APP_ENV[app], seeds=[generation ^ it for it in range(iters)], job_roots=np.arange(g.nb_roots), max_actions=maxa,
                              mode=mode, second_player_random=spr, pendulum_actions=PEND_ACTIONS if app == "pendulum" else None)
    assert nb_eval == [iters] * g.nb_roots
    assert scores.shape == (g.nb_roots,)
    assert np.array_equal(scores.view(np.uint64), ref["score"][:, 0].view(np.uint64)), (scores, ref["score"][:, 0])


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/adapter_driver not built")
def test_errors_surface_as_exceptions_like_the_reference():
    """A missing graph file is reported through std::runtime_error -> exit code 1 and a message, not a crash."""
    out = subprocess.run([DRIVER, "pendulum", "/nonexistent.dot", "0", "0", "1", "10"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "cannot open" in out.stderr


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/adapter_driver not built")
@pytest.mark.parametrize("mode", [A.MODE_TRAINING, A.MODE_VALIDATION, A.MODE_TESTING])
def test_mnist_reference_environment_and_main_set_drive_the_gpu_classifier(oracle_mod, tmp_path, mode):
    """The headline app: the instruction set sliced out of [REF] mnist/src/main.cpp:47-88, the reference's MNIST
    environment (dataset read by its own mnist_reader from IDX files generated at build time), a graph in the
    reference's dot format, tpgx::ClassificationAgent::evaluateAllRoots with params.json's 500 actions per evaluation —
    per-root score, per-class F1 and evaluation counts equal the oracle's on the images MNIST::reset / doAction draw."""
    g = tpgx.synthetic_graph("mnist", roots=25, edges=5, lines=12, depth=2, share_prob=0.1, seed=77)
    dot = os.path.join(str(tmp_path), "mnist.dot")
    tpgx.export_dot(g, dot)
    generation, nb_actions = 5, 500
    out = subprocess.run([DRIVER, "mnist", dot, str(generation), str(mode), "1", str(nb_actions)], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.strip().splitlines() if ln.startswith(("0x", "-0x"))]
    assert len(rows) == g.nb_roots
    score = np.array([float.fromhex(r[0]) for r in rows])
    f1 = np.array([[float.fromhex(v) for v in r[2:]] for r in rows])
    assert all(int(r[1]) == nb_actions for r in rows)
    tr_img, tr_lab = tpgx.synthetic_images(3000, seed=101, label_seed=102)
    te_img, te_lab = tpgx.synthetic_images(1000, seed=202, label_seed=203)
    img, lab = np.concatenate([tr_img, te_img]), np.concatenate([tr_lab, te_lab])
    size, offset = (3000, 0) if mode == A.MODE_TRAINING else (1000, 3000)
    idx = tpgx.mnist_episode_indices(generation ^ 0, mode, size, nb_actions) + offset
    g2 = tpgx.import_dot(dot, 5)
    ref = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g2, nb_constants=5).evaluate_classification(
        img, lab, sample_index=idx.astype(np.int32), want=("score", "class_f1"))
    assert np.array_equal(f1.view(np.uint64), ref["class_f1"].view(np.uint64))
    assert np.allclose(score, ref["score"], rtol=0, atol=1e-15)   # the result object averages the 10 F1 values itself


GRID_MAIN = os.path.join(ROOT, "oracle", "_ref", "gridworld_main")
GRID_PARAMS = os.path.join(ROOT, "oracle", "_ref", "gridworld_app", "params.json")


@pytest.mark.skipif(not (os.path.exists(GRID_MAIN) and os.path.exists(GRID_PARAMS)), reason="oracle/_ref/gridworld_main not built (needs /root/reference at build time)")
def test_the_gridworld_apps_own_main_runs_unmodified_on_the_gpu_agent(tmp_path):
    """[REF] gridworld/src/main.cpp compiled UNMODIFIED (oracle/build_ref.sh) with Learn::ParallelLearningAgent standing for
    tpgx::TrainingAgent: la.init(), params.nbGenerations x la.trainOneGeneration(i), la.getSelector()->keepBestPolicy(),
    getBestRoot() — its params.json as is (500 roots, archive 2000 @ 0.01, maxNbEvaluationPerPolicy 10, 10 generations).
    The logger's per-generation lines (statistics + graph fingerprint) must equal the Python-driven trainer's."""
    import shutil
    shutil.copy(GRID_PARAMS, tmp_path / "params.json")
    out = subprocess.run([GRID_MAIN], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.splitlines() if ln.strip() and ln.split()[0].isdigit()]
    p = tpgx.load_params_json(str(tmp_path / "params.json"), nb_actions=4)
    assert len(rows) == p["nb_generations"] == 10
    tr = tpgx.Trainer(tpgx.INSTRUCTION_SETS["gridworld"], tpgx.APP_SOURCES["gridworld"], p, nb_registers=p["nb_registers"], nb_constants=p["nb_constants"], seed=0)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["gridworld"], tpgx.APP_SOURCES["gridworld"], nb_registers=p["nb_registers"], nb_constants=p["nb_constants"])
    be = tpgx.episodes_backend(ev, A.ENV_GRIDWORLD, nb_iterations=p["nb_iterations"], max_actions=p["max_actions"])
    for gen, row in enumerate(rows):
        st = tr.train_generation(be, gen)
        assert int(row[0]) == gen and int(row[1]) == st["nb_teams"] and int(row[2]) == st["nb_programs"]
        assert float(row[5]) == pytest.approx(st["best_score"], abs=1e-3) and int(row[6]) == st["nb_evaluated_roots"] and int(row[7]) == st["nb_skipped_roots"]
        assert int(row[8], 16) == tr.fingerprint(), gen
    # keepBestPolicy + exports ran: the app's output files exist
    assert (tmp_path / "out_best.dot").exists() and (tmp_path / "out_best_stats.md").read_text().startswith("policy root set")
