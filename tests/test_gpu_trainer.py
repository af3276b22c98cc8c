"""Host generation loop with the GPU backend (tpgx_train_backend_episodes) vs the same loop with the CPU oracle behind the
backend callbacks: scores and archive records are bit-identical, so the whole evolution — populate, behaviour test against
the archive, selection — is identical generation by generation (north_star: "the resulting elite/mutation sequence")."""
import numpy as np
import pytest

import tpgx
from tpgx import abi as A
from test_trainer_cpu import GRID
from trainer_util import oracle_backend

pytestmark = pytest.mark.gpu
PEND_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]  # [REF] pendulum/src/Learn/main.cpp:33


def same_stats(a, b):
    """Generation statistics equal, NaN (no validation pass) matching NaN."""
    return a.keys() == b.keys() and all(a[k] == b[k] or (a[k] != a[k] and b[k] != b[k]) for k in a)


def run_pair(oracle_mod, app, generations, nb_iterations, max_actions, seed, second_player_random=0, pendulum_actions=None, **over):
    p = dict(GRID, nb_actions=tpgx.APP_ACTIONS[app], **over)
    nc = tpgx.APP_CONSTANTS[app]
    mk = lambda: tpgx.Trainer(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], p, nb_constants=nc, seed=seed)
    t_gpu, t_cpu = mk(), mk()
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc)
    be_gpu = tpgx.episodes_backend(ev, tpgx.APP_ENV[app], nb_iterations, max_actions, second_player_random=second_player_random,
                                   pendulum_actions=pendulum_actions)
    be_cpu = oracle_backend(oracle_mod, app, nb_iterations, max_actions, second_player_random=second_player_random,
                            pendulum_actions=pendulum_actions)
    out = []
    for gen in range(generations):
        a, b = t_gpu.train_generation(be_gpu, gen), t_cpu.train_generation(be_cpu, gen)
        assert same_stats(a, b), (gen, a, b)
        assert t_gpu.fingerprint() == t_cpu.fingerprint(), gen
        ra, rb = t_gpu.root_results(), t_cpu.root_results()
        assert np.array_equal(ra[0].view(np.uint64), rb[0].view(np.uint64)) and np.array_equal(ra[1], rb[1]), gen
        out.append(a)
    return out


def test_gridworld_evolution_identical_and_learning(oracle_mod):
    stats = run_pair(oracle_mod, "gridworld", generations=10, nb_iterations=1, max_actions=100, seed=5, nb_roots=60, archive_size=300)
    assert max(s["best_score"] for s in stats) > 0.0
    assert sum(s["behaviour_retries"] for s in stats) > 0, "the behaviour test against the archive ran on the GPU"


def test_stickgame_evolution_identical(oracle_mod):
    """int-typed data sources, random opponent, 20 iterations per job, no program constants."""
    run_pair(oracle_mod, "stickgame", generations=5, nb_iterations=20, max_actions=11, seed=7, second_player_random=1, nb_roots=40,
             archive_size=200, max_outgoing_edges=5, max_init_outgoing_edges=3, p_constant_mutation=0.0)


def test_pendulum_evolution_identical(oracle_mod):
    """libm-heavy environment: identical only because the GPU and the oracle share one libm sequence."""
    run_pair(oracle_mod, "pendulum", generations=3, nb_iterations=2, max_actions=300, seed=9, pendulum_actions=PEND_ACTIONS, nb_roots=30,
             archive_size=200, max_outgoing_edges=5, max_init_outgoing_edges=3)


def test_mnist_evolution_identical(oracle_mod):
    """The mnist app's loop ([REF] mnist/src/main.cpp:106-107,145) at a reduced size: ClassificationLearningAgent semantics
    (F1 scores), the per-generation image draw, the archive of images and the behaviour test on them."""
    from trainer_util import oracle_classification_backend
    img, lab = tpgx.synthetic_images(600, seed=41)
    p = dict(GRID, nb_actions=10, nb_roots=40, archive_size=100, max_outgoing_edges=5, max_init_outgoing_edges=3)
    mk = lambda: tpgx.Trainer(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], p, nb_constants=5, seed=3)
    t_gpu, t_cpu = mk(), mk()
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_observations(img, lab)
    be_gpu = tpgx.classification_backend(ev, img, nb_classes=10, actions_per_evaluation=120)
    be_cpu = oracle_classification_backend(oracle_mod, img, lab, 120)
    for gen in range(4):
        a, b = t_gpu.train_generation(be_gpu, gen), t_cpu.train_generation(be_cpu, gen)
        assert same_stats(a, b), (gen, a, b)
        assert t_gpu.fingerprint() == t_cpu.fingerprint(), gen
    assert a["archive_records"] > 0 and 0.0 <= a["best_score"] <= 1.0


def test_gridworld_params_json_evaluation_cap_and_skipped_roots(oracle_mod):
    """[REF] gridworld/params.json:20 maxNbEvaluationPerPolicy = 10 over 14 generations, and a cap of 2 where many roots reach
    it: [UP] evaluateJob's skip (capped roots are not in the backend's root list, i.e. never launched) and the merge of
    previous results give the same stored results, statistics and evolution with the GPU as with the oracle."""
    stats = run_pair(oracle_mod, "gridworld", generations=14, nb_iterations=1, max_actions=100, seed=5, nb_roots=100, archive_size=300,
                     max_nb_evaluation_per_policy=10, nb_iterations_per_policy_evaluation=1)
    assert all(s["nb_evaluated_roots"] + s["nb_skipped_roots"] == 100 for s in stats)
    stats = run_pair(oracle_mod, "gridworld", generations=12, nb_iterations=1, max_actions=100, seed=6, nb_roots=100, archive_size=300,
                     max_nb_evaluation_per_policy=2, nb_iterations_per_policy_evaluation=1)
    assert sum(s["nb_skipped_roots"] for s in stats) > 0 and all(s["nb_evaluated_roots"] + s["nb_skipped_roots"] == 100 for s in stats)


def test_pendulum_validation_pass_identical(oracle_mod):
    """[REF] pendulum/params.json:14 doValidation: the VALIDATION evaluateAllRoots after the decimation (its own initial
    states: the mode enters the environment's seed) reports the same scores from the GPU as from the oracle."""
    stats = run_pair(oracle_mod, "pendulum", generations=3, nb_iterations=2, max_actions=200, seed=9, pendulum_actions=PEND_ACTIONS, nb_roots=30,
                     archive_size=200, max_outgoing_edges=5, max_init_outgoing_edges=3, do_validation=1, max_nb_evaluation_per_policy=10,
                     nb_iterations_per_policy_evaluation=2)
    assert all(np.isfinite(s["best_validation_score"]) for s in stats)


def test_mnist_loop_with_validation_and_per_class_merge(oracle_mod):
    """[REF] mnist/params.json:11,20 (doValidation true, maxNbEvaluationPerPolicy = a few evaluations' worth of images): the
    ClassificationLearningAgent loop — per-class F1 and per-class evaluation counts combined across generations, skip at the
    cap, validation pass — GPU against oracle."""
    from trainer_util import oracle_classification_backend
    img, lab = tpgx.synthetic_images(600, seed=41)
    p = dict(GRID, nb_actions=10, nb_roots=40, archive_size=100, max_outgoing_edges=5, max_init_outgoing_edges=3, nb_classes=10, do_validation=1,
             max_nb_evaluation_per_policy=240, nb_iterations_per_policy_evaluation=1)
    mk = lambda: tpgx.Trainer(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], p, nb_constants=5, seed=3)
    t_gpu, t_cpu = mk(), mk()
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_observations(img, lab)
    be_gpu = tpgx.classification_backend(ev, img, nb_classes=10, actions_per_evaluation=120)
    be_cpu = oracle_classification_backend(oracle_mod, img, lab, 120)
    skipped = 0
    for gen in range(6):
        a, b = t_gpu.train_generation(be_gpu, gen), t_cpu.train_generation(be_cpu, gen)
        assert same_stats(a, b), (gen, a, b)
        assert t_gpu.fingerprint() == t_cpu.fingerprint(), gen
        ra, rb = t_gpu.root_results(), t_cpu.root_results()
        assert np.array_equal(ra[0].view(np.uint64), rb[0].view(np.uint64)) and np.array_equal(ra[1], rb[1]), gen
        assert ra[1].max() <= 240 and (ra[1] % 120 == 0).all()      # whole evaluations of 120 images, none past the cap
        skipped += a["nb_skipped_roots"]
        assert 0.0 <= a["best_validation_score"] <= 1.0
    assert skipped > 0
