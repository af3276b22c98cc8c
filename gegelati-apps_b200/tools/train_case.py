#!/usr/bin/env python3
"""Full training loops — host mutation + GPU evaluation + archive — with the apps' params.json values, end-to-end generations/s.
  gridworld (BASELINE.json config 5; [REF] gridworld/params.json: 500 roots, 1 iteration x <= 100 actions, archive 2000 @ 0.01)
  mnist     ([REF] mnist/params.json: 500 roots, 500 images per generation drawn from 60 000 synthetic ones, archive 500 @ 0.01)
usage: train_case.py [gridworld|mnist] [generations=200] [nb_roots=500]"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import tpgx  # noqa: E402

app = sys.argv[1] if len(sys.argv) > 1 else "gridworld"
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 200
roots = int(sys.argv[3]) if len(sys.argv) > 3 else 500
COMMON = dict(nb_roots=roots, p_edge_deletion=0.7, p_edge_addition=0.7, p_program_mutation=0.2, p_edge_destination_change=0.1,
              p_edge_destination_is_action=0.5, force_program_behavior_change=1, max_program_size=20, p_delete=0.5, p_add=0.5, p_mutate=1.0,
              p_swap=1.0, p_constant_mutation=0.5, min_const_value=-10, max_const_value=10, ratio_deleted_roots=0.5, archiving_probability=0.01)
if app == "gridworld":
    # [REF] gridworld/params.json:11,20,97: no validation, maxNbEvaluationPerPolicy 10, 1 iteration per evaluation
    P = dict(COMMON, nb_actions=4, max_init_outgoing_edges=4, max_outgoing_edges=10, archive_size=2000, do_validation=0,
             max_nb_evaluation_per_policy=10, nb_iterations_per_policy_evaluation=1)
else:
    # [REF] mnist/params.json:11,20: validation on, maxNbEvaluationPerPolicy 2000 (= 4 evaluations of 500 images), per-class results
    P = dict(COMMON, nb_actions=10, max_init_outgoing_edges=3, max_outgoing_edges=5, archive_size=500, do_validation=1,
             max_nb_evaluation_per_policy=2000, nb_iterations_per_policy_evaluation=1, nb_classes=10)
tr = tpgx.Trainer(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], P, nb_constants=5, seed=0)
ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=5)
if app == "gridworld":
    be = tpgx.episodes_backend(ev, tpgx.abi.ENV_GRIDWORLD, nb_iterations=1, max_actions=100)
else:
    img, lab = tpgx.synthetic_images(60000)
    ev.set_observations(img, lab)
    be = tpgx.classification_backend(ev, img, nb_classes=10, actions_per_evaluation=500)
tr.train_generation(be, 0)  # warm-up (context buffers, first-launch costs)
t0 = time.perf_counter()
per_gen = []
for gen in range(1, gens + 1):
    g0 = time.perf_counter()
    s = tr.train_generation(be, gen)
    per_gen.append(time.perf_counter() - g0)
    if gen % max(1, gens // 10) == 0:
        print("gen %4d  best %8.4f  mean %9.4f  validation best %8.4f  teams %4d  programs %5d  archive %4d  retries %5d  jobs %3d  skipped %3d" %
              (gen, s["best_score"], s["mean_score"], s["best_validation_score"], s["nb_teams"], s["nb_programs"], s["archive_records"],
               s["behaviour_retries"], s["nb_evaluated_roots"], s["nb_skipped_roots"]))
dt = time.perf_counter() - t0
print("%s: %d generations in %.2f s = %.1f generations/s (%d roots, fingerprint %016x)" % (app, gens, dt, gens / dt, roots, tr.fingerprint()))
import numpy as np  # noqa: E402
pg = np.array(per_gen) * 1e3
worst = int(pg.argmax())
print("per generation: median %.2f ms, p99 %.2f ms, slowest %.2f ms at generation %d; without the slowest: %.1f generations/s; CUDA_MODULE_LOADING=%s"
      % (np.median(pg), np.percentile(pg, 99), pg[worst], worst + 1, (gens - 1) / ((pg.sum() - pg[worst]) * 1e-3), os.environ.get("CUDA_MODULE_LOADING", "(default: LAZY)")))
