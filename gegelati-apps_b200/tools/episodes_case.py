#!/usr/bin/env python3
"""The sequential environments at their params.json sizes (BASELINE configs C1, C3, C5): device time of
tpgx_evaluate_episodes, the oracle's threaded evaluateAllRoots on the host cores, and a bit-parity verdict.
Reported in DESIGN.md; not a bench line."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import tpgx  # noqa: E402
from tpgx import abi as A  # noqa: E402
from oracle import oracle  # noqa: E402  (checker + CPU baseline only)

PEND = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
# app, roots, edges per team, iterations, max actions, roots per job, second player random  (Appendix D)
CASES = [("pendulum", 500, 5, 5, 1000, 1, 0), ("gridworld", 500, 10, 1, 100, 1, 0), ("stickgame", 100, 5, 100, 11, 1, 1),
         ("tictactoe", 100, 10, 10, 5, 2, 0)]
threads = os.cpu_count() or 1
for app, R, E, NI, maxa, k, spr in CASES:
    g = tpgx.synthetic_graph(app, roots=R, edges=E, lines=12, depth=3, share_prob=0.1, seed=11)
    nc = tpgx.APP_CONSTANTS[app]
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc)
    ev.set_graph(g)
    o = oracle.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc)
    jobs = np.arange(R) if k == 1 else np.stack([np.arange(R), (np.arange(R) + 1) % R], 1)
    kw = dict(seeds=[1000 * 3 + it for it in range(NI)], job_roots=jobs, max_actions=maxa, roots_per_job=k, second_player_random=spr,
              pendulum_actions=PEND if app == "pendulum" else None)
    ev.evaluate_episodes(tpgx.APP_ENV[app], **kw)
    t0 = time.perf_counter(); dev = ev.evaluate_episodes(tpgx.APP_ENV[app], **kw); t_call = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = o.evaluate_episodes(tpgx.APP_ENV[app], nb_threads=threads, **kw); t_cpu = time.perf_counter() - t0
    same = np.array_equal(dev["episode_score"].view(np.uint64), ref["episode_score"].view(np.uint64)) and np.array_equal(dev["episode_actions"], ref["episode_actions"])
    steps = int(dev["counters"]["evaluations"])
    print("%-10s jobs %4d x %3d episodes, %8d env steps, %10d lines: device %.3f ms (call %.2f ms) | oracle %d threads %.1f ms | "
          "%.2e steps/s vs %.2e | bit-identical: %s" % (app, len(jobs), NI, steps, dev["counters"]["lines_executed"], dev["device_ms"][0],
                                                       t_call * 1e3, threads, t_cpu * 1e3, steps / (dev["device_ms"][0] * 1e-3), steps / t_cpu, same))
