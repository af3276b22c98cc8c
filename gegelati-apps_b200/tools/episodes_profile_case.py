#!/usr/bin/env python3
"""One sequential environment at its params.json size, for ncu captures of the episode kernel.
usage: episodes_profile_case.py [pendulum|gridworld|stickgame|tictactoe] [reps=2]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import numpy as np  # noqa: E402
import tpgx  # noqa: E402

PEND = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
CASES = {"pendulum": (500, 5, 5, 1000, 1, 0), "gridworld": (500, 10, 1, 100, 1, 0), "stickgame": (100, 5, 100, 11, 1, 1),
         "tictactoe": (100, 10, 10, 5, 2, 0)}
app = sys.argv[1] if len(sys.argv) > 1 else "pendulum"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
R, E, NI, maxa, k, spr = CASES[app]
g = tpgx.synthetic_graph(app, roots=R, edges=E, lines=12, depth=3, share_prob=0.1, seed=11)
ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=tpgx.APP_CONSTANTS[app])
ev.set_graph(g)
jobs = np.arange(R) if k == 1 else np.stack([np.arange(R), (np.arange(R) + 1) % R], 1)
kw = dict(seeds=[1000 * 3 + it for it in range(NI)], job_roots=jobs, max_actions=maxa, roots_per_job=k, second_player_random=spr,
          pendulum_actions=PEND if app == "pendulum" else None)
for i in range(reps):
    out = ev.evaluate_episodes(tpgx.APP_ENV[app], **kw)
    print("rep", i, app, "device ms", out["device_ms"][0], "steps", out["counters"]["evaluations"])
