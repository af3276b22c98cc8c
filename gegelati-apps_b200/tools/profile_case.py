#!/usr/bin/env python3
"""Small driver for ncu captures: runs the headline classification evaluation a few times.
usage: profile_case.py [roots] [samples] [reps] [mix] [depth]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import numpy as np  # noqa: E402
import tpgx  # noqa: E402

roots = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mix = sys.argv[4] if len(sys.argv) > 4 else "uniform"
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 1
g = tpgx.synthetic_graph("mnist", roots=roots, edges=5, lines=20, depth=depth, mix=mix, share_prob=0.1 if depth > 1 else 0.0, seed=42)
img, lab = tpgx.synthetic_images(samples)
ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
ev.set_graph(g)
ev.set_observations(img, lab)
for i in range(reps):
    out = ev.evaluate_classification(want=("score", "counters", "device_ms"))
    print("rep", i, "device ms K1/K2/K3", out["device_ms"], "lines/s %.3g" % (out["counters"]["device_lines"] / (out["device_ms"][0] * 1e-3)),
          "checksum", float(out["score"].sum()))
