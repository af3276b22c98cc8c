#!/usr/bin/env python3
"""Times the pieces of the end-to-end (host buffers in, host scores out) evaluateAllRoots call."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import tpgx  # noqa: E402

roots = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
g = tpgx.synthetic_graph("mnist", roots=roots, edges=5, lines=20, depth=1, mix="uniform", seed=42)
img, lab = tpgx.synthetic_images(60000)
pin_img = torch.from_numpy(img).pin_memory().numpy()
pin_lab = torch.from_numpy(lab).pin_memory().numpy()
ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
for rep in range(4):
    t0 = time.perf_counter(); ev.set_observations_pinned(pin_img, pin_lab)
    t1 = time.perf_counter(); ev.set_graph(g)
    t2 = time.perf_counter(); out = ev.evaluate_classification(want=("score", "class_f1", "device_ms"))
    t3 = time.perf_counter()
    print("rep %d: set_observations %.2f ms, set_graph %.2f ms, evaluate %.2f ms (device K1/K2/K3 %s), total %.2f ms" %
          (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, np.round(out["device_ms"], 2), (t3 - t0) * 1e3))
