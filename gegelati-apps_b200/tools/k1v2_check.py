#!/usr/bin/env python3
"""K1 v2 against K1 v1 on the GPU, bid by bid (both are checked against the oracle by tests/; this tool localises a
mismatch to handlers).  usage: k1v2_check.py [roots] [samples] [lines] [seed]"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import numpy as np  # noqa: E402
import tpgx  # noqa: E402

roots = int(sys.argv[1]) if len(sys.argv) > 1 else 200
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lines = int(sys.argv[3]) if len(sys.argv) > 3 else 20
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 7
g = tpgx.synthetic_graph("mnist", roots=roots, edges=5, lines=lines, depth=2, share_prob=0.1, seed=seed, intron_lines=4, line_jitter=6)
img, lab = tpgx.synthetic_images(samples, seed=seed)
res = {}
for v in ("0", "1"):
    os.environ["TPGX_K1_V2"] = v
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    ev.set_graph(g)
    ev.set_observations(img, lab)
    res[v] = ev.evaluate_classification(want=("bid", "score", "action", "counters", "device_ms"))
    res[v + "t"] = ev.evaluate_classification(want=("score", "action"))   # team mode (no bid output)
    ev.close()
b0, b1 = res["0"]["bid"], res["1"]["bid"]
same = (b0.view(np.uint64) == b1.view(np.uint64)) | (np.isnan(b0) & np.isnan(b1))
bad_prog = np.nonzero(~same.all(axis=1))[0]
print("programs", b0.shape[0], "mismatching", len(bad_prog), "| bid-mode scores equal", np.array_equal(res["0"]["score"], res["1"]["score"]),
      "| team-mode scores equal", np.array_equal(res["0t"]["score"], res["1t"]["score"]), "actions equal",
      np.array_equal(res["0t"]["action"], res["1t"]["action"]), "| team vs bid (v2)", np.array_equal(res["1"]["score"], res["1t"]["score"]))
out = tpgx.flatten_programs(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
off = np.concatenate([[0], np.cumsum(out["effective_lines"])])
names = tpgx.k1v2_handler_names()
good_h, bad_h = collections.Counter(), collections.Counter()
for p in range(g.nb_programs):
    enc = tpgx.k1v2_encode_program(out["code"][off[p]:off[p + 1]])
    hs = set(names[h] for h in enc["quads"][:, 0])
    (bad_h if p in set(bad_prog.tolist()) else good_h).update(hs)
if len(bad_prog):
    print("handlers that appear ONLY in mismatching programs:", sorted(h for h in bad_h if h not in good_h))
    print("handler -> share of the programs using it that mismatch:")
    for h in sorted(bad_h, key=lambda k: -bad_h[k] / (bad_h[k] + good_h[k])):
        print("  %-8s %4d / %4d" % (h, bad_h[h], bad_h[h] + good_h[h]))
    p = int(bad_prog[0])
    enc = tpgx.k1v2_encode_program(out["code"][off[p]:off[p + 1]])
    print("first mismatching program", p, [names[h] for h in enc["quads"][:, 0]])
    s = int(np.nonzero(~same[p])[0][0])
    print("  sample", s, "v1", b0[p, s], "v2", b1[p, s], "mismatching samples", int((~same[p]).sum()), "of", same.shape[1])
    sys.exit(1)
print("K1 v2 == K1 v1 on every bid; device ms v1", res["0"]["device_ms"], "v2", res["1"]["device_ms"])
