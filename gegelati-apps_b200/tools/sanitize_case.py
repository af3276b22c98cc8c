#!/usr/bin/env python3
"""Smoke-sized run of every kernel family for compute-sanitizer (memcheck / racecheck): the nine K1 v1 tile-shape variants
and K1 v2, each in team and bid mode, MNIST-only and extended instruction builds, long programs (staging-buffer reloads),
K2 / K3, the archive kernels, and both episode kernels (CTA-per-job and lane-group) on the four sequential environments.
usage: compute-sanitizer --tool memcheck|racecheck python sanitize_case.py [k1|episodes|all]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "python"))
import numpy as np  # noqa: E402
import tpgx  # noqa: E402
from tpgx import abi as A  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
PEND = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]
img, lab = tpgx.synthetic_images(300, seed=3)
n = 0
if what in ("k1", "all"):
    g_mnist = tpgx.synthetic_graph("mnist", roots=10, edges=5, lines=20, depth=2, share_prob=0.1, seed=9, intron_lines=2, line_jitter=4)
    g_long = tpgx.synthetic_graph("mnist", roots=4, edges=3, lines=60, depth=1, seed=10)
    rng = np.random.default_rng(17)
    ops_ext = tpgx.INSTRUCTION_SETS["pendulum"]
    lines, plo, consts = tpgx.random_programs(rng, 30, 10, ops_ext, tpgx.APP_SOURCES["mnist"], 8, 5)
    g_ext = tpgx.TPGraph(lines=lines, program_line_offset=plo, constants=consts, team_edge_offset=(np.arange(7) * 5).astype(np.int32),
                         edge_program=np.arange(30, dtype=np.int32), edge_destination=(-1 - rng.integers(0, 10, 30)).astype(np.int32),
                         root_team=np.arange(6, dtype=np.int32))
    cases = [("v2", "1", "0", tpgx.INSTRUCTION_SETS["mnist"], g_mnist), ("v2 long", "1", "0", tpgx.INSTRUCTION_SETS["mnist"], g_long)]
    cases += [("v1 variant %d" % v, "0", str(v), tpgx.INSTRUCTION_SETS["mnist"], g_mnist) for v in range(9)]
    cases += [("v1 long", "0", "0", tpgx.INSTRUCTION_SETS["mnist"], g_long), ("v1 ext", "0", "0", ops_ext, g_ext), ("v1 ext variant 1", "0", "1", ops_ext, g_ext)]
    for name, v2, variant, ops, g in cases:
        os.environ["TPGX_K1_V2"], os.environ["TPGX_K1_VARIANT"] = v2, variant
        ev = tpgx.Evaluator(ops, tpgx.APP_SOURCES["mnist"], nb_constants=5)
        ev.set_graph(g)
        ev.set_observations(img, lab)
        a = ev.evaluate_classification(want=("score", "bid", "action", "confusion"))     # bid mode + K2 traverse + K3
        b = ev.evaluate_classification(want=("score", "action", "confusion"))            # team mode + K2 walk + K3
        assert np.array_equal(a["score"], b["score"]) and np.array_equal(a["action"], b["action"]), name
        ev.archive_classification(np.arange(g.nb_roots, dtype=np.uint64) + 3, 0.05, 64)
        ev.close()
        n += 1
        print("ok", name, flush=True)
if what in ("episodes", "all"):
    for app, NI, maxa, k, spr in (("pendulum", 3, 40, 1, 0), ("gridworld", 1, 30, 1, 0), ("stickgame", 6, 11, 1, 1), ("tictactoe", 4, 5, 2, 0)):
        g = tpgx.synthetic_graph(app, roots=12, edges=4, lines=10, depth=3, share_prob=0.1, seed=11)
        jobs = np.arange(12) if k == 1 else np.stack([np.arange(12), (np.arange(12) + 1) % 12], 1)
        kw = dict(seeds=[1000 + it for it in range(NI)], job_roots=jobs, max_actions=maxa, roots_per_job=k, second_player_random=spr,
                  pendulum_actions=PEND if app == "pendulum" else None)
        res = []
        for cta in ("", "0"):                                 # CTA-per-job kernel, then the lane-group kernel
            if cta:
                os.environ["TPGX_EP_CTA_WARPS"] = cta
            else:
                os.environ.pop("TPGX_EP_CTA_WARPS", None)
            ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=tpgx.APP_CONSTANTS[app])
            ev.set_graph(g)
            res.append(ev.evaluate_episodes(tpgx.APP_ENV[app], **kw)["episode_score"])
            obs_len = sum(h * w for _, h, w, _ in tpgx.APP_SOURCES[app])
            ev.archive_episodes(tpgx.APP_ENV[app], archive_seeds=np.arange(len(jobs), dtype=np.uint64) + 5, probability=0.05, archive_size=50, obs_len=obs_len, **kw)
            ev.close()
        assert np.array_equal(res[0], res[1]), app
        n += 1
        print("ok episodes", app, flush=True)
print("sanitize_case: %d cases done, %d kernels launched" % (n, tpgx.kernel_launches()))
