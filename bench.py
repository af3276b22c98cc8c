#!/usr/bin/env python3
"""bench.py — headline benchmark of the TPG policy-evaluation hot path (BASELINE.json):
evaluateAllRoots of a MNIST-shaped TPG over 60 000 synthetic 28x28 uint8 samples.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement on the host cores

Workload: N = 1 -> BASELINE config 2 (2000 roots); N > 1 -> BASELINE config 4 (8192 roots IN TOTAL, sharded over the N
GPUs: strong scaling).  One "step" = one evaluateAllRoots of the graph over all samples: per rank K1 bid interpreter
(evaluateTeam fused in) -> K2 decision walk + confusion tables -> K3 per-root F1 records written into the library's
all-gather buffer -> (N > 1) the library's own NCCL all-gather on the evaluation stream.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gegelati-apps_b200", "python"))
sys.path.insert(0, ROOT)

METRIC = "root x sample evaluations per second (evaluateAllRoots, MNIST-shape TPG)"
UNIT = "evals/s"
PEND_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]   # [REF] pendulum/src/Learn/main.cpp:33


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tpgx", choices=["tpgx", "reference"])
    ap.add_argument("--roots", type=int, default=0, help="TOTAL roots of the graph (default: 2000 at N = 1, 8192 at N > 1)")
    ap.add_argument("--samples", type=int, default=60000)
    ap.add_argument("--edges", type=int, default=5)
    ap.add_argument("--lines", type=int, default=20)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--mix", default="uniform", choices=["uniform", "scalar", "arith"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C1 / C3 / C5 figures (N = 1 only)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")) if a.impl == "tpgx" else max(1, a.gpus)
    if a.roots <= 0:
        a.roots = 2000 if max(world, a.gpus) <= 1 else 8192
    return a


def workload_config(a, n_gpus):
    cfg = "C2" if a.roots == 2000 else ("C4" if a.roots == 8192 else "custom")
    return {
        "workload": "%s mnist-shape evaluateAllRoots: G(R=%d roots in total, E=%d, L=%d effective lines, D=%d, mix=%s10, seed=42) x %d synthetic 28x28 uint8 samples, 10 actions"
                    % (cfg, a.roots, a.edges, a.lines, a.depth, a.mix, a.samples),
        "total_roots": a.roots, "roots_per_gpu": -(-a.roots // n_gpus), "samples": a.samples, "edges_per_team": a.edges,
        "lines_per_program": a.lines, "depth": a.depth, "mix": a.mix,
        "sharding": "roots split across GPUs (contiguous slices of the job order), 1 NCCL all-gather of score records enqueued by the library",
        "pipeline": "K1 threaded bid interpreter with fused evaluateTeam (team mode) -> K2 decision walk + confusion tables -> K3 F1 records",
        "l2": "L2 flushed (256 MiB write) before every timed step; decision matrix %.2f GB + Sobel planes 0.65 GB exceed L2 anyway" % (-(-a.roots // n_gpus) * a.samples * 4 / 1e9),
        "tie_break": "last-max (bid >= best)", "libm": "shared (tpgx_math.h)",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_graph(tpgx, a):
    return tpgx.synthetic_graph("mnist", roots=a.roots, edges=a.edges, lines=a.lines, depth=a.depth, mix=a.mix,
                                share_prob=0.1 if a.depth > 1 else 0.0, seed=42)


def cpu_baseline(oracle_mod, tpgx, g, img, lab, budget_s, threads):
    """Times the oracle's threaded evaluateAllRoots (restatement of ParallelLearningAgent) on a bounded
    sample of the same workload: all roots x the first S' samples.  Returns the record, the oracle's result
    (the parity check compares the CUDA path against it) and S'."""
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
    probe = 32
    t = time.perf_counter()
    o.evaluate_classification(img[:probe], lab[:probe], want=("score",), nb_threads=threads)
    dt = max(time.perf_counter() - t, 1e-4)
    n = int(min(len(img), max(probe, probe * budget_s / dt)))
    t = time.perf_counter()
    r = o.evaluate_classification(img[:n], lab[:n], want=("score", "class_f1", "confusion"), nb_threads=threads)
    dt = time.perf_counter() - t
    c = r["counters"]
    rec = {"value": c["evaluations"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "all %d roots x first %d of %d samples, %.1f s, oracle threaded job queue" % (g.nb_roots, n, len(img), dt),
           "instr_per_s": c["lines_executed"] / dt}
    return rec, r, n


def run_reference(a):
    """--impl reference: the CPU restatement of ParallelLearningAgent::evaluateAllRoots on the host cores
    (the GEGELATI library itself is not vendored in the reference tree, so oracle/ is the reference arm), on the same
    graph (same total roots) as the CUDA arm at this N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tpgx
    from oracle import oracle
    threads = os.cpu_count() or 1
    g = make_graph(tpgx, a)
    img, lab = tpgx.synthetic_images(a.samples)
    o = oracle.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
    # size one step (a bounded sample) so that warmup + steps stay within ~2 minutes
    probe = 16
    t = time.perf_counter()
    o.evaluate_classification(img[:probe], lab[:probe], want=("score",), nb_threads=threads)
    dt = max(time.perf_counter() - t, 1e-4)
    per_step_budget = 100.0 / max(1, a.steps + a.warmup)
    n = int(min(a.samples, max(probe, probe * per_step_budget / dt)))
    for _ in range(a.warmup):
        o.evaluate_classification(img[:n], lab[:n], want=("score",), nb_threads=threads)
    t = time.perf_counter()
    lines = 0
    for _ in range(a.steps):
        r = o.evaluate_classification(img[:n], lab[:n], want=("score",), nb_threads=threads)
        lines += r["counters"]["lines_executed"]
    dt = time.perf_counter() - t
    evals = a.roots * n * a.steps
    value = evals / dt
    sample = "each step = all %d roots x first %d of %d samples" % (a.roots, n, a.samples)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong" if a.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a, max(1, a.gpus)), "instr_per_s": lines / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def secondary_configs(tpgx, oracle, A):
    """BASELINE configs 1, 3 and 5 at the apps' params.json sizes (rank 0, N = 1): device time of one evaluateAllRoots, the
    oracle's threaded run on the host cores, bit parity; and the gridworld generation loop (validation off, skip-and-merge
    of previous evaluations on, as [REF] gridworld/params.json:11,20 set them), end to end."""
    threads = os.cpu_count() or 1
    out = {}
    cases = [("c1_pendulum", "pendulum", 500, 5, 5, 1000, 1, 0), ("c3_stickgame", "stickgame", 100, 5, 100, 11, 1, 1),
             ("c3_tictactoe", "tictactoe", 100, 10, 10, 5, 2, 0), ("c5_gridworld_evaluation", "gridworld", 500, 10, 1, 100, 1, 0)]
    for key, app, R, E, NI, maxa, k, spr in cases:
        g = tpgx.synthetic_graph(app, roots=R, edges=E, lines=12, depth=3, share_prob=0.1, seed=11)
        nc = tpgx.APP_CONSTANTS[app]
        ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc)
        ev.set_graph(g)
        o = oracle.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc)
        jobs = np.arange(R) if k == 1 else np.stack([np.arange(R), (np.arange(R) + 1) % R], 1)
        kw = dict(seeds=[1000 * 3 + it for it in range(NI)], job_roots=jobs, max_actions=maxa, roots_per_job=k, second_player_random=spr,
                  pendulum_actions=PEND_ACTIONS if app == "pendulum" else None)
        ev.evaluate_episodes(tpgx.APP_ENV[app], **kw)
        t0 = time.perf_counter()
        dev = ev.evaluate_episodes(tpgx.APP_ENV[app], **kw)
        t_call = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = o.evaluate_episodes(tpgx.APP_ENV[app], nb_threads=threads, **kw)
        t_cpu = time.perf_counter() - t0
        same = bool(np.array_equal(dev["episode_score"].view(np.uint64), ref["episode_score"].view(np.uint64)) and
                    np.array_equal(dev["episode_actions"], ref["episode_actions"]))
        steps = int(dev["counters"]["evaluations"])
        out[key] = {"jobs": int(len(jobs)), "iterations": NI, "max_actions": maxa, "env_steps": steps, "device_ms": float(dev["device_ms"][0]),
                    "call_ms": t_call * 1e3, "oracle_ms": t_cpu * 1e3, "oracle_cores": threads, "steps_per_s": steps / (float(dev["device_ms"][0]) * 1e-3),
                    "bit_exact": same}
        ev.close()
    # C5: the gridworld training loop
    P = dict(nb_actions=4, nb_roots=500, max_init_outgoing_edges=4, max_outgoing_edges=10, p_edge_deletion=0.7, p_edge_addition=0.7,
             p_program_mutation=0.2, p_edge_destination_change=0.1, p_edge_destination_is_action=0.5, force_program_behavior_change=1,
             max_program_size=20, p_delete=0.5, p_add=0.5, p_mutate=1.0, p_swap=1.0, p_constant_mutation=0.5, min_const_value=-10,
             max_const_value=10, ratio_deleted_roots=0.5, archive_size=2000, archiving_probability=0.01, do_validation=0,
             max_nb_evaluation_per_policy=10, nb_iterations_per_policy_evaluation=1)
    tr = tpgx.Trainer(tpgx.INSTRUCTION_SETS["gridworld"], tpgx.APP_SOURCES["gridworld"], P, nb_constants=5, seed=0)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["gridworld"], tpgx.APP_SOURCES["gridworld"], nb_constants=5)
    be = tpgx.episodes_backend(ev, A.ENV_GRIDWORLD, nb_iterations=1, max_actions=100)
    tr.train_generation(be, 0)
    gens, skipped = 200, 0
    t0 = time.perf_counter()
    for gen in range(1, gens + 1):
        st = tr.train_generation(be, gen)
        skipped += st["nb_skipped_roots"]
    dt = time.perf_counter() - t0
    out["c5_gridworld_loop"] = {"generations": gens, "roots": 500, "generations_per_s": gens / dt, "ms_per_generation": dt / gens * 1e3,
                                "best_score": st["best_score"], "roots_skipped_at_the_evaluation_cap": int(skipped), "n_gpus": 1,
                                "note": "host mutation + GPU evaluation + archive, end to end; latency-bound (500 episodes of <= 100 dependent steps): does not shard over GPUs",
                                "fingerprint": "%016x" % tr.fingerprint()}
    return out


def main():
    a = parse()
    # keep stdout clean for the single JSON line: libraries (NCCL prints its version banner on fd 1) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    import tpgx
    from tpgx import abi as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libtpgx has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TPGX_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // world * 2))))
        dist.init_process_group("nccl", device_id=dev)   # plumbing only: id broadcast, barriers, max of the timings

    RT, S, C = a.roots, a.samples, 10
    g = make_graph(tpgx, a)
    img, lab = tpgx.synthetic_images(S)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=local)
    ev.set_graph(g)
    if world > 1:   # the library's own communicator: rank 0's unique id travels through torch.distributed's broadcast
        uid = torch.from_numpy(tpgx.Evaluator.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        ev.comm_init(uid.cpu().numpy(), rank, world)
    _, _, rb, re = ev.comm_shard(RT)
    R = re - rb

    # ---- device-resident arm: observations already in HBM when the timed region starts
    d_img = torch.from_numpy(img).to(dev)
    d_lab = torch.from_numpy(lab).to(dev)
    ev.set_observations_device(d_img.data_ptr(), d_lab.data_ptr(), S)
    gathered = torch.zeros(RT, C + 1, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():   # K1 -> K2 -> K3 -> the library's NCCL all-gather, all enqueued by one C-ABI call; no torch collective here
        ev.evaluate_classification_allgather_device(gathered.data_ptr(), stream.cuda_stream, RT, S, C)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    launches0 = tpgx.kernel_launches()
    wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xff)  # L2 flush, outside the timed events
        starts[i].record()
        step()
        ends[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = tpgx.kernel_launches() - launches0
    sampler.stop_flag = True
    sampler.join()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    evals_per_step = RT * S
    value = evals_per_step * a.steps / (ms * 1e-3)
    all_scores = gathered[:, 0].cpu().numpy()

    # ---- detailed (host-call) pass on this rank's shard: counters, per-kernel device times for the roofline
    det = [ev.evaluate_classification(root_begin=rb, root_end=re, want=("score", "counters", "device_ms")) for _ in range(3)]
    k1_ms = float(np.mean([d["device_ms"][0] for d in det[1:]]))
    k2_ms = float(np.mean([d["device_ms"][1] for d in det[1:]]))
    k3_ms = float(np.mean([d["device_ms"][2] for d in det[1:]]))
    cnt = det[-1]["counters"]
    score_check = float(all_scores.sum())
    assert np.array_equal(all_scores[rb:re], det[-1]["score"]), "device-resident and host-call paths disagree"
    # 1 vs N GPUs: this rank re-evaluates slices of the OTHER ranks' shards on its own GPU; the gathered records must carry the same bits
    multi_gpu_equal, rechecked = True, 0
    if world > 1:
        slot = -(-RT // world)
        for other in range(world):
            if other == rank:
                continue
            ob, oe = other * slot, min(RT, other * slot + 16)
            if oe > ob:
                mine = ev.evaluate_classification(root_begin=ob, root_end=oe, want=("score",))["score"]
                multi_gpu_equal = multi_gpu_equal and bool(np.array_equal(mine.view(np.uint64), all_scores[ob:oe].view(np.uint64)))
                rechecked += oe - ob
        flag = torch.tensor([1 if multi_gpu_equal else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        multi_gpu_equal = bool(flag.item())

    # algorithmic flops of one K1 launch: sum over interpreted lines of the SURVEY 8d per-op flop table x samples
    ops = np.array(tpgx.INSTRUCTION_SETS["mnist"])
    flop_tab = np.array([A.OP_FLOPS[o] for o in ops])
    shard_edges = slice(g.team_edge_offset[rb], g.team_edge_offset[re]) if a.depth == 1 else slice(None)
    progs = np.unique(g.edge_program[shard_edges])
    line_prog = np.repeat(np.arange(g.nb_programs), np.diff(g.program_line_offset))
    shard_ins = g.lines["instruction"][np.isin(line_prog, progs)]
    flops_per_sample = float(flop_tab[shard_ins].sum())
    sobel = np.isin(ops[shard_ins], [A.OP_SOBEL_MAGN, A.OP_SOBEL_DIR])
    flops_per_sample_no_sobel = float(flop_tab[shard_ins][~sobel].sum())
    k1_flops = flops_per_sample * S
    dadd, dfma = ev.measure_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # K1 in team mode (fused evaluateTeam): writes one int32 decision per (team, sample), reads each sample's tile and Sobel planes once
    nb_teams_shard = R if a.depth == 1 else g.nb_teams
    k1_bytes = nb_teams_shard * S * 4.0 + S * (784.0 + 2 * 676 * 8.0)
    k1_kernel = "tpgx_bid2_kernel<true> (K1 v2, threaded bid interpreter, team mode)"
    traffic, ncu = None, None
    try:   # ncu figures of the SAME kernel on the C2 workload (profiles/): only quoted when they belong to the kernel that ran here
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if tj.get("kernel", "").startswith("tpgx_bid2_kernel") and os.environ.get("TPGX_K1_V2", "1") != "0":
            ncu = {k: tj[k] for k in ("fp64_pipe_pct", "issue_active_pct", "warp_instructions", "executed_fp64_warp_instructions",
                                      "executed_fp64_flop_frac_of_peak", "source") if k in tj}
            if RT == 2000 and world == 1 and S == 60000:
                traffic = tj.get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "kernel": k1_kernel, "bound": "fp64", "unit": "TFLOP/s",
        "note": "FP64-issue roofline (north_star): algorithmic flops / CUDA-event time / measured non-FMA FP64 issue rate; the nested hbm object is the same kernel against the HBM roofline",
        "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": dadd / 1e12,
        "frac": (k1_flops / (k1_ms * 1e-3)) / dadd, "traffic": traffic,
        "frac_without_sobel_flops": (flops_per_sample_no_sobel * S / (k1_ms * 1e-3)) / dadd,
        "sobel_note": "sobelMagn / sobelDir lines (20 / 18 algorithmic flops each) are served by K1 as one 32-byte load from planes K0b computes once per upload, outside the timed region: `frac` counts them (SURVEY 8d convention), `frac_without_sobel_flops` does not; `ncu` is what the kernel really executed",
        "ncu": ncu,
        "peak_source": "measured live: non-FMA DADD/DMUL issue rate (tpgx_measure_fp64_peak); DFMA rate %.2f T instr/s" % (dfma / 1e12),
        "flops_convention": "SURVEY.md 8d: transcendental/div/sqrt = 1 flop, sobelMagn 20, sobelDir 18, multByConst 2",
        "k1_ms": k1_ms, "k2_ms": k2_ms, "k3_ms": k3_ms,
        "hbm": {"achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": k1_bytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "k2_achieved": (nb_teams_shard * S * 4.0) / (k2_ms * 1e-3) / 1e9,
                "bytes_convention": "decision matrix write (4 B per team x sample) + uint8 tile + Sobel planes read once per sample"},
        "device_lines_per_s": cnt["device_lines"] / (k1_ms * 1e-3),
    }

    # ---- end-to-end arm through the C ABI with HOST buffers.  Per step: the graph of the generation goes host -> device
    # (tpgx_set_graph: flatten, plan, encode), evaluateAllRoots runs, all roots' records come back to the host.  The
    # observations are uploaded ONCE per context, as the reference loads its dataset once ([REF] mnist/src/mnist.cpp:6: static
    # members); e2e_with_observation_upload re-sends them from page-locked memory every step as well.
    pin_img = torch.from_numpy(img).pin_memory()
    pin_lab = torch.from_numpy(lab).pin_memory()
    pin_img_np, pin_lab_np = pin_img.numpy(), pin_lab.numpy()
    pin_rec = torch.zeros(RT, C + 1, dtype=torch.float64).pin_memory()
    pin_rec_np = pin_rec.numpy()
    ev2 = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=local)
    if world > 1:
        uid = torch.from_numpy(tpgx.Evaluator.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        ev2.comm_init(uid.cpu().numpy(), rank, world)
    graph_bytes = sum(x.nbytes for x in (g.lines, g.program_line_offset, g.constants, g.team_edge_offset, g.edge_program, g.edge_destination, g.root_team))
    ev2.set_observations_pinned(pin_img_np, pin_lab_np)

    def e2e_step(upload_observations, two_calls=False):
        if upload_observations:
            ev2.set_observations_pinned(pin_img_np, pin_lab_np)   # page-locked buffers: copies are enqueued, the flattening below overlaps them
        if two_calls:
            ev2.set_graph_sharded(g)   # every rank holds the whole graph; it flattens / uploads what its root shard reaches
            return ev2.evaluate_classification_allgather(RT, S, nb_classes=C)
        # one call per generation: the host side of the graph is pipelined behind the kernels, slice by slice; the records of
        # all roots land in a page-locked host buffer
        return ev2.evaluate_graph_allgather(g, RT, S, nb_classes=C, records_out=pin_rec_np)

    def time_e2e(upload_observations, two_calls=False):
        n = max(2, min(a.steps, 5))
        e2e_step(upload_observations, two_calls)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            out = e2e_step(upload_observations, two_calls)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert np.array_equal(out["score"].view(np.uint64), all_scores.view(np.uint64)), "end-to-end and device-resident paths disagree"
        return evals_per_step * n / float(dt.item()), float(dt.item()) / n * 1e3

    e_val, e_ms = time_e2e(False)
    f_val, f_ms = time_e2e(True)
    t_val, t_ms = time_e2e(False, two_calls=True)
    e2e = {"value": e_val, "unit": UNIT, "h2d_bytes_per_step": int(graph_bytes), "d2h_bytes_per_step": int(RT * (C + 1) * 8), "ms_per_step": e_ms,
           "note": "per step: tpgx_evaluate_graph_classification_allgather = the generation's graph from host memory (flatten + plan + upload, "
                   "pipelined behind the kernels in slices of the root shard) + evaluateAllRoots + all roots' records to host memory; "
                   "observations uploaded once per context like the reference's static dataset",
           "two_call_form": {"value": t_val, "ms_per_step": t_ms, "note": "tpgx_set_graph_sharded then tpgx_evaluate_classification_allgather (no pipelining)"},
           "with_observation_upload": {"value": f_val, "ms_per_step": f_ms, "h2d_bytes_per_step": int(img.nbytes + lab.nbytes + graph_bytes),
                                       "note": "same, plus tpgx_set_observations_pinned of all 60 000 images every step"}}

    if rank == 0:
        cpu, parity = None, None
        if not a.no_cpu_baseline:
            from oracle import oracle
            threads = os.cpu_count() or 1
            cpu, ref, n_ref = cpu_baseline(oracle, tpgx, g, img, lab, a.cpu_seconds, threads)
            # parity verdict: the CUDA path on exactly what the oracle just evaluated (all roots of the graph x the first n samples)
            devp = ev.evaluate_classification(nb_samples=n_ref, want=("score", "class_f1", "confusion"))
            exact = bool(np.array_equal(devp["score"].view(np.uint64), ref["score"].view(np.uint64)) and
                         np.array_equal(devp["class_f1"].view(np.uint64), ref["class_f1"].view(np.uint64)) and
                         np.array_equal(devp["confusion"], ref["confusion"]))
            parity = {"checked": "scores, per-class F1 and confusion tables of the CUDA path vs the oracle (bitwise)", "roots": int(RT), "samples": int(n_ref),
                      "bit_exact": exact, "multi_gpu_records_equal_single_gpu": multi_gpu_equal if world > 1 else None,
                      "multi_gpu_roots_rechecked_per_rank": int(rechecked) if world > 1 else None}
        secondary = None
        if world == 1 and not a.no_secondary and not a.no_cpu_baseline:
            from oracle import oracle
            secondary = secondary_configs(tpgx, oracle, A)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(a, world),
            "instr_per_s": float(RT) * a.edges * a.lines * S * a.steps / (ms * 1e-3) if a.depth == 1 else None,
            "reference_equivalent_work_of_this_rank": {"evaluations": cnt["evaluations"], "lines_executed": cnt["lines_executed"],
                                                       "program_executions": cnt["program_executions"], "device_lines": cnt["device_lines"]},
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_note": "kernels launched by libtpgx on this rank inside the timed region, counted at the launch sites (K1 + K2 + K3 per step; the all-gather is NCCL's)",
            "secondary": secondary, "clocks": sampler.result(), "wall_s_timed_region": wall, "score_checksum": score_check,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
