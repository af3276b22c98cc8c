#!/usr/bin/env python3
"""bench.py — headline benchmark of the TPG policy-evaluation hot path (BASELINE.json):
evaluateAllRoots of a MNIST-shaped TPG over 60 000 synthetic 28x28 uint8 samples.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement on the host cores

One "step" = one evaluateAllRoots of the whole graph shard over all samples:
K1 bid interpreter (evaluateTeam fused in: per team x sample the winning edge) -> K2 walk (+ confusion tables) -> K3 per-root F1 records
(-> NCCL allgather of the records when N > 1).  Prints ONE JSON line on rank 0.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tpgx", choices=["tpgx", "reference"])
    ap.add_argument("--roots", type=int, default=2000, help="roots per GPU (weak scaling)")
    ap.add_argument("--samples", type=int, default=60000)
    ap.add_argument("--edges", type=int, default=5)
    ap.add_argument("--lines", type=int, default=20)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--mix", default="uniform", choices=["uniform", "scalar", "arith"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": "C2 mnist-shape evaluateAllRoots: G(R=%d per GPU, E=%d, L=%d effective lines, D=%d, mix=%s10, seed=42) x %d synthetic 28x28 uint8 samples, 10 actions"
                    % (a.roots, a.edges, a.lines, a.depth, a.mix, a.samples),
        "roots_per_gpu": a.roots, "total_roots": a.roots * n_gpus, "samples": a.samples, "edges_per_team": a.edges,
        "lines_per_program": a.lines, "depth": a.depth, "mix": a.mix, "sharding": "roots split across GPUs, 1 NCCL allgather of score records", "pipeline": "K1 bid interpreter with fused evaluateTeam (team mode) -> K2 decision walk + confusion tables -> K3 F1 records",
        "l2": "L2 flushed (256 MiB write) before every timed step; decision matrix %.2f GB + Sobel planes 0.65 GB exceed L2 anyway" % (a.roots * a.samples * 4 / 1e9),
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


def cpu_baseline(oracle_mod, tpgx, g, img, lab, budget_s, threads):
    """Times the oracle's threaded evaluateAllRoots (restatement of ParallelLearningAgent) on a bounded
    sample of the same workload: all roots x the first S' samples."""
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
    probe = 32
    t = time.perf_counter()
    o.evaluate_classification(img[:probe], lab[:probe], want=("score",), nb_threads=threads)
    dt = max(time.perf_counter() - t, 1e-4)
    n = int(min(len(img), max(probe, probe * budget_s / dt)))
    t = time.perf_counter()
    r = o.evaluate_classification(img[:n], lab[:n], want=("score",), nb_threads=threads)
    dt = time.perf_counter() - t
    c = r["counters"]
    return {"value": c["evaluations"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "all %d roots x first %d of %d samples, %.1f s, oracle threaded job queue" % (g.nb_roots, n, len(img), dt),
            "instr_per_s": c["lines_executed"] / dt}, dt, c


def run_reference(a):
    """--impl reference: the CPU restatement of ParallelLearningAgent::evaluateAllRoots on the host cores
    (the GEGELATI library itself is not vendored in the reference tree, so oracle/ is the reference arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tpgx
    from oracle import oracle
    threads = os.cpu_count() or 1
    g = tpgx.synthetic_graph("mnist", roots=a.roots, edges=a.edges, lines=a.lines, depth=a.depth, mix=a.mix,
                             share_prob=0.1 if a.depth > 1 else 0.0, seed=42)
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
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a, 1), "instr_per_s": lines / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


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
        dist.init_process_group("nccl", device_id=dev)

    R, S, C = a.roots, a.samples, 10
    # the global graph has world*R roots; this rank evaluates roots [rank*R, (rank+1)*R)
    g = tpgx.synthetic_graph("mnist", roots=R * world, edges=a.edges, lines=a.lines, depth=a.depth, mix=a.mix,
                             share_prob=0.1 if a.depth > 1 else 0.0, seed=42)
    img, lab = tpgx.synthetic_images(S)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=local)
    ev.set_graph(g)
    rb, re = tpgx.shard_range(rank, world, R * world)   # contiguous slice of the job-index order: R roots per rank

    # ---- device-resident arm: observations already in HBM when the timed region starts
    d_img = torch.from_numpy(img).to(dev)
    d_lab = torch.from_numpy(lab).to(dev)
    ev.set_observations_device(d_img.data_ptr(), d_lab.data_ptr(), S)
    records = torch.zeros(R, C + 1, dtype=torch.float64, device=dev)
    gathered = torch.zeros(world * R, C + 1, dtype=torch.float64, device=dev) if world > 1 else records
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        ev.evaluate_classification_device(records.data_ptr(), stream.cuda_stream, S, C, rb, re)
        if world > 1:
            dist.all_gather_into_tensor(gathered, records)   # equal shards: the fixed-size record gather of tpgx.gather_records

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
    wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xff)  # L2 flush, outside the timed events
        starts[i].record()
        step()
        ends[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    evals_per_step = world * R * S
    value = evals_per_step * a.steps / (ms * 1e-3)

    # ---- detailed (host-call) pass: counters, per-kernel device times for the roofline
    det = [ev.evaluate_classification(root_begin=rb, root_end=re, want=("score", "counters", "device_ms")) for _ in range(3)]
    k1_ms = float(np.mean([d["device_ms"][0] for d in det[1:]]))
    k2_ms = float(np.mean([d["device_ms"][1] for d in det[1:]]))
    k3_ms = float(np.mean([d["device_ms"][2] for d in det[1:]]))
    cnt = det[-1]["counters"]
    score_check = float(det[-1]["score"].sum())
    rec_host = records[:, 0].cpu().numpy()
    assert np.array_equal(rec_host, det[-1]["score"]), "device-resident and host-call paths disagree"

    # algorithmic flops of one K1 launch: sum over interpreted lines of the SURVEY 8d per-op flop table x samples
    ops = np.array(tpgx.INSTRUCTION_SETS["mnist"])
    flop_tab = np.array([A.OP_FLOPS[o] for o in ops])
    shard_edges = slice(g.team_edge_offset[rb], g.team_edge_offset[re]) if a.depth == 1 else slice(None)
    progs = np.unique(g.edge_program[shard_edges])
    line_prog = np.repeat(np.arange(g.nb_programs), np.diff(g.program_line_offset))
    flops_per_sample = float(flop_tab[g.lines["instruction"][np.isin(line_prog, progs)]].sum())
    k1_flops = flops_per_sample * S
    dadd, dfma = ev.measure_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # K1 in team mode (fused evaluateTeam): writes one int32 decision per (team, sample), reads each sample's tile and Sobel planes once
    nb_teams_shard = (re - rb) if a.depth == 1 else g.nb_teams
    k1_bytes = nb_teams_shard * S * 4.0 + S * (784.0 + 2 * 676 * 8.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "kernel": "tpgx_bid_kernel (K1 bid interpreter)", "bound": "fp64", "unit": "TFLOP/s",
        "note": "FP64-issue roofline (north_star): algorithmic flops / CUDA-event time / measured non-FMA FP64 issue rate; the nested hbm object is the same kernel against the HBM roofline",
        "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": dadd / 1e12,
        "frac": (k1_flops / (k1_ms * 1e-3)) / dadd, "traffic": traffic,
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

    # ---- end-to-end arm: host buffers in, host scores out, through the C ABI, every step
    pin_img = torch.from_numpy(img).pin_memory()
    pin_lab = torch.from_numpy(lab).pin_memory()
    pin_img_np, pin_lab_np = pin_img.numpy(), pin_lab.numpy()
    graph_bytes = g.lines.nbytes + g.program_line_offset.nbytes + g.constants.nbytes + g.team_edge_offset.nbytes + g.edge_program.nbytes + g.edge_destination.nbytes + g.root_team.nbytes
    ev2 = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=local)
    g_shard = g.subgraph(rb, re) if world > 1 else g   # a rank only ships the part of the graph its roots can reach
    graph_bytes = g_shard.lines.nbytes + g_shard.program_line_offset.nbytes + g_shard.constants.nbytes + g_shard.team_edge_offset.nbytes + g_shard.edge_program.nbytes + g_shard.edge_destination.nbytes + g_shard.root_team.nbytes

    def e2e_step():
        ev2.set_observations_pinned(pin_img_np, pin_lab_np)   # page-locked buffers: copies are enqueued, the flattening below overlaps them
        ev2.set_graph(g_shard)
        return ev2.evaluate_classification(want=("score", "class_f1"))

    e2e_steps = max(2, min(a.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    assert np.array_equal(out["score"], det[-1]["score"])
    e2e = {"value": evals_per_step * e2e_steps / float(e2e_dt.item()), "unit": UNIT,
           "h2d_bytes_per_step": int(img.nbytes + lab.nbytes + graph_bytes), "d2h_bytes_per_step": int(R * (C + 1) * 8),
           "ms_per_step": float(e2e_dt.item()) / e2e_steps * 1e3,
           "note": "tpgx_set_observations_pinned (page-locked host buffers) + tpgx_set_graph + tpgx_evaluate_classification (host outputs) per step"}

    if rank == 0:
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            from oracle import oracle
            g1 = g
            cpu, _, _ = cpu_baseline(oracle, tpgx, g1, img, lab, a.cpu_seconds, os.cpu_count() or 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(a, world),
            "instr_per_s": cnt["lines_executed"] * world * a.steps / (ms * 1e-3),
            "reference_equivalent_work": {"evaluations": cnt["evaluations"], "lines_executed": cnt["lines_executed"],
                                          "program_executions": cnt["program_executions"], "device_lines": cnt["device_lines"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 3 * a.steps,
            "clocks": sampler.result(), "wall_s_timed_region": wall, "score_checksum": score_check,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
