#!/usr/bin/env python3
"""End-to-end time of one evaluateAllRoots of BASELINE config 4 under torchrun (one rank per GPU): two-call form against the one-call
pipeline with 1 / 2 / 3 / 4 / 6 slices per rank (TPGX_PIPELINE).  usage: torchrun --nproc-per-node N e2e_pipeline_case.py [total roots = 8192]"""
import os, sys, time
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
import numpy as np, torch, torch.distributed as dist, tpgx
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = tpgx.synthetic_graph("mnist", roots=R, edges=5, lines=20, depth=1, mix="uniform", seed=42)
img, lab = tpgx.synthetic_images(60000)
ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=local)
if world > 1:
    uid = torch.from_numpy(tpgx.Evaluator.comm_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(uid, 0)
    ev.comm_init(uid.cpu().numpy(), rank, world)
ev.set_observations(img, lab)
def run(mode, n=6):
    for i in range(n + 2):
        if i == 2:
            dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
        if mode == "two":
            ev.set_graph_sharded(g); o = ev.evaluate_classification_allgather(R, 60000)
        else:
            o = ev.evaluate_graph_allgather(g, R, 60000)
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return dt.item() / n * 1e3, float(o["score"].sum())
res = [("two-call", run("two"))]
for k in (1, 2, 3, 4, 6):
    os.environ["TPGX_PIPELINE"] = str(k)
    res.append(("pipeline %d" % k, run("one")))
if rank == 0:
    for name, (ms, chk) in res: print("N=%d R=%d %-12s e2e %.3f ms/step  checksum %.10f" % (world, R, name, ms, chk), flush=True)
dist.barrier(); dist.destroy_process_group()
