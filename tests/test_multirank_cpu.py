"""N > 1 host logic on CPU: two processes (gloo, 127.0.0.1) each evaluate their root shard — with the
oracle standing in for the GPU evaluator, since this container has no GPU — gather the per-root records with
the same helper bench.py uses, and must reproduce the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import tpgx


def test_shard_ranges_partition_the_roots():
    for n in (0, 1, 7, 8, 2000, 8192):
        for world in (1, 2, 3, 8):
            r = [tpgx.shard_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        tpgx.shard_range(2, 2, 10)


def test_slot_shards_of_the_library_partition_the_roots_and_index_the_gather_buffer():
    """tpgx_comm_shard's convention (restated in tpgx.slot_shard): equal slots, so that rank r's records sit at their global
    root numbers in the all-gathered buffer; trailing ranks may be short or empty."""
    for n in (1, 7, 8, 13, 2000, 8192):
        for world in (1, 2, 3, 8):
            r = [tpgx.slot_shard(k, world, n) for k in range(world)]
            slot = -(-n // world)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(b == min(n, k * slot) and e - b <= slot for k, (b, e) in enumerate(r))


def _worker(rank, world, port, nb_roots, out_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gegelati-apps_b200", "python"))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import tpgx as T
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = T.synthetic_graph("mnist", roots=nb_roots, edges=4, lines=10, depth=2, share_prob=0.1, seed=5)
    img, lab = T.synthetic_images(96, seed=8)
    o = oracle.Oracle(T.INSTRUCTION_SETS["mnist"], T.APP_SOURCES["mnist"], g, nb_constants=5)
    b, e = T.shard_range(rank, world, nb_roots)
    r = o.evaluate_classification(img, lab, want=("score", "class_f1"), root_begin=b, root_end=e)
    local = torch.from_numpy(T.records_from(r["score"], r["class_f1"]))
    full = T.gather_records(local, world, nb_roots, dist)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_process(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    nb_roots = 13  # odd on purpose: shards of 7 and 6 roots
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.start_processes(_worker, args=(2, port, nb_roots, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    g = tpgx.synthetic_graph("mnist", roots=nb_roots, edges=4, lines=10, depth=2, share_prob=0.1, seed=5)
    img, lab = tpgx.synthetic_images(96, seed=8)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
    r = o.evaluate_classification(img, lab, want=("score", "class_f1"))
    ref = tpgx.records_from(r["score"], r["class_f1"])
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), rank


def test_subgraph_of_a_root_shard_evaluates_identically(oracle_mod):
    """TPGraph.subgraph (what a rank uploads) keeps exactly the reachable part in reference order: scores of the
    shard's roots are unchanged, shared inner teams and programs included."""
    g = tpgx.synthetic_graph("mnist", roots=14, edges=4, lines=9, depth=3, share_prob=0.2, seed=23)
    img, lab = tpgx.synthetic_images(80, seed=3)
    full = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5).evaluate_classification(
        img, lab, want=("score", "confusion", "counters"))
    for b, e in ((0, 5), (5, 14), (3, 4)):
        sg = g.subgraph(b, e)
        assert sg.nb_roots == e - b and sg.nb_teams <= g.nb_teams and sg.nb_programs <= g.nb_programs
        part = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], sg, nb_constants=5).evaluate_classification(
            img, lab, want=("score", "confusion"))
        assert np.array_equal(part["score"].view(np.uint64), full["score"][b:e].view(np.uint64))
        assert np.array_equal(part["confusion"], full["confusion"][b:e])
