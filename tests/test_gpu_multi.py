"""Multi-GPU on hardware (needs >= 2 devices; the single-GPU test box skips it, `gpurun --gpus 2` runs it): roots sharded
over the devices of ONE process (tpgx_comm_init_all = ncclCommInitAll, the reference's single-process model), per-root
records combined by the library's own NCCL all-gather — scores of 1 GPU vs 2 / 4 / 8 GPUs must be bit-identical
(SURVEY.md §8e acceptance), for even and uneven shards."""
import numpy as np
import pytest

import tpgx

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


def evaluator(g, img, lab, device):
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5, device=device)
    ev.set_graph(g)
    ev.set_observations(img, lab)
    return ev


@pytest.mark.parametrize("roots", [96, 101])
def test_all_gathered_scores_equal_the_single_gpu_scores(oracle_mod, roots):
    n = device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    g = tpgx.synthetic_graph("mnist", roots=roots, edges=5, lines=20, depth=2, share_prob=0.1, seed=13)
    img, lab = tpgx.synthetic_images(1000, seed=3)
    single = evaluator(g, img, lab, 0).evaluate_classification(want=("score", "class_f1"))
    ref = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5).evaluate_classification(
        img, lab, want=("score", "class_f1"))
    assert np.array_equal(single["score"].view(np.uint64), ref["score"].view(np.uint64))
    for world in sorted({2, n} | ({4} if n >= 4 else set())):
        evs = [evaluator(g, img, lab, d) for d in range(world)]
        tpgx.comm_init_all(evs)
        shards = [e.comm_shard(roots) for e in evs]
        assert [s[:2] for s in shards] == [(r, world) for r in range(world)]
        assert shards[0][2] == 0 and shards[-1][3] == roots and all(a[3] == b[2] for a, b in zip(shards, shards[1:]))
        for _ in range(2):   # twice: the gather buffer is reused
            out = tpgx.evaluate_classification_allgather_all(evs, roots, len(img))
            assert np.array_equal(out["score"].view(np.uint64), single["score"].view(np.uint64)), world
            assert np.array_equal(out["class_f1"].view(np.uint64), single["class_f1"].view(np.uint64)), world
        for e in evs:
            e.close()


def test_allgather_without_a_communicator_is_the_single_gpu_evaluation():
    g = tpgx.synthetic_graph("mnist", roots=20, edges=5, lines=10, depth=1, seed=2)
    img, lab = tpgx.synthetic_images(300, seed=3)
    ev = evaluator(g, img, lab, 0)
    a = ev.evaluate_classification(want=("score", "class_f1"))
    b = ev.evaluate_classification_allgather(20)
    assert ev.comm_shard(20) == (0, 1, 0, 20)
    assert np.array_equal(a["score"].view(np.uint64), b["score"].view(np.uint64)) and np.array_equal(a["class_f1"], b["class_f1"])
    assert tpgx.kernel_launches() > 0


def test_sharded_set_graph_flattens_only_the_shard_and_refuses_other_roots():
    n = device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    roots = 64
    g = tpgx.synthetic_graph("mnist", roots=roots, edges=5, lines=20, depth=2, share_prob=0.1, seed=21)
    img, lab = tpgx.synthetic_images(700, seed=4)
    single = evaluator(g, img, lab, 0).evaluate_classification(want=("score", "class_f1"))
    evs = [evaluator(g, img, lab, d) for d in range(2)]
    tpgx.comm_init_all(evs)
    for e in evs:
        e.set_graph_sharded(g)
    out = tpgx.evaluate_classification_allgather_all(evs, roots, len(img))
    assert np.array_equal(out["score"].view(np.uint64), single["score"].view(np.uint64))
    assert np.array_equal(out["class_f1"].view(np.uint64), single["class_f1"].view(np.uint64))
    _, _, rb, re = evs[1].comm_shard(roots)
    own = evs[1].evaluate_classification(root_begin=rb, root_end=re, want=("score",))["score"]
    assert np.array_equal(own.view(np.uint64), single["score"][rb:re].view(np.uint64))
    with pytest.raises(tpgx.TpgxError) as err:
        evs[1].evaluate_classification(root_begin=0, root_end=4, want=("score",))
    assert err.value.status == tpgx.abi.ERR_STATE
    evs[1].set_graph(g)                                  # the plain call lifts the restriction
    assert np.array_equal(evs[1].evaluate_classification(root_begin=0, root_end=4, want=("score",))["score"], single["score"][:4])
