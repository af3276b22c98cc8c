"""Test-only backend for the host generation loop: the CPU oracle behind tpgx_train_backend's two callbacks, with the same
request conventions as the GPU backend (tpgx_train_backend_episodes): job j = root j, iteration seeds 1000 * generation + it."""
import numpy as np

import tpgx
from tpgx import abi as A

SMALL = dict(nb_roots=60, archive_size=300)  # params.json values scaled down so that CPU tests stay fast


def split_obs(sources, obs):
    out, col = [], 0
    for (elem, h, w, _) in sources:
        sp = h * w
        blk = obs[:, col:col + sp]
        out.append(np.ascontiguousarray(blk, dtype=np.float64) if elem == A.ELEM_F64 else np.ascontiguousarray(blk.astype(np.int32)))
        col += sp
    return out


def oracle_backend(oracle_mod, app, nb_iterations, max_actions, second_player_random=0, pendulum_actions=None, log=None):
    ops, src, nc, env = tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], tpgx.APP_CONSTANTS[app], tpgx.APP_ENV[app]

    def evaluate(g, generation, mode, aseeds, prob, cap, obs_len):
        o = oracle_mod.Oracle(ops, src, g, nb_constants=nc)
        kw = dict(seeds=np.arange(nb_iterations, dtype=np.uint64) + np.uint64(1000 * generation), job_roots=np.arange(g.nb_roots),
                  max_actions=max_actions, second_player_random=second_player_random, pendulum_actions=pendulum_actions, mode=mode)
        score = o.evaluate_episodes(env, **kw)["score"][:, 0]
        if log is not None:
            log.append((mode, g.nb_roots, score.copy()))
        if cap == 0 or prob == 0.0 or mode != A.MODE_TRAINING:
            return score, [], [], np.zeros((0, obs_len))
        ar = o.archive_episodes(env, archive_seeds=aseeds, probability=prob, archive_size=cap, obs_len=obs_len, **kw)
        return score, ar["program"], ar["bid"], ar["observation"]

    def execute(g, obs):
        o = oracle_mod.Oracle(ops, src, g, nb_constants=nc)
        return o.execute_programs(len(obs), split_obs(src, obs))

    return tpgx.python_backend(evaluate, execute, nb_constants=nc)


def oracle_classification_backend(oracle_mod, images, labels, actions, log=None):
    """Same conventions as tpgx_train_backend_classification: the images of generation `gen` are those
    tpgx_mnist_episode_indices draws from seed 1000 * gen in TRAINING mode; the archived observation is the image."""
    ops, src, nc = tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], tpgx.APP_CONSTANTS["mnist"]

    def evaluate(g, generation, mode, aseeds, prob, cap, obs_len):
        o = oracle_mod.Oracle(ops, src, g, nb_constants=nc)
        idx = tpgx.mnist_episode_indices(seed=1000 * generation, mode=mode, dataset_size=len(images), nb_actions=actions)
        out = o.evaluate_classification(images, labels, sample_index=idx, want=("score", "class_f1", "confusion"))
        score, cls = out["score"], (out["class_f1"], out["confusion"].sum(axis=2))
        if log is not None:
            log.append((mode, g.nb_roots, score.copy()))
        if cap == 0 or prob == 0.0 or mode != A.MODE_TRAINING:
            return (score, [], [], np.zeros((0, obs_len))) + cls
        ar = o.archive_classification(images, aseeds, prob, cap, sample_index=idx)
        return (score, ar["program"], ar["bid"], images[ar["sample"]].astype(np.float64)) + cls

    def execute(g, obs):
        return oracle_mod.Oracle(ops, src, g, nb_constants=nc).execute_programs(len(obs), [obs])

    return tpgx.python_backend(evaluate, execute, nb_constants=nc)
