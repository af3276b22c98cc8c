"""GPU parity of the instruction table on every app's data-source shape: the generic executor
(tpgx_execute_programs) vs the oracle, bit-exact, on random programs of each app's instruction set."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import tpgx
from tpgx import abi as A

pytestmark = pytest.mark.gpu


def random_obs(app, count, rng):
    obs = []
    for elem, h, w, _ in tpgx.APP_SOURCES[app]:
        if elem == A.ELEM_I32:
            obs.append(rng.integers(-5, 25, (count, h * w)).astype(np.int32))
        elif app == "tictactoe":
            obs.append(rng.integers(-1, 2, (count, h * w)).astype(np.float64))
        elif app == "mnist":
            obs.append(rng.integers(0, 256, (count, h * w)).astype(np.float64) * (rng.random((count, h * w)) < 0.5))
        else:
            x = rng.normal(0, 3, (count, h * w))
            x[::7] *= 1e6     # large trig arguments
            x[::11] = 0.0
            obs.append(x)
    return obs


@pytest.mark.parametrize("app", ["mnist", "pendulum", "gridworld", "stickgame", "tictactoe"])
def test_random_programs_bit_exact(oracle_mod, app):
    rng = np.random.default_rng(hash(app) % 1000)
    g = tpgx.synthetic_graph(app, roots=40, edges=5, lines=16, depth=1, seed=17, intron_lines=4)
    nc = tpgx.APP_CONSTANTS[app]
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc)
    ev.set_graph(g)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc)
    count = 257
    obs = random_obs(app, count, rng)
    dev = ev.execute_programs(count, obs)
    ref = o.execute_programs(count, obs)
    assert np.array_equal(dev.view(np.uint64), ref.view(np.uint64))
    assert np.isfinite(ref).mean() > 0.2


@pytest.mark.parametrize("app,dot", [("pendulum", "Pendulum_out_best.dot"), ("stickgame", "StickGame_out_best.dot"),
                                     ("tictactoe", "TicTacToe_out_best.dot")])
def test_golden_graph_programs(oracle_mod, app, dot):
    """The reference's golden graphs (Appendix C) run through the device executor."""
    nc = tpgx.APP_CONSTANTS[app]
    g = tpgx.import_dot(os.path.join(GOLDEN, dot), nc)
    ev = tpgx.Evaluator(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], nb_constants=nc)
    ev.set_graph(g)
    o = oracle_mod.Oracle(tpgx.INSTRUCTION_SETS[app], tpgx.APP_SOURCES[app], g, nb_constants=nc)
    rng = np.random.default_rng(1)
    obs = random_obs(app, 100, rng)
    if app == "stickgame":
        obs[0][:] = [1, 2, 3, 4]
        obs[1][:, 0] = rng.integers(1, 22, 100)
    assert np.array_equal(ev.execute_programs(100, obs).view(np.uint64), o.execute_programs(100, obs).view(np.uint64))
