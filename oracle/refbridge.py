"""ctypes binding of oracle/_ref/libtpgref.so — the reference's own environment and instruction
sources compiled unmodified (oracle/build_ref.sh).  TEST INFRASTRUCTURE: used by
tests/test_oracle_vs_ref.py and tests/golden/make_fixtures.py to pin the oracle; never by the product."""
import ctypes as C
import os
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libtpgref.so")
APPS = {"mnist": 0, "pendulum": 1, "gridworld": 2, "stickgame": 3, "tictactoe": 4}
KIND_DOUBLE, KIND_INT, KIND_CONSTANT, KIND_WINDOW = 0, 1, 2, 3
ENV_GRIDWORLD, ENV_STICKGAME, ENV_TICTACTOE, ENV_PENDULUM = 0, 1, 2, 3
PENDULUM_ACTIONS = [0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0]  # [REF] pendulum/src/Learn/main.cpp:33
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_set_call.restype = C.c_double
        lib.ref_set_call.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_env_create.restype = C.c_void_p
        lib.ref_env_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.ref_env_clone.restype = C.c_void_p
        lib.ref_env_clone.argtypes = [C.c_void_p]
        lib.ref_env_destroy.argtypes = [C.c_void_p]
        lib.ref_env_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        lib.ref_env_do_action.argtypes = [C.c_void_p, C.c_double]
        lib.ref_env_is_terminal.argtypes = [C.c_void_p]
        lib.ref_env_score.restype = C.c_double
        lib.ref_env_score.argtypes = [C.c_void_p, C.c_int]
        lib.ref_env_nb_actions.argtypes = [C.c_void_p]
        lib.ref_env_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_env_set_pendulum.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib = lib
    return _lib


def set_signature(app):
    """[(nb_operands, [operand kinds])] of the reference's Instructions::Set for `app`."""
    lib, a = load(), APPS[app]
    out = []
    for i in range(lib.ref_set_size(a)):
        n = lib.ref_set_nb_operands(a, i)
        out.append((n, [lib.ref_set_operand_kind(a, i, k) for k in range(n)]))
    return out


def call(app, index, d=(0.0, 0.0), i=(0, 0), c=(0.0, 0.0), win=None):
    """Runs the reference lambda `index` of `app`; every operand slot is supplied in every representation."""
    d = np.asarray(d, dtype=np.float64); iv = np.asarray(i, dtype=np.int32); cv = np.asarray(c, dtype=np.float64)
    w = np.ascontiguousarray(win, dtype=np.float64) if win is not None else np.zeros(9)
    return load().ref_set_call(APPS[app], index, d.ctypes.data, iv.ctypes.data, cv.ctypes.data, w.ctypes.data)


class Env:
    def __init__(self, kind, second_player_random=0, pendulum_actions=None, _h=None):
        self.lib = load()
        self.kind = kind
        if _h is not None:
            self.h = _h
            return
        pa = np.ascontiguousarray(pendulum_actions if pendulum_actions is not None else PENDULUM_ACTIONS, dtype=np.float64)
        self.h = self.lib.ref_env_create(kind, second_player_random, len(pa), pa.ctypes.data)
        if not self.h:
            raise RuntimeError("unknown env")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_env_destroy(self.h); self.h = None

    def clone(self): return Env(self.kind, _h=self.lib.ref_env_clone(self.h))
    def reset(self, seed=0, mode=0): self.lib.ref_env_reset(self.h, seed, mode)
    def do_action(self, a): return self.lib.ref_env_do_action(self.h, float(a))
    def is_terminal(self): return bool(self.lib.ref_env_is_terminal(self.h))
    def score(self, player=0): return self.lib.ref_env_score(self.h, player)
    def nb_actions(self): return self.lib.ref_env_nb_actions(self.h)
    def set_pendulum(self, angle, velocity): return self.lib.ref_env_set_pendulum(self.h, angle, velocity)

    def observe(self, n=16):
        s = np.zeros(n)
        k = self.lib.ref_env_observe(self.h, s.ctypes.data, n)
        return s[:min(k, n)]


# ------------------------------------------------------------------------------------------------
# golden vectors (committed under tests/golden/ so that they travel to boxes without /root/reference)

SPECIALS = [0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 10.0, -10.0, 15.0, 255.0, 1e-300, -1e-300, 5e-324, 1e300, -1e300,
            float("inf"), float("-inf"), float("nan"), 3.141592653589793, 709.78, 710.0, -745.2, 0.4375, 2.4375, 1e-9]


def operand_grid(seed=7, n_random=40):
    rng = np.random.default_rng(seed)
    vals = np.array(SPECIALS + list(rng.normal(0, 50, n_random)) + list(rng.integers(0, 256, 20).astype(float)))
    a, b = np.meshgrid(vals, vals, indexing="ij")
    return a.ravel(), b.ravel()


def window_grid(seed=9, n=400):
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 256, (n, 9)).astype(np.float64)
    w[rng.random((n, 9)) < 0.5] = 0.0
    w[:8] = 0.0
    w[8:16] = 255.0
    return w


def dump_instructions(path):
    a, b = operand_grid()
    ia, ib = np.clip(np.nan_to_num(a, nan=0, posinf=100, neginf=-100), -1000, 1000).astype(np.int32), \
        np.clip(np.nan_to_num(b, nan=0, posinf=100, neginf=-100), -1000, 1000).astype(np.int32)
    wins = window_grid()
    out = {"a": a, "b": b, "ia": ia, "ib": ib, "win": wins}
    for app in APPS:
        for idx, (n, kinds) in enumerate(set_signature(app)):
            if kinds[0] == KIND_WINDOW:
                res = np.array([call(app, idx, win=w) for w in wins])
            else:
                res = np.array([call(app, idx, d=(x, y), i=(p, q), c=(x, y)) for x, y, p, q in zip(a, b, ia, ib)])
            out["%s_%d" % (app, idx)] = res
            out["%s_%d_kinds" % (app, idx)] = np.array(kinds, dtype=np.int32)
    np.savez_compressed(path, **out)


ENV_CASES = [  # (name, kind, second_player_random, nb_actions, max steps)
    ("gridworld", ENV_GRIDWORLD, 0, 4, 100), ("stickgame_random", ENV_STICKGAME, 1, 3, 11), ("stickgame_selfplay", ENV_STICKGAME, 0, 3, 21),
    ("tictactoe_adversarial", ENV_TICTACTOE, 0, 9, 9), ("tictactoe_random", ENV_TICTACTOE, 1, 9, 9), ("pendulum", ENV_PENDULUM, 0, 15, 400),
]


def run_trace(env, seed, mode, actions, max_steps, n_obs=12):
    """reset then apply `actions` until terminal/max; returns per-step rows [obs..., terminal, score0, score1]."""
    env.reset(seed, mode)
    rows = [np.concatenate([_pad(env.observe(n_obs), n_obs), [float(env.is_terminal()), env.score(0), _score1(env)]])]
    for t in range(max_steps):
        if env.is_terminal():
            break
        env.do_action(actions[t])
        rows.append(np.concatenate([_pad(env.observe(n_obs), n_obs), [float(env.is_terminal()), env.score(0), _score1(env)]]))
    return np.array(rows)


def _pad(v, n):
    o = np.zeros(n); o[:len(v)] = v
    return o


def _score1(env):
    try:
        return env.score(1)
    except Exception:  # noqa: BLE001
        return 0.0


def dump_env_traces(path, episodes=12):
    out = {}
    for name, kind, spr, nact, steps in ENV_CASES:
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000 + 11)
        for ep in range(episodes):
            seed, mode = int(rng.integers(0, 2**40)), int(ep % 3)
            actions = rng.integers(0, nact, steps).astype(np.float64)
            env = Env(kind, spr)
            out["%s_%d_meta" % (name, ep)] = np.array([kind, spr, seed, mode], dtype=np.int64)
            out["%s_%d_actions" % (name, ep)] = actions
            out["%s_%d_trace" % (name, ep)] = run_trace(env, seed, mode, actions, steps)
    np.savez_compressed(path, **out)


def dump_golden(directory):
    dump_instructions(os.path.join(directory, "ref_instructions.npz"))
    dump_env_traces(os.path.join(directory, "ref_env_traces.npz"))
    print("wrote ref_instructions.npz, ref_env_traces.npz to", directory)
