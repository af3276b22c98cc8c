"""ctypes binding of the CPU oracle (oracle/libtpgoracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "gegelati-apps_b200", "python"))
from tpgx import _cabi as A  # noqa: E402  (struct layouts of include/tpgx.h)

LIB_PATH = os.path.join(_HERE, "libtpgoracle.so")
MATH_STD, MATH_SHARED = 0, 1
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.POINTER(A.Config), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(A.Graph), C.c_int32, C.c_char_p, C.c_size_t]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_identify_introns.restype = C.c_int64
    lib.orc_identify_introns.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_effective_lines.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_execute_programs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.orc_evaluate_classification.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(A.ClassifRequest), C.POINTER(A.ClassifResult), C.c_int32]
    lib.orc_archive_classification.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(A.ClassifRequest), C.c_void_p, C.c_double, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    lib.orc_evaluate_episodes.argtypes = [C.c_void_p, C.POINTER(A.EpisodeRequest), C.POINTER(A.EpisodeResult), C.c_int32]
    lib.orc_archive_episodes.argtypes = [C.c_void_p, C.POINTER(A.EpisodeRequest), C.c_void_p, C.c_double, C.c_int32] + [C.c_void_p] * 6 + [
        C.c_int32, C.POINTER(C.c_int32)]
    lib.orc_apply_op.restype = C.c_double
    lib.orc_apply_op.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_void_p]
    lib.orc_math_array.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.orc_math_probe.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_env_create.restype = C.c_void_p
    lib.orc_env_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
    lib.orc_env_destroy.argtypes = [C.c_void_p]
    lib.orc_env_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_int32]
    lib.orc_env_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.orc_env_do_action.argtypes = [C.c_void_p, C.c_double]
    lib.orc_env_is_terminal.argtypes = [C.c_void_p]
    lib.orc_env_score.restype = C.c_double
    lib.orc_env_score.argtypes = [C.c_void_p, C.c_int32]
    lib.orc_env_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.orc_rng_u64.argtypes = [C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.orc_rng_f64.argtypes = [C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_void_p]
    lib.orc_rng_raw.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
    lib.orc_rng_restated_u64.argtypes = [C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.orc_rng_restated_f64.argtypes = [C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Oracle:
    def __init__(self, opcodes, sources, graph, nb_registers=8, nb_constants=0, tie_break=A.TIE_LAST_MAX, math=MATH_SHARED):
        self.lib = load()
        cfg = A.Config(device=0, nb_registers=nb_registers, nb_constants=nb_constants, tie_break=tie_break)
        ops = np.ascontiguousarray(opcodes, dtype=np.int32)
        sd = (A.SourceDesc * len(sources))(*[A.SourceDesc(*s) for s in sources])
        cs, self._keep = graph.c_struct()
        err = C.create_string_buffer(512)
        self.h = self.lib.orc_create(C.byref(cfg), len(ops), ops.ctypes.data, len(sources), C.cast(sd, C.c_void_p), C.byref(cs), math, err, 512)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.graph, self.sources = graph, list(sources)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def introns(self):
        flags = np.zeros(len(self.graph.lines), dtype=np.uint8)
        self.lib.orc_identify_introns(self.h, flags.ctypes.data)
        return flags

    def effective_lines(self):
        n = np.zeros(self.graph.nb_programs, dtype=np.int32)
        self.lib.orc_effective_lines(self.h, n.ctypes.data)
        return n

    def execute_programs(self, count, obs, raw=False):
        n = len(self.sources)
        f = (C.c_void_p * n)(); iv = (C.c_void_p * n)(); keep = []
        for k, (elem, h, w, _) in enumerate(self.sources):
            a = np.ascontiguousarray(obs[k], dtype=np.float64 if elem == A.ELEM_F64 else np.int32).reshape(count, h * w)
            keep.append(a)
            if elem == A.ELEM_F64: f[k] = a.ctypes.data
            else: iv[k] = a.ctypes.data
        bid = np.zeros((self.graph.nb_programs, count))
        self.lib.orc_execute_programs(self.h, count, C.cast(f, C.c_void_p), C.cast(iv, C.c_void_p), bid.ctypes.data, int(raw))
        return bid

    def evaluate_classification(self, images, labels, nb_samples=None, sample_index=None, nb_classes=10, root_begin=0,
                                root_end=None, want=("score", "class_f1", "confusion", "counters"), nb_threads=1):
        g = self.graph
        img = np.ascontiguousarray(images, dtype=np.uint8)
        lab = np.ascontiguousarray(labels, dtype=np.uint8)
        root_end = g.nb_roots if root_end is None else root_end
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else img.shape[0]))
        nR = root_end - root_begin
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=nb_classes, root_begin=root_begin, root_end=root_end)
        out, res, cnt = {}, A.ClassifResult(), A.Counters()
        if "score" in want: out["score"] = np.zeros(nR); res.score = out["score"].ctypes.data
        if "class_f1" in want: out["class_f1"] = np.zeros((nR, nb_classes)); res.class_f1 = out["class_f1"].ctypes.data
        if "confusion" in want: out["confusion"] = np.zeros((nR, nb_classes, nb_classes), dtype=np.uint64); res.confusion = out["confusion"].ctypes.data
        if "action" in want: out["action"] = np.zeros((nR, nS), dtype=np.uint8); res.action = out["action"].ctypes.data
        if "bid" in want: out["bid"] = np.zeros((g.nb_programs, nS)); res.bid = out["bid"].ctypes.data
        res.counters = C.pointer(cnt)
        st = self.lib.orc_evaluate_classification(self.h, img.ctypes.data, lab.ctypes.data, img.shape[0], C.byref(req), C.byref(res), nb_threads)
        if st != 0:
            raise RuntimeError("oracle classification failed: %d" % st)
        out["counters"] = cnt.as_dict()
        return out

    def archive_classification(self, images, archive_seeds, probability, archive_size, nb_samples=None, sample_index=None,
                               root_begin=0, root_end=None):
        img = np.ascontiguousarray(images, dtype=np.uint8)
        root_end = self.graph.nb_roots if root_end is None else root_end
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else img.shape[0]))
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=10, root_begin=root_begin, root_end=root_end)
        seeds = np.ascontiguousarray(archive_seeds, dtype=np.uint64)
        prog = np.zeros(archive_size, dtype=np.int32); samp = np.zeros(archive_size, dtype=np.int64); bid = np.zeros(archive_size)
        n = C.c_int32(0)
        st = self.lib.orc_archive_classification(self.h, img.ctypes.data, img.shape[0], C.byref(req), seeds.ctypes.data, float(probability),
                                                 int(archive_size), prog.ctypes.data, samp.ctypes.data, bid.ctypes.data, C.byref(n))
        if st != 0:
            raise RuntimeError("oracle status %d" % st)
        return dict(program=prog[:n.value], sample=samp[:n.value], bid=bid[:n.value])

    def evaluate_episodes(self, env, seeds, job_roots, max_actions, mode=A.MODE_TRAINING, roots_per_job=1,
                          second_player_random=0, pendulum_actions=None, trace_steps=0, nb_threads=1, rng_seeds=None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        rs = np.ascontiguousarray(rng_seeds, dtype=np.uint64) if rng_seeds is not None else None
        jr = np.ascontiguousarray(job_roots, dtype=np.int32).reshape(-1, roots_per_job)
        J, NI = jr.shape[0], len(seeds)
        pa = np.ascontiguousarray(pendulum_actions, dtype=np.float64) if pendulum_actions is not None else None
        req = A.EpisodeRequest(env=env, mode=mode, nb_iterations=NI, max_actions=max_actions, seeds=seeds.ctypes.data, nb_jobs=J,
                               roots_per_job=roots_per_job, job_roots=jr.ctypes.data, second_player_random=second_player_random,
                               nb_pendulum_actions=0 if pa is None else len(pa), pendulum_actions=_ptr(pa), trace_steps=trace_steps,
                               rng_seeds=_ptr(rs))
        out = dict(score=np.zeros((J, roots_per_job)), episode_score=np.zeros((J, NI, roots_per_job)),
                   episode_actions=np.zeros((J, NI), dtype=np.int32), final_state=np.zeros((J, NI, 4)))
        cnt = A.Counters()
        res = A.EpisodeResult(score=out["score"].ctypes.data, episode_score=out["episode_score"].ctypes.data,
                              episode_actions=out["episode_actions"].ctypes.data, final_state=out["final_state"].ctypes.data,
                              counters=C.pointer(cnt))
        if trace_steps > 0:
            out["trace"] = np.zeros((J, NI, trace_steps), dtype=np.int8)
            res.trace = out["trace"].ctypes.data
        st = self.lib.orc_evaluate_episodes(self.h, C.byref(req), C.byref(res), nb_threads)
        if st != 0:
            raise RuntimeError("oracle episodes failed: %d" % st)
        out["counters"] = cnt.as_dict()
        return out


def _archive_episodes(self, env, seeds, job_roots, max_actions, archive_seeds, probability, archive_size, obs_len=0,
                      mode=A.MODE_TRAINING, roots_per_job=1, second_player_random=0, pendulum_actions=None):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    jr = np.ascontiguousarray(job_roots, dtype=np.int32).reshape(-1, roots_per_job)
    J, NI = jr.shape[0], len(seeds)
    pa = np.ascontiguousarray(pendulum_actions, dtype=np.float64) if pendulum_actions is not None else None
    req = A.EpisodeRequest(env=env, mode=mode, nb_iterations=NI, max_actions=max_actions, seeds=seeds.ctypes.data, nb_jobs=J,
                           roots_per_job=roots_per_job, job_roots=jr.ctypes.data, second_player_random=second_player_random,
                           nb_pendulum_actions=0 if pa is None else len(pa), pendulum_actions=_ptr(pa), trace_steps=0)
    aseeds = np.ascontiguousarray(archive_seeds, dtype=np.uint64)
    n = max(1, int(archive_size))
    o = dict(program=np.zeros(n, dtype=np.int32), job=np.zeros(n, dtype=np.int32), iteration=np.zeros(n, dtype=np.int32),
             step=np.zeros(n, dtype=np.int32), bid=np.zeros(n))
    obs = np.zeros((n, max(1, obs_len)))
    cnt = C.c_int32(0)
    st = self.lib.orc_archive_episodes(self.h, C.byref(req), aseeds.ctypes.data, float(probability), int(archive_size), o["program"].ctypes.data,
                                       o["job"].ctypes.data, o["iteration"].ctypes.data, o["step"].ctypes.data, o["bid"].ctypes.data,
                                       obs.ctypes.data if obs_len > 0 else None, int(obs_len), C.byref(cnt))
    if st != 0:
        raise RuntimeError("oracle status %d" % st)
    out = {k: v[:cnt.value] for k, v in o.items()}
    if obs_len > 0:
        out["observation"] = obs[:cnt.value]
    return out


Oracle.archive_episodes = _archive_episodes


def apply_op(opcode, a=0.0, b=0.0, ia=0, ib=0, c=0.0, win=None, math=MATH_SHARED):
    w = np.ascontiguousarray(win, dtype=np.float64) if win is not None else None
    return load().orc_apply_op(opcode, math, a, b, ia, ib, c, _ptr(w))


def math_array(fn, x, math=MATH_SHARED):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    load().orc_math_array(fn, math, x.size, x.ctypes.data, out.ctypes.data)
    return out


def math_probe(fn, a, b=None):
    """Host build of the arithmetic primitive `fn` (tpgx.abi PROBE_*) — the checker for tpgx_math_probe."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
    out = np.empty_like(a)
    load().orc_math_probe(fn, a.size, a.ctypes.data, _ptr(b), out.ctypes.data)
    return out


class Env:
    def __init__(self, kind, second_player_random=0, pendulum_actions=None, math=MATH_SHARED):
        self.lib = load()
        pa = np.ascontiguousarray(pendulum_actions, dtype=np.float64) if pendulum_actions is not None else None
        self.h = self.lib.orc_env_create(kind, second_player_random, 0 if pa is None else len(pa), _ptr(pa), math)
        if not self.h:
            raise RuntimeError("unknown env")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_env_destroy(self.h); self.h = None

    def reset(self, seed=0, mode=0): self.lib.orc_env_reset(self.h, seed, mode)
    def do_action(self, a): return self.lib.orc_env_do_action(self.h, float(a))
    def is_terminal(self): return bool(self.lib.orc_env_is_terminal(self.h))
    def score(self, player=0): return self.lib.orc_env_score(self.h, player)
    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64); self.lib.orc_env_set_state(self.h, s.ctypes.data, len(s))
    def state(self, n=16):
        s = np.zeros(n); k = self.lib.orc_env_state(self.h, s.ctypes.data, n); return s[:min(k, n)]
