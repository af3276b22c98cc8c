"""Host generation loop (include/tpgx.h: tpgx_trainer_*): Learn::LearningAgent::{init, trainOneGeneration} with the GPU
evaluator plugged in — populateTPG -> evaluateAllRoots(TRAINING) + archive -> decimateWorstRoots."""
import ctypes as C
import json
import re

import numpy as np

from . import _cabi as A
from .api import TpgxError, load_library
from .graph import LINE_DTYPE, TPGraph


def load_params_json(path, nb_actions):
    """The keys of an app's params.json the loop uses ([REF] gridworld/params.json; comments as the apps write them are
    stripped — the full File::ParametersParser is out of scope).  Returns the dict Trainer takes."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    txt = re.sub(r",(\s*[}\]])", r"\1", txt)
    j = json.loads(txt)
    mt, mp = j["mutation"]["tpg"], j["mutation"]["prog"]
    return dict(
        nb_actions=nb_actions, nb_roots=mt["nbRoots"], max_init_outgoing_edges=mt["maxInitOutgoingEdges"], max_outgoing_edges=mt["maxOutgoingEdges"],
        p_edge_deletion=mt["pEdgeDeletion"], p_edge_addition=mt["pEdgeAddition"], p_program_mutation=mt["pProgramMutation"],
        p_edge_destination_change=mt["pEdgeDestinationChange"], p_edge_destination_is_action=mt["pEdgeDestinationIsAction"],
        force_program_behavior_change=int(bool(mt.get("forceProgramBehaviorChangeOnMutation", False))),
        max_program_size=mp["maxProgramSize"], p_delete=mp["pDelete"], p_add=mp["pAdd"], p_mutate=mp["pMutate"], p_swap=mp["pSwap"],
        p_constant_mutation=mp.get("pConstantMutation", 0.0), min_const_value=mp.get("minConstValue", 0), max_const_value=mp.get("maxConstValue", 0),
        ratio_deleted_roots=j["ratioDeletedRoots"], archive_size=j["archiveSize"], archiving_probability=j["archivingProbability"],
        # not trainer fields, but what the evaluation request needs
        nb_iterations=j.get("nbIterationsPerPolicyEvaluation", 1), max_actions=j["maxNbActionsPerEval"], nb_generations=j["nbGenerations"],
        nb_registers=j.get("nbRegisters", 8), nb_constants=j.get("nbProgramConstant", 0),
        # [UP] evaluateJob's skip-and-merge of previous results and the validation pass of trainOneGeneration
        max_nb_evaluation_per_policy=j.get("maxNbEvaluationPerPolicy", 0), do_validation=int(bool(j.get("doValidation", False))),
        nb_iterations_per_policy_evaluation=j.get("nbIterationsPerPolicyEvaluation", 1), nb_classes=0)


_TRAINER_FIELDS = [k for k, _ in A.TrainParams._fields_ if k != "reserved"]


def python_backend(evaluate, execute=None, nb_constants=0):
    """Backend from Python callables (tests plug the CPU oracle in here; the product backend is Evaluator-based, see
    episodes_backend).  evaluate(graph, generation, mode, archive_seeds, probability, archive_size, obs_len) ->
    (scores[R], rec_program, rec_bid, rec_obs[n, obs_len]) or, for a classification agent, the same tuple followed by
    (class_scores[R, C], class_counts[R, C]); the graph's root list holds exactly the jobs.
    execute(graph, obs[count, obs_len]) -> bid[P, count]."""
    def graph_of(gp):
        return view_to_graph(gp.contents, nb_constants)

    def ev(_user, gp, generation, mode, aseed, prob, cap, scores, nb_classes, cscores, ccounts, rprog, rbid, robs, obs_len, nrec):
        try:
            g = graph_of(gp)
            seeds = np.ctypeslib.as_array(aseed, shape=(g.nb_roots,)).copy() if aseed else np.zeros(g.nb_roots, dtype=np.uint64)
            res = evaluate(g, int(generation), int(mode), seeds, float(prob), int(cap), int(obs_len))
            sc, p, b, o = res[:4]
            np.ctypeslib.as_array(scores, shape=(g.nb_roots,))[:] = sc
            if nb_classes > 0:
                np.ctypeslib.as_array(cscores, shape=(g.nb_roots * nb_classes,))[:] = np.asarray(res[4], dtype=np.float64).reshape(-1)
                np.ctypeslib.as_array(ccounts, shape=(g.nb_roots * nb_classes,))[:] = np.asarray(res[5], dtype=np.uint64).reshape(-1)
            n = len(p)
            if n:
                np.ctypeslib.as_array(rprog, shape=(n,))[:] = p
                np.ctypeslib.as_array(rbid, shape=(n,))[:] = b
                np.ctypeslib.as_array(robs, shape=(n * obs_len,))[:] = np.asarray(o, dtype=np.float64).reshape(-1)
            nrec[0] = n
            return 0
        except Exception:  # an exception must not cross the C boundary
            import traceback
            traceback.print_exc()
            return A.ERR_STATE

    def ex(_user, gp, count, obs, obs_len, bid):
        try:
            g = graph_of(gp)
            o = np.ctypeslib.as_array(obs, shape=(count, obs_len)).copy()
            np.ctypeslib.as_array(bid, shape=(g.nb_programs * count,))[:] = np.asarray(execute(g, o), dtype=np.float64).reshape(-1)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return A.ERR_STATE

    be = A.TrainBackend(user=None, evaluate=A.EVALUATE_FN(ev), execute=A.EXECUTE_FN(ex if execute else (lambda *a: A.ERR_UNSUPPORTED)))
    be._keep = (ev, ex)
    return be


def view_to_graph(v, nb_constants):
    """Copies a tpgx_graph view (C struct) into a TPGraph."""
    def arr(ptr, ctype, n, dtype):
        if not ptr or n == 0:
            return np.zeros(n, dtype=dtype)
        return np.frombuffer((ctype * n).from_address(ptr), dtype=dtype).copy()
    P, T, R = v.nb_programs, v.nb_teams, v.nb_roots
    plo = arr(v.program_line_offset, C.c_int64, P + 1, np.int64)
    nl = int(plo[-1]) if P else 0
    lines = np.frombuffer((C.c_uint8 * (nl * LINE_DTYPE.itemsize)).from_address(v.lines), dtype=LINE_DTYPE).copy() if nl else np.zeros(0, dtype=LINE_DTYPE)
    teo = arr(v.team_edge_offset, C.c_int32, T + 1, np.int32)
    E = int(teo[-1])
    consts = arr(v.constants, C.c_double, P * nb_constants, np.float64).reshape(P, nb_constants)
    return TPGraph(lines=lines, program_line_offset=plo, constants=consts, team_edge_offset=teo,
                   edge_program=arr(v.edge_program, C.c_int32, E, np.int32), edge_destination=arr(v.edge_destination, C.c_int32, E, np.int32),
                   root_team=arr(v.root_team, C.c_int32, R, np.int32))


class Trainer:
    """Owns one tpgx_trainer (host only; evaluation goes through a backend)."""

    def __init__(self, opcodes, sources, params, nb_registers=8, nb_constants=0, seed=0):
        self.lib = load_library()
        L = self.lib
        L.tpgx_trainer_create.argtypes = [C.POINTER(A.Config), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(A.TrainParams), C.c_uint64,
                                          C.POINTER(C.c_void_p)]
        L.tpgx_trainer_destroy.argtypes = [C.c_void_p]
        L.tpgx_trainer_last_error.argtypes = [C.c_void_p]
        L.tpgx_trainer_last_error.restype = C.c_char_p
        L.tpgx_trainer_graph.argtypes = [C.c_void_p, C.POINTER(A.Graph)]
        L.tpgx_trainer_train_generation.argtypes = [C.c_void_p, C.POINTER(A.TrainBackend), C.c_uint64, C.POINTER(A.TrainStats)]
        L.tpgx_trainer_populate.argtypes = [C.c_void_p, C.POINTER(A.TrainBackend), C.POINTER(A.TrainStats)]
        L.tpgx_trainer_decimate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(A.TrainStats)]
        L.tpgx_trainer_fingerprint.argtypes = [C.c_void_p]
        L.tpgx_trainer_fingerprint.restype = C.c_uint64
        self.nb_constants = nb_constants
        cfg = A.Config(device=0, nb_registers=nb_registers, nb_constants=nb_constants, tie_break=A.TIE_LAST_MAX)
        ops = np.ascontiguousarray(opcodes, dtype=np.int32)
        sd = (A.SourceDesc * len(sources))(*[A.SourceDesc(*s) for s in sources])
        defaults = dict(max_nb_evaluation_per_policy=0, do_validation=0, nb_classes=0,
                        nb_iterations_per_policy_evaluation=params.get("nb_iterations", 1))
        tp = A.TrainParams(**{k: params[k] if k in params else defaults[k] for k in _TRAINER_FIELDS})
        h = C.c_void_p()
        st = L.tpgx_trainer_create(C.byref(cfg), len(ops), ops.ctypes.data, len(sources), C.cast(sd, C.c_void_p), C.byref(tp), int(seed), C.byref(h))
        if st != A.OK:
            raise TpgxError(st, "tpgx_trainer_create")
        self.h = h

    def close(self):
        if self.h:
            self.lib.tpgx_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != A.OK:
            raise TpgxError(st, (self.lib.tpgx_trainer_last_error(self.h) or b"").decode())

    def graph(self):
        v = A.Graph()
        self._check(self.lib.tpgx_trainer_graph(self.h, C.byref(v)))
        return view_to_graph(v, self.nb_constants)

    def populate(self, backend=None):
        s = A.TrainStats()
        self._check(self.lib.tpgx_trainer_populate(self.h, C.byref(backend) if backend is not None else None, C.byref(s)))
        return s.as_dict()

    def decimate(self, root_scores):
        sc = np.ascontiguousarray(root_scores, dtype=np.float64)
        s = A.TrainStats()
        self._check(self.lib.tpgx_trainer_decimate(self.h, sc.ctypes.data, C.byref(s)))
        return s.as_dict()

    def train_generation(self, backend, generation):
        s = A.TrainStats()
        self._check(self.lib.tpgx_trainer_train_generation(self.h, C.byref(backend), int(generation), C.byref(s)))
        return s.as_dict()

    def fingerprint(self):
        return int(self.lib.tpgx_trainer_fingerprint(self.h))

    def root_results(self):
        """[UP] resultsPerRoot of the current roots: (score[R] — NaN where not evaluated yet —, nb_evaluation[R])."""
        L = self.lib
        L.tpgx_trainer_root_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        n = L.tpgx_trainer_root_results(self.h, None, None, 0)
        sc, ne = np.zeros(n), np.zeros(n, dtype=np.uint64)
        L.tpgx_trainer_root_results(self.h, sc.ctypes.data, ne.ctypes.data, n)
        return sc, ne

    def best_root(self):
        """(index in the current root list, stored score) of [UP] getBestRoot()."""
        L = self.lib
        L.tpgx_trainer_best_root.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        i, sc = C.c_int32(-1), C.c_double(0.0)
        self._check(L.tpgx_trainer_best_root(self.h, C.byref(i), C.byref(sc)))
        return i.value, sc.value

    def keep_best_policy(self):
        self.lib.tpgx_trainer_keep_best_policy.argtypes = [C.c_void_p]
        self._check(self.lib.tpgx_trainer_keep_best_policy(self.h))


def episodes_backend(evaluator, env, nb_iterations, max_actions, mode=A.MODE_TRAINING, second_player_random=0, pendulum_actions=None):
    """The GPU backend (tpgx_train_backend_episodes) on an Evaluator's context."""
    lib = evaluator.lib
    pa = np.ascontiguousarray(pendulum_actions, dtype=np.float64) if pendulum_actions is not None else None
    req = A.EpisodeRequest(env=env, mode=mode, nb_iterations=nb_iterations, max_actions=max_actions, seeds=None, nb_jobs=0, roots_per_job=1,
                           job_roots=None, second_player_random=second_player_random, nb_pendulum_actions=0 if pa is None else len(pa),
                           pendulum_actions=pa.ctypes.data if pa is not None else None, trace_steps=0)
    be = A.TrainBackend()
    lib.tpgx_train_backend_episodes.argtypes = [C.c_void_p, C.POINTER(A.EpisodeRequest), C.POINTER(A.TrainBackend)]
    st = lib.tpgx_train_backend_episodes(evaluator.h, C.byref(req), C.byref(be))
    if st != A.OK:
        raise TpgxError(st, (lib.tpgx_last_error(evaluator.h) or b"").decode())
    be._keep = (pa, req, evaluator)
    return be


def classification_backend(evaluator, images, nb_classes=10, actions_per_evaluation=500):
    """The GPU backend (tpgx_train_backend_classification) on an Evaluator whose observations are `images` (uint8 [S, H*W])."""
    lib = evaluator.lib
    img = np.ascontiguousarray(images, dtype=np.uint8)
    be = A.TrainBackend()
    lib.tpgx_train_backend_classification.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(A.TrainBackend)]
    st = lib.tpgx_train_backend_classification(evaluator.h, img.ctypes.data, img.shape[0], int(nb_classes), int(actions_per_evaluation), C.byref(be))
    if st != A.OK:
        raise TpgxError(st, (lib.tpgx_last_error(evaluator.h) or b"").decode())
    be._keep = (img, evaluator)
    return be
