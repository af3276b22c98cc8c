"""ctypes mirror of include/tpgx.h (struct layouts shared by libtpgx.so and the test oracle)."""
import ctypes as C

OK = 0
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_GRAPH_INVALID, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_STATE = -1, -2, -3, -4, -5, -6, -7
STATUS_NAME = {0: "OK", -1: "BAD_ARG", -2: "UNSUPPORTED", -3: "GRAPH_INVALID", -4: "CUDA", -5: "NCCL", -6: "OOM", -7: "STATE"}

(OP_SUB, OP_ADD, OP_MUL, OP_DIV, OP_MAX, OP_EXP, OP_LOG, OP_COS, OP_SIN, OP_TAN, OP_PI, OP_MULC, OP_FMODG, OP_SUB_II,
 OP_CAST_I, OP_EQ0, OP_EQM1, OP_EQ1, OP_GE15, OP_COND, OP_SOBEL_MAGN, OP_SOBEL_DIR) = range(22)
OP_COUNT = 22
TIE_LAST_MAX, TIE_FIRST_MAX = 0, 1
ELEM_F64, ELEM_I32 = 0, 1
STORE_U8, STORE_F64, STORE_I32 = 0, 1, 2
ENV_GRIDWORLD, ENV_STICKGAME, ENV_TICTACTOE, ENV_PENDULUM = 0, 1, 2, 3
(PROBE_DIV, PROBE_DIV_K, PROBE_EXP, PROBE_EXP_K, PROBE_LOG, PROBE_LOG_K, PROBE_ATAN, PROBE_SQRT_NONNEG, PROBE_DIV_SMALL_INT,
 PROBE_SIN, PROBE_COS, PROBE_TAN, PROBE_DIV10, PROBE_DIV10_K) = range(14)
MODE_TRAINING, MODE_VALIDATION, MODE_TESTING = 0, 1, 2

# operand signature per opcode: D double, I int, C constant, W 3x3 window
OP_SIGNATURE = {
    OP_SUB: "DD", OP_ADD: "DD", OP_MUL: "DD", OP_DIV: "DD", OP_MAX: "DD", OP_EXP: "D", OP_LOG: "D", OP_COS: "D",
    OP_SIN: "D", OP_TAN: "D", OP_PI: "D", OP_MULC: "DC", OP_FMODG: "DD", OP_SUB_II: "II", OP_CAST_I: "I", OP_EQ0: "D",
    OP_EQM1: "D", OP_EQ1: "D", OP_GE15: "D", OP_COND: "DD", OP_SOBEL_MAGN: "W", OP_SOBEL_DIR: "W",
}
# algorithmic FP64 operations per executed line (SURVEY.md section 8d convention)
OP_FLOPS = {
    OP_SUB: 1, OP_ADD: 1, OP_MUL: 1, OP_DIV: 1, OP_MAX: 1, OP_EXP: 1, OP_LOG: 1, OP_COS: 1, OP_SIN: 1, OP_TAN: 1,
    OP_PI: 0, OP_MULC: 2, OP_FMODG: 2, OP_SUB_II: 1, OP_CAST_I: 0, OP_EQ0: 1, OP_EQM1: 1, OP_EQ1: 1, OP_GE15: 1,
    OP_COND: 2, OP_SOBEL_MAGN: 20, OP_SOBEL_DIR: 18,
}


class SourceDesc(C.Structure):
    _fields_ = [("elem", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("is_2d", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("nb_registers", C.c_int32), ("nb_constants", C.c_int32),
                ("tie_break", C.c_int32), ("reserved", C.c_int32 * 4)]


class Line(C.Structure):
    _fields_ = [("instruction", C.c_uint32), ("destination", C.c_uint32), ("src", C.c_uint32 * 2), ("loc", C.c_uint32 * 2)]


class Graph(C.Structure):
    _fields_ = [("nb_programs", C.c_int32), ("program_line_offset", C.c_void_p), ("lines", C.c_void_p),
                ("intron", C.c_void_p), ("constants", C.c_void_p), ("nb_teams", C.c_int32),
                ("team_edge_offset", C.c_void_p), ("edge_program", C.c_void_p), ("edge_destination", C.c_void_p),
                ("nb_roots", C.c_int32), ("root_team", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("evaluations", C.c_uint64), ("team_visits", C.c_uint64), ("program_executions", C.c_uint64),
                ("lines_executed", C.c_uint64), ("device_lines", C.c_uint64), ("device_program_executions", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class ClassifRequest(C.Structure):
    _fields_ = [("nb_samples", C.c_int64), ("sample_index", C.c_void_p), ("nb_classes", C.c_int32),
                ("root_begin", C.c_int32), ("root_end", C.c_int32), ("reserved", C.c_int32)]


class ClassifResult(C.Structure):
    _fields_ = [("score", C.c_void_p), ("class_f1", C.c_void_p), ("confusion", C.c_void_p), ("action", C.c_void_p),
                ("bid", C.c_void_p), ("counters", C.POINTER(Counters)), ("device_ms", C.c_void_p)]


class ArchiveRequest(C.Structure):
    _fields_ = [("archive_seed", C.c_void_p), ("probability", C.c_double), ("archive_size", C.c_int32), ("reserved", C.c_int32)]


class ArchiveResult(C.Structure):
    _fields_ = [("program", C.c_void_p), ("sample", C.c_void_p), ("bid", C.c_void_p), ("nb_records", C.c_int32), ("reserved", C.c_int32)]


class EpisodeArchiveResult(C.Structure):
    _fields_ = [("program", C.c_void_p), ("job", C.c_void_p), ("iteration", C.c_void_p), ("step", C.c_void_p), ("bid", C.c_void_p),
                ("observation", C.c_void_p), ("obs_len", C.c_int32), ("nb_records", C.c_int32)]


class EpisodeRequest(C.Structure):
    _fields_ = [("env", C.c_int32), ("mode", C.c_int32), ("nb_iterations", C.c_int32), ("max_actions", C.c_int32),
                ("seeds", C.c_void_p), ("nb_jobs", C.c_int32), ("roots_per_job", C.c_int32), ("job_roots", C.c_void_p),
                ("second_player_random", C.c_int32), ("nb_pendulum_actions", C.c_int32), ("pendulum_actions", C.c_void_p),
                ("trace_steps", C.c_int32), ("reserved", C.c_int32), ("rng_seeds", C.c_void_p)]


class EpisodeResult(C.Structure):
    _fields_ = [("score", C.c_void_p), ("episode_score", C.c_void_p), ("episode_actions", C.c_void_p),
                ("trace", C.c_void_p), ("final_state", C.c_void_p), ("counters", C.POINTER(Counters)),
                ("device_ms", C.c_void_p)]


class TrainParams(C.Structure):
    _fields_ = [("nb_actions", C.c_int32), ("nb_roots", C.c_int32), ("max_init_outgoing_edges", C.c_int32), ("max_outgoing_edges", C.c_int32),
                ("p_edge_deletion", C.c_double), ("p_edge_addition", C.c_double), ("p_program_mutation", C.c_double),
                ("p_edge_destination_change", C.c_double), ("p_edge_destination_is_action", C.c_double),
                ("force_program_behavior_change", C.c_int32), ("max_program_size", C.c_int32),
                ("p_delete", C.c_double), ("p_add", C.c_double), ("p_mutate", C.c_double), ("p_swap", C.c_double), ("p_constant_mutation", C.c_double),
                ("min_const_value", C.c_int32), ("max_const_value", C.c_int32), ("ratio_deleted_roots", C.c_double),
                ("archive_size", C.c_int32), ("reserved", C.c_int32), ("archiving_probability", C.c_double),
                ("max_nb_evaluation_per_policy", C.c_int32), ("do_validation", C.c_int32), ("nb_iterations_per_policy_evaluation", C.c_int32),
                ("nb_classes", C.c_int32)]


EVALUATE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(Graph), C.c_uint64, C.c_int32, C.POINTER(C.c_uint64), C.c_double, C.c_int32,
                          C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                          C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_int32))
EXECUTE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(Graph), C.c_int64, C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double))


class TrainBackend(C.Structure):
    _fields_ = [("user", C.c_void_p), ("evaluate", EVALUATE_FN), ("execute", EXECUTE_FN)]


class TrainStats(C.Structure):
    _fields_ = [("nb_teams", C.c_int32), ("nb_programs", C.c_int32), ("nb_roots", C.c_int32), ("nb_new_programs", C.c_int32),
                ("nb_removed_roots", C.c_int32), ("archive_records", C.c_int32), ("behaviour_retries", C.c_int32), ("reserved", C.c_int32),
                ("best_score", C.c_double), ("mean_score", C.c_double), ("worst_score", C.c_double),
                ("nb_evaluated_roots", C.c_int32), ("nb_skipped_roots", C.c_int32),
                ("best_validation_score", C.c_double), ("mean_validation_score", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
