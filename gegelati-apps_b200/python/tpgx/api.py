"""ctypes binding of libtpgx.so — the C ABI declared in include/tpgx.h.  Mirrors the reference's
agent-level calls: one Evaluator ~ one Learn::LearningAgent's evaluation engine.  There is no
CPU fallback: if the CUDA library is missing or no GPU is present this raises."""
import ctypes as C
import os
import numpy as np

from . import _cabi as A
from .graph import TPGraph

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TPGX_LIB") or os.path.normpath(os.path.join(_PKG, "..", "..", "lib", "libtpgx.so"))

EXPORTS = [
    "tpgx_create", "tpgx_destroy", "tpgx_last_error", "tpgx_abi_version", "tpgx_set_instruction_table", "tpgx_k1v2_encode_program", "tpgx_k1v2_handler_name", "tpgx_comm_unique_id", "tpgx_comm_init", "tpgx_comm_init_all", "tpgx_comm_shard",
    "tpgx_evaluate_classification_allgather", "tpgx_evaluate_graph_classification_allgather", "tpgx_evaluate_classification_allgather_device", "tpgx_evaluate_classification_allgather_all", "tpgx_kernel_launches",
    "tpgx_set_data_sources", "tpgx_set_graph", "tpgx_set_graph_sharded", "tpgx_set_observations", "tpgx_set_observations_pinned", "tpgx_set_observations_device",
    "tpgx_evaluate_classification", "tpgx_evaluate_classification_device", "tpgx_archive_classification", "tpgx_evaluate_episodes", "tpgx_archive_episodes",
    "tpgx_execute_programs", "tpgx_measure_fp64_peak", "tpgx_math_probe", "tpgx_mnist_episode_indices", "tpgx_flatten_programs", "tpgx_slice_graph",
    "tpgx_trainer_create", "tpgx_trainer_destroy", "tpgx_trainer_last_error", "tpgx_trainer_graph", "tpgx_trainer_train_generation",
    "tpgx_trainer_populate", "tpgx_trainer_decimate", "tpgx_trainer_fingerprint", "tpgx_trainer_root_results", "tpgx_trainer_best_root", "tpgx_trainer_keep_best_policy", "tpgx_train_backend_episodes", "tpgx_train_backend_classification", "tpgx_train_backend_release",
]

_lib = None


class TpgxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("tpgx %s: %s" % (A.STATUS_NAME.get(status, status), message))
        self.status = status


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libtpgx.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C gegelati-apps_b200` — there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.tpgx_last_error.restype = C.c_char_p
    lib.tpgx_last_error.argtypes = [C.c_void_p]
    lib.tpgx_create.argtypes = [C.POINTER(A.Config), C.POINTER(C.c_void_p)]
    lib.tpgx_destroy.argtypes = [C.c_void_p]
    lib.tpgx_set_instruction_table.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.tpgx_set_data_sources.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.tpgx_set_graph.argtypes = [C.c_void_p, C.POINTER(A.Graph)]
    lib.tpgx_set_observations.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    lib.tpgx_set_observations_device.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    lib.tpgx_set_observations_pinned.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    lib.tpgx_evaluate_classification.argtypes = [C.c_void_p, C.POINTER(A.ClassifRequest), C.POINTER(A.ClassifResult)]
    lib.tpgx_evaluate_classification_device.argtypes = [C.c_void_p, C.POINTER(A.ClassifRequest), C.c_void_p, C.c_void_p]
    lib.tpgx_archive_classification.argtypes = [C.c_void_p, C.POINTER(A.ClassifRequest), C.POINTER(A.ArchiveRequest), C.POINTER(A.ArchiveResult)]
    lib.tpgx_evaluate_episodes.argtypes = [C.c_void_p, C.POINTER(A.EpisodeRequest), C.POINTER(A.EpisodeResult)]
    lib.tpgx_execute_programs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tpgx_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tpgx_math_probe.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tpgx_mnist_episode_indices.argtypes = [C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data if a is not None else None


def flatten_programs(opcodes, sources, graph, nb_registers=8, nb_constants=0):
    """Host flattener alone (no GPU): returns dict(intron [lines] uint8, effective_lines [P] int32, code uint32[]).
    Raises TpgxError with the status tpgx_set_graph would return."""
    lib = load_library()
    lib.tpgx_flatten_programs.argtypes = [C.POINTER(A.Config), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(A.Graph), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
    cfg = A.Config(device=0, nb_registers=nb_registers, nb_constants=nb_constants, tie_break=A.TIE_LAST_MAX)
    ops = np.ascontiguousarray(opcodes, dtype=np.int32)
    sd = (A.SourceDesc * len(sources))(*[A.SourceDesc(*s) for s in sources])
    gs, keep = graph.c_struct()
    intron = np.zeros(len(graph.lines), dtype=np.uint8)
    eff = np.zeros(graph.nb_programs, dtype=np.int32)
    code = np.zeros(max(1, len(graph.lines)), dtype=np.uint32)
    n = C.c_int64(0)
    err = C.create_string_buffer(512)
    st = lib.tpgx_flatten_programs(C.byref(cfg), len(ops), ops.ctypes.data, len(sources), C.cast(sd, C.c_void_p), C.byref(gs), intron.ctypes.data,
                                   eff.ctypes.data, code.ctypes.data, len(code), C.byref(n), err, 512)
    if st != A.OK:
        raise TpgxError(st, err.value.decode())
    return dict(intron=intron, effective_lines=eff, code=code[:n.value])


def slice_graph(graph, root_begin, root_end):
    """tpgx_slice_graph (host only): the renumbered sub-graph roots [root_begin, root_end) reach, as the library sets up each
    pipeline slice of evaluate_graph_allgather.  Returns dict(team_ids, program_ids, team_edge_offset, edge_program,
    edge_destination, root_team); raises TpgxError on a malformed graph."""
    lib = load_library()
    I32 = C.POINTER(C.c_int32)
    lib.tpgx_slice_graph.argtypes = [C.POINTER(A.Graph), C.c_int32, C.c_int32, I32, I32, I32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    gs, keep = graph.c_struct()
    nt, ne, npg = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    st = lib.tpgx_slice_graph(C.byref(gs), root_begin, root_end, C.byref(nt), C.byref(ne), C.byref(npg), None, None, None, None, None, None)
    if st != A.OK:
        raise TpgxError(st, "tpgx_slice_graph")
    out = dict(team_ids=np.zeros(nt.value, dtype=np.int32), team_edge_offset=np.zeros(nt.value + 1, dtype=np.int32),
               edge_program=np.zeros(ne.value, dtype=np.int32), edge_destination=np.zeros(ne.value, dtype=np.int32),
               program_ids=np.zeros(npg.value, dtype=np.int32), root_team=np.zeros(root_end - root_begin, dtype=np.int32))
    st = lib.tpgx_slice_graph(C.byref(gs), root_begin, root_end, C.byref(nt), C.byref(ne), C.byref(npg), out["team_ids"].ctypes.data,
                              out["team_edge_offset"].ctypes.data, out["edge_program"].ctypes.data, out["edge_destination"].ctypes.data,
                              out["program_ids"].ctypes.data, out["root_team"].ctypes.data)
    if st != A.OK:
        raise TpgxError(st, "tpgx_slice_graph")
    return out


def k1v2_handler_names():
    """Handler names of the K1 v2 interpreter, indexed by handler id."""
    lib = load_library()
    lib.tpgx_k1v2_handler_name.restype = C.c_char_p
    lib.tpgx_k1v2_handler_name.argtypes = [C.c_int32]
    names = []
    while True:
        n = lib.tpgx_k1v2_handler_name(len(names))
        if n is None:
            return names
        names.append(n.decode())


def k1v2_encode_program(words, width=28, hw=784, nwin=676):
    """K1 v2 entry stream of one program (host only): dict(quads uint32[entries][4], slots) or None when the program is
    not encodable for K1 v2."""
    lib = load_library()
    lib.tpgx_k1v2_encode_program.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    w = np.ascontiguousarray(words, dtype=np.uint32)
    ne, ns = C.c_int32(0), C.c_int32(0)
    st = lib.tpgx_k1v2_encode_program(w.ctypes.data, len(w), width, hw, nwin, None, 0, C.byref(ne), C.byref(ns))
    if st == A.ERR_UNSUPPORTED:
        return None
    if st != A.OK:
        raise TpgxError(st, "tpgx_k1v2_encode_program")
    q = np.zeros((ne.value, 4), dtype=np.uint32)
    st = lib.tpgx_k1v2_encode_program(w.ctypes.data, len(w), width, hw, nwin, q.ctypes.data, len(q), C.byref(ne), C.byref(ns))
    if st != A.OK:
        raise TpgxError(st, "tpgx_k1v2_encode_program")
    return dict(quads=q, slots=ns.value)


def kernel_launches():
    """Kernels launched by libtpgx in this process so far."""
    lib = load_library()
    lib.tpgx_kernel_launches.restype = C.c_uint64
    return int(lib.tpgx_kernel_launches())


def comm_init_all(evaluators):
    """Single-process multi-GPU: one NCCL communicator over the evaluators' contexts (one per device)."""
    lib = load_library()
    arr = (C.c_void_p * len(evaluators))(*[e.h for e in evaluators])
    lib.tpgx_comm_init_all.argtypes = [C.c_void_p, C.c_int32]
    st = lib.tpgx_comm_init_all(C.cast(arr, C.c_void_p), len(evaluators))
    if st != A.OK:
        raise TpgxError(st, (lib.tpgx_last_error(evaluators[0].h) or b"").decode())


def evaluate_classification_allgather_all(evaluators, total_roots, nb_samples, nb_classes=10):
    """The single-process collective over all contexts; returns dict(score, class_f1) of all roots."""
    lib = load_library()
    arr = (C.c_void_p * len(evaluators))(*[e.h for e in evaluators])
    req = A.ClassifRequest(nb_samples=int(nb_samples), sample_index=None, nb_classes=nb_classes, root_begin=0, root_end=0)
    rec = np.zeros((int(total_roots), nb_classes + 1))
    lib.tpgx_evaluate_classification_allgather_all.argtypes = [C.c_void_p, C.c_int32, C.POINTER(A.ClassifRequest), C.c_int32, C.c_void_p]
    st = lib.tpgx_evaluate_classification_allgather_all(C.cast(arr, C.c_void_p), len(evaluators), C.byref(req), int(total_roots), rec.ctypes.data)
    if st != A.OK:
        raise TpgxError(st, (lib.tpgx_last_error(evaluators[0].h) or b"").decode())
    return dict(score=rec[:, 0].copy(), class_f1=rec[:, 1:].copy())


def mnist_episode_indices(seed, mode, dataset_size, nb_actions):
    """sample_index list of one MNIST evaluation ([REF] mnist/src/mnist.cpp:8-34,58-69); host-only."""
    out = np.zeros(int(nb_actions), dtype=np.int32)
    st = load_library().tpgx_mnist_episode_indices(int(seed), int(mode), int(dataset_size), int(nb_actions), out.ctypes.data)
    if st != A.OK:
        raise TpgxError(st, "tpgx_mnist_episode_indices: bad argument")
    return out


def load_idx(path):
    """IDX file (the MNIST distribution format read by [REF] mnist/src/mnist_reader/mnist_reader_common.hpp:24-78):
    big-endian magic 0x0000 08 <ndim>, ndim big-endian uint32 extents, then uint8 data.  Returns a uint8 array;
    images come back as [count, rows*cols] — the layout tpgx_set_observations takes."""
    import gzip
    import struct
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 4 or raw[0] != 0 or raw[1] != 0 or raw[2] != 0x08:
        raise ValueError("%s: not an unsigned-byte IDX file" % path)
    ndim = raw[3]
    dims = struct.unpack(">%dI" % ndim, raw[4:4 + 4 * ndim])
    data = np.frombuffer(raw, dtype=np.uint8, offset=4 + 4 * ndim)
    if data.size != int(np.prod(dims)):
        raise ValueError("%s: %d data bytes, header says %s" % (path, data.size, dims))
    return data.reshape(dims[0], -1).copy() if ndim > 1 else data.copy()


class Evaluator:
    """Owns one tpgx_ctx (one GPU)."""

    def __init__(self, opcodes, sources, nb_registers=8, nb_constants=0, tie_break=A.TIE_LAST_MAX, device=0):
        self.lib = load_library()
        cfg = A.Config(device=device, nb_registers=nb_registers, nb_constants=nb_constants, tie_break=tie_break)
        h = C.c_void_p()
        st = self.lib.tpgx_create(C.byref(cfg), C.byref(h))
        if st != A.OK:
            raise TpgxError(st, (self.lib.tpgx_last_error(None) or b"").decode())
        self.h = h
        self.nb_constants = nb_constants
        self.sources = list(sources)
        ops = np.ascontiguousarray(opcodes, dtype=np.int32)
        self._check(self.lib.tpgx_set_instruction_table(self.h, len(ops), ops.ctypes.data))
        sd = (A.SourceDesc * len(sources))(*[A.SourceDesc(*s) for s in sources])
        self._check(self.lib.tpgx_set_data_sources(self.h, len(sources), C.cast(sd, C.c_void_p)))
        self.graph = None
        self.nb_obs = 0

    def _check(self, st):
        if st != A.OK:
            raise TpgxError(st, (self.lib.tpgx_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.tpgx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_graph(self, g: TPGraph):
        cs, keep = g.c_struct()
        self._check(self.lib.tpgx_set_graph(self.h, C.byref(cs)))
        self.graph = g

    def set_graph_sharded(self, g: TPGraph):
        """tpgx_set_graph_sharded: only what this rank's root shard reaches is flattened and uploaded."""
        self.lib.tpgx_set_graph_sharded.argtypes = [C.c_void_p, C.POINTER(A.Graph)]
        cs, self._graph_keep = g.c_struct()
        self._check(self.lib.tpgx_set_graph_sharded(self.h, C.byref(cs)))
        self.graph = g

    def set_observations(self, images_u8, labels=None):
        img = np.ascontiguousarray(images_u8, dtype=np.uint8)
        lab = np.ascontiguousarray(labels, dtype=np.uint8) if labels is not None else None
        self._check(self.lib.tpgx_set_observations(self.h, 0, A.STORE_U8, img.ctypes.data, _ptr(lab), img.shape[0]))
        self.nb_obs = img.shape[0]

    def set_observations_pinned(self, images_u8, labels=None):
        """Page-locked host arrays (e.g. torch pin_memory().numpy()): copies are only enqueued; the arrays must stay alive and
        unchanged until the next evaluate_* returns."""
        assert images_u8.dtype == np.uint8 and images_u8.flags.c_contiguous
        self._pinned_keepalive = (images_u8, labels)
        self._check(self.lib.tpgx_set_observations_pinned(self.h, 0, A.STORE_U8, images_u8.ctypes.data, _ptr(labels), images_u8.shape[0]))
        self.nb_obs = images_u8.shape[0]

    def set_observations_device(self, data_ptr, labels_ptr, count):
        self._check(self.lib.tpgx_set_observations_device(self.h, 0, A.STORE_U8, data_ptr, labels_ptr, count))
        self.nb_obs = count

    def evaluate_classification(self, nb_samples=None, sample_index=None, nb_classes=10, root_begin=0, root_end=None,
                                want=("score", "class_f1", "confusion", "counters", "device_ms")):
        g = self.graph
        root_end = g.nb_roots if root_end is None else root_end
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else self.nb_obs))
        nR = root_end - root_begin
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=nb_classes, root_begin=root_begin, root_end=root_end)
        out = {}
        res = A.ClassifResult()
        cnt = A.Counters()
        if "score" in want: out["score"] = np.zeros(nR); res.score = out["score"].ctypes.data
        if "class_f1" in want: out["class_f1"] = np.zeros((nR, nb_classes)); res.class_f1 = out["class_f1"].ctypes.data
        if "confusion" in want: out["confusion"] = np.zeros((nR, nb_classes, nb_classes), dtype=np.uint64); res.confusion = out["confusion"].ctypes.data
        if "action" in want: out["action"] = np.zeros((nR, nS), dtype=np.uint8); res.action = out["action"].ctypes.data
        if "bid" in want: out["bid"] = np.zeros((g.nb_programs, nS)); res.bid = out["bid"].ctypes.data
        if "counters" in want: res.counters = C.pointer(cnt)
        if "device_ms" in want: out["device_ms"] = np.zeros(3, dtype=np.float32); res.device_ms = out["device_ms"].ctypes.data
        self._check(self.lib.tpgx_evaluate_classification(self.h, C.byref(req), C.byref(res)))
        if "counters" in want: out["counters"] = cnt.as_dict()
        return out

    def evaluate_classification_device(self, records_ptr, stream_ptr, nb_samples, nb_classes=10, root_begin=0, root_end=None):
        root_end = self.graph.nb_roots if root_end is None else root_end
        req = A.ClassifRequest(nb_samples=int(nb_samples), sample_index=None, nb_classes=nb_classes, root_begin=root_begin, root_end=root_end)
        self._check(self.lib.tpgx_evaluate_classification_device(self.h, C.byref(req), records_ptr, stream_ptr))

    # ---- multi-GPU: the library's own NCCL communicator (include/tpgx.h, "multi-GPU") ----
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (rank 0 creates it, the caller hands it to every rank)."""
        lib = load_library()
        lib.tpgx_comm_unique_id.argtypes = [C.c_void_p]
        buf = np.zeros(128, dtype=np.uint8)
        st = lib.tpgx_comm_unique_id(buf.ctypes.data)
        if st != A.OK:
            raise TpgxError(st, (lib.tpgx_last_error(None) or b"").decode())
        return buf

    def comm_init(self, unique_id, rank, world):
        self.lib.tpgx_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._check(self.lib.tpgx_comm_init(self.h, uid.ctypes.data, int(rank), int(world)))

    def comm_shard(self, total_roots):
        """(rank, world, root_begin, root_end) of this context for a graph of total_roots roots."""
        self.lib.tpgx_comm_shard.argtypes = [C.c_void_p, C.c_int32] + [C.POINTER(C.c_int32)] * 4
        v = [C.c_int32(0) for _ in range(4)]
        self._check(self.lib.tpgx_comm_shard(self.h, int(total_roots), *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def evaluate_classification_allgather(self, total_roots, nb_samples=None, sample_index=None, nb_classes=10):
        """Collective: this rank's shard evaluated, all ranks' records gathered by the library's NCCL all-gather.
        Returns dict(score [total_roots], class_f1 [total_roots, C])."""
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else self.nb_obs))
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=nb_classes, root_begin=0, root_end=0)
        rec = np.zeros((int(total_roots), nb_classes + 1))
        self.lib.tpgx_evaluate_classification_allgather.argtypes = [C.c_void_p, C.POINTER(A.ClassifRequest), C.c_int32, C.c_void_p]
        self._check(self.lib.tpgx_evaluate_classification_allgather(self.h, C.byref(req), int(total_roots), rec.ctypes.data))
        return dict(score=rec[:, 0].copy(), class_f1=rec[:, 1:].copy())

    def evaluate_graph_allgather(self, graph, total_roots=None, nb_samples=None, sample_index=None, nb_classes=10, records_out=None):
        """One generation in one call (tpgx_evaluate_graph_classification_allgather): the graph from host memory, this rank's
        shard evaluated with the host side pipelined behind the device work, all ranks' records gathered.
        Returns dict(score [total_roots], class_f1 [total_roots, C]); with records_out (a C-contiguous float64 array
        [total_roots, 1 + C], e.g. page-locked) the records land there and the result holds views of it."""
        total_roots = graph.nb_roots if total_roots is None else int(total_roots)
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else self.nb_obs))
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=nb_classes, root_begin=0, root_end=0)
        if records_out is None:
            rec = np.zeros((total_roots, nb_classes + 1))
        else:
            rec = records_out
            assert rec.dtype == np.float64 and rec.shape == (total_roots, nb_classes + 1) and rec.flags["C_CONTIGUOUS"]
        gs, keep = graph.c_struct()
        self.lib.tpgx_evaluate_graph_classification_allgather.argtypes = [C.c_void_p, C.POINTER(A.Graph), C.POINTER(A.ClassifRequest), C.c_int32, C.c_void_p]
        self._check(self.lib.tpgx_evaluate_graph_classification_allgather(self.h, C.byref(gs), C.byref(req), total_roots, rec.ctypes.data))
        del keep
        if records_out is not None:
            return dict(score=rec[:, 0], class_f1=rec[:, 1:])
        return dict(score=rec[:, 0].copy(), class_f1=rec[:, 1:].copy())

    def evaluate_classification_allgather_device(self, records_all_ptr, stream_ptr, total_roots, nb_samples, nb_classes=10):
        req = A.ClassifRequest(nb_samples=int(nb_samples), sample_index=None, nb_classes=nb_classes, root_begin=0, root_end=0)
        self.lib.tpgx_evaluate_classification_allgather_device.argtypes = [C.c_void_p, C.POINTER(A.ClassifRequest), C.c_int32, C.c_void_p, C.c_void_p]
        self._check(self.lib.tpgx_evaluate_classification_allgather_device(self.h, C.byref(req), int(total_roots), int(records_all_ptr), stream_ptr))

    def archive_classification(self, archive_seeds, probability, archive_size, nb_samples=None, sample_index=None, root_begin=0, root_end=None):
        """Archive left by a TRAINING evaluateAllRoots ([UP] Archive::addRecording): dict(program, sample, bid) in archive order."""
        root_end = self.graph.nb_roots if root_end is None else root_end
        idx = np.ascontiguousarray(sample_index, dtype=np.int32) if sample_index is not None else None
        nS = int(nb_samples if nb_samples is not None else (len(idx) if idx is not None else self.nb_obs))
        req = A.ClassifRequest(nb_samples=nS, sample_index=_ptr(idx), nb_classes=10, root_begin=root_begin, root_end=root_end)
        seeds = np.ascontiguousarray(archive_seeds, dtype=np.uint64)
        assert len(seeds) == root_end - root_begin
        prog = np.zeros(max(1, archive_size), dtype=np.int32); samp = np.zeros(max(1, archive_size), dtype=np.int64); bid = np.zeros(max(1, archive_size))
        ar = A.ArchiveRequest(archive_seed=seeds.ctypes.data, probability=float(probability), archive_size=int(archive_size))
        res = A.ArchiveResult(program=prog.ctypes.data, sample=samp.ctypes.data, bid=bid.ctypes.data)
        self._check(self.lib.tpgx_archive_classification(self.h, C.byref(req), C.byref(ar), C.byref(res)))
        n = res.nb_records
        return dict(program=prog[:n], sample=samp[:n], bid=bid[:n])

    def execute_programs(self, count, obs):
        """obs: list per data source of arrays [count, space] (float64 or int32). Returns bid [P, count]."""
        n = len(self.sources)
        f = (C.c_void_p * n)(); iv = (C.c_void_p * n)(); keep = []
        for k, (elem, h, w, _) in enumerate(self.sources):
            a = np.ascontiguousarray(obs[k], dtype=np.float64 if elem == A.ELEM_F64 else np.int32).reshape(count, h * w)
            keep.append(a)
            if elem == A.ELEM_F64: f[k] = a.ctypes.data
            else: iv[k] = a.ctypes.data
        bid = np.zeros((self.graph.nb_programs, count))
        self._check(self.lib.tpgx_execute_programs(self.h, count, C.cast(f, C.c_void_p), C.cast(iv, C.c_void_p), bid.ctypes.data))
        return bid

    def evaluate_episodes(self, env, seeds, job_roots, max_actions, mode=A.MODE_TRAINING, roots_per_job=1,
                          second_player_random=0, pendulum_actions=None, trace_steps=0, rng_seeds=None):
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
                   episode_actions=np.zeros((J, NI), dtype=np.int32), final_state=np.zeros((J, NI, 4)),
                   device_ms=np.zeros(1, dtype=np.float32))
        cnt = A.Counters()
        res = A.EpisodeResult(score=out["score"].ctypes.data, episode_score=out["episode_score"].ctypes.data,
                              episode_actions=out["episode_actions"].ctypes.data, final_state=out["final_state"].ctypes.data,
                              counters=C.pointer(cnt), device_ms=out["device_ms"].ctypes.data)
        if trace_steps > 0:
            out["trace"] = np.zeros((J, NI, trace_steps), dtype=np.int8)
            res.trace = out["trace"].ctypes.data
        self._check(self.lib.tpgx_evaluate_episodes(self.h, C.byref(req), C.byref(res)))
        out["counters"] = cnt.as_dict()
        return out

    def archive_episodes(self, env, seeds, job_roots, max_actions, archive_seeds, probability, archive_size, obs_len=0,
                         mode=A.MODE_TRAINING, roots_per_job=1, second_player_random=0, pendulum_actions=None):
        """Archive left by a TRAINING evaluateAllRoots of a sequential environment ([UP] Archive::addRecording):
        dict(program, job, iteration, step, bid[, observation]) in archive order, plus the evaluation's scores."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        jr = np.ascontiguousarray(job_roots, dtype=np.int32).reshape(-1, roots_per_job)
        J, NI = jr.shape[0], len(seeds)
        pa = np.ascontiguousarray(pendulum_actions, dtype=np.float64) if pendulum_actions is not None else None
        req = A.EpisodeRequest(env=env, mode=mode, nb_iterations=NI, max_actions=max_actions, seeds=seeds.ctypes.data, nb_jobs=J,
                               roots_per_job=roots_per_job, job_roots=jr.ctypes.data, second_player_random=second_player_random,
                               nb_pendulum_actions=0 if pa is None else len(pa), pendulum_actions=_ptr(pa), trace_steps=0)
        aseeds = np.ascontiguousarray(archive_seeds, dtype=np.uint64)
        assert len(aseeds) == J
        n = max(1, int(archive_size))
        o = dict(program=np.zeros(n, dtype=np.int32), job=np.zeros(n, dtype=np.int32), iteration=np.zeros(n, dtype=np.int32),
                 step=np.zeros(n, dtype=np.int32), bid=np.zeros(n))
        if obs_len > 0:
            o["observation"] = np.zeros((n, obs_len))
        score = np.zeros((J, roots_per_job))
        ev = A.EpisodeResult(score=score.ctypes.data)
        ar = A.ArchiveRequest(archive_seed=aseeds.ctypes.data, probability=float(probability), archive_size=int(archive_size))
        res = A.EpisodeArchiveResult(program=o["program"].ctypes.data, job=o["job"].ctypes.data, iteration=o["iteration"].ctypes.data,
                                     step=o["step"].ctypes.data, bid=o["bid"].ctypes.data,
                                     observation=o["observation"].ctypes.data if obs_len > 0 else None, obs_len=int(obs_len))
        self._check(self.lib.tpgx_archive_episodes(self.h, C.byref(req), C.byref(ar), C.byref(ev), C.byref(res)))
        out = {k: v[:res.nb_records] for k, v in o.items()}
        out["score"] = score
        return out

    def math_probe(self, fn, a, b=None):
        """Device arithmetic primitive `fn` (PROBE_*) elementwise — diagnostic for parity tests."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
        out = np.empty_like(a)
        self._check(self.lib.tpgx_math_probe(self.h, fn, a.size, a.ctypes.data, _ptr(b), out.ctypes.data))
        return out

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.tpgx_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value
