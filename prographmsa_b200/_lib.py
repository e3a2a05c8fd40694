"""ctypes binding of libpgm_b200.so (include/pgm_b200.h).  There is no CPU fallback: if the
shared library is missing, or no B200 is present when a context is created, this raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGM_LIB_PATH") or os.path.join(_HERE, "libpgm_b200.so")   # PGM_LIB_PATH: A/B runs of two builds (tools/)
_lib = None


class PgmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libpgm_b200 error %d: %s" % (code, msg))
        self.code = code


class MLParams(C.Structure):
    _fields_ = [("n_states", C.c_int), ("Q", C.c_void_p), ("pi", C.c_void_p),
                ("mldist_flag", C.c_int), ("mldist_gap_flag", C.c_int),
                ("indel_rate", C.c_double), ("min_dist", C.c_double), ("max_dist", C.c_double),
                ("cutoff_dist", C.c_double), ("DIST_MAX", C.c_double), ("VAR_MAX", C.c_double), ("VAR_MIN", C.c_double)]


class Graph(C.Structure):
    _fields_ = [("N", C.c_int), ("sites", C.c_void_p), ("e_rowptr", C.c_void_p), ("e_col", C.c_void_p),
                ("e_val", C.c_void_p), ("r_rowptr", C.c_void_p), ("r_col", C.c_void_p), ("r_units", C.c_void_p)]


class DPModel(C.Structure):
    _fields_ = [("M", C.c_void_p), ("pi", C.c_void_p), ("distance", C.c_double), ("epsilon", C.c_double),
                ("delta", C.c_double)]


class DPOpts(C.Structure):
    _fields_ = [("indel_rate", C.c_double), ("end_indel_prob", C.c_double), ("repeat_rate", C.c_double),
                ("repeatext_prob", C.c_double), ("emission_mode", C.c_int)]


class MergeJob(C.Structure):
    _fields_ = [("N1", C.c_int), ("N2", C.c_int), ("sites1", C.c_void_p), ("sites2", C.c_void_p), ("P1", C.c_void_p), ("P2", C.c_void_p),
                ("len", C.c_int), ("rm1", C.c_void_p), ("rm2", C.c_void_p), ("matched", C.c_void_p), ("out", C.c_void_p)]


class Alignment(C.Structure):
    _fields_ = [("score", C.c_float), ("n_tr_indels", C.c_uint32), ("mapping1", C.c_void_p), ("mapping2", C.c_void_p),
                ("capacity", C.c_int32), ("length", C.c_int32), ("status", C.c_int32), ("cell_updates", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("nw_forward_ms", C.c_float), ("nw_walk_ms", C.c_float), ("ml_ms", C.c_float),
                ("emission_ms", C.c_float), ("graph_fill_ms", C.c_float), ("graph_trace_ms", C.c_float),
                ("nw_forward_launches", C.c_uint32), ("nw_walk_launches", C.c_uint32), ("ml_launches", C.c_uint32),
                ("emission_launches", C.c_uint32), ("graph_fill_launches", C.c_uint32), ("graph_trace_launches", C.c_uint32),
                ("nw_cells", C.c_uint64), ("ml_evaluations", C.c_uint64), ("graph_cell_updates", C.c_uint64),
                ("graph_cells", C.c_uint64), ("tb_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/pgm_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pgm_merge_profiles_batch", "pgm_create", "pgm_destroy", "pgm_sync", "pgm_last_error", "pgm_abi_version", "pgm_set_scratch_bytes", "pgm_stream",
    "pgm_enable_timing", "pgm_get_timing",
    "pgm_nw_ml_pair_range", "pgm_nw_ml_pair_range_dev", "pgm_nw_ml_all_pairs",
    "pgm_prealigned_ml_pair_range", "pgm_prealigned_ml_pair_range_dev", "pgm_prealigned_ml_all_pairs",
    "pgm_ml_distances", "pgm_angle_all_pairs",
    "pgm_align_graphs_batch", "pgm_graph_batch_create", "pgm_graph_batch_run", "pgm_graph_batch_fetch",
    "pgm_graph_batch_destroy", "pgm_dynprog_scores", "pgm_emission_scores",
    "pgm_host_clean_graph", "pgm_host_merge_plan", "pgm_host_merge_edges", "pgm_bionj_joins", "pgm_probe_fp64",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpgm_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C prographmsa_b200/csrc`); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.pgm_last_error.restype = C.c_char_p
        _lib.pgm_stream.restype = C.c_void_p
        _lib.pgm_destroy.restype = None
        _lib.pgm_graph_batch_destroy.restype = None
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        _lib.pgm_host_clean_graph.argtypes = [vp, C.c_double, vp, ip, vp, vp, vp, vp, vp, vp]
        _lib.pgm_host_merge_plan.argtypes = [vp, vp, C.c_int, vp, vp, vp, ip]
        _lib.pgm_host_merge_edges.argtypes = [vp, vp, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_double,
                                              vp, vp, vp, ip, vp, vp, vp, ip]
    return _lib


def check(rc):
    if rc != 0:
        raise PgmError(rc, lib().pgm_last_error().decode("utf-8", "replace"))


class Context:
    """One pgm_ctx (one GPU).  Not re-entrant."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().pgm_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().pgm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().pgm_sync(self._h))

    def set_scratch_bytes(self, b):
        check(lib().pgm_set_scratch_bytes(self._h, C.c_uint64(int(b))))

    def enable_timing(self, on=True):
        check(lib().pgm_enable_timing(self._h, int(on)))

    def timing(self):
        t = Timing()
        check(lib().pgm_get_timing(self._h, C.byref(t)))
        return t.as_dict()

    def stream(self):
        return lib().pgm_stream(self._h)


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
