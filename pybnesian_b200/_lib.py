"""ctypes binding of libpbn_cuda.so (C ABI in include/pbn_cuda.h).

There is no CPU fallback: if the CUDA library is missing or no sm_100 GPU is usable,
the first compute call raises (RuntimeError), it never silently computes elsewhere.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBN_CUDA_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PBN_CUDA_LIB") or os.path.join(_HERE, "libpbn_cuda.so")

PBN_OK, PBN_ERR_CUDA, PBN_ERR_ARG, PBN_ERR_SINGULAR, PBN_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
PBN_F64, PBN_F32 = 0, 1
BW_NORMAL_REFERENCE, BW_SCOTT = 0, 1

EXPORTS = [
    "pbn_last_error", "pbn_version", "pbn_device_count", "pbn_ctx_create", "pbn_ctx_create_multi", "pbn_ctx_num_devices",
    "pbn_ctx_device", "pbn_ctx_warmup", "pbn_cv_score_jobs", "pbn_ctx_set_skipping", "pbn_ctx_skip_stats", "pbn_ctx_destroy", "pbn_ctx_set_stream",
    "pbn_ctx_stream", "pbn_ctx_synchronize", "pbn_ctx_sm_count", "pbn_ctx_counters", "pbn_table_upload",
    "pbn_table_free", "pbn_table_rows", "pbn_table_cols", "pbn_table_download", "pbn_table_moments", "pbn_bandwidth",
    "pbn_diag_bandwidth", "pbn_kde_fit", "pbn_ckde_fit", "pbn_product_kde_fit", "pbn_kde_free", "pbn_kde_num_instances", "pbn_kde_lognorm",
    "pbn_kde_logl", "pbn_kde_logl_device", "pbn_ctx_last_fallback_rows", "pbn_ctx_last_row_kernel_rows", "pbn_device_alloc", "pbn_device_free",
    "pbn_device_read", "pbn_ctx_set_timing", "pbn_ctx_pair_kernel_time", "pbn_ucv_create", "pbn_ucv_free",
    "pbn_ucv_score", "pbn_ucv_pair_sums", "pbn_ucv_pairs", "pbn_ucv_bandwidth",
    "pbn_ucv_score_from_sums", "pbn_lg_fit", "pbn_lg_logl", "pbn_cv_split", "pbn_holdout_split", "pbn_cv_create", "pbn_cv_free", "pbn_cv_table",
    "pbn_cv_folds", "pbn_cv_train_moments", "pbn_cv_scores", "pbn_sort_desc", "pbn_intset_new", "pbn_intset_clone",
    "pbn_intset_free", "pbn_intset_insert", "pbn_intset_erase", "pbn_intset_clear", "pbn_intset_contains",
    "pbn_intset_size", "pbn_intset_list", "pbn_discrete_slices", "pbn_table_take", "pbn_kde_logl_multi",
    "pbn_ckde_cdf", "pbn_ckde_sample_indices", "pbn_ckde_sample", "pbn_lg_sample", "pbn_uniform_real",
]
PBN_MAX_DIM = 32
FACTOR_CKDE, FACTOR_LINEAR_GAUSSIAN = 0, 1


class SingularCovarianceData(ValueError):
    """Mirror of pybnesian.SingularCovarianceData (pybindings_kde.cpp:114), a ValueError."""


class Rows(ctypes.Structure):
    _fields_ = [("b0", ctypes.c_int64), ("e0", ctypes.c_int64), ("b1", ctypes.c_int64), ("e1", ctypes.c_int64)]

    @staticmethod
    def single(begin, end):
        return Rows(begin, end, 0, 0)


class CVItem(ctypes.Structure):
    """pbn_cv_item."""
    _fields_ = [("factor", ctypes.c_int), ("rule", ctypes.c_int), ("n_vars", ctypes.c_int),
                ("vars", ctypes.c_int * PBN_MAX_DIM)]


_lib = None
_lock = threading.Lock()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pybnesian_b200: CUDA extension %s is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C pybnesian_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, ci, dp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.pbn_last_error.restype = ctypes.c_char_p
        L.pbn_version.restype = ctypes.c_char_p
        L.pbn_device_count.argtypes = [ip]
        L.pbn_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
        L.pbn_ctx_create_multi.argtypes = [ip, ci, ctypes.POINTER(vp)]
        L.pbn_ctx_num_devices.argtypes = [vp]
        L.pbn_ctx_warmup.argtypes = [vp]
        L.pbn_ctx_device.argtypes = [vp, ci]
        L.pbn_ctx_destroy.argtypes = [vp]
        L.pbn_ctx_set_stream.argtypes = [vp, vp]
        L.pbn_ctx_stream.argtypes = [vp]
        L.pbn_ctx_stream.restype = vp
        L.pbn_ctx_synchronize.argtypes = [vp]
        L.pbn_ctx_sm_count.argtypes = [vp]
        L.pbn_ctx_counters.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.pbn_table_upload.argtypes = [vp, ctypes.POINTER(vp), ci, i64, ci, ctypes.POINTER(vp)]
        L.pbn_table_free.argtypes = [vp]
        L.pbn_table_rows.argtypes = [vp]
        L.pbn_table_rows.restype = i64
        L.pbn_table_cols.argtypes = [vp]
        L.pbn_table_download.argtypes = [vp, vp, ci, Rows, vp]
        L.pbn_table_moments.argtypes = [vp, vp, ip, ci, Rows, dp, dp]
        L.pbn_bandwidth.argtypes = [vp, vp, ip, ci, Rows, ci, dp]
        L.pbn_diag_bandwidth.argtypes = [vp, vp, ip, ci, Rows, ci, dp]
        L.pbn_kde_fit.argtypes = [vp, vp, ip, ci, Rows, dp, ctypes.POINTER(vp)]
        L.pbn_ckde_fit.argtypes = [vp, vp, ip, ci, Rows, dp, ctypes.POINTER(vp)]
        L.pbn_product_kde_fit.argtypes = [vp, vp, ip, ci, Rows, dp, ctypes.POINTER(vp)]
        L.pbn_kde_free.argtypes = [vp]
        L.pbn_kde_num_instances.argtypes = [vp]
        L.pbn_kde_num_instances.restype = i64
        L.pbn_kde_lognorm.argtypes = [vp]
        L.pbn_kde_lognorm.restype = ctypes.c_double
        L.pbn_kde_logl.argtypes = [vp, vp, vp, ip, Rows, dp, dp]
        L.pbn_kde_logl_device.argtypes = [vp, vp, vp, ip, Rows, vp, vp]
        L.pbn_ctx_last_fallback_rows.argtypes = [vp, ctypes.POINTER(i64)]
        L.pbn_ctx_last_row_kernel_rows.argtypes = [vp, ctypes.POINTER(i64)]
        L.pbn_ctx_set_skipping.argtypes = [vp, ci]
        L.pbn_ctx_skip_stats.argtypes = [vp] + [ctypes.POINTER(i64)] * 4 + [ci]
        L.pbn_device_alloc.argtypes = [vp, i64, ctypes.POINTER(vp)]
        L.pbn_device_free.argtypes = [vp, vp]
        L.pbn_device_read.argtypes = [vp, vp, i64, vp]
        L.pbn_ucv_create.argtypes = [vp, vp, ip, ci, Rows, ctypes.POINTER(vp)]
        L.pbn_ucv_free.argtypes = [vp]
        L.pbn_ucv_score.argtypes = [vp, dp, ci, dp]
        L.pbn_ucv_pair_sums.argtypes = [vp, dp, ci, ci, ci, dp, dp]
        L.pbn_ucv_score_from_sums.argtypes = [vp, dp, ci, ctypes.c_double, ctypes.c_double, dp]
        L.pbn_ucv_pairs.argtypes = [vp]
        L.pbn_ucv_pairs.restype = i64
        L.pbn_ucv_bandwidth.argtypes = [vp, vp, ip, ci, Rows, ci, dp, ip]
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.pbn_lg_fit.argtypes = [vp, vp, ip, ci, Rows, dp, dp]
        L.pbn_lg_logl.argtypes = [vp, vp, ip, ci, Rows, dp, ctypes.c_double, dp, dp]
        L.pbn_cv_split.argtypes = [i32p, i64, ci, ctypes.c_uint32, i32p]
        L.pbn_holdout_split.argtypes = [i32p, i64, ctypes.c_double, ctypes.c_uint32, i32p]
        L.pbn_cv_create.argtypes = [vp, vp, i32p, i64, i32p, ci, ctypes.POINTER(vp)]
        L.pbn_cv_free.argtypes = [vp]
        L.pbn_cv_table.argtypes = [vp]
        L.pbn_cv_table.restype = vp
        L.pbn_cv_folds.argtypes = [vp]
        L.pbn_cv_train_moments.argtypes = [vp, ci, ip, ci, dp, dp]
        L.pbn_cv_scores.argtypes = [vp, vp, ctypes.POINTER(CVItem), ci, ci, ci, dp, ip]
        L.pbn_cv_score_jobs.argtypes = [vp, vp, ctypes.POINTER(CVItem), ci, i32p, i32p, ci, dp, ip]
        L.pbn_sort_desc.argtypes = [i32p, i64, dp]
        L.pbn_intset_new.argtypes = [ctypes.POINTER(vp)]
        L.pbn_intset_clone.argtypes = [vp, ctypes.POINTER(vp)]
        L.pbn_intset_free.argtypes = [vp]
        L.pbn_intset_insert.argtypes = [vp, ci]
        L.pbn_intset_erase.argtypes = [vp, ci]
        L.pbn_intset_clear.argtypes = [vp]
        L.pbn_intset_contains.argtypes = [vp, ci]
        L.pbn_intset_size.argtypes = [vp]
        L.pbn_intset_list.argtypes = [vp, ip]
        L.pbn_discrete_slices.argtypes = [ctypes.POINTER(i32p), i32p, ci, i64, ctypes.POINTER(ctypes.c_uint8), ci, i32p,
                                          ctypes.POINTER(i64)]
        L.pbn_table_take.argtypes = [vp, vp, i32p, i64, ctypes.POINTER(vp)]
        L.pbn_kde_logl_multi.argtypes = [vp, ctypes.POINTER(vp), ci, vp, ip, ctypes.POINTER(Rows), dp, dp]
        L.pbn_ckde_cdf.argtypes = [vp, vp, vp, ip, Rows, dp]
        L.pbn_ckde_sample_indices.argtypes = [vp, vp, vp, ip, Rows, vp, i32p]
        L.pbn_ckde_sample.argtypes = [vp, vp, dp, vp, ip, Rows, vp, ip, ctypes.POINTER(vp), i64, ctypes.c_uint32, vp, i32p]
        L.pbn_uniform_real.argtypes = [i64, ctypes.c_uint32, ci, vp]
        L.pbn_lg_sample.argtypes = [dp, ctypes.c_double, ci, ctypes.POINTER(vp), ci, i64, ctypes.c_uint32, dp]
        L.pbn_ctx_set_timing.argtypes = [vp, ci]
        L.pbn_ctx_pair_kernel_time.argtypes = [vp, dp, ctypes.POINTER(i64), ctypes.POINTER(i64), ci]
        _lib = L
    return _lib


def raise_for(rc, msg):
    if rc == PBN_ERR_SINGULAR:
        raise SingularCovarianceData(msg)
    if rc in (PBN_ERR_ARG, PBN_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)


def check(rc):
    if rc == PBN_OK:
        return
    msg = lib().pbn_last_error().decode("utf-8", "replace")
    if rc == PBN_ERR_SINGULAR:
        raise SingularCovarianceData(msg)
    if rc in (PBN_ERR_ARG, PBN_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)


class Context:
    """One GPU (pbn_ctx) or, with a list of devices, one context over several GPUs of this process
    (pbn_ctx_create_multi): tables and fitted models are then replicated and logl / cdf / CV scores / UCV shard over
    all of them without torchrun.

    The default context is chosen from the environment:
      PBN_CUDA_DEVICES=0,1,2 | all   the listed (all visible) devices in ONE context
      PBN_CUDA_DEVICE=k              device k only (what bench.py and every torchrun rank set)
      LOCAL_RANK=k (torchrun)        device k only: one process per GPU, pybnesian_b200.parallel sums the scalars
      nothing set                    every visible device in one context
    """

    def __init__(self, device=0):
        self.handle = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = [int(d) for d in device]
            check(lib().pbn_ctx_create_multi(int_array(devs), len(devs), ctypes.byref(self.handle)))
            self.devices = devs
        else:
            check(lib().pbn_ctx_create(int(device), ctypes.byref(self.handle)))
            self.devices = [int(device)]
        self.device = self.devices[0]
        # load the kernel modules in the background (ctypes releases the GIL): the first fit / logl call then finds them ready
        if os.environ.get("PBN_CUDA_WARMUP", "1") != "0":
            threading.Thread(target=lib().pbn_ctx_warmup, args=(self.handle,), daemon=True).start()

    def warmup(self):
        """Blocks until the kernel modules are loaded on every device of the context."""
        check(lib().pbn_ctx_warmup(self.handle))

    @property
    def num_devices(self):
        return len(self.devices)

    def synchronize(self):
        check(lib().pbn_ctx_synchronize(self.handle))

    def set_stream(self, cuda_stream_ptr):
        check(lib().pbn_ctx_set_stream(self.handle, ctypes.c_void_p(cuda_stream_ptr or 0)))

    @property
    def stream(self):
        return lib().pbn_ctx_stream(self.handle)

    @property
    def sm_count(self):
        return lib().pbn_ctx_sm_count(self.handle)

    def counters(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().pbn_ctx_counters(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def set_timing(self, on):
        check(lib().pbn_ctx_set_timing(self.handle, 1 if on else 0))

    def pair_kernel_time(self, reset=False):
        """(total ms, launches, pair evaluations) of the timed pair-kernel launches."""
        ms, n, pe = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().pbn_ctx_pair_kernel_time(self.handle, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(pe),
                                             1 if reset else 0))
        return ms.value, n.value, pe.value

    def last_fallback_rows(self):
        v = ctypes.c_int64()
        check(lib().pbn_ctx_last_fallback_rows(self.handle, ctypes.byref(v)))
        return v.value

    def set_skipping(self, on):
        """Tile skipping on (default) / off = every (train, test) pair evaluated (include/pbn_cuda.h)."""
        check(lib().pbn_ctx_set_skipping(self.handle, 1 if on else 0))

    def skip_stats(self, reset=False):
        """(test tile x train tile) units: of the last call (total, evaluated) and of the timed launches."""
        a, b, c, d = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().pbn_ctx_skip_stats(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d),
                                       1 if reset else 0))
        return {"last_total": a.value, "last_evaluated": b.value, "timed_total": c.value, "timed_evaluated": d.value}

    def last_row_kernel_rows(self):
        v = ctypes.c_int64()
        check(lib().pbn_ctx_last_row_kernel_rows(self.handle, ctypes.byref(v)))
        return v.value


_default_ctx = None


def _default_devices():
    env = os.environ
    spec = env.get("PBN_CUDA_DEVICES")
    if spec:
        if spec.strip().lower() == "all":
            n = ctypes.c_int()
            check(lib().pbn_device_count(ctypes.byref(n)))
            return list(range(n.value))
        return [int(x) for x in spec.split(",") if x.strip() != ""]
    if "PBN_CUDA_DEVICE" in env:
        return [int(env["PBN_CUDA_DEVICE"])]
    if "LOCAL_RANK" in env:
        return [int(env["LOCAL_RANK"])]
    n = ctypes.c_int()
    check(lib().pbn_device_count(ctypes.byref(n)))
    return list(range(max(n.value, 1)))


def default_context():
    global _default_ctx
    if _default_ctx is None:
        devs = _default_devices()
        _default_ctx = Context(devs if len(devs) > 1 else devs[0])
    return _default_ctx


def set_default_context(ctx):
    """Replace the process-wide default context (e.g. `Context([0, 1, 2, 3])`); objects created earlier keep theirs."""
    global _default_ctx
    _default_ctx = ctx
    return ctx


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])
