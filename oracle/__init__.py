"""CPU oracle — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so`` (``pbn_oracle.cpp``), the plain C++
restatement of the reference's KDE/CKDE log-likelihood path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this
package; ``pybnesian_b200`` never does.

All matrices are passed column-major (``order='F'``), null rows already removed —
the layout ``DataFrame::to_eigen`` hands to the reference's kernels
(/root/reference/pybnesian/dataset/dataset.hpp:236-338).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_dp = ctypes.POINTER(ctypes.c_double)
_c_vp = ctypes.c_void_p


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    src = os.path.join(_HERE, "pbn_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
    if os.path.isdir("/root/reference/pybnesian/kde/opencl_kernels") and os.path.isdir(os.path.join(_HERE, "ref_shim")):
        subprocess.check_call(["make", "-C", _HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = _c_int
    return _lib


def _dtype_code(a):
    if a.dtype == np.float64:
        return 0
    if a.dtype == np.float32:
        return 1
    raise ValueError("oracle: only float64 / float32 data")


def _fmat(a):
    a = np.asarray(a)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def _dptr(a):
    return a.ctypes.data_as(_c_dp)


def num_threads():
    return lib().orc_num_threads()


def use_all_threads():
    """All host threads for the OpenMP loops of the oracle (torchrun sets OMP_NUM_THREADS=1 for its workers, which
    made the round-1 reference arm single-threaded at N >= 2).  Returns the thread count now in use."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for L in (lib(), _fast_lib):
        if L is not None:
            L.orc_set_num_threads(_c_int(n))
    return num_threads()


# Timing build of the same source: -O3 -march=native (SURVEY.md 8d / BASELINE.md 3), compiled ON the machine that
# runs it (a -march=native object must not travel between hosts) into oracle/_fast/ (git- and gpurun-ignored).
# The parity build above stays -O2 -ffp-contract=off: it is the one pinned bit for bit to the reference kernels.
_fast_lib = None


def fast_lib():
    global _fast_lib
    if _fast_lib is None:
        import hashlib
        tag = "generic"
        try:
            with open("/proc/cpuinfo") as f:
                flags = [ln for ln in f if ln.startswith("flags") or ln.startswith("model name")][:2]
            tag = hashlib.sha1("".join(flags).encode()).hexdigest()[:12]
        except OSError:
            pass
        out_dir = os.path.join(_HERE, "_fast")
        path = os.path.join(out_dir, "liboracle_fast_%s.so" % tag)
        src = os.path.join(_HERE, "pbn_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            os.makedirs(out_dir, exist_ok=True)
            cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
            subprocess.check_call([cxx, "-O3", "-march=native", "-std=c++17", "-fPIC", "-fopenmp", "-w", "-shared",
                                   "-o", path, src])
        _fast_lib = ctypes.CDLL(path)
        _fast_lib.orc_num_threads.restype = _c_int
    return _fast_lib


def cov(X):
    X = _fmat(X)
    n, d = X.shape
    out = np.empty((d, d), order="F")
    lib().orc_cov(_c_vp(X.ctypes.data), _c_i64(n), _c_int(d), _c_int(_dtype_code(X)), _dptr(out))
    return out


class SingularCovariance(ValueError):
    pass


def bandwidth(X, rule="normal_reference"):
    """NormalReferenceRule / ScottsBandwidth bandwidth matrix (d x d float64)."""
    X = _fmat(X)
    n, d = X.shape
    H = np.empty((d, d), order="F")
    st = lib().orc_bandwidth(_c_vp(X.ctypes.data), _c_i64(n), _c_int(d), _c_int(_dtype_code(X)),
                             _c_int(0 if rule == "normal_reference" else 1), _dptr(H))
    if st:
        raise SingularCovariance("status %d" % st)
    return H


def kde_prepare(H, N):
    H = np.asfortranarray(np.asarray(H, dtype=np.float64))
    d = H.shape[0]
    L = np.empty((d, d), order="F")
    ln = ctypes.c_double()
    st = lib().orc_kde_prepare(_dptr(H), _c_int(d), _c_i64(N), _dptr(L), ctypes.byref(ln))
    if st:
        raise SingularCovariance("bandwidth not positive definite")
    return L, ln.value


def _logl_call(fn, train, test, H):
    train, test = _fmat(train), _fmat(test)
    assert train.dtype == test.dtype and train.shape[1] == test.shape[1]
    N, d = train.shape
    m = test.shape[0]
    H = np.asfortranarray(np.asarray(H, dtype=np.float64))
    out = np.empty(m)
    s = ctypes.c_double()
    fn(_c_vp(train.ctypes.data), _c_i64(N), _c_vp(test.ctypes.data), _c_i64(m), _c_int(d),
       _c_int(_dtype_code(train)), _dptr(H), _dptr(out), ctypes.byref(s))
    return out, s.value


def kde_logl(train, test, H):
    """(logl[m], slogl) of a KDE with bandwidth H, reference arithmetic in the data's dtype."""
    return _logl_call(lib().orc_kde_logl, train, test, H)


def ckde_logl(train, test, Hjoint, fast=False):
    """(logl[m], slogl) of CKDE(col 0 | cols 1..): joint - marginal with H[1:,1:].  `fast` runs the -O3 -march=native
    timing build (bench.py's CPU arms); parity checks use the default build."""
    return _logl_call((fast_lib() if fast else lib()).orc_ckde_logl, train, test, Hjoint)


def kde_logl_ld(train, test, H):
    """Long-double direct evaluation (independent cross-check of the restatement)."""
    train, test = _fmat(train), _fmat(test)
    N, d = train.shape
    m = test.shape[0]
    H = np.asfortranarray(np.asarray(H, dtype=np.float64))
    out = np.empty(m)
    st = lib().orc_kde_logl_ld(_c_vp(train.ctypes.data), _c_i64(N), _c_vp(test.ctypes.data), _c_i64(m), _c_int(d),
                               _c_int(_dtype_code(train)), _dptr(H), _dptr(out))
    if st:
        raise SingularCovariance("bandwidth not positive definite")
    return out


def ucv_score_unconstrained(X, H):
    X = _fmat(X)
    N, d = X.shape
    H = np.asfortranarray(np.asarray(H, dtype=np.float64).reshape(d, d))
    out = ctypes.c_double()
    lib().orc_ucv_score_unconstrained(_c_vp(X.ctypes.data), _c_i64(N), _c_int(d), _c_int(_dtype_code(X)), _dptr(H),
                                      ctypes.byref(out))
    return out.value


def ucv_score_diagonal(X, hdiag):
    X = _fmat(X)
    N, d = X.shape
    h = np.ascontiguousarray(np.asarray(hdiag, dtype=np.float64))
    out = ctypes.c_double()
    lib().orc_ucv_score_diagonal(_c_vp(X.ctypes.data), _c_i64(N), _c_int(d), _c_int(_dtype_code(X)), _dptr(h),
                                 ctypes.byref(out))
    return out.value


def cv_indices(valid_rows, k, seed):
    """Shuffled indices + fold limits (CrossValidationProperties)."""
    idx = np.ascontiguousarray(np.asarray(valid_rows, dtype=np.int32)).copy()
    limits = np.empty(k + 1, dtype=np.int32)
    lib().orc_cv_indices(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _c_i64(idx.size), _c_int(k),
                         ctypes.c_uint32(seed), limits.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return idx, limits


def holdout_indices(valid_rows, test_ratio, seed):
    idx = np.ascontiguousarray(np.asarray(valid_rows, dtype=np.int32)).copy()
    ntr = ctypes.c_int32()
    lib().orc_holdout_indices(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _c_i64(idx.size),
                              ctypes.c_double(test_ratio), ctypes.c_uint32(seed), ctypes.byref(ntr))
    return idx[:ntr.value], idx[ntr.value:]


def _colptrs(cols):
    arr = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    return arr


def lg_fit(y, parents):
    """MLE of a LinearGaussianCPD: (beta[p+1], variance)."""
    cols = [np.ascontiguousarray(y)] + [np.ascontiguousarray(p) for p in parents]
    p = len(cols) - 1
    beta = np.empty(p + 1)
    var = ctypes.c_double()
    lib().orc_lg_fit(_colptrs(cols), _c_i64(cols[0].size), _c_int(p), _c_int(_dtype_code(cols[0])), _dptr(beta),
                     ctypes.byref(var))
    return beta, var.value


def lg_logl(y, parents, beta, variance):
    cols = [np.ascontiguousarray(y)] + [np.ascontiguousarray(p) for p in parents]
    p = len(cols) - 1
    m = cols[0].size
    beta = np.ascontiguousarray(np.asarray(beta, dtype=np.float64))
    out = np.empty(m)
    s = ctypes.c_double()
    lib().orc_lg_logl(_colptrs(cols), _c_i64(m), _c_int(p), _c_int(_dtype_code(cols[0])), _dptr(beta),
                      ctypes.c_double(variance), _dptr(out), ctypes.byref(s))
    return out, s.value


def cv_score(X, indices, limits, factor="ckde", rule="normal_reference"):
    """CVLikelihood.local_score for column 0 given columns 1.. (X: all rows, column-major)."""
    X = _fmat(X)
    n, d = X.shape
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    limits = np.ascontiguousarray(limits, dtype=np.int32)
    out = ctypes.c_double()
    st = lib().orc_cv_score(_c_vp(X.ctypes.data), _c_i64(n), _c_int(d), _c_int(_dtype_code(X)),
                            indices.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                            limits.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _c_int(limits.size - 1),
                            _c_int(0 if factor == "ckde" else 1), _c_int(0 if rule == "normal_reference" else 1),
                            ctypes.byref(out))
    if st:
        raise SingularCovariance("status %d" % st)
    return out.value


# ---------------------------------------------------------------------------------------
# oracle/_ref: the reference's own OpenCL-C kernels compiled as C++ (ref_shim/), when built.
# Exists only where /root/reference was present at build time (this container); the
# prebuilt .so travels to the GPU box.  Used by tests to pin the restatement above.
# ---------------------------------------------------------------------------------------
_REF_PATH = os.path.join(_HERE, "_ref", "libref_kernels.so")
_ref = None


def ref_available():
    return os.path.exists(_REF_PATH)


def _reflib():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_PATH)
    return _ref


def ref_kde_logl(train, test, H):
    """KDE logl/slogl computed by the reference's kernels (host logic restated in ref_driver.cpp)."""
    train, test = _fmat(train), _fmat(test)
    N, d = train.shape
    m = test.shape[0]
    L, lognorm = kde_prepare(H, N)
    Lt = np.asfortranarray(L.astype(train.dtype))
    out = np.empty(m)
    s = ctypes.c_double()
    _reflib().ref_kde_logl(_c_vp(train.ctypes.data), _c_int(N), _c_vp(test.ctypes.data), _c_int(m), _c_int(d),
                           _c_int(_dtype_code(train)), _c_vp(Lt.ctypes.data), ctypes.c_double(lognorm), _dptr(out),
                           ctypes.byref(s))
    return out, s.value


def ref_ckde_logl(train, test, Hjoint):
    train, test = _fmat(train), _fmat(test)
    Hjoint = np.asarray(Hjoint, dtype=np.float64)
    joint, sj = ref_kde_logl(train, test, Hjoint)
    if train.shape[1] == 1:
        return joint, sj
    marg, _ = ref_kde_logl(train[:, 1:], test[:, 1:], Hjoint[1:, 1:])
    m = test.shape[0]
    out = np.empty(m)
    s = ctypes.c_double()
    _reflib().ref_ckde_combine(_dptr(np.ascontiguousarray(joint)), _dptr(np.ascontiguousarray(marg)), _c_int(m),
                               _c_int(_dtype_code(train)), _dptr(out), ctypes.byref(s))
    return out, s.value


def ref_ucv_score_unconstrained(X, H):
    """UCVScorer::score_unconstrained with the pair sums computed by the reference's kernels
    (the scalar epilogue restates kde/UCV.cpp:296-304, 357)."""
    X = _fmat(X)
    N, d = X.shape
    T = X.dtype.type
    Ht = np.asarray(H, dtype=np.float64).reshape(d, d).astype(X.dtype)
    L = np.linalg.cholesky(Ht.astype(np.float64)).astype(X.dtype)
    Lt = np.asfortranarray(L)
    pi_t = float(T(np.pi))
    lognorm_H = T(-np.sum(np.log(np.diag(Lt).astype(np.float64))) - 0.5 * d * np.log(2 * pi_t))
    lognorm_2H = float(lognorm_H) - 0.5 * d * np.log(2.0)
    s2h, sh = ctypes.c_double(), ctypes.c_double()
    _reflib().ref_ucv_sums(_c_vp(X.ctypes.data), _c_int(N), _c_int(d), _c_int(_dtype_code(X)), _c_vp(Lt.ctypes.data),
                           ctypes.c_double(lognorm_2H), ctypes.c_double(float(lognorm_H)), ctypes.byref(s2h),
                           ctypes.byref(sh))
    s2, s1 = T(s2h.value), T(sh.value)
    return float(np.exp(lognorm_2H) + float(T(2) * s2 / T(N)) - float(T(4) * s1 / T(N - 1)))


def diag_bandwidth(X, rule="normal_reference"):
    """NormalReferenceRule / ScottsBandwidth diag_bandwidth (d variances, float64)."""
    X = _fmat(X)
    n, d = X.shape
    h = np.empty(d)
    st = lib().orc_diag_bandwidth(_c_vp(X.ctypes.data), _c_i64(n), _c_int(d), _c_int(_dtype_code(X)),
                                  _c_int(0 if rule == "normal_reference" else 1), _dptr(h))
    if st:
        raise SingularCovariance("status %d" % st)
    return h


def product_kde_logl(train, test, hdiag):
    """(logl[m], slogl) of a ProductKDE with diagonal bandwidth `hdiag` (variances)."""
    train, test = _fmat(train), _fmat(test)
    assert train.dtype == test.dtype and train.shape[1] == test.shape[1]
    N, d = train.shape
    m = test.shape[0]
    h = np.ascontiguousarray(np.asarray(hdiag, dtype=np.float64).ravel())
    out = np.empty(m)
    s = ctypes.c_double()
    lib().orc_product_kde_logl(_c_vp(train.ctypes.data), _c_i64(N), _c_vp(test.ctypes.data), _c_i64(m), _c_int(d),
                               _c_int(_dtype_code(train)), _dptr(h), _dptr(out), ctypes.byref(s))
    return out, s.value


def ref_product_kde_logl(train, test, hdiag):
    """ProductKDE logl/slogl computed by the reference's kernels `logl_values_1d_mat` /
    `add_logl_values_1d_mat` (host logic of kde/ProductKDE.hpp:233-296 restated in ref_driver.cpp)."""
    train, test = _fmat(train), _fmat(test)
    N, d = train.shape
    m = test.shape[0]
    h = np.asarray(hdiag, dtype=np.float64).ravel()
    sd = np.sqrt(h) if train.dtype == np.float64 else np.sqrt(h.astype(np.float32))
    sd = np.ascontiguousarray(sd.astype(train.dtype))
    lognorm = -0.5 * d * np.log(2 * np.pi) - 0.5 * np.sum(np.log(h)) - np.log(N)
    out = np.empty(m)
    s = ctypes.c_double()
    _reflib().ref_product_kde_logl(_c_vp(train.ctypes.data), _c_int(N), _c_vp(test.ctypes.data), _c_int(m), _c_int(d),
                                   _c_int(_dtype_code(train)), _c_vp(sd.ctypes.data), ctypes.c_double(lognorm),
                                   _dptr(out), ctypes.byref(s))
    return out, s.value


# ---- CKDE::cdf / CKDE::sample, LinearGaussianCPD::cdf / sample ----------------------
def ckde_cdf(train, test, Hjoint):
    """CKDE(col 0 | cols 1..)::cdf of every test row (float64 out; reference arithmetic in the data's dtype)."""
    train, test = _fmat(train), _fmat(test)
    assert train.dtype == test.dtype and train.shape[1] == test.shape[1]
    N, d = train.shape
    m = test.shape[0]
    H = np.asfortranarray(np.asarray(Hjoint, dtype=np.float64).reshape(d, d))
    out = np.empty(m)
    lib().orc_ckde_cdf(_c_vp(train.ctypes.data), _c_i64(N), _c_vp(test.ctypes.data), _c_i64(m), _c_int(d),
                       _c_int(_dtype_code(train)), _dptr(H), _dptr(out))
    return out


def ckde_sample(train, Hjoint, evidence, n, seed):
    """CKDE::sample(n, evidence, seed): (samples in the data's dtype, sampled training-row indices)."""
    train = _fmat(train)
    N, d = train.shape
    H = np.asfortranarray(np.asarray(Hjoint, dtype=np.float64).reshape(d, d))
    out = np.empty(n, dtype=train.dtype)
    idx = np.empty(n, dtype=np.int32)
    ev = None
    if d > 1:
        ev = _fmat(evidence)
        assert ev.dtype == train.dtype and ev.shape == (n, d - 1)
    lib().orc_ckde_sample(_c_vp(train.ctypes.data), _c_i64(N), _c_int(d), _c_int(_dtype_code(train)), _dptr(H),
                          _c_vp(ev.ctypes.data) if ev is not None else None, _c_i64(n), ctypes.c_uint32(seed),
                          _c_vp(out.ctypes.data), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out, idx


def uniform_real(n, seed, dtype):
    """n draws of std::uniform_real_distribution<T>(0, 1) from std::mt19937{seed}."""
    out = np.empty(n, dtype=dtype)
    lib().orc_uniform_real(_c_i64(n), ctypes.c_uint32(seed), _c_int(_dtype_code(out)), _c_vp(out.ctypes.data))
    return out


def ckde_sample_indices(mtrain, etest, Hmarg, random_prob):
    """Training-row index drawn for every evidence row from the marginal kernel weights."""
    mtrain, etest = _fmat(mtrain), _fmat(etest)
    N, p = mtrain.shape
    n = etest.shape[0]
    H = np.asfortranarray(np.asarray(Hmarg, dtype=np.float64).reshape(p, p))
    rp = np.ascontiguousarray(random_prob, dtype=mtrain.dtype)
    out = np.empty(n, dtype=np.int32)
    st = lib().orc_ckde_sample_indices(_c_vp(mtrain.ctypes.data), _c_i64(N), _c_vp(etest.ctypes.data), _c_i64(n),
                                       _c_int(p), _c_int(_dtype_code(mtrain)), _dptr(H), _c_vp(rp.ctypes.data),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if st:
        raise SingularCovariance("bandwidth not positive definite")
    return out


def lg_sample(beta, variance, evidence_cols, n, seed):
    """LinearGaussianCPD::sample: float64 samples; evidence_cols = list of n-vectors (float64 or float32)."""
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    cols = [np.ascontiguousarray(c) for c in evidence_cols]
    code = _dtype_code(cols[0]) if cols else 0
    ptrs = (_c_vp * max(1, len(cols)))(*[c.ctypes.data for c in cols])
    out = np.empty(n)
    lib().orc_lg_sample(_dptr(beta), ctypes.c_double(variance), _c_int(len(cols)), ptrs, _c_int(code), _c_i64(n),
                        ctypes.c_uint32(seed), _dptr(out))
    return out


def _cond_params(Hjoint):
    """Host-side quantities of CKDE::_cdf_multivariate / _sample_multivariate (CKDE.hpp:346-360, 594-616)."""
    H = np.asarray(Hjoint, dtype=np.float64)
    Lm = np.linalg.cholesky(H[1:, 1:])
    inv = np.linalg.solve(Lm, np.eye(Lm.shape[0]))
    R = inv @ H[1:, 0]
    return Lm, R @ inv, H[0, 0] - R @ R


def ref_ckde_cdf(train, test, Hjoint):
    """CKDE::cdf computed by the reference's kernels (host enqueue logic restated in ref_driver.cpp)."""
    train, test = _fmat(train), _fmat(test)
    N, d = train.shape
    m = test.shape[0]
    T = train.dtype
    H = np.asarray(Hjoint, dtype=np.float64).reshape(d, d)
    out = np.empty(m)
    x = np.ascontiguousarray(test[:, 0])
    if d == 1:
        _reflib().ref_ckde_cdf(_c_vp(train.ctypes.data), _c_int(N), _c_vp(x.ctypes.data), None, _c_int(m), _c_int(1),
                               _c_int(_dtype_code(train)), None, ctypes.c_double(0.0), None,
                               ctypes.c_double(float(T.type(1.0 / np.sqrt(H[0, 0])))), _dptr(out))
        return out
    Lm, transform, cond_var = _cond_params(H)
    _, lognorm = kde_prepare(H[1:, 1:], N)
    et = np.asfortranarray(test[:, 1:])
    Lt = np.asfortranarray(Lm.astype(T))
    tr = np.ascontiguousarray(transform.astype(T))
    _reflib().ref_ckde_cdf(_c_vp(train.ctypes.data), _c_int(N), _c_vp(x.ctypes.data), _c_vp(et.ctypes.data), _c_int(m),
                           _c_int(d), _c_int(_dtype_code(train)), _c_vp(Lt.ctypes.data),
                           ctypes.c_double(lognorm + np.log(float(N))), _c_vp(tr.ctypes.data),
                           ctypes.c_double(float(T.type(1.0 / np.sqrt(cond_var)))), _dptr(out))
    return out


def ref_ckde_sample_indices(mtrain, etest, Hmarg, random_prob):
    mtrain, etest = _fmat(mtrain), _fmat(etest)
    N, p = mtrain.shape
    n = etest.shape[0]
    L, lognorm = kde_prepare(np.asarray(Hmarg, dtype=np.float64).reshape(p, p), N)
    Lt = np.asfortranarray(L.astype(mtrain.dtype))
    rp = np.ascontiguousarray(random_prob, dtype=mtrain.dtype)
    out = np.empty(n, dtype=np.int32)
    _reflib().ref_ckde_sample_indices(_c_vp(mtrain.ctypes.data), _c_int(N), _c_vp(etest.ctypes.data), _c_int(n),
                                      _c_int(p), _c_int(_dtype_code(mtrain)), _c_vp(Lt.ctypes.data),
                                      ctypes.c_double(lognorm), _c_vp(rp.ctypes.data),
                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out
