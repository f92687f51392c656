"""TEST INFRASTRUCTURE — ctypes front end of the CPU oracle (oracle/zel_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may
import this module, and only as the checker.  The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzel_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "zeldovich_ref")

ICFORMATS = {"Zeldovich": 0, "RVZel": 1, "RVdoubleZel": 2, "ZelSimple": 3}

RECORD_DTYPES = {
    0: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3)]),
    1: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f4", 3), ("vel", "<f4", 3)]),
    2: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3), ("vel", "<f8", 3)]),
    3: np.dtype([("displ", "<f4", 3)]),
}


class Config(C.Structure):
    _fields_ = [
        ("ppd", C.c_int64),
        ("boxsize", C.c_double),
        ("seed", C.c_int64),
        ("k_cutoff", C.c_double),
        ("corner_modes", C.c_int),
        ("qonemode", C.c_int),
        ("one_mode", C.c_int * 3),
        ("qPLT", C.c_int),
        ("qPLTrescale", C.c_int),
        ("PLT_target_z", C.c_double),
        ("z_initial", C.c_double),
        ("f_cluster", C.c_double),
        ("fixed_power", C.c_int),
        ("is_powerlaw", C.c_int),
        ("powerlaw_index", C.c_double),
        ("Pk_norm", C.c_double),
        ("Pk_sigma", C.c_double),
        ("Pk_sigma_ratio", C.c_double),
        ("Pk_smooth", C.c_double),
        ("Pk_scale", C.c_double),
        ("icformat", C.c_int),
        ("f_NL", C.c_double),
        ("n_s", C.c_double),
        ("Omega_M", C.c_double),
    ]


def make_config(ppd, boxsize=720.0, seed=12346, k_cutoff=1.0, corner_modes=0, qonemode=0, one_mode=(0, 0, 0), qPLT=0,
                qPLTrescale=0, PLT_target_z=0.0, z_initial=49.0, f_cluster=1.0, fixed_power=0, is_powerlaw=0,
                powerlaw_index=1000.0, Pk_norm=8.0, Pk_sigma=0.0210839935761, Pk_sigma_ratio=0.0, Pk_smooth=0.0,
                Pk_scale=1.0, icformat="RVZel", f_NL=0.0, n_s=1.0, Omega_M=1.0):
    c = Config()
    c.ppd, c.boxsize, c.seed, c.k_cutoff = ppd, boxsize, seed, k_cutoff
    c.corner_modes, c.qonemode = corner_modes, qonemode
    c.one_mode[:] = list(one_mode)
    c.qPLT, c.qPLTrescale, c.PLT_target_z, c.z_initial = qPLT, qPLTrescale, PLT_target_z, z_initial
    c.f_cluster, c.fixed_power = f_cluster, fixed_power
    c.is_powerlaw, c.powerlaw_index = is_powerlaw, powerlaw_index
    c.Pk_norm, c.Pk_sigma, c.Pk_sigma_ratio, c.Pk_smooth, c.Pk_scale = Pk_norm, Pk_sigma, Pk_sigma_ratio, Pk_smooth, Pk_scale
    c.icformat = ICFORMATS[icformat] if isinstance(icformat, str) else int(icformat)
    c.f_NL, c.n_s, c.Omega_M = f_NL, n_s, Omega_M
    return c


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.zo_pcg_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.POINTER(C.c_uint64)]
        L.zo_one_rand.argtypes = [C.c_uint64]
        L.zo_one_rand.restype = C.c_double
        L.zo_power_scalars.argtypes = [C.POINTER(Config), C.c_int, dp, dp, dp]
        L.zo_power_table.argtypes = [C.POINTER(Config), C.c_int, dp, dp, C.c_int64, dp]
        L.zo_infer_Tk.argtypes = [C.POINTER(Config), C.c_int, dp, dp, C.c_double]
        L.zo_infer_Tk.restype = C.c_double
        L.zo_spectral_cube.argtypes = [C.POINTER(Config), C.c_int, dp, dp, C.c_int64, dp, dp]
        L.zo_fft3_backward.argtypes = [dp, C.c_int64]
        L.zo_emit.argtypes = [C.POINTER(Config), dp, C.c_void_p, dp]
        L.zo_run.argtypes = [C.POINTER(Config), C.c_int, dp, dp, C.c_int64, dp, C.c_void_p, dp]
        L.zo_run.restype = C.c_int
        L.zo_planes.argtypes = [C.POINTER(Config), C.c_int, dp, dp, C.c_int64, dp, C.c_int, C.POINTER(C.c_int64), C.c_void_p, dp]
        L.zo_planes.restype = C.c_int
        L.zo_record_bytes.argtypes = [C.c_int]
        L.zo_record_bytes.restype = C.c_size_t
        L.zo_narray.argtypes = [C.POINTER(Config)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _table(pk):
    if pk is None:
        z = np.zeros(1)
        return 0, z, z
    k = np.ascontiguousarray(pk[0], dtype=np.float64)
    p = np.ascontiguousarray(pk[1], dtype=np.float64)
    return len(k), k, p


def _eig(eig):
    if eig is None:
        return 0, np.zeros(4)
    ppd_e, tab = eig
    return int(ppd_e), np.ascontiguousarray(tab, dtype=np.float64).reshape(-1)


def set_threads(n):
    """OpenMP threads of the oracle (torchrun pins its ranks to OMP_NUM_THREADS=1)."""
    lib().zo_set_threads(int(n))


def pcg_draws(seed, offset, n):
    out = np.empty(n, dtype=np.uint64)
    seed &= (1 << 64) - 1
    lib().zo_pcg_draws(seed, (offset >> 64) & ((1 << 64) - 1), offset & ((1 << 64) - 1), n,
                       out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def one_rand(r):
    return lib().zo_one_rand(int(r))


def power_scalars(cfg, pk):
    n, k, p = _table(pk)
    out = np.zeros(3)
    lib().zo_power_scalars(C.byref(cfg), n, _dp(k), _dp(p), _dp(out))
    return dict(normalization=out[0], Pk_smooth2=out[1], sigma_check=out[2])


def infer_Tk(cfg, pk, k):
    n, kk, p = _table(pk)
    return lib().zo_infer_Tk(C.byref(cfg), n, _dp(kk), _dp(p), float(k))


def power_table(cfg, pk, count):
    n, k, p = _table(pk)
    out = np.zeros(count)
    lib().zo_power_table(C.byref(cfg), n, _dp(k), _dp(p), count, _dp(out))
    return out


def spectral_cube(cfg, pk, eig=None):
    """complex128 array [narray][z][y][x] before any FFT."""
    n, k, p = _table(pk)
    pe, tab = _eig(eig)
    na = 4 if cfg.qPLT else 2
    N = cfg.ppd
    out = np.zeros((na, N, N, N), dtype=np.complex128)
    lib().zo_spectral_cube(C.byref(cfg), n, _dp(k), _dp(p), pe, _dp(tab), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def fft3_backward(cube):
    a = np.ascontiguousarray(cube, dtype=np.complex128).copy()
    N = a.shape[-1]
    for sub in a.reshape(-1, N, N, N):
        lib().zo_fft3_backward(sub.ctypes.data_as(C.POINTER(C.c_double)), N)
    return a


def run(cfg, pk, eig=None):
    """Full hot path.  Returns (records structured array [N^3], stats dict)."""
    n, k, p = _table(pk)
    pe, tab = _eig(eig)
    N = cfg.ppd
    dt = RECORD_DTYPES[cfg.icformat]
    rec = np.zeros(N * N * N, dtype=dt)
    stats = np.zeros(4)
    rc = lib().zo_run(C.byref(cfg), n, _dp(k), _dp(p), pe, _dp(tab), rec.ctypes.data_as(C.c_void_p), _dp(stats))
    if rc:
        raise RuntimeError(f"zo_run failed rc={rc}")
    return rec, dict(density_variance=stats[0], max_disp=stats[1:4].copy())


def planes(cfg, pk, zs, eig=None):
    """Records of the chosen z planes at any size (zo_planes: direct z summation, no cube).
    Returns (records [len(zs), N, N], list of per-plane stats dicts)."""
    n, k, p = _table(pk)
    pe, tab = _eig(eig)
    N = cfg.ppd
    zs = np.ascontiguousarray(zs, dtype=np.int64)
    dt = RECORD_DTYPES[cfg.icformat]
    rec = np.zeros((len(zs), N, N), dtype=dt)
    stats = np.zeros((len(zs), 4))
    rc = lib().zo_planes(C.byref(cfg), n, _dp(k), _dp(p), pe, _dp(tab), len(zs), zs.ctypes.data_as(C.POINTER(C.c_int64)),
                         rec.ctypes.data_as(C.c_void_p), _dp(stats))
    if rc:
        raise RuntimeError(f"zo_planes failed rc={rc}")
    return rec, [dict(density_variance=s[0], max_disp=s[1:4].copy()) for s in stats]


# ---------------------------------------------------------------- reference binary


def read_ic_dir(path, ppd, cpd, icformat):
    """Concatenate ic_* files in ascending-z order (src/output.cpp:208: plane z -> ic_{z*cpd/ppd})."""
    fmt = ICFORMATS[icformat] if isinstance(icformat, str) else icformat
    dt = RECORD_DTYPES[fmt]
    seen, parts = set(), []
    for z in range(ppd):
        n = z * cpd // ppd
        if n in seen:
            continue
        seen.add(n)
        parts.append(np.fromfile(os.path.join(path, f"ic_{n}"), dtype=dt))
    rec = np.concatenate(parts)
    assert rec.size == ppd**3, (rec.size, ppd**3)
    return rec


def run_reference(param_path, cwd, threads=None):
    """Run oracle/_ref/zeldovich_ref on a parameter file; returns its stderr text."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run([REF_BIN, param_path], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed ({r.returncode}):\n{r.stderr[-2000:]}")
    return r.stderr


def field_rel_err(a, b):
    """max|a-b| / max|b| per component field (BASELINE.md §4 definition of the 1e-10 tolerance)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))
