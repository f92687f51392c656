"""zeldovich-plt_b200 — Python binding (ctypes) of libzeldovich_b200.so.

The product is the native library: CUDA kernels for sm_100a behind the C ABI declared in
``include/zeldovich_b200.h`` plus the C++ host code that mirrors the reference's
``Parameters`` / ``PowerSpectrum`` classes.  This module only marshals arguments; it does
no arithmetic of its own and has no CPU fallback — if the shared library is missing or no
B200 is present, the calls raise.

The directory name contains a hyphen, so load it with ``__graft_entry__.load_package()``
(which registers it as ``zeldovich_plt_b200`` in ``sys.modules``).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZPLT_LIB") or os.path.join(HERE, "libzeldovich_b200.so")
CLI_PATH = os.path.join(HERE, "bin", "zeldovich")

OK, EINVAL, ECUDA, ESTATE, ENOMEM = 0, 1, 2, 3, 4
ICFORMATS = {"Zeldovich": 0, "RVZel": 1, "RVdoubleZel": 2, "ZelSimple": 3}
RECORD_DTYPES = {
    0: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3)]),
    1: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f4", 3), ("vel", "<f4", 3)]),
    2: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3), ("vel", "<f8", 3)]),
    3: np.dtype([("displ", "<f4", 3)]),
}


class ZpltError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libzeldovich_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """``zplt_config`` (include/zeldovich_b200.h)."""

    _fields_ = [
        ("ppd", C.c_int64),
        ("boxsize", C.c_double),
        ("seed", C.c_int64),
        ("k_cutoff", C.c_double),
        ("corner_modes", C.c_int32),
        ("qonemode", C.c_int32),
        ("one_mode", C.c_int32 * 3),
        ("qPLT", C.c_int32),
        ("qPLTrescale", C.c_int32),
        ("fixed_power", C.c_int32),
        ("PLT_target_z", C.c_double),
        ("z_initial", C.c_double),
        ("f_cluster", C.c_double),
        ("icformat", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
        ("f_NL", C.c_double),
        ("n_s", C.c_double),
        ("Omega_M", C.c_double),
    ]


class Params(C.Structure):
    """``zplt_params``: reference class Parameters after setup()."""

    _fields_ = (
        [(n, C.c_double) for n in (
            "boxsize", "Pk_scale", "separation", "fundamental", "nyquist", "k_cutoff", "Pk_norm", "Pk_sigma", "Pk_sigma_ratio",
            "f_cluster", "Pk_smooth", "Pk_powerlaw_index", "z_initial", "PLT_target_z", "f_NL", "n_s", "Omega_M")]
        + [("ppd", C.c_int64), ("np", C.c_int64)]
        + [(n, C.c_int32) for n in ("cpd", "numblock", "qdensity", "qascii", "qoneslab", "seed", "qPk_fix_to_mean", "qonemode")]
        + [("one_mode", C.c_int32 * 3)]
        + [(n, C.c_int32) for n in ("qPLT", "qPLTrescale", "AllowDirectIO", "version", "CornerModes")]
        + [("Pk_filename", C.c_char * 1024), ("output_dir", C.c_char * 1024), ("density_filename", C.c_char * 1024),
           ("PLT_filename", C.c_char * 1024), ("ICFormat", C.c_char * 64)]
    )


class RunReport(C.Structure):
    _fields_ = [
        ("density_variance", C.c_double), ("rms_density", C.c_double), ("max_disp", C.c_double * 3),
        ("input_sigma", C.c_double), ("sigma_prediction", C.c_double),
        ("seconds_total", C.c_double), ("seconds_preamble", C.c_double), ("seconds_device", C.c_double),
        ("seconds_write", C.c_double), ("stage_ms", C.c_double * 4),
        ("ppd", C.c_int64), ("files_written", C.c_int64), ("bytes_written", C.c_int64),
        ("ooc_passes", C.c_int64), ("ooc_disk", C.c_int64), ("ooc_bytes", C.c_int64), ("ooc_part", C.c_int64), ("seconds_blocks", C.c_double),
    ]


# every symbol include/zeldovich_b200.h declares
EXPORTS = [
    "zplt_create", "zplt_destroy", "zplt_last_error", "zplt_record_bytes", "zplt_narray", "zplt_set_power_spline",
    "zplt_set_power_law", "zplt_set_primordial", "zplt_set_eigenmodes", "zplt_workspace_bytes", "zplt_set_workspace", "zplt_set_stream",
    "zplt_generate", "zplt_emit_planes", "zplt_fetch_planes", "zplt_emit_planes_density", "zplt_fetch_planes_density",
    "zplt_write_outputs", "zplt_reset_stats", "zplt_get_stats", "zplt_synchronize",
    "zplt_get_timings", "zplt_set_option", "zplt_dbg_set_peers", "zplt_dbg_spectral_hot", "zplt_dbg_hot_draws", "zplt_dbg_fft_variant",
    "zplt_exchange_info", "zplt_exchange_done", "zplt_slab_set_rank", "zplt_exchange_adopt", "zplt_ipc_export", "zplt_ipc_import", "zplt_ipc_close", "zplt_potential_begin", "zplt_potential_exchange", "zplt_slab_owner", "zplt_slab_offset", "zplt_dbg_pcg_draws", "zplt_dbg_mode_draws", "zplt_dbg_power_table", "zplt_dbg_spectral",
    "zplt_dbg_after_generate", "zplt_dbg_fft", "zplt_params_load", "zplt_icformat_code", "zplt_config_from_params",
    "zplt_power_create", "zplt_power_destroy", "zplt_power_info", "zplt_power_arrays", "zplt_power_eval",
    "zplt_power_sigmaR", "zplt_power_infer_Tk", "zplt_power_primordial_norm", "zplt_power_apply", "zplt_load_eigenmodes_file", "zplt_write_ic_files", "zplt_run_param_file",
]

_lib = None


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZpltError(-1, f"{LIB_PATH} not found — run __graft_entry__.build() (make -C zeldovich-plt_b200/csrc)")
    L = C.CDLL(LIB_PATH)
    vp, dp, i32, i64, u64 = C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_int64, C.c_uint64
    L.zplt_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.zplt_destroy.argtypes = [vp]
    L.zplt_destroy.restype = None
    L.zplt_last_error.restype = C.c_char_p
    L.zplt_record_bytes.argtypes = [i32]
    L.zplt_record_bytes.restype = C.c_size_t
    L.zplt_narray.argtypes = [vp]
    L.zplt_set_power_spline.argtypes = [vp, i32, dp, dp, dp, C.c_double, C.c_double]
    L.zplt_set_power_law.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.zplt_set_primordial.argtypes = [vp, C.c_double]
    L.zplt_set_eigenmodes.argtypes = [vp, i32, dp]
    L.zplt_workspace_bytes.argtypes = [vp]
    L.zplt_workspace_bytes.restype = C.c_size_t
    L.zplt_set_workspace.argtypes = [vp, vp, C.c_size_t]
    L.zplt_set_stream.argtypes = [vp, vp]
    L.zplt_generate.argtypes = [vp]
    L.zplt_emit_planes.argtypes = [vp, i64, i64, vp]
    L.zplt_fetch_planes.argtypes = [vp, i64, i64, vp]
    L.zplt_emit_planes_density.argtypes = [vp, i64, i64, vp, vp]
    L.zplt_fetch_planes_density.argtypes = [vp, i64, i64, vp, vp]
    L.zplt_write_outputs.argtypes = [vp, C.c_char_p, i32, i32, C.c_char_p, i32]
    L.zplt_reset_stats.argtypes = [vp]
    L.zplt_get_stats.argtypes = [vp, dp, dp]
    L.zplt_synchronize.argtypes = [vp]
    L.zplt_get_timings.argtypes = [vp, dp]
    L.zplt_exchange_info.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.zplt_exchange_done.argtypes = [vp]
    L.zplt_slab_set_rank.argtypes = [vp, i32]
    L.zplt_exchange_adopt.argtypes = [vp]
    L.zplt_ipc_export.argtypes = [vp, C.c_char_p]
    L.zplt_ipc_import.argtypes = [vp, i32, C.c_char_p]
    L.zplt_ipc_close.argtypes = [vp]
    L.zplt_potential_begin.argtypes = [vp]
    L.zplt_potential_exchange.argtypes = [vp]
    L.zplt_slab_owner.argtypes = [i64, i32, i64, C.POINTER(i32), C.POINTER(i32)]
    L.zplt_slab_offset.argtypes = [i64, i32, i32, i32, i32, i32, i64, i64]
    L.zplt_slab_offset.restype = i64
    L.zplt_dbg_pcg_draws.argtypes = [u64, u64, u64, i64, C.POINTER(u64)]
    L.zplt_dbg_mode_draws.argtypes = [vp, i64, C.POINTER(i32), C.POINTER(u64), dp]
    L.zplt_dbg_power_table.argtypes = [vp, i64, dp]
    L.zplt_dbg_spectral.argtypes = [vp, dp]
    L.zplt_dbg_after_generate.argtypes = [vp, dp]
    L.zplt_dbg_fft.argtypes = [i32, i64, i32, dp]
    L.zplt_dbg_fft_variant.argtypes = [i32, i64, i32, i32, dp]
    L.zplt_dbg_spectral_hot.argtypes = [vp, dp]
    L.zplt_dbg_hot_draws.argtypes = [vp, C.POINTER(u64)]
    L.zplt_dbg_set_peers.argtypes = [vp, i32, C.POINTER(vp)]
    L.zplt_set_option.argtypes = [vp, C.c_char_p, i32]
    L.zplt_params_load.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.zplt_icformat_code.argtypes = [C.c_char_p]
    L.zplt_config_from_params.argtypes = [C.POINTER(Params), C.POINTER(Config)]
    L.zplt_power_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.zplt_power_destroy.argtypes = [vp]
    L.zplt_power_destroy.restype = None
    L.zplt_power_info.argtypes = [vp, C.POINTER(i32), dp, dp]
    L.zplt_power_arrays.argtypes = [vp, dp, dp, dp]
    L.zplt_power_eval.argtypes = [vp, C.c_double]
    L.zplt_power_eval.restype = C.c_double
    L.zplt_power_sigmaR.argtypes = [vp, C.c_double]
    L.zplt_power_sigmaR.restype = C.c_double
    L.zplt_power_infer_Tk.argtypes = [vp, C.c_double]
    L.zplt_power_infer_Tk.restype = C.c_double
    L.zplt_power_primordial_norm.argtypes = [vp]
    L.zplt_power_primordial_norm.restype = C.c_double
    L.zplt_power_apply.argtypes = [vp, vp]
    L.zplt_load_eigenmodes_file.argtypes = [vp, C.c_char_p]
    L.zplt_write_ic_files.argtypes = [vp, C.c_char_p, i32]
    L.zplt_run_param_file.argtypes = [C.c_char_p, i32, i32, C.POINTER(RunReport)]
    _lib = L
    return L


def _ck(rc):
    if rc != OK:
        raise ZpltError(rc, lib().zplt_last_error().decode(errors="replace"))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------ host mirrors ----
class Parameters:
    """Mirror of reference ``Parameters(param_file)`` (src/parameters.cpp:11-57)."""

    def __init__(self, param_file):
        self.pod = Params()
        _ck(lib().zplt_params_load(os.fsencode(param_file), C.byref(self.pod)))

    def __getattr__(self, name):
        v = getattr(self.pod, name)
        if isinstance(v, bytes):
            return v.decode()
        if hasattr(v, "__len__"):
            return list(v)
        return v

    def config(self, device=-1):
        c = Config()
        _ck(lib().zplt_config_from_params(C.byref(self.pod), C.byref(c)))
        c.device = device
        return c


class PowerSpectrum:
    """Mirror of reference ``PowerSpectrum`` after InitFromFile/InitFromPowerLaw (host side)."""

    def __init__(self, params: Parameters):
        self._h = C.c_void_p()
        _ck(lib().zplt_power_create(C.byref(params.pod), C.byref(self._h)))
        n, norm, sm2 = C.c_int32(), C.c_double(), C.c_double()
        _ck(lib().zplt_power_info(self._h, C.byref(n), C.byref(norm), C.byref(sm2)))
        self.n, self.normalization, self.Pk_smooth2 = n.value, norm.value, sm2.value

    def arrays(self):
        x, y, y2 = (np.zeros(self.n) for _ in range(3))
        _ck(lib().zplt_power_arrays(self._h, _dp(x), _dp(y), _dp(y2)))
        return x, y, y2

    def power(self, k):
        return lib().zplt_power_eval(self._h, float(k))

    def sigmaR(self, R):
        return lib().zplt_power_sigmaR(self._h, float(R))

    def infer_Tk(self, k):
        return lib().zplt_power_infer_Tk(self._h, float(k))

    @property
    def primordial_norm(self):
        return lib().zplt_power_primordial_norm(self._h)

    def apply(self, ctx):
        _ck(lib().zplt_power_apply(self._h, ctx._h))

    def close(self):
        if self._h:
            lib().zplt_power_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_config(ppd, boxsize=720.0, seed=12346, k_cutoff=1.0, corner_modes=0, qonemode=0, one_mode=(0, 0, 0), qPLT=0,
                qPLTrescale=0, PLT_target_z=0.0, z_initial=49.0, f_cluster=1.0, fixed_power=0, icformat="RVZel", device=-1,
                rank=0, nranks=1, f_NL=0.0, n_s=1.0, Omega_M=1.0):
    """``zplt_config`` from keywords; unknown keywords raise (nothing is silently dropped)."""
    c = Config()
    c.f_NL, c.n_s, c.Omega_M = f_NL, n_s, Omega_M
    c.ppd, c.boxsize, c.seed, c.k_cutoff = ppd, boxsize, seed, k_cutoff
    c.corner_modes, c.qonemode = corner_modes, qonemode
    c.one_mode[:] = list(one_mode)
    c.qPLT, c.qPLTrescale, c.fixed_power = qPLT, qPLTrescale, fixed_power
    c.PLT_target_z, c.z_initial, c.f_cluster = PLT_target_z, z_initial, f_cluster
    c.icformat = ICFORMATS[icformat] if isinstance(icformat, str) else int(icformat)
    c.device, c.rank, c.nranks = device, rank, nranks
    return c


# ------------------------------------------------------------------ device context --
class Context:
    """One IC-generation context on one GPU (``zplt_ctx``)."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self._h = C.c_void_p()
        _ck(lib().zplt_create(C.byref(cfg), C.byref(self._h)))
        self.ppd = int(cfg.ppd)
        self.narray = lib().zplt_narray(self._h)
        self.record_bytes = lib().zplt_record_bytes(cfg.icformat)
        self.record_dtype = RECORD_DTYPES[cfg.icformat]

    def close(self):
        if self._h:
            lib().zplt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # inputs
    def set_power_spline(self, x, y, y2, normalization, Pk_smooth2):
        x, y, y2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, y2))
        _ck(lib().zplt_set_power_spline(self._h, len(x), _dp(x), _dp(y), _dp(y2), normalization, Pk_smooth2))

    def set_power_law(self, index, normalization, Pk_smooth2):
        _ck(lib().zplt_set_power_law(self._h, index, normalization, Pk_smooth2))

    def set_primordial(self, primordial_norm):
        _ck(lib().zplt_set_primordial(self._h, float(primordial_norm)))

    def set_eigenmodes(self, ppd_e, table):
        t = np.ascontiguousarray(table, dtype=np.float64).reshape(-1)
        assert t.size == ppd_e * ppd_e * (ppd_e // 2 + 1) * 4
        _ck(lib().zplt_set_eigenmodes(self._h, ppd_e, _dp(t)))

    def load_eigenmodes_file(self, path):
        _ck(lib().zplt_load_eigenmodes_file(self._h, os.fsencode(path)))

    # resources
    def workspace_bytes(self):
        return lib().zplt_workspace_bytes(self._h)

    def set_workspace(self, device_ptr, nbytes):
        _ck(lib().zplt_set_workspace(self._h, C.c_void_p(device_ptr), nbytes))

    def set_stream(self, cuda_stream):
        _ck(lib().zplt_set_stream(self._h, C.c_void_p(cuda_stream)))

    # hot path
    def generate(self):
        _ck(lib().zplt_generate(self._h))

    def emit_planes(self, z0, nz, device_ptr):
        _ck(lib().zplt_emit_planes(self._h, z0, nz, C.c_void_p(device_ptr)))

    def fetch_planes(self, z0, nz, out=None):
        n = nz * self.ppd * self.ppd
        if out is None:
            out = np.empty(n, dtype=self.record_dtype)
        assert out.nbytes >= n * self.record_bytes
        _ck(lib().zplt_fetch_planes(self._h, z0, nz, C.c_void_p(out.ctypes.data)))
        return out

    def fetch_planes_density(self, z0, nz, records=True):
        """(records or None, float32 density [nz, ppd, ppd]) of planes z0..z0+nz-1 (ZD_qdensity)."""
        n = nz * self.ppd * self.ppd
        rec = np.empty(n, dtype=self.record_dtype) if records else None
        dens = np.empty(n, dtype=np.float32)
        _ck(lib().zplt_fetch_planes_density(self._h, z0, nz, C.c_void_p(rec.ctypes.data) if records else None,
                                            C.c_void_p(dens.ctypes.data)))
        return rec, dens.reshape(nz, self.ppd, self.ppd)

    def fetch_planes_ptr(self, z0, nz, host_ptr):
        _ck(lib().zplt_fetch_planes(self._h, z0, nz, C.c_void_p(host_ptr)))

    def exchange_info(self):
        """(send_ptr, recv_ptr, bytes_per_peer) of a slab-decomposed context."""
        send, recv, nb = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _ck(lib().zplt_exchange_info(self._h, C.byref(send), C.byref(recv), C.byref(nb)))
        return send.value, recv.value, nb.value

    def ipc_export(self):
        buf = C.create_string_buffer(64)
        _ck(lib().zplt_ipc_export(self._h, buf))
        return buf.raw

    def ipc_import(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        _ck(lib().zplt_ipc_import(self._h, len(handles), blob))

    def potential_begin(self):
        """ZD_f_NL on slab ranks, stage 1 of the potential pass (a barrier across ranks follows)."""
        _ck(lib().zplt_potential_begin(self._h))

    def potential_exchange(self):
        """ZD_f_NL on slab ranks, stage 2 of the potential pass (a barrier across ranks follows, then generate())."""
        _ck(lib().zplt_potential_exchange(self._h))

    def ipc_close(self):
        _ck(lib().zplt_ipc_close(self._h))

    def exchange_done(self):
        _ck(lib().zplt_exchange_done(self._h))

    def slab_set_rank(self, rank):
        """Out of core: this context is slab rank ``rank`` from now on (zplt_slab_set_rank)."""
        _ck(lib().zplt_slab_set_rank(self._h, rank))
        self.cfg.rank = rank

    def exchange_adopt(self):
        """Out of core: the caller has filled the receive buffer with this rank's blocks (zplt_exchange_adopt)."""
        _ck(lib().zplt_exchange_adopt(self._h))

    def reset_stats(self):
        _ck(lib().zplt_reset_stats(self._h))

    def stats(self):
        var = C.c_double()
        md = np.zeros(3)
        _ck(lib().zplt_get_stats(self._h, C.byref(var), _dp(md)))
        return dict(density_variance=var.value, max_disp=md)

    def synchronize(self):
        _ck(lib().zplt_synchronize(self._h))

    def timings(self):
        """Device milliseconds of the last generate / emit calls: generation + x FFT (one kernel), z FFT, y FFT + emission."""
        t = np.zeros(8)
        _ck(lib().zplt_get_timings(self._h, _dp(t)))
        return dict(gen_xfft_ms=t[0], zfft_ms=t[1], yfft_emit_ms=t[3], launches=[int(v) for v in t[4:8]])

    def set_option(self, name, value):
        """Tuning / diagnostic switch (struct Tuning, csrc/zplt_internal.h)."""
        _ck(lib().zplt_set_option(self._h, name.encode(), int(value)))

    def dbg_set_peers(self, recv_ptrs):
        """Same-device stand-ins for the peers' receive buffers (None = discard that rank's share)."""
        arr = (C.c_void_p * len(recv_ptrs))(*[C.c_void_p(p) if p else C.c_void_p(None) for p in recv_ptrs])
        _ck(lib().zplt_dbg_set_peers(self._h, len(recv_ptrs), arr))

    def write_ic_files(self, output_dir, cpd):
        _ck(lib().zplt_write_ic_files(self._h, os.fsencode(output_dir), cpd))

    # introspection
    def mode_draws(self, kvecs):
        k = np.ascontiguousarray(kvecs, dtype=np.int32).reshape(-1, 3)
        raw = np.zeros(2 * len(k), dtype=np.uint64)
        u = np.zeros(2 * len(k))
        _ck(lib().zplt_dbg_mode_draws(self._h, len(k), k.ctypes.data_as(C.POINTER(C.c_int32)),
                                      raw.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(u)))
        return raw.reshape(-1, 2), u.reshape(-1, 2)

    def power_table(self, count):
        out = np.zeros(count)
        _ck(lib().zplt_dbg_power_table(self._h, count, _dp(out)))
        return out

    def spectral(self):
        N = self.ppd
        out = np.zeros((self.narray, N, N, N), dtype=np.complex128)
        _ck(lib().zplt_dbg_spectral(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def spectral_hot(self):
        """The packed arrays as the HOT generation kernel forms them (its transform skipped)."""
        N = self.ppd
        out = np.zeros((self.narray, N, N, N), dtype=np.complex128)
        _ck(lib().zplt_dbg_spectral_hot(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def hot_draws(self):
        """Raw 64-bit draws consumed by the hot generation kernel: uint64 [z][y < ppd/2][x][2] (skipped rows stay 0)."""
        N = self.ppd
        out = np.zeros((N, N // 2, N, 2), dtype=np.uint64)
        _ck(lib().zplt_dbg_hot_draws(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def after_generate(self):
        N = self.ppd
        out = np.zeros((self.narray, N, N, N), dtype=np.complex128)
        _ck(lib().zplt_dbg_after_generate(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out


def slab_owner(ppd, nranks, y):
    r, s = C.c_int32(), C.c_int32()
    _ck(lib().zplt_slab_owner(ppd, nranks, y, C.byref(r), C.byref(s)))
    return r.value, s.value


def slab_offset(ppd, nranks, narray, stage, rank, a, z, y):
    return lib().zplt_slab_offset(ppd, nranks, narray, stage, rank, a, z, y)


def pcg_draws(seed, offset, n):
    out = np.zeros(n, dtype=np.uint64)
    m = (1 << 64) - 1
    _ck(lib().zplt_dbg_pcg_draws(seed & m, (offset >> 64) & m, offset & m, n, out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def fft_backward(data, row_mode=True, variant=0):
    """Unnormalised backward FFT along the last (row_mode) or first axis of a 2-D complex array, on the GPU.
    variant 0: the kernels a default context uses; 1: plain one-tile-per-CTA kernels; 2: 8-pencil decimation (n = 2048)."""
    a = np.ascontiguousarray(data, dtype=np.complex128).copy()
    if row_mode:
        batch, n = a.shape
    else:
        n, batch = a.shape
    _ck(lib().zplt_dbg_fft_variant(n, batch, 1 if row_mode else 0, variant, a.ctypes.data_as(C.POINTER(C.c_double))))
    return a


def format_density_name(pattern, ppd):
    """ZD_density_filename with ppd filled in, as the reference's fmt::format does it (src/output.cpp:283)."""
    L = lib()
    L.zplt_format_density_name_.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(2048)
    L.zplt_format_density_name_(pattern.encode(), int(ppd), buf, 2048)
    return buf.value.decode()


def run_param_file(param_file, device=-1, write_files=True):
    """What ``./zeldovich <param_file>`` does, in-process.  Returns the run report."""
    rep = RunReport()
    _ck(lib().zplt_run_param_file(os.fsencode(param_file), device, 1 if write_files else 0, C.byref(rep)))
    return rep
