"""CPU tests of the host side of the boundary: the C-ABI library loads and exports every
declared symbol, the ParseHeader-compatible reader, Parameters checks, and the host
PowerSpectrum (spline + sigma(R) normalisation) against the oracle.  No device calls."""
import os
import re
import tempfile

import numpy as np
import pytest

import helpers
from __graft_entry__ import ROOT, load_package, load_synth


@pytest.fixture(scope="module")
def pkg():
    p = load_package()
    p.lib()
    return p


def write_case(tmp, pk=None, **over):
    synth = load_synth()
    k, p = pk if pk is not None else helpers.wmap_pk()
    synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
    over.setdefault("ZD_Pk_filename", '"%s"' % os.path.join(tmp, "pk.pow"))
    return synth.write_param(os.path.join(tmp, "c.par"), **over)


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "zeldovich_b200.h")).read()
    declared = set(re.findall(r"\b(zplt_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.EXPORTS), declared ^ set(pkg.EXPORTS)
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_header(pkg):
    # sizes the C compiler produced for the same declarations
    import ctypes as C
    import subprocess

    src = '#include "zeldovich_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n",sizeof(zplt_config),sizeof(zplt_params),sizeof(zplt_run_report));return 0;}\n'
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "s.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(tmp, "s.c"), "-o", os.path.join(tmp, "s")])
        out = subprocess.check_output([os.path.join(tmp, "s")], text=True).split()
    assert [int(v) for v in out] == [C.sizeof(pkg.Config), C.sizeof(pkg.Params), C.sizeof(pkg.RunReport)]


def test_no_device_means_error_not_fallback(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ZpltError) as e:
        pkg.Context(pkg.make_config(32))
    assert e.value.code == pkg.ECUDA


def test_parameters_example_values(pkg):
    with tempfile.TemporaryDirectory() as tmp:
        path = write_case(tmp, NP=64**3)
        P = pkg.Parameters(path)
    assert P.ppd == 64 and P.np == 262144 and P.cpd == 375 and P.numblock == 4
    assert P.boxsize == 720.0 and P.seed == 12346 and P.version == 2
    assert P.ICFormat == "RVZel" and P.output_dir == "./ic_out"
    assert P.separation == 720.0 / 64 and P.fundamental == 2.0 * np.pi / 720.0 and P.nyquist == np.pi / (720.0 / 64)
    # the reference scanner's own decimal conversion: digits / 10^n, not strtod
    assert P.Pk_sigma == 210839935761.0 / 1e13
    assert P.qoneslab == -1 and P.f_cluster == 1.0 and P.Pk_powerlaw_index == 1000


def test_parser_grammar(pkg):
    text = (
        "# leading comment\n"
        "##\nBoxSize = 1   # inside a block comment\n##\n"
        'BoxSize = 2.5D2   # Fortran exponent\n'
        "CPD = 11\nNP = 4096\nZD_NumBlock = 2\nZD_Pk_scale = 1\nZD_Seed = -5\nZD_Pk_norm = 0\n"
        "ZD_Pk_sigma = 1.0\nZD_Pk_smooth = 0\nInitialRedshift = 9\nICFormat = ZelSimple\n"
        "InitialConditionsDirectory = 'out dir'\nZD_Version = 2.0\n"
        "ZD_Pk_powerlaw_index = -1.5\n"
        "ZD_one_mode = 1 \\\n   -2 3\nZD_qonemode = 1\n"
        "SomeUnknownKey = 3 4 5\n"
        "\x02\nBoxSize = 999\n"
    )
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "g.par")
        open(path, "w").write(text)
        P = pkg.Parameters(path)
    assert P.boxsize == 250.0 and P.seed == -5 and P.ICFormat == "ZelSimple" and P.output_dir == "out dir"
    assert P.version == 2 and P.one_mode == [1, -2, 3] and P.Pk_powerlaw_index == -1.5 and P.ppd == 16


@pytest.mark.parametrize("bad,msg", [
    (dict(ZD_Version=None), "ZD_Version"),
    (dict(NP=1000001), "perfect cube"),
    (dict(ZD_Pk_sigma=0), "exactly one of Pk_sigma"),
    (dict(ZD_k_cutoff="0.5"), "k_cutoff"),
    (dict(ZD_qPLT=1), "ZD_PLT_filename"),
    (dict(ZD_NumBlock=3), "NumBlock"),
    (dict(ZD_f_cluster="1.5"), "f_cluster"),
    (dict(ZD_Pk_powerlaw_index="-1"), "exactly one of ZD_Pk_filename"),
])
def test_parameter_checks(pkg, bad, msg):
    with tempfile.TemporaryDirectory() as tmp:
        over = dict(bad)
        drop = [k for k, v in over.items() if v is None]
        path = write_case(tmp, **{k: v for k, v in over.items() if v is not None})
        if drop:
            lines = [l for l in open(path) if not any(l.startswith(k + " ") for k in drop)]
            open(path, "w").writelines(lines)
        with pytest.raises(pkg.ZpltError) as e:
            pkg.Parameters(path)
    assert msg in str(e.value)


def test_host_power_spectrum_matches_oracle(pkg, oracle):
    with tempfile.TemporaryDirectory() as tmp:
        P = pkg.Parameters(write_case(tmp, NP=64**3))
        pk = pkg.PowerSpectrum(P)
        s = oracle.power_scalars(oracle.make_config(64), helpers.wmap_pk())
        assert abs(pk.normalization / s["normalization"] - 1) < 1e-14
        assert pk.n == 176
        # sigma(8) comes back as ZD_Pk_sigma; the unnormalised value is what the reference prints
        assert abs(pk.sigmaR(8.0) - 0.0210839935761 / 720.0**1.5) < 1e-15
        want = oracle.power_table(oracle.make_config(64), helpers.wmap_pk(), 200)
        fund = 2 * np.pi / 720.0
        got = np.array([pk.power(np.sqrt(m * fund * fund)) for m in range(200)])
        assert got[0] == 0 and np.max(np.abs(got[1:] / want[1:] - 1)) < 1e-14
        x, y, y2 = pk.arrays()
        assert np.all(np.diff(x) > 0) and y2[0] == 0 and y2[-1] == 0


def test_host_transfer_function_for_f_nl(pkg, oracle):
    """PowerSpectrum::infer_Tk / primordial_norm (reference src/power_spectrum.cpp:221-222, 263-274): T(k) = 1 at the
    smallest k of the table, sqrt(P / (primordial_norm k^n_s)) elsewhere — the host scalar the f_NL kernels consume."""
    with tempfile.TemporaryDirectory() as tmp:
        P = pkg.Parameters(write_case(tmp, NP=64**3, ZD_f_NL="100", ZD_n_s="0.96", Omega_M="0.3"))
        assert P.pod.f_NL == 100.0 and P.pod.n_s == 0.96 and P.pod.Omega_M == 0.3
        cfg = P.config()
        assert cfg.f_NL == 100.0 and cfg.n_s == 0.96 and cfg.Omega_M == 0.3
        pk = pkg.PowerSpectrum(P)
        k, p = helpers.wmap_pk()
        kmin = k[k > 0].min()
        assert abs(pk.infer_Tk(kmin) - 1.0) < 1e-14 and pk.infer_Tk(0.0) == 1.0
        assert abs(pk.primordial_norm / (pk.power(kmin) / kmin**0.96) - 1) < 1e-13
        ocfg = oracle.make_config(64, f_NL=100.0, n_s=0.96, Omega_M=0.3)
        for kk in (1e-3, 0.01, 0.0873, 0.5, 2.0, 7.5):
            assert abs(pk.infer_Tk(kk) / oracle.infer_Tk(ocfg, (k, p), kk) - 1) < 1e-13


def test_host_power_law_and_sigma_ratio(pkg, oracle):
    with tempfile.TemporaryDirectory() as tmp:
        path = write_case(tmp, NP=32**3, ZD_Pk_filename='""', ZD_Pk_powerlaw_index="-2.0", ZD_Pk_sigma=0, ZD_Pk_sigma_ratio="0.5",
                          ZD_Pk_smooth="1.5")
        P = pkg.Parameters(path)
        pk = pkg.PowerSpectrum(P)
    cfg = oracle.make_config(32, is_powerlaw=1, powerlaw_index=-2.0, Pk_sigma=0.0, Pk_sigma_ratio=0.5, Pk_smooth=1.5)
    s = oracle.power_scalars(cfg, None)
    assert pk.normalization == s["normalization"] and pk.Pk_smooth2 == 2.25 and pk.n == 0
    want = oracle.power_table(cfg, None, 50)
    fund = 2 * np.pi / 720.0
    got = np.array([pk.power(np.sqrt(m * fund * fund)) for m in range(50)])
    assert np.max(np.abs(got[1:] / want[1:] - 1)) < 1e-14


def test_cli_usage_and_bad_file(pkg):
    import subprocess

    r = subprocess.run([pkg.CLI_PATH], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([pkg.CLI_PATH, "/nonexistent.par"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1


def test_ic_file_reader_round_trip(tmp_path):
    """The consumer-side reader (zeldovich-plt_b200/icfiles.py) on files laid out as the reference lays them out: a golden
    case's records split into ic_<z*CPD//PPD> files in ascending z."""
    import importlib.util

    import numpy as np

    import helpers

    spec = importlib.util.spec_from_file_location("zplt_icfiles", os.path.join(helpers.ROOT, "zeldovich-plt_b200", "icfiles.py"))
    icf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(icf)
    case, raw, _, _ = helpers.load_golden("plt16_direct_rvzel")  # CPD = 5 < PPD = 16: several planes per file
    ppd, cpd = 16, int(case["params"]["CPD"])
    rec = raw.view(icf.RECORD_DTYPES[1])
    planes = icf.ic_file_planes(ppd, cpd)
    assert sorted(planes) == sorted({z * cpd // ppd for z in range(ppd)})
    for n, (z0, z1) in planes.items():
        rec[z0 * ppd * ppd:z1 * ppd * ppd].tofile(str(tmp_path / f"ic_{n}"))
    got, (z0, z1) = icf.read_ic_files(str(tmp_path), ppd, cpd, "RVZel")
    assert (z0, z1) == (0, ppd) and np.array_equal(got.view(np.uint8), rec.view(np.uint8))
    part, (a, b) = icf.read_ic_files(str(tmp_path), ppd, cpd, "RVZel", files=[1, 2])
    assert np.array_equal(part["ijk"][:, 0], np.repeat(np.arange(a, b), ppd * ppd))
    pos = icf.global_positions(got, ppd, 720.0)
    assert pos.shape == (ppd**3, 3) and pos.min() >= 0.0 and pos.max() < 720.0
    site = icf.lattice_indices(got, ppd) * (720.0 / ppd)
    d = (pos - site + 360.0) % 720.0 - 360.0  # displacement back out of the wrapped position
    assert np.allclose(d, got["displ"], atol=1e-4)
    assert np.array_equal(icf.velocities(got), got["vel"].astype(np.float64))
    simple = np.zeros(ppd**3, dtype=icf.RECORD_DTYPES[3])
    assert np.array_equal(icf.lattice_indices(simple, ppd)[-1], [ppd - 1, ppd - 1, ppd - 1])
    with pytest.raises(ValueError):
        icf.read_ic_files(str(tmp_path), ppd + 16, cpd, "RVZel")  # wrong ppd: sizes do not match


def test_density_file_name_formatting(pkg):
    """ZD_density_filename goes through fmt::format(name, ppd) in the reference (src/output.cpp:283)."""
    f = pkg.format_density_name
    assert f("density{:d}", 256) == "density256"
    assert f("density{}", 64) == "density64" and f("d{0}.bin", 64) == "d64.bin" and f("d{0:d}", 7) == "d7"
    assert f("rho_{:05d}.f32", 128) == "rho_00128.f32" and f("rho_{:5d}", 128) == "rho_  128"
    assert f("plain", 32) == "plain" and f("{{x}}{:d}", 32) == "{x}32"


def test_bench_line_helpers():
    """The pure parts of bench.py (roofline block, workload description, reference sample size) on CPU."""
    import argparse
    import importlib.util

    spec = importlib.util.spec_from_file_location("zplt_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, na, rb = 1024, 4, 32
    r = bench.roofline_block(N, na, rb, 1, [30.5, 27.5, 24.9], 82.9, False)
    assert r["kernel"].startswith("gen_xfft_kernel") and set(r["kernels"]) == set(bench.KERNELS)
    assert abs(r["kernels"]["z-FFT"]["achieved"] - 32 * na * N**3 / 27.5e-3 / 1e9) < 1e-6
    assert abs(r["step_algorithmic_bytes"] - (64 * na + rb) * N**3) == 0 and 0.5 < r["step_frac"] < 0.65
    r8 = bench.roofline_block(N, na, rb, 8, [8.0, 3.7, 4.0], 15.9, True)
    assert set(r8["kernels"]) == {"generate+x-FFT", "y-FFT+emit"} and abs(r8["kernels"]["generate+x-FFT"]["ms"] - 11.7) < 1e-9
    cfg = bench.workload_config(N, True, "RVZel", 8, "p2p")
    assert cfg["ppd"] == N and cfg["narray"] == 4 and "8 GPUs" in cfg["parallelism"]
    args = argparse.Namespace(ref_ppd=0, ppd=1024)
    ppd, why = bench.ref_ppd_for(args, True)
    assert ppd in (128, 256, 512, 1024) and why
    assert bench.ref_ppd_for(argparse.Namespace(ref_ppd=256, ppd=1024), True)[0] == 256


# ------------------------------------------------------------------ out-of-core scheduling ----
@pytest.fixture(scope="module")
def mocklib():
    """The UNMODIFIED host half (host/*.cpp) linked against tests/mock_device.cpp instead of the CUDA half: the pass
    scheduling, the block store and the ic file writer of zplt_run_param_file run on this machine, no GPU involved."""
    import ctypes as C
    import subprocess

    pkg = load_package()
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libzplt_mockdev.so")
    host = os.path.join(ROOT, "zeldovich-plt_b200", "host")
    srcs = [os.path.join(host, f) for f in ("ParseHeader.cpp", "host_parameters.cpp", "host_power.cpp", "host_api.cpp")]
    srcs.append(os.path.join(ROOT, "tests", "mock_device.cpp"))
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=gnu++17", "-O1", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", *srcs, "-o", so])
    L = C.CDLL(so)
    L.zplt_run_param_file.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(pkg.RunReport)]
    L.zplt_last_error.restype = C.c_char_p
    return L, pkg


def _mock_run(mocklib, tmp, env, write_files=1, **over):
    import ctypes as C

    L, pkg = mocklib
    out = os.path.join(tmp, "ic")
    par = write_case(tmp, InitialConditionsDirectory='"%s"' % out, **over)
    rep = pkg.RunReport()
    keys = ("ZPLT_OOC_PASSES", "ZPLT_OOC_STORE", "ZPLT_MOCK_FREE_BYTES", "ZPLT_MOCK_FAIL_COPY", "ZPLT_OOC_PART")
    L.zplt_mock_reset_copies()
    saved = {k: os.environ.pop(k, None) for k in keys}
    os.environ.update(env)
    try:
        rc = L.zplt_run_param_file(os.fsencode(par), 0, write_files, C.byref(rep))
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    return rc, rep, out, L.zplt_last_error().decode()


def _check_mock_files(out, ppd, cpd, rb, planes=None):
    """Every ic file holds exactly its planes (z*cpd/ppd == file number) in ascending z, each with the mock's pattern."""
    planes = list(range(ppd)) if planes is None else planes
    words = ppd * ppd * rb // 4
    j = np.arange(words, dtype=np.uint32) & 0xFFFFF
    want = {}
    for z in planes:
        want.setdefault(z * cpd // ppd, []).append(z)
    got = sorted(f for f in os.listdir(out) if f.startswith("ic_"))
    assert got == sorted(f"ic_{k}" for k in want), got
    for k, zs in want.items():
        data = np.fromfile(os.path.join(out, f"ic_{k}"), dtype=np.uint32)
        assert data.size == len(zs) * words, (k, data.size)
        for i, z in enumerate(zs):
            assert np.array_equal(data[i * words:(i + 1) * words], (np.uint32(z) << np.uint32(20)) ^ j), (k, z)


@pytest.mark.parametrize("chunk_planes", [0, 3])
@pytest.mark.parametrize("store", ["ram", "pageable", "disk"])
@pytest.mark.parametrize("passes,ppd,cpd,fmt,qplt", [(4, 32, 5, "RVZel", 0), (2, 16, 16, "RVdoubleZel", 1), (8, 32, 3, "ZelSimple", 0), (16, 32, 32, "Zeldovich", 0),
                                                     (2, 128, 7, "RVZel", 1), (2, 256, 9, "ZelSimple", 0)])
def test_out_of_core_pass_scheduling(mocklib, monkeypatch, store, passes, ppd, cpd, fmt, qplt, chunk_planes):
    """zplt_run_param_file out of core (reference -DDISK, src/block_array.cpp:129-382): block (s, d) must reach rank d's receive
    buffer at position s (the mock's zplt_exchange_adopt checks every value), planes must reach the ic files in ascending z,
    block files carry the reference's names while they exist and are gone afterwards."""
    synth = load_synth()
    rb = {"RVZel": 32, "RVdoubleZel": 56, "ZelSimple": 12, "Zeldovich": 32}[fmt]
    if chunk_planes:
        chunk_planes = max(chunk_planes, ppd // 5)  # a handful of chunks whatever the grid
        monkeypatch.setenv("ZPLT_IC_CHUNK_BYTES", str(chunk_planes * ppd * ppd * rb))  # chunks that straddle pass boundaries or not
    with tempfile.TemporaryDirectory() as tmp:
        over = dict(NP=ppd**3, CPD=cpd, ICFormat='"%s"' % fmt)
        if qplt:
            synth.write_eigmodes(os.path.join(tmp, "eig"), 8)
            over.update(ZD_qPLT=1, ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": str(passes), "ZPLT_OOC_STORE": store}, **over)
        assert rc == 0, err
        _check_mock_files(out, ppd, cpd, rb)
        assert rep.ooc_passes == passes and rep.ooc_disk == (store == "disk")
        assert rep.ooc_bytes == 16 * (4 if qplt else 2) * ppd**3  # the whole cube went out once
        assert rep.density_variance == ppd  # (mock) planes that went through the emission: each exactly once
        assert rep.max_disp[0] == passes  # (mock) one generation per pass
        assert rep.bytes_written == ppd**3 * rb and rep.files_written == len({z * cpd // ppd for z in range(ppd)})
        assert not [f for f in os.listdir(out) if f.startswith("zeldovich.")]  # block directories removed


def test_out_of_core_options_and_auto_passes(mocklib):
    ppd = 32
    cube = 16 * 2 * ppd**3
    with tempfile.TemporaryDirectory() as tmp:
        # the cube fits: resident run, no passes
        rc, rep, out, err = _mock_run(mocklib, tmp, {}, NP=ppd**3, CPD=4)
        assert rc == 0 and rep.ooc_passes == 0, err
        _check_mock_files(out, ppd, 4, 32)
    with tempfile.TemporaryDirectory() as tmp:
        # free HBM for a quarter of the cube twice over (+ the fixed 1 GiB allowance and the padding room): 8 passes, not 4
        free = (1 << 30) + 2 * cube // 8 + (ppd // 8) * 8192 * 16
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_MOCK_FREE_BYTES": str(free)}, NP=ppd**3, CPD=4, ZD_qdensity=1,
                                      ZD_density_filename='"dens{:d}"')
        assert rc == 0 and rep.ooc_passes == 8 and rep.ooc_disk == 0, (err, rep.ooc_passes)
        _check_mock_files(out, ppd, 4, 32)
        dens = np.fromfile(os.path.join(out, f"dens{ppd}"), dtype=np.float32).reshape(ppd, ppd * ppd)
        assert np.array_equal(dens, np.repeat(np.arange(ppd, dtype=np.float32)[:, None], ppd * ppd, axis=1))
    with tempfile.TemporaryDirectory() as tmp:
        # ZD_qoneslab: one plane, from the pass that owns it; statistics-only runs (write_files = 0) still emit every plane
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": "4"}, NP=ppd**3, CPD=4, ZD_qoneslab=21)
        assert rc == 0, err
        _check_mock_files(out, ppd, 4, 32, planes=[21])
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": "4", "ZPLT_OOC_STORE": "disk"}, write_files=0, NP=ppd**3, CPD=4)
        assert rc == 0 and rep.density_variance == ppd and rep.files_written == 0, err
    with tempfile.TemporaryDirectory() as tmp:
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": "3"}, NP=ppd**3, CPD=4)
        assert rc != 0 and "ZPLT_OOC_PASSES" in err
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_MOCK_FREE_BYTES": str(1 << 20)}, NP=ppd**3, CPD=4)
        assert rc != 0 and "16 passes" in err
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": "2", "ZPLT_OOC_STORE": "tape"}, NP=ppd**3, CPD=4)
        assert rc != 0 and "ZPLT_OOC_STORE" in err


@pytest.mark.parametrize("ppd,cpd,fmt,chunk_planes", [(32, 5, "RVZel", 0), (256, 256, "ZelSimple", 0), (64, 7, "RVZel", 5), (32, 32, "RVZel", 1)])
def test_ic_writer_resident_options(mocklib, monkeypatch, ppd, cpd, fmt, chunk_planes):
    """zplt_write_outputs on a resident context (mock device half): ZD_qdensity 0/1/2 and ZD_qoneslab, the writer's chunking
    (ppd=256 ZelSimple is 201 MB of records in one chunk of planes; ppd=32 shares files between planes), stale files removed."""
    import ctypes as C

    L, pkg = mocklib
    rb = {"RVZel": 32, "ZelSimple": 12}[fmt]
    if chunk_planes:  # many chunks: both staging buffers in use, writers overlapped with the next fetch
        monkeypatch.setenv("ZPLT_IC_CHUNK_BYTES", str(chunk_planes * ppd * ppd * rb))
    cfg = pkg.make_config(ppd, icformat=fmt)
    ctx = C.c_void_p()
    assert L.zplt_create(C.byref(cfg), C.byref(ctx)) == 0
    L.zplt_set_power_law.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.zplt_write_outputs.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    assert L.zplt_set_power_law(ctx, 1.0, 1.0, 0.0) == 0
    L.zplt_generate.argtypes = [C.c_void_p]
    assert L.zplt_generate(ctx) == 0
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "ic")
        dens = os.path.join(tmp, "density")
        os.makedirs(out)
        open(os.path.join(out, "ic_9999"), "w").write("stale")
        open(os.path.join(out, "zeldovich.0.0"), "w").write("stale")
        open(os.path.join(out, "keep.txt"), "w").write("not ours")
        assert L.zplt_write_outputs(ctx, os.fsencode(out), cpd, 0, None, -1) == 0, L.zplt_last_error()
        assert os.path.exists(os.path.join(out, "keep.txt")) and not os.path.exists(os.path.join(out, "zeldovich.0.0"))
        os.remove(os.path.join(out, "keep.txt"))
        _check_mock_files(out, ppd, cpd, rb)
        # records + density
        assert L.zplt_write_outputs(ctx, os.fsencode(out), cpd, 1, os.fsencode(dens), -1) == 0, L.zplt_last_error()
        _check_mock_files(out, ppd, cpd, rb)
        d = np.fromfile(dens, dtype=np.float32).reshape(ppd, ppd * ppd)
        assert np.array_equal(d[:, 0], np.arange(ppd, dtype=np.float32)) and np.all(d == d[:, :1])
        # density only: no ic files at all
        assert L.zplt_write_outputs(ctx, os.fsencode(out), cpd, 2, os.fsencode(dens), -1) == 0, L.zplt_last_error()
        assert not [f for f in os.listdir(out) if f.startswith("ic_")]
        assert os.path.getsize(dens) == 4 * ppd**3
        # one plane; a plane number beyond the grid writes nothing (the reference's loop never matches)
        assert L.zplt_write_outputs(ctx, os.fsencode(out), cpd, 0, None, ppd - 3) == 0, L.zplt_last_error()
        _check_mock_files(out, ppd, cpd, rb, planes=[ppd - 3])
        assert L.zplt_write_outputs(ctx, os.fsencode(out), cpd, 0, None, ppd + 5) == 0, L.zplt_last_error()
        assert not [f for f in os.listdir(out) if f.startswith("ic_")]
    L.zplt_destroy.argtypes = [C.c_void_p]
    L.zplt_destroy(ctx)


@pytest.mark.parametrize("store,nth", [("ram", 3), ("ram", 20), ("pageable", 2), ("pageable", 9), ("disk", 5), ("disk", 21)])
def test_out_of_core_copy_failure_is_reported(mocklib, store, nth):
    """A failing block copy (either pass, any store, also inside a copy thread of the pageable store) ends the run with the
    device half's message in the caller's thread, and leaves no block files behind."""
    with tempfile.TemporaryDirectory() as tmp:
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": "4", "ZPLT_OOC_STORE": store, "ZPLT_MOCK_FAIL_COPY": str(nth)},
                                      NP=32**3, CPD=4)
        assert rc == 2 and "mock: cudaMemcpy" in err, (rc, err)  # ZPLT_ECUDA
        assert not [f for f in os.listdir(out) if f.startswith("zeldovich.")]


def test_out_of_core_in_two_invocations(mocklib):
    """ZPLT_OOC_PART=1 then =2 (the reference's -DPART1 / -DPART2 builds, src/zeldovich.cpp:938-979): the first run leaves
    passes^2 block files with the reference's names and no ic files, the second turns them into the ic files and removes them."""
    ppd, cpd, G = 32, 5, 4
    with tempfile.TemporaryDirectory() as tmp:
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": str(G), "ZPLT_OOC_PART": "1"}, NP=ppd**3, CPD=cpd)
        assert rc == 0 and rep.ooc_part == 1 and rep.ooc_disk == 1, err
        blk = 16 * 2 * ppd**3 // G**2
        for s in range(G):
            assert sorted(os.listdir(os.path.join(out, f"zeldovich.{s}"))) == [f"zeldovich.{s}.{d}" for d in range(G)]
            assert all(os.path.getsize(os.path.join(out, f"zeldovich.{s}", f"zeldovich.{s}.{d}")) == blk for d in range(G))
        assert not [f for f in os.listdir(out) if f.startswith("ic_")] and rep.density_variance == 0
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": str(G), "ZPLT_OOC_PART": "2"}, NP=ppd**3, CPD=cpd)
        assert rc == 0 and rep.ooc_part == 2, err
        _check_mock_files(out, ppd, cpd, 32)  # the mock checks every value of every block it adopts
        assert rep.density_variance == ppd and rep.max_disp[0] == 0  # every plane emitted, nothing generated in this invocation
        assert not [f for f in os.listdir(out) if f.startswith("zeldovich.")]
        # pass 2 without the blocks of a pass 1, and the split without an explicit blocking, are errors
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PASSES": str(G), "ZPLT_OOC_PART": "2"}, NP=ppd**3, CPD=cpd)
        assert rc != 0 and "cannot read block file" in err
        rc, rep, out, err = _mock_run(mocklib, tmp, {"ZPLT_OOC_PART": "1", "ZPLT_MOCK_FREE_BYTES": str((1 << 30) + 16 * 2 * ppd**3 - 1)}, NP=ppd**3, CPD=cpd)
        assert rc != 0 and "ZPLT_OOC_PASSES" in err
