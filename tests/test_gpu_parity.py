"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
golden fixtures produced by the unmodified reference.

Bars (BASELINE.json north_star): RNG draws, mode ordering and particle ids bit-exact;
displacements and velocities within 1e-10, defined field-relative (max|a-b|/max|b| per
component); float32 ICFormats are compared after the cast and must agree
to one float ulp of the field scale (2e-7) — in practice they are almost always bit-equal.
"""
import os
import tempfile

import numpy as np
import pytest

import helpers
from __graft_entry__ import load_package, load_synth

pytestmark = pytest.mark.gpu

TOL = 1e-10
M = 65536


@pytest.fixture(scope="module")
def pkg():
    return load_package()


def ctx_from(pkg, P, power, rank=0, nranks=1):
    """Another context (e.g. another slab rank) from already parsed parameters and an already normalised spectrum."""
    cfg = P.config(device=0)
    cfg.rank, cfg.nranks = rank, nranks
    ctx = pkg.Context(cfg)
    power.apply(ctx)
    if P.qPLT:
        ctx.load_eigenmodes_file(P.PLT_filename)
    return ctx


def write_case_files(kw, pk, eig, **extra):
    """Parameter file (+ P(k) table, eigenmode file) of a keyword case in a fresh directory; returns the file's path."""
    synth = load_synth()
    tmp = tempfile.mkdtemp(prefix="zplt_")
    synth.write_power_table(os.path.join(tmp, "pk.pow"), pk[0], pk[1])
    fmtname = kw["icformat"]
    over = dict(NP=kw["ppd"] ** 3, BoxSize=repr(float(kw["boxsize"])), ZD_Seed=kw["seed"], ZD_k_cutoff=repr(float(kw["k_cutoff"])),
                ZD_CornerModes=kw.get("corner_modes", 0), ZD_qPLT=kw["qPLT"], ZD_qPLT_rescale=kw["qPLTrescale"],
                ZD_PLT_target_z=repr(float(kw["PLT_target_z"])), InitialRedshift=repr(float(kw["z_initial"])),
                ZD_f_cluster=repr(float(kw["f_cluster"])), ZD_qPk_fix_to_mean=kw["fixed_power"],
                ZD_Pk_norm=repr(float(kw["Pk_norm"])), ZD_Pk_smooth=repr(float(kw["Pk_smooth"])),
                ZD_Pk_scale=repr(float(kw["Pk_scale"])), ICFormat='"%s"' % fmtname, ZD_qonemode=kw.get("qonemode", 0),
                ZD_one_mode=" ".join(str(v) for v in kw.get("one_mode", (0, 0, 0))),
                ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"))
    if kw.get("f_NL", 0.0) != 0.0:
        over.update(ZD_f_NL=repr(float(kw["f_NL"])), ZD_n_s=repr(float(kw.get("n_s", 1.0))), Omega_M=repr(float(kw.get("Omega_M", 1.0))))
    if kw["Pk_sigma"] > 0:
        over["ZD_Pk_sigma"] = repr(float(kw["Pk_sigma"]))
    else:
        over["ZD_Pk_sigma"] = 0
        over["ZD_Pk_sigma_ratio"] = repr(float(kw["Pk_sigma_ratio"]))
    if eig is not None:
        synth_path = os.path.join(tmp, "eig.bin")
        with open(synth_path, "wb") as f:
            f.write(np.int32(eig[0]).tobytes())
            f.write(np.ascontiguousarray(eig[1], dtype=np.float64).tobytes())
        over["ZD_PLT_filename"] = '"%s"' % synth_path
    over.update(extra)
    return synth.write_param(os.path.join(tmp, "c.par"), **over)


def make_ctx(pkg, kw, pk, eig, rank=0, nranks=1):
    """Product context from keyword parameters; the spline comes from the product's own host code."""
    P = pkg.Parameters(write_case_files(kw, pk, eig))
    power = pkg.PowerSpectrum(P)
    cfg = P.config(device=0)
    cfg.rank, cfg.nranks = rank, nranks
    ctx = pkg.Context(cfg)
    power.apply(ctx)
    if eig is not None:
        ctx.load_eigenmodes_file(P.PLT_filename)
    return ctx, P, power


def default_kw(**over):
    kw = dict(ppd=32, boxsize=720.0, seed=12346, k_cutoff=1.0, corner_modes=0, qPLT=0, qPLTrescale=0, PLT_target_z=0.0,
              z_initial=49.0, f_cluster=1.0, fixed_power=0, Pk_norm=8.0, Pk_sigma=0.0210839935761, Pk_sigma_ratio=0.0,
              Pk_smooth=0.0, Pk_scale=1.0, icformat="RVZel")
    kw.update(over)
    return kw


def compare_records(oracle, got, want, tol64=TOL, tol32=2e-7):
    if "ijk" in want.dtype.names:
        assert np.array_equal(got["ijk"], want["ijk"]), "particle ids differ"
        assert np.all(got["pad"] == 0)
    worst = 0.0
    for f in ("displ", "vel"):
        if f in want.dtype.names:
            tol = tol32 if want[f].dtype == np.float32 else tol64
            for c in range(3):
                err = oracle.field_rel_err(got[f][:, c], want[f][:, c])
                assert err < tol, (f, c, err)
                worst = max(worst, err)
    return worst


# ---------------------------------------------------------------- RNG ---------------
def test_pcg64_device_known_answers(pkg):
    d = pkg.pcg_draws(12346, 0, 4)
    assert list(d) == [13376226141762278320, 13264298068723250620, 14189328008317063736, 6008591607947420752]
    assert pkg.pcg_draws(12346, 2 * M * M, 1)[0] == 14931042480954944222
    assert pkg.pcg_draws(12346, 2 * (3 * M * M + (M - 5) * M + 7), 1)[0] == 7757910958070640359


def test_pcg64_device_vs_oracle(pkg, oracle):
    for seed, off in ((0, 0), (12346, 123456789012345), (-7, (1 << 70) + 12345), ((1 << 63) + 5, 2 * 511 * M * M + 17)):
        assert np.array_equal(pkg.pcg_draws(seed, off, 64), oracle.pcg_draws(seed, off, 64))


@pytest.mark.parametrize("ppd", [16, 64, 256])
def test_mode_draws_bit_exact(pkg, oracle, ppd):
    """Closed-form RNG position of every mode == the reference's sequential nskip walk (via the oracle)."""
    ctx, P, power = make_ctx(pkg, default_kw(ppd=ppd), helpers.wmap_pk(), None)
    rng = np.random.RandomState(1)
    h = ppd // 2
    k = np.stack([rng.randint(-h + 1, h + 1, 4000), rng.randint(0, h, 4000), rng.randint(-h + 1, h + 1, 4000)], axis=1)
    k[:8] = [[0, 0, 0], [h, 0, 0], [-h + 1, h - 1, -h + 1], [h, h - 1, h], [-1, 0, -1], [1, 1, 1], [0, h - 1, 0], [-1, 3, h]]
    raw, u = ctx.mode_draws(k)
    for i in range(len(k)):
        kx, ky, kz = (int(v) for v in k[i])
        off = 2 * (ky * M * M + (kz % M) * M + (kx % M))
        want = oracle.pcg_draws(12346, off, 2)
        assert raw[i, 0] == want[0] and raw[i, 1] == want[1], (k[i], raw[i], want)
        assert u[i, 0] == oracle.one_rand(want[0]) and u[i, 1] == oracle.one_rand(want[1])
    ctx.close()


# ---------------------------------------------------------------- host boundary -----
def test_host_scalars_match_oracle(pkg, oracle):
    kw = default_kw(ppd=64)
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), None)
    s = oracle.power_scalars(oracle.make_config(**kw), helpers.wmap_pk())
    assert abs(power.normalization / s["normalization"] - 1) < 1e-14
    assert P.fundamental == 2.0 * np.pi / 720.0
    n = 3 * 32 * 32 + 1
    got = ctx.power_table(n)
    want = oracle.power_table(oracle.make_config(**kw), helpers.wmap_pk(), n)
    assert got[0] == 0.0
    assert np.max(np.abs(got[1:] / want[1:] - 1)) < 1e-13
    ctx.close()


# ---------------------------------------------------------------- FFT ---------------
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("row_mode", [True, False])
@pytest.mark.parametrize("variant", [0, 1])
def test_fft_matches_numpy(pkg, n, row_mode, variant):
    """variant 0: the kernels a default context launches (TMA-ring strided passes where they exist); 1: the plain kernels."""
    rng = np.random.RandomState(n)
    batch = 64
    shape = (batch, n) if row_mode else (n, batch)
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    got = pkg.fft_backward(a, row_mode, variant)
    want = np.fft.ifft(a, axis=1 if row_mode else 0) * n  # unnormalised backward, sign +1
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-14


@pytest.mark.parametrize("batch", [64, 8 * 37])
def test_fft2048_decimation_kernel_matches_numpy(pkg, batch):
    """The 8-pencil decimation-in-time kernel for the N = 2048 strided passes (csrc/zplt_fft2048_kernels.cu)."""
    n = 2048
    rng = np.random.RandomState(batch)
    a = rng.standard_normal((n, batch)) + 1j * rng.standard_normal((n, batch))
    got = pkg.fft_backward(a, False, 2)
    want = np.fft.ifft(a, axis=0) * n
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-14


def test_fft_linearity_and_impulse(pkg):
    n, batch = 1024, 32
    a = np.zeros((batch, n), dtype=np.complex128)
    for b in range(batch):
        a[b, (7 * b + 1) % n] = 1.0
    got = pkg.fft_backward(a, True)
    j = np.arange(n)
    for b in range(batch):
        want = np.exp(2j * np.pi * ((j * ((7 * b + 1) % n)) % n) / n)  # reduce the phase first
        assert np.max(np.abs(got[b] - want)) < 1e-13


# ---------------------------------------------------------------- spectral arrays ---
@pytest.mark.parametrize("case", [
    dict(ppd=32),
    dict(ppd=32, k_cutoff=2.0, corner_modes=1),
    dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, f_cluster=0.97, icformat="RVdoubleZel", eig=16),
    dict(ppd=32, qPLT=1, icformat="RVZel", eig=64),
    dict(ppd=32, fixed_power=1, seed=-3),
    # ZD_f_NL: the density of every mode comes from the transformed potential (no masked site stays zero)
    dict(ppd=32, f_NL=3000.0, n_s=0.96, Omega_M=0.3),
    dict(ppd=32, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16, f_NL=-1500.0, n_s=0.96, Omega_M=0.3),
])
def test_spectral_arrays_before_fft(pkg, oracle, case):
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    got = ctx.spectral()
    want = oracle.spectral_cube(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) / scale < 1e-13
    # masked sites are exactly zero in both
    assert np.array_equal(got == 0, want == 0)
    ctx.close()


# ---------------------------------------------------------------- the hot kernel ----
@pytest.mark.parametrize("case", [
    dict(ppd=32),
    dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, f_cluster=0.97, icformat="RVdoubleZel", eig=16),
    dict(ppd=128, qPLT=1, icformat="RVZel", eig=128),
    dict(ppd=256, k_cutoff=2.0, seed=-3),
    dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128),
    dict(ppd=64, f_NL=3000.0, n_s=0.96, Omega_M=0.3),
])
def test_hot_generation_kernel_before_fft(pkg, oracle, case):
    """The packed arrays as gen_xfft_kernel — the product's generation kernel — forms them (its x transform skipped):
    run walk of the generator, eigenmode cell sharing, zero-row skip, pencil builder, twin rows, ky = 0 plane."""
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    got = ctx.spectral_hot()
    want = oracle.spectral_cube(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) / scale < 1e-13
    assert np.array_equal(got == 0, want == 0)
    ctx.close()


@pytest.mark.parametrize("ppd,kc,na4", [(64, 1.0, 0), (256, 1.0, 0), (128, 2.0, 0), (512, 1.0, 1)])
def test_hot_generation_kernel_draws_bit_exact(pkg, oracle, ppd, kc, na4):
    """The raw 64-bit draws gen_xfft_kernel consumes, site by site, against the reference's sequential walk of the plane
    generators (oracle.pcg_draws at the closed-form position): bit-exact, in the kernel that ships.  512 with 4 arrays
    is the 8-pencil / runs-of-2 instantiation."""
    synth = load_synth()
    eig = (16, synth.make_eigmodes(16)) if na4 else None
    kw = default_kw(ppd=ppd, k_cutoff=kc, qPLT=na4, icformat="RVZel")
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    raw = ctx.hot_draws()  # [z][y][x][2]
    ctx.close()
    N, h = ppd, ppd // 2
    k1 = np.where(np.arange(N) > h, np.arange(N) - N, np.arange(N))
    kmax = int(h / kc + 0.5)
    fund2 = (2 * np.pi / 720.0) ** 2
    k2cut = (np.pi / (720.0 / N)) ** 2 / (kc * kc)
    rng = np.random.RandomState(ppd)
    rows = [(0, 0), (0, h), (1, N - 1), (h - 1, h + 1), (3, 5)] + [(int(rng.randint(0, h)), int(rng.randint(0, N))) for _ in range(40)]
    checked = 0
    for y, z in rows:
        if y == 0 and z > h:
            # these entries are the twins of row (0, N - z) (reference src/zeldovich.cpp:485-503 overwrites what it drew here)
            assert not raw[z, y].any()
            continue
        kz = int(k1[z])
        n2 = k1.astype(np.int64) ** 2 + y * y + kz * kz
        masked = (np.abs(k1) == kmax) | (abs(kz) == kmax) | (y == kmax) | (n2 * fund2 >= k2cut)
        if masked.all():
            assert not raw[z, y].any()  # the kernel skips all-masked rows without drawing
            continue
        # the reference's walk of this row: kx = 0..N/2 are consecutive draw pairs, then kx = -N/2+1..-1 at (kx mod 65536)
        base = 2 * (y * M * M + (kz % M) * M)
        want = np.empty((N, 2), dtype=np.uint64)
        want[:h + 1] = oracle.pcg_draws(12346, base, 2 * (h + 1)).reshape(-1, 2)
        want[h + 1:] = oracle.pcg_draws(12346, base + 2 * (M - h + 1), 2 * (h - 1)).reshape(-1, 2)
        # every site of a run that holds an unmasked site consumes its draws; runs are 16/NP sites (2, 4 or 8) — compare the unmasked ones
        assert np.array_equal(raw[z, y][~masked], want[~masked]), (y, z)
        checked += int((~masked).sum())
    assert checked > 500


# ---------------------------------------------------------------- benchmark sizes ---
_PLANES = {}


def oracle_planes(oracle, kw, zs, eig):
    """oracle.planes, remembered per configuration (the large sizes cost the CPU a minute)."""
    key = (tuple(sorted((k, str(v)) for k, v in kw.items())), tuple(zs), eig[0] if eig else None)
    if key not in _PLANES:
        _PLANES.clear()  # one entry: the records of four PPD=2048 planes are 0.5 GB
        _PLANES[key] = oracle.planes(oracle.make_config(**kw), helpers.wmap_pk(), zs, eig)
    return _PLANES[key]


def plane_check(pkg, oracle, ctx, kw, eig, zs, zlocal0=0, tol64=TOL):
    """Selected planes of a generated context against the plane oracle (direct z summation, oracle.planes)."""
    want, wst = oracle_planes(oracle, kw, zs, eig)
    worst = 0.0
    for i, z in enumerate(zs):
        ctx.reset_stats()
        got = ctx.fetch_planes(z - zlocal0, 1)
        worst = max(worst, compare_records(oracle, got, want[i].reshape(-1), tol64=tol64))
        st = ctx.stats()
        assert abs(st["density_variance"] / wst[i]["density_variance"] - 1) < 1e-9
        assert np.allclose(st["max_disp"], wst[i]["max_disp"], rtol=1e-9)
    return worst


@pytest.mark.parametrize("case", [
    # BASELINE configs[3] exactly — the bench.py workload: PPD=1024 qPLT + rescale, RVZel
    dict(ppd=1024, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128),
    # the same arrays through the double-precision records (configs[1]'s format)
    dict(ppd=1024, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=128),
    # PPD=512 qPLT: the 8-pencil / runs-of-2 generation kernel, ring kernels at N=512
    dict(ppd=512, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128),
    # configs[2]'s large run: PPD=512 with ZD_k_cutoff=2, ZA
    dict(ppd=512, k_cutoff=2.0, icformat="RVdoubleZel"),
])
def test_benchmark_configs_vs_plane_oracle(pkg, oracle, case):
    """The benchmark configurations against the CPU oracle: planes 0, 1, N/2-1 and N-1 of the full-size runs.
    ids exact; float32 fields to one ulp of the field scale (2e-7), double fields to 1e-10."""
    import torch

    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    N = kw["ppd"]
    need = (4 if kw["qPLT"] else 2) * 16 * N**3 + (8 << 30)
    if torch.cuda.mem_get_info()[0] < need:
        pytest.skip("not enough free device memory")
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    ctx.generate()
    worst = plane_check(pkg, oracle, ctx, kw, eig, [0, 1, N // 2 - 1, N - 1])
    print(case, "worst field-relative error", worst)
    ctx.close()


# ---------------------------------------------------------------- full path ---------
@pytest.mark.parametrize("name", sorted(helpers.golden_cases().keys()))
def test_full_path_matches_golden(pkg, oracle, name):
    """CUDA path vs ic_* records written by the unmodified reference."""
    case, raw, eig, _ = helpers.load_golden(name)
    kw = default_kw(**helpers.params_to_kwargs(case["params"]))
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    ctx.generate()
    got = ctx.fetch_planes(0, kw["ppd"])
    gold = raw.view(ctx.record_dtype)
    compare_records(oracle, got, gold, tol64=1e-10)
    st = ctx.stats()
    N = kw["ppd"]
    assert abs(np.sqrt(st["density_variance"] / N**3) - case["stderr"]["rms_density"]) < 1e-6
    assert np.allclose(st["max_disp"], case["stderr"]["max_disp"], rtol=2e-6)
    ctx.close()


@pytest.mark.parametrize("case", [
    dict(ppd=64, icformat="RVZel"),
    dict(ppd=128, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=128),
    dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=128),
    dict(ppd=128, k_cutoff=2.0, icformat="Zeldovich"),
    dict(ppd=64, icformat="ZelSimple", boxsize=250.0, seed=77),
    # local primordial non-Gaussianity (reference src/zeldovich.cpp:699-790, 945-960), interpolated and direct eigenmodes
    dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16, f_NL=2500.0, n_s=0.96, Omega_M=0.3),
    dict(ppd=128, icformat="RVdoubleZel", f_NL=-4000.0, n_s=0.97, Omega_M=0.31, k_cutoff=2.0),
    dict(ppd=256, qPLT=1, icformat="RVZel", eig=256, f_NL=1000.0, n_s=0.96, Omega_M=0.3),
])
def test_full_path_matches_oracle(pkg, oracle, case):
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    ctx.generate()
    got = ctx.fetch_planes(0, kw["ppd"])
    want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    worst = compare_records(oracle, got, want)
    st = ctx.stats()
    assert abs(st["density_variance"] / wst["density_variance"] - 1) < 1e-10
    assert np.allclose(st["max_disp"], wst["max_disp"], rtol=1e-10)
    print(case, "worst field-relative error", worst)
    ctx.close()


@pytest.mark.parametrize("small", [64, 256])
def test_oversampled_pair_phase_matched(pkg, oracle, small):
    """BASELINE config 3 (PPD=256 then PPD=512 with ZD_k_cutoff=2): PPD=N and PPD=2N with ZD_k_cutoff=2 carry the same
    modes with the same phases, so the ZA displacements of the oversampled run at even sites are those of the small run.
    (With qPLT only the phases match: the eigenmodes are a function of k relative to each run's own lattice.)"""
    pk = helpers.wmap_pk()
    big = 2 * small
    recs = {}
    for ppd, kc in ((small, 1.0), (big, 2.0)):
        ctx, P, power = make_ctx(pkg, default_kw(ppd=ppd, k_cutoff=kc, icformat="Zeldovich"), pk, None)
        ctx.generate()
        recs[ppd] = ctx.fetch_planes(0, ppd)
        ctx.close()
    b = recs[big].reshape(big, big, big)[::2, ::2, ::2].reshape(-1)
    for c in range(3):
        assert oracle.field_rel_err(b["displ"][:, c], recs[small]["displ"][:, c]) < 1e-11


def test_plane_ranges_and_file_writer(pkg, oracle):
    kw = default_kw(ppd=32, icformat="RVZel")
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), None)
    ctx.generate()
    whole = ctx.fetch_planes(0, 32)
    part = ctx.fetch_planes(5, 9)
    assert np.array_equal(part.view(np.uint8), whole.reshape(32, -1)[5:14].reshape(-1).view(np.uint8))
    with tempfile.TemporaryDirectory() as tmp:
        ctx.write_ic_files(tmp, 5)  # cpd < ppd: several planes share a file, ascending-z append order
        rec = oracle.read_ic_dir(tmp, 32, 5, "RVZel")
        assert np.array_equal(rec.view(np.uint8), whole.view(np.uint8))
        assert sorted(os.listdir(tmp)) == sorted({"ic_%d" % (z * 5 // 32) for z in range(32)})
    ctx.close()


def test_cli_matches_reference_layout(pkg, oracle):
    """./zeldovich <param_file> writes the same files the reference writes (golden case)."""
    import subprocess

    case, raw, eig, text = helpers.load_golden("plt16_interp_rvdouble")
    synth = load_synth()
    k, p = helpers.wmap_pk()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
        with open(os.path.join(tmp, "eig.bin"), "wb") as f:
            f.write(np.int32(eig[0]).tobytes())
            f.write(np.ascontiguousarray(eig[1]).tobytes())
        with open(os.path.join(tmp, "case.par"), "w") as f:
            f.write(text)
        r = subprocess.run([pkg.CLI_PATH, "case.par"], cwd=tmp, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        rec = oracle.read_ic_dir(os.path.join(tmp, "ic_out"), 16, int(case["params"]["CPD"]), "RVdoubleZel")
        gold = raw.view(rec.dtype)
        compare_records(oracle, rec, gold)
        assert "rms density variation of the pixels is 0.007389" in r.stderr
    r = subprocess.run([pkg.CLI_PATH], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr


def test_errors_are_reported(pkg):
    with pytest.raises(pkg.ZpltError):
        pkg.Context(pkg.make_config(45))  # odd: the reference asserts an even ppd (src/block_array.cpp:38)
    with pytest.raises(pkg.ZpltError):
        pkg.Context(pkg.make_config(1536))  # not a power of two and beyond the general path's range
    ctx = pkg.Context(pkg.make_config(32))
    with pytest.raises(pkg.ZpltError):
        ctx.generate()  # power spectrum not set
    ctx.close()


# ---------------------------------------------------------------- ppd not a power of two
@pytest.mark.parametrize("case", [
    dict(ppd=24, icformat="RVZel"),
    dict(ppd=80, k_cutoff=2.0, corner_modes=1, icformat="Zeldovich", seed=-11),
    dict(ppd=96, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, f_cluster=0.97, icformat="RVdoubleZel", eig=128),
    dict(ppd=48, qPLT=1, icformat="RVZel", eig=16),
    dict(ppd=20, icformat="ZelSimple"),
])
def test_general_ppd_full_path(pkg, oracle, case):
    """Even ppd that is not a power of two (the reference takes any even ppd, src/block_array.cpp:38-40; its production
    grids are 2^a 3^b): the general path — plain generation kernel, Bluestein transforms on the power-of-two kernels, unfused
    emission — against the oracle (direct DFT).  Same bars as everywhere: ids exact, fields 1e-10 / one float ulp."""
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    got_spec = ctx.spectral()
    want_spec = oracle.spectral_cube(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    assert np.max(np.abs(got_spec - want_spec)) / np.max(np.abs(want_spec)) < 1e-13
    ctx.generate()
    got = ctx.fetch_planes(0, kw["ppd"])
    want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    worst = compare_records(oracle, got, want)
    st = ctx.stats()
    assert abs(st["density_variance"] / wst["density_variance"] - 1) < 1e-10
    assert np.allclose(st["max_disp"], wst["max_disp"], rtol=1e-10)
    print(case, "worst field-relative error", worst)
    ctx.close()


def test_general_ppd_768_planes(pkg, oracle):
    """ppd = 768 = 2^8 * 3 (an Abacus-style grid) with qPLT + rescale, RVZel: planes against the plane oracle."""
    import torch

    if torch.cuda.mem_get_info()[0] < (40 << 30):
        pytest.skip("needs ~30 GB of free device memory")
    kw = default_kw(ppd=768, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel")
    synth = load_synth()
    eig = (128, synth.make_eigmodes(128))
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    ctx.generate()
    worst = plane_check(pkg, oracle, ctx, kw, eig, [0, 383, 767])
    print("ppd=768 worst field-relative error", worst)
    ctx.close()


# ---------------------------------------------------------------- slab decomposition
@pytest.mark.parametrize("G,case", [
    (2, dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16)),
    (4, dict(ppd=64, icformat="RVZel")),
    (8, dict(ppd=128, qPLT=1, icformat="RVZel", eig=128)),
])
def test_slab_decomposition_single_process(pkg, oracle, G, case):
    """All G ranks of a slab-decomposed run emulated on one GPU (device copies stand in for the all-to-all):
    the concatenated planes must equal the oracle's records."""
    import torch

    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    N = kw["ppd"]
    ctxs, bufs = [], []
    for r in range(G):
        ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig, rank=r, nranks=G)
        buf = torch.empty(ctx.workspace_bytes() // 8, dtype=torch.float64, device="cuda:0")
        ctx.set_workspace(buf.data_ptr(), buf.numel() * 8)
        ctx.generate()
        ctx.synchronize()
        ctxs.append(ctx)
        bufs.append(buf)
    half = ctxs[0].narray * N**3 // G * 2  # float64 elements of the stage-1 buffer; the receive buffer follows it
    blk = half // G
    for dst in range(G):
        for src in range(G):
            bufs[dst][half + src * blk: half + (src + 1) * blk] = bufs[src][dst * blk:(dst + 1) * blk]
    torch.cuda.synchronize()
    parts = []
    var, md = 0.0, np.zeros(3)
    for r in range(G):
        ctxs[r].exchange_done()
        parts.append(ctxs[r].fetch_planes(0, N // G))
        st = ctxs[r].stats()
        var += st["density_variance"]
        md = np.where(np.abs(st["max_disp"]) > np.abs(md), st["max_disp"], md)
        ctxs[r].close()
    got = np.concatenate(parts)
    want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    compare_records(oracle, got, want)
    assert abs(var / wst["density_variance"] - 1) < 1e-10
    assert np.allclose(md, wst["max_disp"], rtol=1e-10)


@pytest.mark.parametrize("passes,store,case", [
    (4, "ram", dict(ppd=64, icformat="RVZel")),
    (2, "disk", dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16)),
    (8, "ram", dict(ppd=128, qPLT=1, icformat="RVZel", eig=128)),
    (16, "disk", dict(ppd=256, k_cutoff=2.0, icformat="Zeldovich")),
    (2, "pageable", dict(ppd=128, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),  # 32 MiB chunks, 2 copy threads
])
def test_out_of_core_run_matches_oracle(pkg, oracle, monkeypatch, passes, store, case):
    """zplt_run_param_file out of core (the reference's -DDISK mode, src/block_array.cpp:129-382): one context plays the slab
    ranks one after the other (zplt_slab_set_rank / zplt_exchange_adopt), the blocks wait in host memory or in files; HBM
    holds 2/passes of the cube.  The ic files must hold the oracle's records."""
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    N, cpd = kw["ppd"], 5
    out = tempfile.mkdtemp(prefix="zplt_ooc_")
    par = write_case_files(kw, helpers.wmap_pk(), eig, InitialConditionsDirectory='"%s"' % out, CPD=cpd)
    monkeypatch.setenv("ZPLT_OOC_PASSES", str(passes))
    monkeypatch.setenv("ZPLT_OOC_STORE", store)
    rep = pkg.run_param_file(par, device=0)
    assert rep.ooc_passes == passes and rep.ooc_disk == (store == "disk")
    assert rep.ooc_bytes == 16 * (4 if kw["qPLT"] else 2) * N**3
    got = oracle.read_ic_dir(out, N, cpd, kw["icformat"])
    want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    compare_records(oracle, got, want)
    assert abs(rep.density_variance / wst["density_variance"] - 1) < 1e-10
    assert np.allclose(np.array(rep.max_disp[:]), wst["max_disp"], rtol=1e-10)
    assert sorted(os.listdir(out)) == sorted({"ic_%d" % (z * cpd // N) for z in range(N)})  # no block files left behind
    # the same parameter file with the cube resident gives the same bytes
    monkeypatch.setenv("ZPLT_OOC_PASSES", "0")
    rep1 = pkg.run_param_file(par, device=0)
    assert rep1.ooc_passes == 0
    again = oracle.read_ic_dir(out, N, cpd, kw["icformat"])
    compare_records(oracle, again, want)
    assert np.array_equal(again.view(np.uint8), got.view(np.uint8))  # same kernels, same arithmetic: not one bit differs
    import shutil

    shutil.rmtree(out, ignore_errors=True)


@pytest.mark.parametrize("G,opts,case", [
    (2, {}, dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16)),
    (4, {}, dict(ppd=64, icformat="RVZel")),
    (8, {}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    (8, {"slab_ring": 0, "yring": 0, "slab_groups": 3}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    (8, {"p2p_resident": 0}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    (4, {"p2p_resident": 0, "slab_ring": 0}, dict(ppd=128, icformat="RVdoubleZel")),
    (4, {"slab_ring": 2}, dict(ppd=512, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    # receive layouts of the fused exchange: per-source blocks, padded planes
    (4, {"b2_layout": 1}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    (4, {"b2_layout": 0}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    (2, {"b2_layout": 1}, dict(ppd=128, qPLT=1, icformat="RVdoubleZel", eig=16)),
    (8, {"b2_pad": 520}, dict(ppd=256, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=128)),
    (2, {"b2_pad": 8192, "p2p_resident": 0}, dict(ppd=128, icformat="RVZel")),
    (2, {"slab_groups": 1, "p2p_ctas": 0}, dict(ppd=128, k_cutoff=2.0, icformat="Zeldovich")),
    (4, {}, dict(ppd=512, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel", eig=128)),
    # ZD_f_NL on slab ranks: the potential pass with its own two exchanges (reference src/zeldovich.cpp:699-790, 945-960)
    (2, {}, dict(ppd=64, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16, f_NL=2500.0, n_s=0.96, Omega_M=0.3)),
    (4, {}, dict(ppd=128, icformat="RVdoubleZel", f_NL=-4000.0, n_s=0.97, Omega_M=0.31, k_cutoff=2.0)),
    (8, {"slab_ring": 0}, dict(ppd=256, qPLT=1, icformat="RVZel", eig=256, f_NL=1000.0, n_s=0.96, Omega_M=0.3)),
])
def test_fused_exchange_single_gpu(pkg, oracle, G, opts, case):
    """The product's multi-GPU stage 1 on one GPU: every rank of a G-rank run executes the grouped, overlapped generation
    and fft_tile_p2p(_ring)_kernel — the z pass that stores straight into the owners' receive buffers — with the other
    ranks' buffers on the same device standing in for NVLink peers (zplt_dbg_set_peers).  Records vs the oracle."""
    import torch

    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    N = kw["ppd"]
    ctx0, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig, rank=0, nranks=G)
    ctxs = [ctx0] + [ctx_from(pkg, P, power, r, G) for r in range(1, G)]
    bufs = [torch.empty(c.workspace_bytes() // 8, dtype=torch.float64, device="cuda:0") for c in ctxs]
    half_bytes = 16 * ctxs[0].narray * N**3 // G  # the receive buffer follows the stage-1 buffer
    for c, b in zip(ctxs, bufs):
        for k, v in opts.items():
            c.set_option(k, v)
        c.set_workspace(b.data_ptr(), b.numel() * 8)
        b.fill_(float("nan"))  # every row of every receive buffer must be written by somebody
    torch.cuda.synchronize()  # the fills run on torch's stream, the library on its own (non-blocking) streams
    for c in ctxs:
        c.dbg_set_peers([b.data_ptr() + half_bytes for b in bufs])
    if kw.get("f_NL", 0.0) != 0.0:
        # the potential pass of slab ranks: two more exchanges, a barrier (here: a device synchronisation) after each stage
        for c in ctxs:
            c.potential_begin()
        torch.cuda.synchronize()
        for c in ctxs:
            c.potential_exchange()
        torch.cuda.synchronize()
    for c in ctxs:
        # one rank at a time: the z pass of a rank is resident on half of the SMs until its generation kernels are done, and
        # several contexts doing that on ONE GPU at once (only this emulation does) would leave no SM to the generation kernels
        c.generate()
        c.synchronize()
    torch.cuda.synchronize()
    parts, var, md = [], 0.0, np.zeros(3)
    for c in ctxs:
        c.exchange_done()
        parts.append(c.fetch_planes(0, N // G))
        st = c.stats()
        var += st["density_variance"]
        md = np.where(np.abs(st["max_disp"]) > np.abs(md), st["max_disp"], md)
    got = np.concatenate(parts)
    if N <= 256:
        want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
        compare_records(oracle, got, want)
        assert abs(var / wst["density_variance"] - 1) < 1e-10
        assert np.allclose(md, wst["max_disp"], rtol=1e-10)
    else:
        zs = [0, N // G - 1, N // G, N // 2 + 1, N - 1]
        want, _ = oracle_planes(oracle, kw, zs, eig)
        g3 = got.reshape(N, N * N)
        for i, z in enumerate(zs):
            compare_records(oracle, g3[z], want[i].reshape(-1))
    for c in ctxs:
        c.close()
    del bufs
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dit,dit_emit", [(1, 0), (0, 0), (1, 1)])
def test_c5_rank_of_eight_at_ppd2048(pkg, oracle, dit, dit_emit):
    """BASELINE configs[4] (the north-star size): PPD=2048 qPLT + rescale RVZel over 8 slab ranks.  One GPU cannot hold
    the run, but it can hold ONE rank's buffers: the 8 ranks run their stage 1 one after the other in the same workspace,
    each storing only the share of the rank under test (the other peers are NULL = discarded), which then runs its
    stage 2.  Every N = 2048 kernel of the product (generation, z pass + exchange — 4-pencil and 8-pencil decimation
    forms —, y pass + emission) faces the oracle: planes 0 and 255 of rank 0, 1792 and 2047 of rank 7."""
    import torch

    N, G = 2048, 8
    kw = default_kw(ppd=N, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel")
    synth = load_synth()
    eig = (128, synth.make_eigmodes(128))
    ctx0, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig, rank=0, nranks=G)
    ws = ctx0.workspace_bytes()
    if torch.cuda.mem_get_info()[0] < ws + (6 << 30):
        ctx0.close()
        pytest.skip("needs ~145 GB of free device memory")
    W = torch.empty(ws // 8, dtype=torch.float64, device="cuda:0")
    open_ctxs = []
    try:
        slab_bytes = 16 * 4 * N**3 // G  # the stage-1 buffer; the receive buffer follows it
        recv = W.data_ptr() + slab_bytes
        zs = [0, N // G - 1, N - N // G, N - 1]
        want, _ = oracle_planes(oracle, kw, zs, eig)
        ctx0.close()
        worst = 0.0
        for target, planes in ((0, (0, 1)), (G - 1, (2, 3))):
            W[slab_bytes // 8:].fill_(float("nan"))
            torch.cuda.synchronize()  # the fill runs on torch's stream, the library on its own (non-blocking) streams
            tctx = None
            for src in range(G):
                c = ctx_from(pkg, P, power, src, G)
                open_ctxs.append(c)
                c.set_option("dit2048", dit)  # (1, 0) are the defaults
                c.set_option("dit2048_emit", dit_emit)
                c.set_workspace(W.data_ptr(), ws)
                c.dbg_set_peers([recv if r == target else None for r in range(G)])
                c.generate()
                c.synchronize()
                if src == target:
                    tctx = c
                else:
                    c.close()
            tctx.exchange_done()
            for i in planes:
                got = tctx.fetch_planes(zs[i] - target * (N // G), 1)
                worst = max(worst, compare_records(oracle, got, want[i].reshape(-1)))
            tctx.close()
    finally:
        for c in open_ctxs:
            c.close()
        W = None
        torch.cuda.empty_cache()
    print("PPD=2048 qPLT+rescale RVZel, 8 ranks, dit2048 =", dit, "dit2048_emit =", dit_emit, ": worst field-relative error", worst)


# ---------------------------------------------------------------- option coverage ---
@pytest.mark.parametrize("case", [
    dict(ppd=32, qonemode=1, one_mode=(3, 2, -5)),
    dict(ppd=32, qonemode=1, one_mode=(0, 0, 7), qPLT=1, icformat="RVdoubleZel", eig=16),
    dict(ppd=64, k_cutoff=2.0, corner_modes=1, icformat="Zeldovich"),
    dict(ppd=32, corner_modes=1, qPLT=1, qPLTrescale=1, PLT_target_z=2.0, f_cluster=0.9, icformat="RVZel", eig=32),
    dict(ppd=64, Pk_smooth=2.5, Pk_scale=0.7, icformat="ZelSimple"),
    dict(ppd=16, icformat="RVdoubleZel"),
])
def test_options_full_path(pkg, oracle, case):
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    kw = default_kw(**case)
    synth = load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    ctx.generate()
    got = ctx.fetch_planes(0, kw["ppd"])
    want, wst = oracle.run(oracle.make_config(**kw), helpers.wmap_pk(), eig)
    compare_records(oracle, got, want)
    assert abs(ctx.stats()["density_variance"] - wst["density_variance"]) <= 1e-10 * max(wst["density_variance"], 1e-300)
    ctx.close()


def test_power_law_full_path(pkg, oracle):
    synth = load_synth()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_param(os.path.join(tmp, "c.par"), NP=32**3, ZD_Pk_filename='""', ZD_Pk_powerlaw_index="-1.5", ZD_Pk_sigma=0,
                          ZD_Pk_sigma_ratio="0.01", ZD_Pk_norm="4.0", ICFormat='"RVdoubleZel"', ZD_Seed=5)
        P = pkg.Parameters(os.path.join(tmp, "c.par"))
        power = pkg.PowerSpectrum(P)
        ctx = pkg.Context(P.config(device=0))
        power.apply(ctx)
        ctx.generate()
        got = ctx.fetch_planes(0, 32)
        ctx.close()
    cfg = oracle.make_config(32, seed=5, is_powerlaw=1, powerlaw_index=-1.5, Pk_sigma=0.0, Pk_sigma_ratio=0.01, Pk_norm=4.0,
                             icformat="RVdoubleZel")
    want, _ = oracle.run(cfg, None)
    compare_records(oracle, got, want)


def test_cli_qoneslab(pkg, oracle):
    import subprocess

    synth = load_synth()
    k, p = helpers.wmap_pk()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
        synth.write_param(os.path.join(tmp, "c.par"), NP=16**3, ZD_Pk_filename='"pk.pow"', ZD_qoneslab=5, CPD=16)
        r = subprocess.run([pkg.CLI_PATH, "c.par"], cwd=tmp, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        assert os.listdir(os.path.join(tmp, "ic_out")) == ["ic_5"]
        rec = np.fromfile(os.path.join(tmp, "ic_out", "ic_5"), dtype=pkg.RECORD_DTYPES[1])
    want, _ = oracle.run(oracle.make_config(16), (k, p))
    compare_records(oracle, rec, want.reshape(16, -1)[5])


def test_density_planes_qdensity(pkg, oracle):
    """ZD_qdensity: float32 Re(A0) planes next to (1) or instead of (2) the records, through API and CLI."""
    import subprocess

    synth = load_synth()
    kw = default_kw(ppd=32, icformat="RVZel")
    pk = helpers.wmap_pk()
    ctx, P, power = make_ctx(pkg, kw, pk, None)
    ctx.generate()
    rec, dens = ctx.fetch_planes_density(0, 32)
    _, dens_only = ctx.fetch_planes_density(3, 5, records=False)
    ctx.close()
    cube = oracle.fft3_backward(oracle.spectral_cube(oracle.make_config(**kw), pk))
    want = cube[0].real
    assert np.max(np.abs(dens - want.astype(np.float32))) <= 2e-7 * np.max(np.abs(want))
    assert np.array_equal(dens_only, dens[3:8])
    wrec, _ = oracle.run(oracle.make_config(**kw), pk)
    compare_records(oracle, rec, wrec)
    for qd in (1, 2):
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_power_table(os.path.join(tmp, "pk.pow"), pk[0], pk[1])
            synth.write_param(os.path.join(tmp, "c.par"), NP=32**3, ZD_Pk_filename='"pk.pow"', ZD_qdensity=qd)
            r = subprocess.run([pkg.CLI_PATH, "c.par"], cwd=tmp, stderr=subprocess.PIPE, text=True)
            assert r.returncode == 0, r.stderr
            files = sorted(os.listdir(os.path.join(tmp, "ic_out")))
            assert "density32" in files
            assert (len(files) == 1) == (qd == 2)  # qdensity = 2 writes no ic_* files
            d = np.fromfile(os.path.join(tmp, "ic_out", "density32"), dtype=np.float32).reshape(32, 32, 32)
            assert np.array_equal(d, dens)
            assert ("maximum component-wise" in r.stderr) == (qd == 1)


def test_full_size_oversampling_property(pkg, oracle):
    """BASELINE full size on one GPU: PPD=1024 with ZD_k_cutoff=2 sampled at even lattice sites carries exactly the
    modes of PPD=512 (size-independent property; the oracle is too slow at this size).  Also: ids, zero padding,
    emission is idempotent (the cube is not modified by zplt_emit_planes)."""
    import torch

    if torch.cuda.mem_get_info()[0] < (60 << 30):
        pytest.skip("needs ~40 GB of free device memory")
    pk = helpers.wmap_pk()
    big, P, _ = make_ctx(pkg, default_kw(ppd=1024, k_cutoff=2.0), pk, None)
    big.generate()
    small, P2, _ = make_ctx(pkg, default_kw(ppd=512), pk, None)
    small.generate()
    worst = 0.0
    for z in (0, 2, 510, 1022):
        a = big.fetch_planes(z, 1).reshape(1024, 1024)
        assert np.array_equal(a["ijk"][..., 0], np.full((1024, 1024), z))
        assert np.array_equal(a["ijk"][..., 1], np.arange(1024)[:, None].repeat(1024, 1))
        assert np.array_equal(a["ijk"][..., 2], np.arange(1024)[None, :].repeat(1024, 0))
        assert np.all(a["pad"] == 0)
        again = big.fetch_planes(z, 1)
        assert np.array_equal(again.view(np.uint8), a.reshape(-1).view(np.uint8))
        b = small.fetch_planes(z // 2, 1).reshape(512, 512)
        sub = a[::2, ::2]
        for f in ("displ", "vel"):
            worst = max(worst, float(np.abs(sub[f].astype(np.float64) - b[f].astype(np.float64)).max() / np.abs(b[f]).max()))
    assert worst < 2e-7, worst
    big.close()
    small.close()


def test_kernel_variants_agree_at_full_size(pkg):
    """BASELINE full size, the benchmark configuration (PPD=1024 qPLT+rescale RVZel): the ring-prefetched kernels and the
    256-bit record stores must produce what the plain one-tile-per-CTA kernels produce (same transform code, different data
    movement) — ids byte for byte, fields to one float32 ulp of the field scale."""
    import torch

    if torch.cuda.mem_get_info()[0] < (100 << 30):
        pytest.skip("needs ~75 GB of free device memory")
    synth = load_synth()
    eig = (128, synth.make_eigmodes(128))
    kw = default_kw(ppd=1024, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVZel")
    ctx, P, power = make_ctx(pkg, kw, helpers.wmap_pk(), eig)
    planes = (0, 1, 511, 1023)
    defaults = {"zring": 12, "yring": 12, "wide_records": -1}

    def run(opts):
        for k, v in {**defaults, **opts}.items():
            ctx.set_option(k, v)
        ctx.generate()
        return [ctx.fetch_planes(z, 1).copy() for z in planes]

    ref = run({"zring": 0, "yring": 0, "wide_records": 0})
    for opts in ({}, {"wide_records": 0}, {"yring": 0}, {"zring": 0}):
        got = run(opts)
        for a, b in zip(got, ref):
            assert np.array_equal(a["ijk"], b["ijk"]) and np.all(a["pad"] == 0), opts
            for f in ("displ", "vel"):
                scale = float(np.abs(b[f]).max())
                assert float(np.abs(a[f].astype(np.float64) - b[f].astype(np.float64)).max()) <= 2e-7 * scale, (opts, f)
    ctx.close()
