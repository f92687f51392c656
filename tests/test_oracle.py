"""CPU tests that pin the oracle (oracle/zel_oracle.c) to the reference.

The reference repository has no tests or golden vectors of its own, so
the pins are: the PCG64 known answers recorded in SURVEY.md §4, the golden ic_* records
in tests/golden/ produced by the unmodified reference sources (make_golden.py), and —
when oracle/_ref/zeldovich_ref is present — a live run of that binary.
"""
import os
import tempfile

import numpy as np
import pytest

import helpers

M = 65536


def test_pcg64_known_answers(oracle):
    d = oracle.pcg_draws(12346, 0, 4)
    assert list(d) == [13376226141762278320, 13264298068723250620, 14189328008317063736, 6008591607947420752]
    assert oracle.pcg_draws(12346, 2 * M * M, 1)[0] == 14931042480954944222
    assert oracle.pcg_draws(12346, 2 * (3 * M * M + (M - 5) * M + 7), 1)[0] == 7757910958070640359


def test_pcg64_matches_numpy(oracle):
    # numpy's PCG64 is the same setseq_xsl_rr_128_64 generator; inject the seeded state
    mult = (2549297995355413924 << 64) | 4865540595714422341
    inc = (6364136223846793005 << 64) | 1442695040888963407
    mask = (1 << 128) - 1
    for seed in (0, 1, 12346, (1 << 64) - 7):
        bg = np.random.PCG64(0)
        st = bg.state
        st["state"] = {"state": ((seed + inc) * mult + inc) & mask, "inc": inc}
        bg.state = st
        bg.advance(123456789012345)
        want = bg.random_raw(16)
        got = oracle.pcg_draws(seed, 123456789012345, 16)
        assert np.array_equal(want, got)


def test_one_rand_range(oracle):
    assert oracle.one_rand((1 << 64) - 1) == 1.0
    assert oracle.one_rand(0) == 2.0**-64
    assert oracle.one_rand((1 << 63) - 1) == 0.5
    assert oracle.one_rand((1 << 64) - 2) == 1.0  # rounds to nearest even


def test_normalisation_scalar(oracle):
    cfg = oracle.make_config(64)
    s = oracle.power_scalars(cfg, helpers.wmap_pk())
    # sigma(8) after normalisation must come back as ZD_Pk_sigma
    assert abs(s["sigma_check"] - 0.0210839935761) < 1e-12
    # reference prints "Input sigma(8.000000) = 0.0781753" (golden stderr)
    sig_in = 0.0210839935761 / np.sqrt(s["normalization"] * 720.0**3)
    assert abs(sig_in - 0.0781753) < 5e-8


@pytest.mark.parametrize("name", sorted(helpers.golden_cases().keys()))
def test_oracle_matches_golden(oracle, name):
    case, raw, eig, _ = helpers.load_golden(name)
    kw = helpers.params_to_kwargs(case["params"])
    cfg = oracle.make_config(**kw)
    rec, st = oracle.run(cfg, helpers.wmap_pk(), eig)
    gold = raw.view(oracle.RECORD_DTYPES[cfg.icformat])
    if "ijk" in gold.dtype.names:
        assert np.array_equal(rec["ijk"], gold["ijk"])
    for f in ("displ", "vel"):
        if f in gold.dtype.names:
            for c in range(3):
                err = oracle.field_rel_err(rec[f][:, c], gold[f][:, c])
                tol = 1e-6 if gold[f].dtype == np.float32 else 1e-12
                assert err < tol, (name, f, c, err)
    N = cfg.ppd
    assert abs(np.sqrt(st["density_variance"] / N**3) - case["stderr"]["rms_density"]) < 1e-6
    assert np.allclose(st["max_disp"], case["stderr"]["max_disp"], rtol=2e-6)


def test_oracle_oversampling_identity(oracle):
    # PPD=2N with k_cutoff=2, sampled at even sites, equals PPD=N
    pk = helpers.wmap_pk()
    a, _ = oracle.run(oracle.make_config(16, icformat="Zeldovich"), pk)
    b, _ = oracle.run(oracle.make_config(32, k_cutoff=2.0, icformat="Zeldovich"), pk)
    b = b.reshape(32, 32, 32)[::2, ::2, ::2].reshape(-1)
    for c in range(3):
        assert oracle.field_rel_err(b["displ"][:, c], a["displ"][:, c]) < 1e-13


@pytest.mark.parametrize("case", [
    dict(ppd=64, icformat="RVZel"),
    dict(ppd=32, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, f_cluster=0.97, icformat="RVdoubleZel", eig=16),
    dict(ppd=32, qPLT=1, icformat="RVZel", eig=64, seed=-5),
    dict(ppd=32, k_cutoff=2.0, corner_modes=1, icformat="Zeldovich"),
    dict(ppd=16, qonemode=1, one_mode=(-3, 2, 5), icformat="ZelSimple"),
    dict(ppd=32, k_cutoff=2.0, fixed_power=1, icformat="RVdoubleZel"),
    dict(ppd=24, qPLT=1, qPLTrescale=1, PLT_target_z=5.0, icformat="RVdoubleZel", eig=16),  # not a power of two: direct DFT sums
])
def test_plane_oracle_matches_full_oracle(oracle, case):
    """zo_planes (selected planes by direct z summation — what faces the benchmark sizes) against zo_run, itself pinned to
    the reference-generated goldens above: ids and float32 casts identical, double fields to 1e-13."""
    case = dict(case)
    eig_ppd = case.pop("eig", None)
    synth = helpers.load_synth()
    eig = (eig_ppd, synth.make_eigmodes(eig_ppd)) if eig_ppd else None
    cfg = oracle.make_config(**case)
    N = cfg.ppd
    full, _ = oracle.run(cfg, helpers.wmap_pk(), eig)
    full = full.reshape(N, N, N)
    zs = [0, 1, N // 2 - 1, N // 2, N - 1]
    rec, st = oracle.planes(cfg, helpers.wmap_pk(), zs, eig)
    for i, z in enumerate(zs):
        want = full[z]
        if "ijk" in want.dtype.names:
            assert np.array_equal(rec[i]["ijk"], want["ijk"]) and np.all(rec[i]["pad"] == 0)
        for f in ("displ", "vel"):
            if f in want.dtype.names:
                tol = 2e-7 if want[f].dtype == np.float32 else 1e-13
                for c in range(3):
                    # scale of the whole field, not of the plane
                    err = np.max(np.abs(rec[i][f][..., c].astype(np.float64) - want[f][..., c])) / np.max(np.abs(full[f][..., c]))
                    assert err < tol, (z, f, c, err)
        if want.dtype.names[0] == "ijk" and "displ" in want.dtype.names:
            md = st[i]["max_disp"]  # (pos0, pos1, pos2) = (displ[2], displ[1], displ[0])
            for j in range(3):
                assert abs(abs(md[j]) - np.max(np.abs(want["displ"][..., 2 - j]))) <= 2e-7 * abs(md[j])


@pytest.mark.skipif(not os.path.exists(os.path.join(helpers.ROOT, "oracle", "_ref", "zeldovich_ref")), reason="reference binary not built")
def test_oracle_matches_live_reference(oracle):
    synth = helpers.load_synth()
    k, p = helpers.wmap_pk()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
        synth.write_eigmodes(os.path.join(tmp, "eig.bin"), 16)
        over = dict(NP=32**3, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="3.0", ICFormat='"RVdoubleZel"',
                    ZD_Pk_filename='"pk.pow"', ZD_PLT_filename='"eig.bin"', ZD_Seed=99, ZD_k_cutoff="1.0")
        synth.write_param(os.path.join(tmp, "c.par"), **over)
        oracle.run_reference("c.par", cwd=tmp, threads=4)
        ref = oracle.read_ic_dir(os.path.join(tmp, "ic_out"), 32, 375, "RVdoubleZel")
    cfg = oracle.make_config(32, seed=99, qPLT=1, qPLTrescale=1, PLT_target_z=3.0, icformat="RVdoubleZel")
    rec, _ = oracle.run(cfg, (k, p), (16, synth.make_eigmodes(16)))
    assert np.array_equal(rec["ijk"], ref["ijk"])
    for f in ("displ", "vel"):
        for c in range(3):
            assert oracle.field_rel_err(rec[f][:, c], ref[f][:, c]) < 1e-12


@pytest.mark.skipif(not os.path.exists(os.path.join(helpers.ROOT, "oracle", "_ref", "zeldovich_ref")), reason="reference binary not built")
def test_oracle_f_nl_matches_live_reference(oracle):
    """ZD_f_NL (reference src/zeldovich.cpp:699-790, 945-960): the restatement against a live run of the reference binary."""
    synth = helpers.load_synth()
    k, p = helpers.wmap_pk()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
        synth.write_eigmodes(os.path.join(tmp, "eig.bin"), 16)
        over = dict(NP=32**3, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="3.0", ICFormat='"RVdoubleZel"', ZD_Pk_filename='"pk.pow"',
                    ZD_PLT_filename='"eig.bin"', ZD_Seed=4242, ZD_f_NL="-3000", ZD_n_s="0.965", Omega_M="0.315", ZD_k_cutoff="2.0")
        synth.write_param(os.path.join(tmp, "c.par"), **over)
        oracle.run_reference("c.par", cwd=tmp, threads=4)
        ref = oracle.read_ic_dir(os.path.join(tmp, "ic_out"), 32, 375, "RVdoubleZel")
    cfg = oracle.make_config(32, seed=4242, qPLT=1, qPLTrescale=1, PLT_target_z=3.0, icformat="RVdoubleZel", f_NL=-3000.0, n_s=0.965,
                             Omega_M=0.315, k_cutoff=2.0)
    rec, _ = oracle.run(cfg, (k, p), (16, synth.make_eigmodes(16)))
    assert np.array_equal(rec["ijk"], ref["ijk"])
    for f in ("displ", "vel"):
        for c in range(3):
            assert oracle.field_rel_err(rec[f][:, c], ref[f][:, c]) < 1e-12


# option matrix against live runs of the reference binary: every switch of the mode loop and of the writer the
# restatement implements (reference src/zeldovich.cpp:350-358, 404-438; src/power_spectrum.cpp:225-261, 349-352; src/output.cpp:78-82)
LIVE_CASES = {
    "one_mode": (dict(NP=16**3, ZD_qonemode=1, ZD_one_mode="2 3 -1", ICFormat='"Zeldovich"'), dict(qonemode=1, one_mode=(2, 3, -1), icformat="Zeldovich")),
    "corner_modes_cutoff": (dict(NP=32**3, ZD_CornerModes=1, ZD_k_cutoff="2.0", ICFormat='"RVZel"'), dict(ppd=32, corner_modes=1, k_cutoff=2.0)),
    "f_cluster_velocity": (dict(NP=16**3, ZD_f_cluster="0.9", ICFormat='"RVdoubleZel"'), dict(f_cluster=0.9, icformat="RVdoubleZel")),
    "power_law_smoothed": (dict(NP=16**3, ZD_Pk_filename='""', ZD_Pk_powerlaw_index="-1.5", ZD_Pk_smooth="3.0", ICFormat='"RVdoubleZel"'),
                           dict(is_powerlaw=1, powerlaw_index=-1.5, Pk_smooth=3.0, icformat="RVdoubleZel")),
    "sigma_ratio_scale": (dict(NP=16**3, ZD_Pk_sigma=0, ZD_Pk_sigma_ratio="0.02", ZD_Pk_scale="1.3", BoxSize="500", ICFormat='"RVdoubleZel"'),
                          dict(Pk_sigma=0.0, Pk_sigma_ratio=0.02, Pk_scale=1.3, boxsize=500.0, icformat="RVdoubleZel")),
    "negative_seed_fixed": (dict(NP=16**3, ZD_Seed=-123456, ZD_qPk_fix_to_mean=1, ICFormat='"RVdoubleZel"'),
                            dict(seed=-123456, fixed_power=1, icformat="RVdoubleZel")),
    "f_nl_power_law": (dict(NP=16**3, ZD_Pk_filename='""', ZD_Pk_powerlaw_index="-2.0", ZD_f_NL="800", ZD_n_s="1.0", Omega_M="1.0",
                            ICFormat='"RVdoubleZel"'),
                       dict(is_powerlaw=1, powerlaw_index=-2.0, f_NL=800.0, n_s=1.0, Omega_M=1.0, icformat="RVdoubleZel")),
}


@pytest.mark.skipif(not os.path.exists(os.path.join(helpers.ROOT, "oracle", "_ref", "zeldovich_ref")), reason="reference binary not built")
@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_oracle_options_match_live_reference(oracle, name):
    synth = helpers.load_synth()
    over, kw = LIVE_CASES[name]
    over, kw = dict(over), dict(kw)
    k, p = helpers.wmap_pk()
    ppd = kw.pop("ppd", 16)
    fmt = kw.get("icformat", "RVZel")
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_power_table(os.path.join(tmp, "pk.pow"), k, p)
        over.setdefault("ZD_Pk_filename", '"pk.pow"')
        synth.write_param(os.path.join(tmp, "c.par"), **over)
        oracle.run_reference("c.par", cwd=tmp, threads=2)
        ref = oracle.read_ic_dir(os.path.join(tmp, "ic_out"), ppd, 375, fmt)
    rec, _ = oracle.run(oracle.make_config(ppd, **kw), None if kw.get("is_powerlaw") else (k, p))
    assert np.array_equal(rec["ijk"], ref["ijk"])
    for f in ("displ", "vel"):
        if f in ref.dtype.names:
            tol = 1e-6 if ref[f].dtype == np.float32 else 1e-12
            for c in range(3):
                assert oracle.field_rel_err(rec[f][:, c], ref[f][:, c]) < tol, (name, f, c)
