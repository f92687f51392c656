"""Shared helpers for the parity tests: golden-case loading and config translation."""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")


def load_synth():
    spec = importlib.util.spec_from_file_location("zplt_synth", os.path.join(ROOT, "zeldovich-plt_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def wmap_pk():
    d = np.load(os.path.join(GOLDEN, "wmap1_pk.npy"))
    return d[:, 0].copy(), d[:, 1].copy()


def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def load_golden(name):
    case = golden_cases()[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    eig = None
    if case["eig_ppd"]:
        eig = (case["eig_ppd"], z["eig"])
    return case, z["records"], eig, str(z["param_text"])


def params_to_kwargs(p):
    """Parameter-file dictionary (strings) -> keyword arguments common to oracle and product configs."""
    g = lambda k, d=None: p.get(k, d)
    np_ = int(g("NP"))
    ppd = round(np_ ** (1 / 3))
    return dict(
        ppd=ppd,
        boxsize=float(g("BoxSize")),
        seed=int(g("ZD_Seed")),
        k_cutoff=float(g("ZD_k_cutoff", 1.0)),
        corner_modes=int(g("ZD_CornerModes", 0)),
        qPLT=int(g("ZD_qPLT", 0)),
        qPLTrescale=int(g("ZD_qPLT_rescale", 0)),
        PLT_target_z=float(g("ZD_PLT_target_z", 0.0)),
        z_initial=float(g("InitialRedshift")),
        f_cluster=float(g("ZD_f_cluster", 1.0)),
        fixed_power=int(g("ZD_qPk_fix_to_mean", 0)),
        Pk_norm=float(g("ZD_Pk_norm")),
        Pk_sigma=float(g("ZD_Pk_sigma", 0.0)),
        Pk_sigma_ratio=float(g("ZD_Pk_sigma_ratio", 0.0)),
        Pk_smooth=float(g("ZD_Pk_smooth", 0.0)),
        Pk_scale=float(g("ZD_Pk_scale", 1.0)),
        icformat=g("ICFormat").strip('"'),
        f_NL=float(g("ZD_f_NL", 0.0)),
        n_s=float(g("ZD_n_s", 1.0)),
        Omega_M=float(g("Omega_M", 1.0)),
    )
