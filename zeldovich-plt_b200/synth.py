"""Synthetic inputs for the IC hot path: PLT eigenmode tables and parameter files.

The reference ships its 34 MB ``eigmodes128`` table out of tree (reference
``.MISSING_LARGE_BLOBS``), so every qPLT configuration here runs on a synthetic table
written in the exact on-disk format the reference reads
(reference src/zeldovich.cpp:794-830): ``int32 ppd_e`` followed by
``double[ppd_e][ppd_e][ppd_e/2+1][4]`` indexed ``[ikx][iky][ikz][ex,ey,ez,lambda]``
(EIGMODE macro, reference src/zeldovich.cpp:155-158).  Index ``ppd_e/2`` means the
+Nyquist wavenumber; only ``kz >= 0`` is stored.

The table is smooth (so that trilinear interpolation is well behaved) but is
deliberately NOT symmetric under ``k -> -k``: an implementation that re-interpolates
the eigenvector at ``-k`` for the Hermitian-conjugate entries, instead of reusing
the primary mode's eigenvector (reference src/zeldovich.cpp:460-466), fails parity.
Physical fidelity is not claimed.
"""
import os

import numpy as np


def make_eigmodes(ppd_e: int = 128, seed: int = 7) -> np.ndarray:
    """Return the ``[ppd_e][ppd_e][ppd_e//2+1][4]`` float64 table."""
    rng = np.random.RandomState(seed)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=6)
    h = ppd_e // 2
    idx = np.arange(ppd_e)
    kw = np.where(idx > h, idx - ppd_e, idx).astype(np.float64)  # index h is +Nyquist
    kx = kw[:, None, None]
    ky = kw[None, :, None]
    kz = np.arange(h + 1, dtype=np.float64)[None, None, :]
    kx, ky, kz = np.broadcast_arrays(kx, ky, kz)
    k2 = kx * kx + ky * ky + kz * kz
    kmag = np.sqrt(np.where(k2 > 0, k2, 1.0))
    q = kmag / h  # |k| / k_Nyquist, up to sqrt(3)
    u = np.pi * kx / h
    v = np.pi * ky / h
    w = np.pi * kz / h
    amp = 0.06 * np.minimum(q, 1.5) ** 2
    # smooth perturbation, odd+even mix so that it has no k -> -k symmetry
    px = amp * (np.sin(v + ph[0]) + 0.5 * np.cos(w + ph[1]))
    py = amp * (np.sin(w + ph[2]) + 0.5 * np.cos(u + ph[3]))
    pz = amp * (np.sin(u + ph[4]) + 0.5 * np.cos(v + ph[5]))
    ex = kx / kmag + px
    ey = ky / kmag + py
    ez = kz / kmag + pz
    nrm = np.sqrt(ex * ex + ey * ey + ez * ez)
    ex, ey, ez = ex / nrm, ey / nrm, ez / nrm
    lam = 1.0 - 0.11 * q * q * (1.0 + 0.2 * np.sin(u + 0.3) * np.cos(v - 0.2) + 0.1 * np.sin(w + 0.1))
    tab = np.stack([ex, ey, ez, lam], axis=-1).astype(np.float64)
    tab[0, 0, 0] = (0.0, 0.0, 1.0, 1.0)
    return np.ascontiguousarray(tab)


def write_eigmodes(path: str, ppd_e: int = 128, seed: int = 7) -> str:
    tab = make_eigmodes(ppd_e, seed)
    with open(path, "wb") as f:
        f.write(np.int32(ppd_e).tobytes())
        f.write(tab.tobytes())
    return path


def read_eigmodes(path: str):
    with open(path, "rb") as f:
        ppd_e = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
        tab = np.frombuffer(f.read(), dtype=np.float64)
    assert tab.size == ppd_e * ppd_e * (ppd_e // 2 + 1) * 4
    return ppd_e, tab.reshape(ppd_e, ppd_e, ppd_e // 2 + 1, 4)


# Parameter-file template: reference example.par with the keys SURVEY.md §8d fixes.
_BASE = dict(
    BoxSize=720,
    CPD=375,
    ICFormat='"RVZel"',
    InitialConditionsDirectory='"./ic_out"',
    InitialRedshift=49,
    NP=262144,
    ZD_NumBlock=4,
    ZD_Pk_filename='"wmap1new.pow"',
    ZD_Pk_norm="8.0",
    ZD_Pk_scale="1.0",
    ZD_Pk_sigma="0.0210839935761",
    ZD_Pk_smooth="0.0",
    ZD_Seed=12346,
    ZD_k_cutoff="1.0",
    ZD_qPLT=0,
    ZD_qPLT_rescale=0,
    ZD_qPk_fix_to_mean=0,
    ZD_Version=2,
    ZD_f_NL=0,
)


def param_text(**over) -> str:
    """Text of a parameter file; keyword overrides replace or add keys.

    String-valued keys must be passed already quoted if they need quotes.
    """
    d = dict(_BASE)
    d.update(over)
    return "# synthetic zeldovich parameter file\n" + "".join(f"{k} = {v}\n" for k, v in d.items())


def write_param(path: str, **over) -> str:
    with open(path, "w") as f:
        f.write(param_text(**over))
    return path


def make_power_table(n: int = 256, kmin: float = 1e-5, kmax: float = 50.0, gamma: float = 0.2, n_s: float = 0.97):
    """Synthetic linear P(k): BBKS transfer function, ``n`` log-spaced rows (k, P).

    Stands in for the reference's ``wmap1new.pow`` in throughput runs; reaches
    k=50 so that no configuration extrapolates the spline (reference
    src/power_spectrum.cpp:239-254 warns when it does).  Amplitude is arbitrary:
    ``ZD_Pk_norm``/``ZD_Pk_sigma`` renormalise it.
    """
    k = np.exp(np.linspace(np.log(kmin), np.log(kmax), n))
    q = k / gamma
    t = np.log(1.0 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    p = 2.0e4 * k**n_s * t * t
    return k, p


def write_power_table(path: str, k=None, p=None) -> str:
    """Write a two-column text table that ``sscanf("%lf %lf")`` reads back exactly."""
    if k is None:
        k, p = make_power_table()
    with open(path, "w") as f:
        f.write("# k  P(k)\n")
        for a, b in zip(k, p):
            f.write(f"{float(a)!r} {float(b)!r}\n")
    return path
