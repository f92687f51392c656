"""Consumer side of the ``ic_<n>`` files (the first step of it): read what ``zeldovich <param_file>`` wrote and
form the particles an N-body code starts from.

File semantics are the reference's (reference src/output.cpp:208-212, README.md "ICFormat"): plane ``z`` of the particle
lattice is appended to ``ic_{z*CPD//PPD}`` in ascending ``z``, ``y`` outer and ``x`` inner inside a plane; a record carries the
lattice index ``(i, j, k) = (z, y, x)`` and the comoving displacement (and velocity) in ``BoxSize`` units, components in the
order of ``(i, j, k)``.  Global positions are ``(i, j, k)/PPD * BoxSize + displ`` (reference README.md:406-412).
ZelSimple records carry no lattice index; it is implied by the order.
"""
import os

import numpy as np

ICFORMATS = {"Zeldovich": 0, "RVZel": 1, "RVdoubleZel": 2, "ZelSimple": 3}
RECORD_DTYPES = {
    0: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3)]),
    1: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f4", 3), ("vel", "<f4", 3)]),
    2: np.dtype([("ijk", "<u2", 3), ("pad", "<u2"), ("displ", "<f8", 3), ("vel", "<f8", 3)]),
    3: np.dtype([("displ", "<f4", 3)]),
}


def ic_file_planes(ppd, cpd):
    """{file number: (first plane, one past the last plane)} — which lattice planes each ``ic_<n>`` holds."""
    out = {}
    for z in range(ppd):
        n = z * cpd // ppd
        lo, hi = out.get(n, (z, z))
        out[n] = (min(lo, z), z + 1)
    return out


def read_ic_files(directory, ppd, cpd, icformat, files=None):
    """Records of the whole run (or of the listed file numbers) as one structured array in (z, y, x) order, plus the plane
    range ``(z0, z1)`` they cover.  Sizes are checked against what the file numbering implies."""
    fmt = ICFORMATS[icformat] if isinstance(icformat, str) else int(icformat)
    dt = RECORD_DTYPES[fmt]
    planes = ic_file_planes(ppd, cpd)
    want = sorted(planes) if files is None else sorted(files)
    parts = []
    for n in want:
        if n not in planes:
            raise ValueError(f"ic_{n} holds no plane of a ppd={ppd}, cpd={cpd} run")
        rec = np.fromfile(os.path.join(directory, f"ic_{n}"), dtype=dt)
        z0, z1 = planes[n]
        if rec.size != (z1 - z0) * ppd * ppd:
            raise ValueError(f"ic_{n}: {rec.size} records, expected {(z1 - z0) * ppd * ppd} (planes {z0}..{z1 - 1})")
        if "ijk" in dt.names and rec.size and not (rec["ijk"][0, 0] == z0 and rec["ijk"][-1, 0] == z1 - 1):
            raise ValueError(f"ic_{n}: lattice indices do not match planes {z0}..{z1 - 1}")
        parts.append(rec)
    z0 = planes[want[0]][0] if want else 0
    z1 = planes[want[-1]][1] if want else 0
    return (np.concatenate(parts) if parts else np.zeros(0, dtype=dt)), (z0, z1)


def lattice_indices(records, ppd, z0=0):
    """(n, 3) integer lattice indices (i, j, k) of the records; from the records themselves, or from their order (ZelSimple)."""
    if "ijk" in records.dtype.names:
        return records["ijk"].astype(np.int64)
    n = np.arange(records.size, dtype=np.int64)
    return np.stack([z0 + n // (ppd * ppd), (n // ppd) % ppd, n % ppd], axis=1)


def global_positions(records, ppd, boxsize, z0=0, wrap=True):
    """Comoving positions in BoxSize units: lattice site + displacement (reference README.md:406-412), wrapped into [0, BoxSize)."""
    pos = lattice_indices(records, ppd, z0).astype(np.float64) * (float(boxsize) / ppd) + records["displ"].astype(np.float64)
    return np.mod(pos, float(boxsize)) if wrap else pos


def velocities(records, f_growth=None):
    """Comoving redshift-space displacements (multiply by a*H(z) for proper velocities, reference README "ICFormat"): the
    ``vel`` field of the RV formats; for the displacement-only formats the Zel'dovich velocity ``f * displ``."""
    if "vel" in records.dtype.names:
        return records["vel"].astype(np.float64)
    if f_growth is None:
        raise ValueError("this ICFormat stores no velocities: pass the growth rate f (1 in an EdS universe)")
    return float(f_growth) * records["displ"].astype(np.float64)
