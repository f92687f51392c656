#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (where /root/reference exists) after `make -C oracle ref`:

    python tests/golden/make_golden.py

For every case below it writes a parameter file, the P(k) text table and (for qPLT)
a small synthetic eigenmode table into a scratch directory, runs
oracle/_ref/zeldovich_ref on it, and stores under tests/golden/:

  <case>.npz   records (raw bytes of the concatenated ic_* files, z-ascending), the
               eigenmode table used, and the configuration
  cases.json   the configuration of every case plus the scalars the reference
               printed on stderr (sigma lines, rms density, max displacements)

The reference's own repository holds no golden vectors, so these
files, produced by its sources, are what pins the oracle.  wmap1_pk.npy holds the
(k, P) rows of the reference's wmap1new.pow as float64.
"""
import importlib.util
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zel_oracle as zo  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "zeldovich-plt_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

CASES = {
    # name: (param overrides, eig ppd or None)
    "za16_rvzel": (dict(NP=16**3, ICFormat='"RVZel"'), None),
    "za32_kc2_zelsimple": (dict(NP=32**3, ZD_k_cutoff="2.0", ICFormat='"ZelSimple"'), None),
    "plt16_interp_rvdouble": (
        dict(NP=16**3, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ICFormat='"RVdoubleZel"', ZD_f_cluster="0.97"), 8),
    "plt16_direct_rvzel": (dict(NP=16**3, ZD_qPLT=1, ZD_qPLT_rescale=0, ICFormat='"RVZel"', CPD=5), 16),
    "za16_fixed_zeldovich": (dict(NP=16**3, ZD_qPk_fix_to_mean=1, ICFormat='"Zeldovich"', ZD_Seed=-7, BoxSize="250.5"), None),
    # local primordial non-Gaussianity (reference src/zeldovich.cpp:699-790, 945-960); large f_NL so that the phi^2 term is visible
    "za16_fnl_rvdouble": (dict(NP=16**3, ICFormat='"RVdoubleZel"', ZD_f_NL="5000", ZD_n_s="0.96", Omega_M="0.3"), None),
    "plt16_fnl_rvzel": (dict(NP=16**3, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ICFormat='"RVZel"', ZD_f_NL="-2000",
                             ZD_n_s="0.96", Omega_M="0.3"), 8),
    # non-power-of-two particle grids (the reference takes any even ppd: src/block_array.cpp:38-40; its shim FFT sums them directly)
    "za24_rvdouble": (dict(NP=24**3, ICFormat='"RVdoubleZel"', ZD_NumBlock=2), None),
    "plt48_rvzel": (dict(NP=48**3, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ICFormat='"RVZel"', ZD_NumBlock=4, CPD=7), 16),
}


def main():
    pk = np.load(os.path.join(HERE, "wmap1_pk.npy"))
    out = {}
    only = set(sys.argv[1:])  # optional: regenerate just the named cases, keep the others' entries
    if only and os.path.exists(os.path.join(HERE, "cases.json")):
        with open(os.path.join(HERE, "cases.json")) as f:
            out = json.load(f)
    for name, (over, eig_ppd) in CASES.items():
        if only and name not in only:
            continue
        tmp = tempfile.mkdtemp(prefix="zgold_")
        try:
            synth.write_power_table(os.path.join(tmp, "pk.pow"), pk[:, 0], pk[:, 1])
            over = dict(over)
            over["ZD_Pk_filename"] = '"pk.pow"'
            eig = np.zeros(0)
            if eig_ppd:
                synth.write_eigmodes(os.path.join(tmp, "eig.bin"), eig_ppd)
                over["ZD_PLT_filename"] = '"eig.bin"'
                eig = synth.make_eigmodes(eig_ppd)
            text = synth.param_text(**over)
            with open(os.path.join(tmp, "case.par"), "w") as f:
                f.write(text)
            err = zo.run_reference("case.par", cwd=tmp)
            cfg = dict(synth._BASE)
            cfg.update(over)
            ppd = round(int(cfg["NP"]) ** (1 / 3))
            fmt = cfg["ICFormat"].strip('"')
            rec = zo.read_ic_dir(os.path.join(tmp, "ic_out"), ppd, int(cfg["CPD"]), fmt)
            scal = {}
            m = re.search(r"Input sigma\(([-0-9.e+]+)\) = ([-0-9.e+]+)", err)
            if m:
                scal["input_sigma"] = float(m.group(2))
            m = re.search(r"rms density variation of the pixels is ([-0-9.e+]+)", err)
            scal["rms_density"] = float(m.group(1))
            m = re.search(r"displacements are \(([-0-9.e+]+), ([-0-9.e+]+), ([-0-9.e+]+)\)", err)
            scal["max_disp"] = [float(m.group(i)) for i in (1, 2, 3)]
            np.savez_compressed(os.path.join(HERE, name + ".npz"), records=rec.view(np.uint8), eig=eig, param_text=np.array(text))
            out[name] = dict(params={k: str(v) for k, v in cfg.items()}, eig_ppd=eig_ppd, ppd=ppd, icformat=fmt, stderr=scal)
            print(name, ppd, fmt, rec.size, scal)
        finally:
            shutil.rmtree(tmp)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
