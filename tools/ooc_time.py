#!/usr/bin/env python
"""python tools/ooc_time.py --ppd 1024 --passes 4 [--store ram|disk] [--files] — wall clock of zplt_run_param_file out of core
(blocks through host memory or files), next to the same parameter file with the cube resident (--passes 0)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppd", type=int, default=1024)
    ap.add_argument("--passes", type=int, nargs="+", default=[4])
    ap.add_argument("--store", default="ram")
    ap.add_argument("--za", action="store_true")
    ap.add_argument("--files", action="store_true")
    a = ap.parse_args()
    pkg, synth = load_package(), load_synth()
    N = a.ppd
    tmp = tempfile.mkdtemp(prefix="zooc_")
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"),
                InitialConditionsDirectory='"%s"' % os.path.join(tmp, "ic"))
    if not a.za:
        synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
    par = synth.write_param(os.path.join(tmp, "c.par"), **over)
    os.environ["ZPLT_OOC_STORE"] = a.store
    for G in a.passes:
        os.environ["ZPLT_OOC_PASSES"] = str(G)
        t0 = time.perf_counter()
        rep = pkg.run_param_file(par, device=0, write_files=a.files)
        wall = time.perf_counter() - t0
        print(json.dumps({"ppd": N, "qPLT": not a.za, "passes": int(rep.ooc_passes), "store": a.store, "files": a.files, "wall_s": wall,
                          "seconds_total": rep.seconds_total, "seconds_blocks": rep.seconds_blocks, "seconds_device_and_d2h": rep.seconds_device,
                          "seconds_write": rep.seconds_write, "block_bytes_out": int(rep.ooc_bytes), "mpart_s": N**3 / wall / 1e6,
                          "density_variance": rep.density_variance}), flush=True)


if __name__ == "__main__":
    main()
