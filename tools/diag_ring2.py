#!/usr/bin/env python
"""python tools/diag_ring2.py — which emission after zplt_generate goes wrong?  Single GPU, PPD=1024 qPLT RVZel: the FIRST emission
after generate with the ring kernel (with/without the ring z pass before it, with/without a synchronisation in between) against
the one-tile kernel's records of the same cube."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402

pkg, synth = load_package(), load_synth()
N = 1024
tmp = tempfile.mkdtemp(prefix="zdiag_")
synth.write_power_table(os.path.join(tmp, "pk.pow"))
synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"), ZD_qPLT=1, ZD_qPLT_rescale=1,
            ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
power = pkg.PowerSpectrum(P)
ctx = pkg.Context(P.config(device=0))
power.apply(ctx)
ctx.load_eigenmodes_file(P.PLT_filename)


def fetch(yring):
    ctx.set_option("yring", yring)
    return ctx.fetch_planes(0, N).view(np.uint8).reshape(N, -1)


def report(name, got, base):
    bad = got != base
    planes = np.nonzero(bad.any(axis=1))[0]
    print(f"{name}: {int(bad.sum())} bytes differ in {len(planes)} planes {planes[:16]}", flush=True)


for zring, sync in ((12, False), (12, True), (0, False), (12, False)):
    ctx.set_option("zring", zring)
    ctx.generate()
    if sync:
        ctx.synchronize()
    first = fetch(12).copy()
    base = fetch(0).copy()
    second = fetch(12)
    tag = f"zring={zring} sync={sync}"
    report(tag + " first emission (ring) vs one-tile", first, base)
    report(tag + " second emission (ring) vs one-tile", second, base)
ctx.close()
