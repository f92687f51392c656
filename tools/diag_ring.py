#!/usr/bin/env python
"""python tools/diag_ring.py — ring-prefetched emission against the one-tile-per-CTA emission on the SAME transformed cube
(single GPU, PPD=1024 qPLT RVZel): every record, several variants, each run twice (a race shows up as run-to-run noise)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402

pkg, synth = load_package(), load_synth()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
ctx.set_option("zring", 0)
ctx.generate()
ctx.set_option("yring", 0)
base = ctx.fetch_planes(0, N).view(np.uint8).reshape(N, -1).copy()
again = ctx.fetch_planes(0, N).view(np.uint8).reshape(N, -1)
print("one-tile kernel, run-to-run: planes differing", int((base != again).any(axis=1).sum()), flush=True)
for name, opts in (("ring", {"yring": 12, "emit_prefetch": 1}), ("ring + proxy fence", {"yring": 12, "emit_prefetch": 2}),
                   ("ring, slices re-read by plain loads", {"yring": 12, "emit_prefetch": 5}), ("ring, narrow records", {"yring": 12, "emit_prefetch": 1, "wide_records": 0})):
    for k, v in {"wide_records": -1, **opts}.items():
        ctx.set_option(k, v)
    for rep in range(2):
        got = ctx.fetch_planes(0, N).view(np.uint8).reshape(N, -1)
        bad = (got != base)
        planes = np.nonzero(bad.any(axis=1))[0]
        print(f"{name} (run {rep}): {int(bad.sum())} bytes differ in {len(planes)} planes {planes[:12]}", flush=True)
ctx.close()
