#!/usr/bin/env python
"""python tools/diag_slab2.py [ppd] [ranks] — slab ranks emulated on one GPU: ring emission variants against the one-tile-per-CTA
emission on the SAME receive buffers (every record of every rank)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402

pkg, synth = load_package(), load_synth()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tmp = tempfile.mkdtemp(prefix="zdiag_")
synth.write_power_table(os.path.join(tmp, "pk.pow"))
synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"), ZD_qPLT=1, ZD_qPLT_rescale=1,
            ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
power = pkg.PowerSpectrum(P)


def make(rank):
    cfg = P.config(device=0)
    cfg.rank, cfg.nranks = rank, G
    c = pkg.Context(cfg)
    power.apply(c)
    c.load_eigenmodes_file(P.PLT_filename)
    return c


ctxs = [make(r) for r in range(G)]
bufs = [torch.empty(c.workspace_bytes() // 8, dtype=torch.float64, device="cuda:0") for c in ctxs]
half = 16 * ctxs[0].narray * N**3 // G
for c, b in zip(ctxs, bufs):
    c.set_workspace(b.data_ptr(), b.numel() * 8)
for c in ctxs:
    c.dbg_set_peers([b.data_ptr() + half for b in bufs])
for c in ctxs:
    c.generate()
    c.synchronize()  # one resident z pass at a time on this GPU
torch.cuda.synchronize()
for c in ctxs:
    c.exchange_done()


def fetch_all(opts):
    out = []
    for c in ctxs:
        for k, v in {"yring": 12, "slab_ring": 2, "emit_prefetch": 1, "wide_records": -1, **opts}.items():
            c.set_option(k, v)
        out.append(c.fetch_planes(0, N // G).view(np.uint8).reshape(N // G, -1))
    return np.concatenate(out)


base = fetch_all({"yring": 0}).copy()
for name, opts in (("one-tile again", {"yring": 0}), ("ring", {}), ("ring", {}), ("ring + proxy fence", {"emit_prefetch": 2}),
                   ("ring, slices re-read by plain loads", {"emit_prefetch": 5}), ("ring, narrow records", {"wide_records": 0})):
    got = fetch_all(opts)
    bad = got != base
    planes = np.nonzero(bad.any(axis=1))[0]
    print(f"{name}: {int(bad.sum())} bytes differ in {len(planes)} planes {planes[:16]}", flush=True)
    if len(planes):
        z = planes[0]
        cols = np.nonzero(bad[z].reshape(N, N, 32).any(axis=(0, 2)))[0]
        offs = np.nonzero(bad[z].reshape(N, N, 32).any(axis=(0, 1)))[0]
        print(f"   plane {z}: x {cols[:24]} ({len(cols)}), record bytes {offs}", flush=True)
for c in ctxs:
    c.close()
