#!/usr/bin/env python
"""python tools/diag_slab3.py — what makes the records of an emulated slab run differ from the single-GPU run (PPD=1024, 2 ranks)?
Variants: workspace filled with NaN first or not; first emission with the ring kernel or with the one-tile kernel."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402

pkg, synth = load_package(), load_synth()
N, G = 1024, 2
tmp = tempfile.mkdtemp(prefix="zdiag_")
synth.write_power_table(os.path.join(tmp, "pk.pow"))
synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"), ZD_qPLT=1, ZD_qPLT_rescale=1,
            ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
power = pkg.PowerSpectrum(P)


def make(rank, nranks):
    cfg = P.config(device=0)
    cfg.rank, cfg.nranks = rank, nranks
    c = pkg.Context(cfg)
    power.apply(c)
    c.load_eigenmodes_file(P.PLT_filename)
    return c


def report(name, got, base):
    bad = got != base
    planes = np.nonzero(bad.any(axis=1))[0]
    print(f"{name}: {int(bad.sum())} bytes differ in {len(planes)} planes {planes[:16]}", flush=True)


ref = make(0, 1)
ref.generate()
want = ref.fetch_planes(0, N).view(np.uint8).reshape(N, -1).copy()
ref.set_option("yring", 0)
report("single GPU: one-tile emission vs ring emission", ref.fetch_planes(0, N).view(np.uint8).reshape(N, -1), want)
ref.close()

for nanfill, first_ring, sync_each in ((True, True, False), (False, True, False), (True, False, False), (True, True, True)):
    ctxs = [make(r, G) for r in range(G)]
    bufs = [torch.empty(c.workspace_bytes() // 8, dtype=torch.float64, device="cuda:0") for c in ctxs]
    half = 16 * ctxs[0].narray * N**3 // G
    for c, b in zip(ctxs, bufs):
        c.set_workspace(b.data_ptr(), b.numel() * 8)
        if nanfill:
            b.fill_(float("nan"))
        c.set_option("slab_ring", 2)
    torch.cuda.synchronize()  # the NaN fills run on torch's stream, the library on its own (non-blocking) streams
    for c in ctxs:
        c.dbg_set_peers([b.data_ptr() + half for b in bufs])
    for c in ctxs:
        c.generate()
        c.synchronize()  # one resident z pass at a time on this GPU
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    for c in ctxs:
        c.exchange_done()
    tag = f"slab nanfill={nanfill} first={'ring' if first_ring else 'one-tile'} sync_each_rank={sync_each}"
    for it in range(2):
        ring = first_ring if it == 0 else not first_ring
        parts = []
        for c in ctxs:
            c.set_option("yring", 12 if ring else 0)
            parts.append(c.fetch_planes(0, N // G).view(np.uint8).reshape(N // G, -1))
        report(f"{tag}: emission {it} ({'ring' if ring else 'one-tile'}) vs single GPU", np.concatenate(parts), want)
    for c in ctxs:
        c.close()
    c = b = None
    del bufs, ctxs, parts
    torch.cuda.empty_cache()
