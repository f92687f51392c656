#!/usr/bin/env python
"""python tools/diag_slab.py --ppd 1024 --ranks 2 [--opt k=v ...] — every record of a slab run (all ranks emulated on one
GPU through zplt_dbg_set_peers) against a single-GPU run of the same parameters; reports where the differences are."""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppd", type=int, default=1024)
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--za", action="store_true")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    pkg, synth = load_package(), load_synth()
    N, G = args.ppd, args.ranks
    tmp = tempfile.mkdtemp(prefix="zdiag_")
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"))
    if not args.za:
        synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
    P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
    power = pkg.PowerSpectrum(P)

    def make(rank, nranks):
        cfg = P.config(device=0)
        cfg.rank, cfg.nranks = rank, nranks
        c = pkg.Context(cfg)
        power.apply(c)
        if not args.za:
            c.load_eigenmodes_file(P.PLT_filename)
        for o in args.opt:
            k, v = o.split("=")
            c.set_option(k, int(v))
        return c

    ctxs = [make(r, G) for r in range(G)]
    bufs = [torch.empty(c.workspace_bytes() // 8, dtype=torch.float64, device="cuda:0") for c in ctxs]
    half = 16 * ctxs[0].narray * N**3 // G
    for c, b in zip(ctxs, bufs):
        c.set_workspace(b.data_ptr(), b.numel() * 8)
        b.fill_(float("nan"))
    torch.cuda.synchronize()  # the NaN fills run on torch's stream, the library on its own (non-blocking) streams
    for c in ctxs:
        c.dbg_set_peers([b.data_ptr() + half for b in bufs])
    for c in ctxs:
        c.generate()
        c.synchronize()  # one resident z pass at a time on this GPU
    torch.cuda.synchronize()
    parts = []
    for c in ctxs:
        c.exchange_done()
        parts.append(c.fetch_planes(0, N // G))
        c.close()
    del bufs
    torch.cuda.empty_cache()
    got = np.concatenate(parts).reshape(N, N, N)
    del parts
    ref = make(0, 1)
    ref.generate()
    want = ref.fetch_planes(0, N).reshape(N, N, N)
    ref.close()
    print("ids identical:", np.array_equal(got["ijk"], want["ijk"]))
    for f in ("displ", "vel"):
        for comp in range(3):
            a, b = got[f][..., comp], want[f][..., comp]
            bad = a != b
            nb = int(bad.sum())
            scale = float(np.abs(b).max())
            big = np.abs(a.astype(np.float64) - b.astype(np.float64)) > 2e-7 * scale
            nbig = int(big.sum())
            print(f"{f}[{comp}]: {nb} differ, {nbig} beyond one float ulp of the field scale, nan {int(np.isnan(a).sum())}")
            if nbig:
                z, y, x = np.nonzero(big)
                print("   z:", np.unique(z)[:20], len(np.unique(z)), " y:", np.unique(y)[:20], len(np.unique(y)), " x:", np.unique(x)[:20], len(np.unique(x)))
                print("   worst", float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / scale),
                      " sample (z,y,x,got,want):", [(int(z[i]), int(y[i]), int(x[i]), float(a[z[i], y[i], x[i]]), float(b[z[i], y[i], x[i]])) for i in range(0, len(z), max(1, len(z) // 6))][:6])


if __name__ == "__main__":
    main()
