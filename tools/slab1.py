#!/usr/bin/env python
"""python tools/slab1.py --ppd 2048 --ranks 8 [--opt name=value ...] — ONE rank of a slab run timed on one GPU.

The rank's own receive buffer stands in for every peer (zplt_dbg_set_peers), so the peer stores of the z pass + exchange
kernel become local HBM stores: the kernels' own cost (loads, transform, store pattern) without NVLink.  Stage 2 of a
slab rank is local anyway, so its time here is what a real run sees.  The records are not meaningful (every source
overwrites the same rows); this is a timing harness for kernel work, parity lives in tests/.
"""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package, load_synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppd", type=int, default=2048)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--rank", type=int, default=1)
    ap.add_argument("--za", action="store_true")
    ap.add_argument("--icformat", default="RVZel")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="name=value tuning switch (zplt_set_option), repeatable")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    pkg, synth = load_package(), load_synth()
    N, G = args.ppd, args.ranks
    tmp = tempfile.mkdtemp(prefix="zslab1_")
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=N**3, ICFormat='"%s"' % args.icformat, ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"))
    if not args.za:
        synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
    P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
    power = pkg.PowerSpectrum(P)
    cfg = P.config(device=0)
    cfg.rank, cfg.nranks = args.rank, G
    ctx = pkg.Context(cfg)
    power.apply(ctx)
    if not args.za:
        ctx.load_eigenmodes_file(P.PLT_filename)
    opts = {}
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
        opts[k] = int(v)
    ws = ctx.workspace_bytes()
    W = torch.empty(ws // 8, dtype=torch.float64, device="cuda:0")
    ctx.set_workspace(W.data_ptr(), ws)
    recv = W.data_ptr() + 16 * ctx.narray * N**3 // G  # the receive buffer follows the stage-1 buffer
    ctx.dbg_set_peers([recv] * G)
    nloc = N // G
    rb = ctx.record_bytes
    free_b = torch.cuda.mem_get_info()[0]
    out_planes = max(1, min(nloc, int((free_b - (4 << 30)) // (N * N * rb))))
    out = torch.empty(out_planes * N * N * rb, dtype=torch.uint8, device="cuda:0")
    res = []
    for it in range(args.steps + 1):
        ctx.generate()
        ctx.synchronize()
        ctx.exchange_done()
        for z0 in range(0, nloc, out_planes):
            ctx.emit_planes(z0, min(out_planes, nloc - z0), out.data_ptr())
        ctx.synchronize()
        t = ctx.timings()
        if it:
            res.append((t["gen_xfft_ms"] + t["zfft_ms"], t["yfft_emit_ms"]))
    na = 2 if args.za else 4
    st1 = min(r[0] for r in res)
    st2 = min(r[1] for r in res)
    b1 = 48 * na * N**3 / G  # generation writes 16*na, the z pass reads and writes 16*na each
    b2 = (16 * na + rb) * N**3 / G
    print(json.dumps({"tag": args.tag, "ppd": N, "ranks": G, "rank": args.rank, "opts": opts, "stage1_ms": st1, "stage2_ms": st2,
                      "stage1_gbs": b1 / st1 / 1e6, "stage2_gbs": b2 / st2 / 1e6, "stage2_frac_of_6463": b2 / st2 / 1e6 / 6463.3}))
    ctx.close()


if __name__ == "__main__":
    main()
