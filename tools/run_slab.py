#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/run_slab.py [--ppd N] — slab-decomposed IC generation over NCCL,
checked on rank 0 against a single-GPU run of the same parameter file (bit-identical records expected)."""
import argparse
import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import PKG_DIR, load_package, load_synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppd", type=int, default=256)
    ap.add_argument("--za", action="store_true")
    ap.add_argument("--oversample-check", action="store_true",
                    help="run PPD with ZD_k_cutoff=2 and compare even lattice sites with a single-GPU PPD/2 run (ZA only)")
    ap.add_argument("--p2p", action="store_true", help="fused exchange: z-pass kernel stores into peer memory over NVLink")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg, synth = load_package(), load_synth()
    spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(PKG_DIR, "distributed.py"))
    zd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(zd)

    N = args.ppd
    tmp = tempfile.mkdtemp(prefix=f"zslab{rank}_")
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=N**3, ICFormat='"RVZel"', ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"))
    if args.oversample_check:
        assert args.za, "the oversampling identity holds for ZA only (PLT eigenmodes depend on the lattice)"
        over["ZD_k_cutoff"] = "2.0"
    if not args.za:
        synth.write_eigmodes(os.path.join(tmp, "eig"), 128)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
    par = synth.write_param(os.path.join(tmp, "c.par"), **over)
    P = pkg.Parameters(par)
    power = pkg.PowerSpectrum(P)

    def make(rank_, world_):
        cfg = P.config(device=local)
        cfg.rank, cfg.nranks = rank_, world_
        ctx = pkg.Context(cfg)
        power.apply(ctx)
        if not args.za:
            ctx.load_eigenmodes_file(P.PLT_filename)
        return ctx

    ctx = make(rank, world)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    if args.p2p:
        ws = zd.PeerExchange(ctx)
        ws.begin()
        ctx.generate()
        ws.exchange()
    else:
        ws = zd.SlabWorkspace(ctx, dev)
        with torch.cuda.stream(stream):
            ctx.generate()
            ws.exchange()
        stream.synchronize()
    mine = ctx.fetch_planes(0, N // world)
    st = ctx.stats()
    if args.p2p:
        ws.close()
    ctx.close()
    del ws
    if args.oversample_check:
        # SURVEY §4 identity: PPD=N with k_cutoff=2 sampled at even sites == PPD=N/2 (same modes, same phases)
        torch.cuda.empty_cache()
        par2 = synth.write_param(os.path.join(tmp, "half.par"), **dict(over, NP=(N // 2) ** 3, ZD_k_cutoff="1.0"))
        P2 = pkg.Parameters(par2)
        pw2 = pkg.PowerSpectrum(P2)
        c2 = pkg.Context(P2.config(device=local))
        pw2.apply(c2)
        c2.generate()
        np_loc = N // world
        worst, nchk = 0.0, 0
        rec = mine.reshape(np_loc, N, N)
        for zl in range(0, np_loc, max(2, np_loc // 8 // 2 * 2)):
            zg = rank * np_loc + zl
            if zg % 2:
                continue
            ref = c2.fetch_planes(zg // 2, 1).reshape(N // 2, N // 2)
            sub = rec[zl, ::2, ::2]
            assert np.array_equal(sub["ijk"] // 2, ref["ijk"]) and np.all(sub["ijk"] % 2 == 0)
            for f in ("displ", "vel"):
                d = np.abs(sub[f].astype(np.float64) - ref[f].astype(np.float64)).max() / np.abs(ref[f]).max()
                worst = max(worst, float(d))
            nchk += 1
        c2.close()
        t = torch.tensor([worst], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"oversample check PPD={N} (k_cutoff=2) vs PPD={N // 2}: {nchk} planes per rank, worst field-relative "
                  f"difference {t.item():.3e} (float32 records)")
        assert t.item() < 2e-7
        dist.barrier()
        dist.destroy_process_group()
        return
    # gather on rank 0 and compare with a single-GPU run
    parts = [None] * world
    dist.gather_object((mine.view(np.uint8), st), parts if rank == 0 else None, dst=0)
    if rank == 0:
        got = np.concatenate([p[0] for p in parts])
        var = sum(p[1]["density_variance"] for p in parts)
        ref_ctx = make(0, 1)
        ref_ctx.generate()
        wantr = ref_ctx.fetch_planes(0, N)
        want = wantr.view(np.uint8)
        rst = ref_ctx.stats()
        same = np.array_equal(got, want)
        # the z pass of a slab rank is another instantiation of the same transform (different FMA contraction choices by the
        # compiler): ids must be identical, fields may differ in the last bit of a double, i.e. rarely by one float32 ulp
        gotr = got.view(wantr.dtype)
        ids = np.array_equal(gotr["ijk"], wantr["ijk"])
        worst, ndiff = 0.0, int((gotr["displ"] != wantr["displ"]).sum() + (gotr["vel"] != wantr["vel"]).sum())
        for f in ("displ", "vel"):
            worst = max(worst, float(np.abs(gotr[f].astype(np.float64) - wantr[f].astype(np.float64)).max() / np.abs(wantr[f]).max()))
        print(f"slab run PPD={N} world={world} p2p={args.p2p}: records byte-identical to single-GPU run: {same}; ids identical: {ids}; "
              f"{ndiff} of {6 * N**3} float32 fields differ, worst {worst:.2e} of the field scale; "
              f"density variance {var:.12g} vs {rst['density_variance']:.12g}")
        assert ids and worst < 2e-7
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
