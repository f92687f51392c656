#!/usr/bin/env python
"""Slab-decomposed ``zeldovich <param_file>`` across the GPUs of one node.

    torchrun --nnodes=1 --nproc-per-node G zeldovich-plt_b200/run_distributed.py <param_file>

Same parameter file, same ``ic_<n>`` files as the single-GPU CLI: rank r generates the z planes
[r*PPD/G, (r+1)*PPD/G) and appends them to ``InitialConditionsDirectory/ic_{z*CPD/PPD}``
(reference src/output.cpp:208).  When CPD < PPD several planes share a file and a file can
straddle two ranks; ascending-z append order is kept by letting a rank write the planes of a file
it shares with the previous rank only after that rank has finished (SURVEY.md trap T8).
"""
import argparse
import importlib.util
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from __graft_entry__ import load_package  # noqa: E402


def _load_distributed():
    spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(HERE, "distributed.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("param_file")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg, zd = load_package(), _load_distributed()
    t0 = time.perf_counter()

    P = pkg.Parameters(args.param_file)
    power = pkg.PowerSpectrum(P)
    cfg = P.config(device=local)
    cfg.rank, cfg.nranks = rank, world
    ctx = pkg.Context(cfg)
    power.apply(ctx)
    if P.qPLT:
        ctx.load_eigenmodes_file(P.PLT_filename)
    N, cpd, rb = P.ppd, P.cpd, ctx.record_bytes
    outdir = P.output_dir
    if rank == 0:  # SetupOutputDir (reference src/output.cpp:236-251)
        os.makedirs(outdir, exist_ok=True)
        for fn in os.listdir(outdir):
            if (fn.startswith("ic_") or fn.startswith("zeldovich.")) and os.path.isfile(os.path.join(outdir, fn)):
                os.remove(os.path.join(outdir, fn))
    dist.barrier()

    if world == 1:
        ctx.generate()
    elif args.exchange == "p2p":
        ex = zd.PeerExchange(ctx)
        ex.generate()  # with ZD_f_NL != 0 this includes the potential pass and its two barriers
    else:
        stream = torch.cuda.Stream(device=dev)
        ctx.set_stream(stream.cuda_stream)
        ws = zd.SlabWorkspace(ctx, dev)
        with torch.cuda.stream(stream):
            ctx.generate()
            ws.exchange()
        stream.synchronize()
    t1 = time.perf_counter()

    z0, z1 = zd.plane_range(N, rank, world)
    fileno = lambda z: z * cpd // N
    qdensity, qoneslab = P.qdensity, P.qoneslab
    records = qdensity != 2  # ZD_qdensity = 2: density only, no ic_* files (reference src/zeldovich.cpp:871-876, output.cpp:207-224)
    # ZD_qoneslab >= 0: only that plane is written (reference src/zeldovich.cpp:669), by the rank that owns it
    w0, w1 = (z0, z1) if qoneslab < 0 else (max(z0, qoneslab), min(z1, qoneslab + 1))
    shared = records and w0 < w1 and w0 > 0 and qoneslab < 0 and fileno(w0) == fileno(w0 - 1)
    nshared = 0
    if shared:
        while w0 + nshared < w1 and fileno(w0 + nshared) == fileno(w0):
            nshared += 1
    plane = N * N * rb
    dplane = N * N * 4
    chunk = max(1, min(max(1, w1 - w0), (1 << 30) // plane))
    pinned = torch.empty(chunk * plane, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    # density planes (float32 Re A0, reference src/output.cpp:196,217-224) are kept per rank and appended in rank order below
    dens_parts = []

    def write_planes(za, zb):
        for c0 in range(za, zb, chunk):
            n = min(chunk, zb - c0)
            if qdensity:
                rec, dens = ctx.fetch_planes_density(c0 - z0, n, records=records)
                dens_parts.append((c0, dens.copy()))
                if records:
                    host[:n * plane] = rec.view(np.uint8)
            else:
                ctx.fetch_planes_ptr(c0 - z0, n, pinned.data_ptr())
            for i in range(n if records else 0):
                with open(os.path.join(outdir, f"ic_{fileno(c0 + i)}"), "ab") as f:
                    f.write(host[i * plane:(i + 1) * plane].tobytes())

    if w0 < w1:
        write_planes(w0 + nshared, w1)  # files that start inside this rank's range: no ordering constraint
    for turn in range(world):       # the leading planes that continue the previous rank's last file, in rank order
        if rank == turn and nshared:
            write_planes(w0, w0 + nshared)
        dist.barrier()
    if qdensity:
        # one density file, planes in ascending z: rank 0 creates it ("wb", reference src/output.cpp:282-288), the others append in turn
        name = pkg.format_density_name(P.density_filename, N)  # the reference's fmt::format(name, ppd), src/output.cpp:283
        dpath = os.path.join(outdir, name)
        for turn in range(world):
            if rank == turn:
                with open(dpath, "wb" if rank == 0 else "ab") as f:
                    for _, d in sorted(dens_parts, key=lambda t: t[0]):
                        f.write(d.tobytes())
            dist.barrier()
    t2 = time.perf_counter()
    st = ctx.stats()
    stats = torch.tensor([st["density_variance"], *st["max_disp"]], dtype=torch.float64, device=dev)
    allst = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allst, stats)
    if rank == 0:
        var = sum(float(s[0]) for s in allst)
        md = np.zeros(3)
        for s in allst:
            v = s[1:].cpu().numpy()
            md = np.where(np.abs(v) > np.abs(md), v, md)
        print(f"The rms density variation of the pixels is {np.sqrt(var / N**3):f}", file=sys.stderr)
        if qdensity != 2:  # reference src/zeldovich.cpp:998
            print(f"The maximum component-wise displacements are ({md[0]:g}, {md[1]:g}, {md[2]:g}), same units as BoxSize.",
                  file=sys.stderr)
        print(f"zeldovich took {t2 - t0:.4g} sec for ppd {N} on {world} GPUs ==> {N**3 / 1e6 / (t2 - t0):.3g} Mpart/sec "
              f"(device {t1 - t0:.3g} s, writing {t2 - t1:.3g} s)", file=sys.stderr)
    if world > 1 and args.exchange == "p2p":
        ex.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
