#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/sweep.py [--ppd 1024 2048] — one launch, many tuning variants.

The tuning switches are runtime options of a context (zplt_set_option), so a sweep over the stage-1 SM split
(p2p_ctas), the number of row groups (slab_groups) and the kernel forms costs one process start per problem size
instead of one per variant.  Prints one JSON line per variant on rank 0 (device-timed, max over ranks).
"""
import argparse
import importlib.util
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import PKG_DIR, load_package, load_synth  # noqa: E402

VARIANTS = {
    "default (per-source receive layout)": {},
    "rows at their true y": {"b2_layout": 0},
    "rows at their true y, pad 4104": {"b2_layout": 0, "b2_pad": 4104},
    "per-source, 96 CTAs": {"p2p_ctas": 96},
    "per-source, per-group launches": {"p2p_resident": 0},
}
VARIANTS_2048 = {
    "default (per-source receive layout)": {},
    "rows at their true y": {"b2_layout": 0},
    "per-source, 64 CTAs": {"p2p_ctas": 64},
    "per-source, 98 CTAs": {"p2p_ctas": 98},
}
DEFAULTS = {"p2p_ctas": -1, "slab_groups": 16, "slab_ring": 1, "dit2048": 1, "dit2048_emit": 0, "p2p_resident": 1, "p2p_helper": 1, "b2_layout": -1, "b2_pad": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppd", type=int, nargs="+", default=[1024])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg, synth = load_package(), load_synth()
    spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(PKG_DIR, "distributed.py"))
    zd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(zd)
    args = argparse.Namespace(icformat="RVZel", exchange="p2p")
    tmp = tempfile.mkdtemp(prefix="zsweep_")
    for N in a.ppd:
        pb = bench.Problem(pkg, synth, zd, torch, dist, args, N, True, rank, world, local, dev, tmp)
        table = VARIANTS_2048 if N == 2048 else VARIANTS
        for name, opts in table.items():
            if a.only and name not in a.only:
                continue
            for k, v in {**DEFAULTS, **opts}.items():
                pb.ctx.set_option(k, v)
            ms, a2a, stage, _ = pb.measure(a.steps, 2)
            if rank == 0:
                sent = 16 * 4 * N**3 // world * (world - 1) // world
                print(json.dumps({"ppd": N, "gpus": world, "variant": name, "opts": opts, "ms_per_step": ms, "stage1_ms": stage[0] + stage[1],
                                  "stage2_ms": stage[2], "sync_ms": a2a, "nvlink_gbs_per_gpu": sent / ((stage[0] + stage[1]) * 1e-3) / 1e9,
                                  "gpart_s": N**3 / ms / 1e6}), flush=True)
        pb.close()
        del pb
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
