#!/usr/bin/env python
"""bench.py — IC particles/sec of the zeldovich-PLT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle/_ref)

One "step" = one pass of the hot path over one synthetic problem: mode generation ->
z-FFT -> y-FFT -> x-FFT + record emission, for PPD^3 particles (default workload: BASELINE.json
configs[3], PPD=1024 qPLT + rescale, single-precision RVZel records).

`value`  : particles/s with all inputs resident in HBM and the records left in HBM (CUDA events).
`e2e`    : particles/s through the C ABI from host buffers: power spline + eigenmode table copied
           host->device, records copied device->host (pinned), every step, wall clock.
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak.
`cpu_baseline`: the unmodified reference (oracle/_ref, shim FFT) on this box's host cores, bounded sample.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL's version banner goes to stdout by default; stdout must carry exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "IC particles/sec"
UNIT = "particles/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ppd", type=int, default=1024)
    ap.add_argument("--za", action="store_true", help="ZA instead of qPLT+rescale (2 packed arrays)")
    ap.add_argument("--icformat", default="RVZel")
    ap.add_argument("--ref-ppd", type=int, default=0, help="PPD of the CPU-reference sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: z-pass kernel stores into peer memory (p2p) or NCCL all_to_all_single")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks ----------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smmax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ inputs ----------
def write_inputs(tmp, ppd, qplt, icformat, synth, outdir="ic_out", numblock=4):
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=ppd**3, ICFormat='"%s"' % icformat, ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"),
                InitialConditionsDirectory='"%s"' % os.path.join(tmp, outdir), ZD_NumBlock=numblock)
    if qplt:
        eigp = os.path.join(tmp, "eigmodes128")
        if not os.path.exists(eigp):
            synth.write_eigmodes(eigp, 128)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % eigp)
    return synth.write_param(os.path.join(tmp, f"bench_{ppd}.par"), **over)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ------------------------------------------------------------------ CPU reference ---
def run_reference_sample(ppd, qplt, icformat, synth, threads):
    """One run of oracle/_ref/zeldovich_ref (unmodified reference sources + shim FFT) on a bounded sample."""
    ref = os.path.join(ROOT, "oracle", "_ref", "zeldovich_ref")
    if not os.path.exists(ref):
        return None
    with tempfile.TemporaryDirectory(prefix="zref_") as tmp:
        par = write_inputs(tmp, ppd, qplt, icformat, synth, numblock=4)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        t0 = time.perf_counter()
        r = subprocess.run([ref, par], cwd=tmp, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        err = r.stderr

        def grab(pat, n=1):
            m = re.search(pat, err)
            return [float(m.group(i + 1)) for i in range(n)] if m else [float("nan")] * n

        pre = grab(r"Preamble took ([0-9.e+-]+) seconds")[0]
        gen = grab(r"Computing, Saving the Planes took ([0-9.e+-]+) ([0-9.e+-]+) sec", 2)
        xy = grab(r"Loading, FFTs, Writing took ([0-9.e+-]+) ([0-9.e+-]+) ([0-9.e+-]+) seconds", 3)
        hot = gen[0] + gen[1] + xy[0] + xy[1] + xy[2]
        return dict(ppd=ppd, wall_s=wall, preamble_s=pre, hot_path_s=hot, gen_zfft_s=gen[0], block_copy_s=gen[1] + xy[0],
                    fft2d_s=xy[1], write_s=xy[2])


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def auto_ref_ppd():
    try:
        kb = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1))
    except Exception:
        kb = 16 << 20
    return 512 if kb > (48 << 20) else 256


def cpu_baseline(args, synth, qplt):
    threads = os.cpu_count() or 1
    ppd = args.ref_ppd or auto_ref_ppd()
    s = run_reference_sample(ppd, qplt, args.icformat, synth, threads)
    if s is None:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref/zeldovich_ref not built"}
    if "error" in s:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "reference run failed: " + s["error"]}
    return {
        "value": ppd**3 / s["hot_path_s"], "unit": UNIT, "cores": threads, "kind": "reference",
        "sample": (f"unmodified reference sources (oracle/_ref, shim radix-4 FFT, not FFTW) PPD={ppd} "
                   f"{'qPLT+rescale' if qplt else 'ZA'} {args.icformat}, OMP_NUM_THREADS={threads} on {cpu_model()}; "
                   f"hot path {s['hot_path_s']:.2f}s (gen+zFFT {s['gen_zfft_s']:.2f}, block copies {s['block_copy_s']:.2f}, "
                   f"2-D FFT {s['fft2d_s']:.2f}, WriteParticlesSlab {s['write_s']:.2f}); preamble {s['preamble_s']:.2f}s excluded; "
                   f"whole process {s['wall_s']:.2f}s = {ppd**3 / s['wall_s']:.3g} particles/s"),
        "whole_run_value": ppd**3 / s["wall_s"],
    }


# ------------------------------------------------------------------ reference arm ---
def main_reference(args, rank, world):
    if rank != 0:
        return
    from __graft_entry__ import load_synth

    synth = load_synth()
    qplt = not args.za
    threads = os.cpu_count() or 1
    ppd = args.ref_ppd or auto_ref_ppd()
    times = []
    last = None
    for i in range(args.warmup + args.steps):
        s = run_reference_sample(ppd, qplt, args.icformat, synth, threads)
        if s is None or "error" in s:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/zeldovich_ref missing or failed"}))
            return
        if i >= args.warmup:
            times.append(s["hot_path_s"])
        last = s
        if sum(times) > 150:  # bounded: keep the whole arm within a few minutes
            break
    t = sum(times) / len(times)
    val = ppd**3 / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"PPD={ppd} {'qPLT+rescale' if qplt else 'ZA'} {args.icformat} (bounded CPU sample of the "
                               f"PPD={args.ppd} workload)", "ppd": ppd},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"oracle/_ref (reference sources + shim FFT), hot path only (preamble {last['preamble_s']:.2f}s "
                                   f"excluded), OMP_NUM_THREADS={threads}, {cpu_model()}"},
        "e2e": {"value": ppd**3 / last["wall_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ CUDA arm --------
def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from __graft_entry__ import load_package, load_synth

    pkg = load_package()
    synth = load_synth()
    pkg.lib()  # raises if the native library is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N = args.ppd
    qplt = not args.za
    na = 4 if qplt else 2
    tmp = tempfile.mkdtemp(prefix="zbench_")
    par = write_inputs(tmp, N, qplt, args.icformat, synth)
    P = pkg.Parameters(par)
    power = pkg.PowerSpectrum(P)
    cfg = P.config(device=local_rank)
    cfg.rank, cfg.nranks = rank, world  # world > 1: slab decomposition, one all-to-all per step
    ctx = pkg.Context(cfg)
    rb = ctx.record_bytes
    nloc = N // world  # z planes this rank emits

    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        import importlib.util

        from __graft_entry__ import PKG_DIR

        spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(PKG_DIR, "distributed.py"))
        zd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(zd)
        ws = zd.PeerExchange(ctx) if args.exchange == "p2p" else zd.SlabWorkspace(ctx, dev)
    else:
        work = torch.empty(ctx.workspace_bytes(), dtype=torch.uint8, device=dev)
        ctx.set_workspace(work.data_ptr(), work.numel())
    # records stay in HBM; if the full set does not fit beside the slabs, planes cycle through a smaller buffer
    free_b = torch.cuda.mem_get_info(dev)[0]
    plane_b = N * N * rb
    out_planes = max(1, min(nloc, int((free_b - (6 << 30)) // plane_b)))
    out = torch.empty(out_planes * plane_b, dtype=torch.uint8, device=dev)
    power.apply(ctx)
    if qplt:
        ctx.load_eigenmodes_file(P.PLT_filename)
    a2a_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step():
        ctx.generate()
        if world > 1:
            a2a_ev[0].record(stream)
            ws.exchange()
            a2a_ev[1].record(stream)
        for z0 in range(0, nloc, out_planes):
            ctx.emit_planes(z0, min(out_planes, nloc - z0), out.data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
    # per-stage CUDA-event times of the last step (events recorded by the library on the same stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    tm = ctx.timings()
    stage = [tm["generate_ms"], tm["zfft_ms"], tm["xfft_emit_ms"]]
    a2a_ms = a2a_ev[0].elapsed_time(a2a_ev[1]) if world > 1 else 0.0
    launches = sum(tm["launches"]) * args.steps
    clocks = sampler.stop()
    red = torch.tensor([total_ms / args.steps, a2a_ms] + stage, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    ms_per_step = float(red[0].item())
    a2a_ms = float(red[1].item())
    stage = [float(v) for v in red[2:].tolist()]

    # ---- end to end through the C ABI with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        plane = N * N * rb
        # one pinned host buffer for all of this rank's records when it is affordable (<= 40 GB), else 2 GiB chunks
        chunk = nloc if nloc * plane <= (40 << 30) else max(1, min(nloc, (2 << 30) // plane))
        try:
            pinned = torch.empty(chunk * plane, dtype=torch.uint8, pin_memory=True)
        except RuntimeError:
            chunk = max(1, min(nloc, (2 << 30) // plane))
            pinned = torch.empty(chunk * plane, dtype=torch.uint8, pin_memory=True)
        x, y, y2 = power.arrays()
        eig_tab = synth.read_eigmodes(P.PLT_filename)[1] if qplt else None

        def pinned_copy(a):
            # the step's host inputs live in pinned memory (numpy views of pinned torch tensors); pageable if pinning fails
            try:
                import numpy as np

                t = torch.from_numpy(np.array(a, order="C", copy=True)).pin_memory()
                keep_alive.append(t)
                return t.numpy()
            except Exception:
                return a

        keep_alive = []
        x, y, y2 = pinned_copy(x), pinned_copy(y), pinned_copy(y2)
        if qplt:
            eig_tab = pinned_copy(eig_tab)
        h2d = x.nbytes * 3 + (eig_tab.nbytes if qplt else 0)
        nsteps = max(1, min(args.steps, 3))

        def e2e_step():
            ctx.set_power_spline(x, y, y2, power.normalization, power.Pk_smooth2)  # H2D + table kernel
            if qplt:
                ctx.set_eigenmodes(128, eig_tab)  # H2D
            with torch.cuda.stream(stream):
                ctx.generate()
                if world > 1:
                    ws.exchange()
            for z0 in range(0, nloc, chunk):
                ctx.fetch_planes_ptr(z0, min(chunk, nloc - z0), pinned.data_ptr())  # D2H of every record

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / nsteps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": N**3 / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(N**3 * rb), "steps": nsteps, "seconds_per_step": float(dt.item()),
               "note": f"host spline+eigenmode tables -> device, every record -> pinned host ({chunk} planes per fetch call), wall clock"}

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel = the slower of the two strided FFT passes (K2: read + write 16*narray B per particle each way)
        names = ["generate+x-FFT", "z-FFT", "y-FFT+emit"]
        alg_bytes = [v // world for v in (16 * na * N**3, 32 * na * N**3, (16 * na + rb) * N**3)]  # per GPU: write; r+w; read+records
        # N=1: the in-place strided pass (z axis), reads and writes every array once.  N>1: that pass is fused with the
        # NVLink exchange and overlapped with generation, so the local HBM-bound kernel is the y pass + emission.
        dom = 1 if world == 1 else 2
        achieved = alg_bytes[dom] / (stage[dom] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1:  # the captures are single-GPU launches; a slab rank's launch moves 1/world of it
            try:
                traffic = json.load(open(tp)).get(f"{names[dom]}@{N}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": N**3 / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PPD={N} {'qPLT+rescale' if qplt else 'ZA'} {args.icformat}, synthetic BBKS P(k) + synthetic eigmodes128",
                       "ppd": N, "narray": na, "record_bytes": rb,
                       "parallelism": "single GPU" if world == 1 else (
                           f"slab decomposition over {world} GPUs: y-row pairs -> "
                           + ("z-FFT kernel storing into peer memory over NVLink (CUDA IPC)" if args.exchange == "p2p" else "NCCL all_to_all_single")
                           + " -> z planes"),
                       "l2": f"inputs larger than L2 ({16 * na * N**3 / 1e9:.1f} GB spectral arrays streamed per pass)"},
            "stage_ms": dict(zip(names, stage)),
            "stage_gbs": {n: alg_bytes[i] / (stage[i] * 1e-3) / 1e9 for i, n in enumerate(names)},
            "roofline": {"bound": "hbm", "kernel": ("fft_tile_ring_kernel" if dom == 1 else "fft_emit_strided_kernel") + f" ({names[dom]})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes[dom]},
            "clocks": clocks, "gpu_launches": launches,
            "all_to_all": None if world == 1 else {
                "how": args.exchange,
                # p2p: the exchange IS the z-FFT kernel, run in row groups on a second stream while the next group is
                # generated; stage_ms["generate+x-FFT"] + stage_ms["z-FFT"] is the whole overlapped stage 1, "ms" only the
                # final sync + barrier.  The NVLink rate is therefore a lower bound (bytes / whole stage-1 time).
                "ms": a2a_ms, "bytes_sent_per_gpu": int(16 * na * N**3 // world * (world - 1) // world),
                "nvlink_gbs_per_gpu": 16 * na * N**3 / world * (world - 1) / world
                / (((stage[0] + stage[1]) if args.exchange == "p2p" else a2a_ms) * 1e-3) / 1e9,
                "reference_peer_copy_gbs": 770.0},
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, synth, qplt)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
