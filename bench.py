#!/usr/bin/env python
"""bench.py — IC particles/sec of the zeldovich-PLT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle/_ref)

One "step" = one pass of the hot path over one synthetic problem: mode generation + x-FFT ->
z-FFT (fused with the NVLink exchange on N > 1 GPUs) -> y-FFT + record emission, for PPD^3 particles
(default workload: BASELINE.json configs[3], PPD=1024 qPLT + rescale, single-precision RVZel records;
N > 1 shards the same problem: strong scaling).

`value`   : particles/s with all inputs resident in HBM and the records left in HBM (CUDA events).
`e2e`     : particles/s through the C ABI from host buffers: power spline + eigenmode table copied
            host->device, records copied device->host (pinned), every step, wall clock.
`e2e_files`: wall clock of the drop-in call itself, `zeldovich <param_file>` (zplt_run_param_file), writing every ic_* file.
`roofline`: every kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak; `kernel` = the longest one.
`parity`  : planes of the very problem being timed, fetched after the timed region and compared with the CPU oracle
            (oracle/zel_oracle.c as the checker only, outside every timed region).
`cpu_baseline`: the unmodified reference (oracle/_ref, shim FFT) on this box's host cores, bounded sample.
"""
import argparse
import json
import os
import re
import shutil
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
KERNELS = ["generate+x-FFT", "z-FFT", "y-FFT+emit"]


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
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip e2e_files (the run that writes every ic_* file)")
    ap.add_argument("--no-ppd2048", action="store_true", help="8 GPUs: skip the PPD=2048 sub-record (BASELINE configs[4])")
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


def workload_config(ppd, qplt, icformat, world, exchange):
    na = 4 if qplt else 2
    return {"workload": f"PPD={ppd} {'qPLT+rescale' if qplt else 'ZA'} {icformat}, synthetic BBKS P(k) + synthetic eigmodes128",
            "ppd": ppd, "narray": na,
            "parallelism": "single GPU" if world == 1 else (
                f"slab decomposition over {world} GPUs: y-row pairs -> "
                + ("z-FFT kernel storing into peer memory over NVLink (CUDA IPC)" if exchange == "p2p" else "NCCL all_to_all_single")
                + " -> z planes"),
            "l2": f"inputs larger than L2 ({16 * na * ppd**3 / 1e9:.1f} GB spectral arrays streamed per pass)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def mem_available_gb():
    try:
        return int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) / (1 << 20)
    except Exception:
        return 16.0


def bind_to_gpu_numa(index):
    """Run this process (and place the pages it pins from now on) on the NUMA node the GPU hangs off.  Best effort."""
    try:
        import torch

        p = torch.cuda.get_device_properties(index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "GPU reports no NUMA node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return f"no usable CPU on NUMA node {node}"
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


# ------------------------------------------------------------------ CPU reference ---
def run_reference_sample(ppd, qplt, icformat, synth, threads, tmp_root=None):
    """One run of oracle/_ref/zeldovich_ref (unmodified reference sources + shim FFT) on a bounded sample."""
    ref = os.path.join(ROOT, "oracle", "_ref", "zeldovich_ref")
    if not os.path.exists(ref):
        return None
    with tempfile.TemporaryDirectory(prefix="zref_", dir=tmp_root) as tmp:
        par = write_inputs(tmp, ppd, qplt, icformat, synth, numblock=8 if ppd >= 1024 else 4)
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


def ref_ppd_for(args, qplt):
    """The reference runs the bench workload itself when the host can hold it (in-RAM BlockArray: 16*narray*N^3*(1+2/NB) B plus
    the output buffer and the ic files in /tmp), otherwise a smaller sample; the line says which."""
    if args.ref_ppd:
        return args.ref_ppd, "as requested"
    na = 4 if qplt else 2
    need_gb = (16 * na * args.ppd**3 * 1.25 + 2 * 32 * args.ppd**3) / 2**30 + 8  # arrays + records in the page cache + slack
    avail = mem_available_gb()
    if avail >= need_gb:
        return args.ppd, f"the bench workload itself (MemAvailable {avail:.0f} GB >= {need_gb:.0f} GB needed)"
    ppd = args.ppd
    while ppd > 128 and (16 * na * ppd**3 * 1.5 + 2 * 32 * ppd**3) / 2**30 + 4 > avail:
        ppd //= 2
    return ppd, f"bounded sample: MemAvailable {avail:.0f} GB < {need_gb:.0f} GB the PPD={args.ppd} reference run needs"


def cpu_baseline(args, synth, qplt):
    threads = os.cpu_count() or 1
    ppd, why = ref_ppd_for(args, qplt)
    s = run_reference_sample(ppd, qplt, args.icformat, synth, threads)
    if s is None:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref/zeldovich_ref not built"}
    if "error" in s:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "reference run failed: " + s["error"]}
    return {
        "value": ppd**3 / s["hot_path_s"], "unit": UNIT, "cores": threads, "kind": "reference",
        "sample": (f"unmodified reference sources (oracle/_ref, shim radix-4 FFT, not FFTW) PPD={ppd} ({why}) "
                   f"{'qPLT+rescale' if qplt else 'ZA'} {args.icformat}, OMP_NUM_THREADS={threads} on {cpu_model()}; "
                   f"hot path {s['hot_path_s']:.2f}s (gen+zFFT {s['gen_zfft_s']:.2f}, block copies {s['block_copy_s']:.2f}, "
                   f"2-D FFT {s['fft2d_s']:.2f}, WriteParticlesSlab {s['write_s']:.2f}); preamble {s['preamble_s']:.2f}s excluded; "
                   f"whole process {s['wall_s']:.2f}s = {ppd**3 / s['wall_s']:.3g} particles/s"),
        "whole_run_value": ppd**3 / s["wall_s"], "ppd": ppd,
    }


# ------------------------------------------------------------------ reference arm ---
def main_reference(args, rank, world):
    if rank != 0:
        return
    from __graft_entry__ import load_synth

    synth = load_synth()
    qplt = not args.za
    threads = os.cpu_count() or 1
    ppd, why = ref_ppd_for(args, qplt)
    # a step of the reference at PPD=1024 is ~half a minute of all host cores: bound the arm to ~2-3 minutes
    budget_s = 120.0
    times, last, nwarm = [], None, 0
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        s = run_reference_sample(ppd, qplt, args.icformat, synth, threads)
        if s is None or "error" in s:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/zeldovich_ref missing or failed"}))
            return
        last = s
        # long runs (PPD=1024: most of a minute each) get one warm-up run, short ones all the requested ones
        warm = i < (args.warmup if s["wall_s"] < 10 else min(args.warmup, 1))
        if warm and not times:
            nwarm += 1
        else:
            times.append(s["hot_path_s"])
        if len(times) >= args.steps or (times and time.perf_counter() - t_start > budget_s):
            break
    t = sum(times) / len(times)
    val = ppd**3 / t
    cfg = workload_config(args.ppd, qplt, args.icformat, 1, args.exchange)
    cfg["parallelism"] = f"reference CPU path, OpenMP over {threads} host threads"
    cfg["sample"] = f"PPD={ppd}: {why}; {len(times)} timed runs after {nwarm} warm-up (bounded to ~{budget_s:.0f} s of host time)"
    cfg["sample_ppd"] = ppd
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": nwarm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"oracle/_ref (reference sources + shim FFT, not FFTW) at PPD={ppd}, hot path only (preamble "
                                   f"{last['preamble_s']:.2f}s excluded; its WriteParticlesSlab fwrite of every ic file, "
                                   f"{last['write_s']:.2f}s, included), OMP_NUM_THREADS={threads}, {cpu_model()}"},
        "e2e": {"value": ppd**3 / last["wall_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "whole process wall clock, ic_* files written"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ CUDA arm --------
class Problem:
    """One context of the product on this rank plus what the measurement needs around it."""

    def __init__(self, pkg, synth, zd, torch, dist, args, N, qplt, rank, world, local_rank, dev, tmp):
        self.pkg, self.torch, self.dist, self.world, self.rank, self.dev, self.N = pkg, torch, dist, world, rank, dev, N
        self.qplt, self.na = qplt, 4 if qplt else 2
        self.par = write_inputs(tmp, N, qplt, args.icformat, synth)
        self.P = pkg.Parameters(self.par)
        self.power = pkg.PowerSpectrum(self.P)
        cfg = self.P.config(device=local_rank)
        cfg.rank, cfg.nranks = rank, world  # world > 1: slab decomposition, one exchange per step
        self.ctx = ctx = pkg.Context(cfg)
        self.rb = ctx.record_bytes
        self.nloc = N // world  # z planes this rank emits
        self.stream = torch.cuda.Stream(device=dev)
        ctx.set_stream(self.stream.cuda_stream)
        self.p2p = world > 1 and args.exchange == "p2p"
        if world > 1:
            self.ws = zd.PeerExchange(ctx) if self.p2p else zd.SlabWorkspace(ctx, dev)
        else:
            self.work = torch.empty(ctx.workspace_bytes(), dtype=torch.uint8, device=dev)
            ctx.set_workspace(self.work.data_ptr(), self.work.numel())
        # records stay in HBM; if the full set does not fit beside the slabs, planes cycle through a smaller buffer
        free_b = torch.cuda.mem_get_info(dev)[0]
        plane_b = N * N * self.rb
        self.out_planes = max(1, min(self.nloc, int((free_b - (6 << 30)) // plane_b)))
        self.out = torch.empty(self.out_planes * plane_b, dtype=torch.uint8, device=dev)
        self.power.apply(ctx)
        if qplt:
            ctx.load_eigenmodes_file(self.P.PLT_filename)
        self.a2a_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step(self):
        ctx = self.ctx
        if self.world > 1:
            self.ws.begin()  # p2p: everybody has finished emitting from the buffers the coming z pass stores into
        ctx.generate()
        if self.world > 1:
            self.a2a_ev[0].record(self.stream)
            self.ws.exchange()
            self.a2a_ev[1].record(self.stream)
        for z0 in range(0, self.nloc, self.out_planes):
            ctx.emit_planes(z0, min(self.out_planes, self.nloc - z0), self.out.data_ptr())

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def measure(self, steps, warmup, sampler=None):
        """(ms_per_step, all-to-all ms, [stage ms], launches): CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.step()
        self.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(steps):
                self.step()
            e1.record(self.stream)
        self.barrier()
        total_ms = e0.elapsed_time(e1)
        tm = self.ctx.timings()  # per-stage CUDA-event times of the last step (events recorded by the library on the same stream)
        stage = [tm["gen_xfft_ms"], tm["zfft_ms"], tm["yfft_emit_ms"]]
        a2a_ms = self.a2a_ev[0].elapsed_time(self.a2a_ev[1]) if self.world > 1 else 0.0
        red = torch.tensor([total_ms / steps, a2a_ms] + stage, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(red, op=self.dist.ReduceOp.MAX)
        r = [float(v) for v in red.tolist()]
        return r[0], r[1], r[2:], sum(tm["launches"]) * steps

    def parity(self, max_planes=3):
        """Planes of the problem just timed against the CPU oracle: rank 0 computes the expected records (oracle.planes — the
        checker, outside every timed region), every owner compares its own plane."""
        import numpy as np

        torch, dist, N, world = self.torch, self.dist, self.N, self.world
        zs = sorted({0, (world // 2) * self.nloc + min(1, self.nloc - 1), N - 1})[:max_planes]
        t0 = time.perf_counter()
        dt = self.ctx.record_dtype
        failed = ""
        if self.rank == 0:
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import zel_oracle as zo

                from __graft_entry__ import load_synth

                synth = load_synth()
                zo.set_threads(os.cpu_count() or 1)  # torchrun sets OMP_NUM_THREADS=1 for its ranks; the other ranks only wait here
                k, p = synth.make_power_table()
                kw = dict(ppd=N, icformat=self.ctx.cfg.icformat)
                if self.qplt:
                    kw.update(qPLT=1, qPLTrescale=1, PLT_target_z=5.0)
                want, _ = zo.planes(zo.make_config(**kw), (k, p), zs, (128, synth.make_eigmodes(128)) if self.qplt else None)
                want_t = torch.from_numpy(want.view(np.uint8).reshape(len(zs), -1).copy())
            except Exception as e:  # noqa: BLE001 — the checker must not take the measurement down with it
                failed = f"oracle unavailable: {type(e).__name__}: {e}"[:200]
                want_t = torch.zeros((len(zs), N * N * self.rb), dtype=torch.uint8)
        else:
            want_t = torch.empty((len(zs), N * N * self.rb), dtype=torch.uint8)
        if world > 1:
            flag = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device=self.dev)
            dist.broadcast(flag, src=0)
            if flag.item() != 0.0:
                return {"ok": None, "error": failed or "oracle unavailable on rank 0", "planes": zs}
        elif failed:
            return {"ok": None, "error": failed, "planes": zs}
        if world > 1:
            want_d = want_t.to(self.dev)
            dist.broadcast(want_d, src=0)
            want_t = want_d.cpu()
            del want_d
        worst, ids_ok, checked = 0.0, True, 0
        for i, z in enumerate(zs):
            if z // self.nloc != self.rank:
                continue
            got = self.ctx.fetch_planes(z - self.rank * self.nloc, 1)
            w = want_t[i].numpy().view(dt)
            checked += 1
            if "ijk" in dt.names:
                ids_ok = ids_ok and bool(np.array_equal(got["ijk"], w["ijk"])) and bool(np.all(got["pad"] == 0))
            for f in ("displ", "vel"):
                if f in dt.names:
                    for c in range(3):
                        a, b = got[f][:, c].astype(np.float64), w[f][:, c].astype(np.float64)
                        worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(b))))
        red = torch.tensor([worst, 0.0 if ids_ok else 1.0, float(checked)], dtype=torch.float64, device=self.dev)
        if world > 1:
            tot = red.clone()
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            checked = int(tot[2].item())
        worst, ids_ok = float(red[0].item()), red[1].item() == 0.0
        tol = 2e-7 if self.ctx.cfg.icformat in (1, 3) else 1e-10
        return {"ok": bool(ids_ok and worst < tol and checked == len(zs)), "planes": zs, "planes_checked": checked, "ids_exact": bool(ids_ok),
                "max_field_rel_err": worst, "tolerance": tol,
                "how": "planes of the timed problem fetched after the timed region vs oracle/zel_oracle.c zo_planes (CPU, checker only)",
                "seconds": time.perf_counter() - t0}

    def close(self):
        if self.p2p:
            self.ws.close()  # unmap the peers' buffers everywhere before anybody frees its own
        self.ctx.close()
        for name in ("ws", "work", "out"):
            if hasattr(self, name):
                delattr(self, name)
        self.torch.cuda.empty_cache()


def roofline_block(N, na, rb, world, stage, ms_per_step, p2p):
    peak, peak_src = peaks()
    # per GPU and per launch set: K1 writes the arrays once, K2 reads and writes them, K3 reads them and writes the records
    alg = [16 * na * N**3 // world, 32 * na * N**3 // world, (16 * na + rb) * N**3 // world]
    names = {0: "gen_xfft_kernel", 1: "fft_tile_p2p_ring_kernel (z pass + NVLink exchange)" if p2p else "fft_tile_ring_kernel",
             2: "fft_emit_ring_kernel"}
    kern = {}
    for i, n in enumerate(KERNELS):
        if stage[i] > 0:
            a = alg[i] / (stage[i] * 1e-3) / 1e9
            kern[n] = {"kernel": names[i], "ms": stage[i], "algorithmic_bytes": alg[i], "achieved": a, "frac": a / peak}
    if p2p:
        # the z pass runs on a second stream inside the generation stage (it is NVLink-bound; see all_to_all): one entry for both
        kern.pop("z-FFT", None)
        st1 = stage[0] + stage[1]
        a = (alg[0] + alg[1]) / (st1 * 1e-3) / 1e9
        kern["generate+x-FFT"] = {"kernel": "gen_xfft_kernel overlapped with fft_tile_p2p_ring_kernel (stage 1 of a slab rank)", "ms": st1,
                                  "algorithmic_bytes": alg[0] + alg[1], "achieved": a, "frac": a / peak}
    dom = max(kern, key=lambda n: kern[n]["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:  # the captures are single-GPU launches
        try:
            traffic = json.load(open(tp)).get(f"{dom}@{N}")
        except Exception:
            traffic = None
    total = sum(alg)
    step_a = total / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": f"{kern[dom]['kernel']} ({dom}) — the longest kernel of the step", "achieved": kern[dom]["achieved"],
            "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"], "kernels": kern,
            "step_algorithmic_bytes": total, "step_achieved": step_a, "step_frac": step_a / peak}


def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from __graft_entry__ import PKG_DIR, load_package, load_synth

    pkg = load_package()
    synth = load_synth()
    pkg.lib()  # raises if the native library is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)  # before anything is pinned
    zd = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        import importlib.util

        spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(PKG_DIR, "distributed.py"))
        zd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(zd)

    N = args.ppd
    qplt = not args.za
    na = 4 if qplt else 2
    tmp = tempfile.mkdtemp(prefix="zbench_")
    pb = Problem(pkg, synth, zd, torch, dist, args, N, qplt, rank, world, local_rank, dev, tmp)
    ctx, rb, nloc, stream = pb.ctx, pb.rb, pb.nloc, pb.stream

    sampler = ClockSampler(local_rank)
    ms_per_step, a2a_ms, stage, launches = pb.measure(args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    parity = None if args.no_parity else pb.parity()

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
        x, y, y2 = pb.power.arrays()
        eig_tab = synth.read_eigmodes(pb.P.PLT_filename)[1] if qplt else None
        keep_alive = []

        def pinned_copy(a):
            # the step's host inputs live in pinned memory (numpy views of pinned torch tensors); pageable if pinning fails
            try:
                import numpy as np

                t = torch.from_numpy(np.array(a, order="C", copy=True)).pin_memory()
                keep_alive.append(t)
                return t.numpy()
            except Exception:
                return a

        x, y, y2 = pinned_copy(x), pinned_copy(y), pinned_copy(y2)
        if qplt:
            eig_tab = pinned_copy(eig_tab)
        h2d = x.nbytes * 3 + (eig_tab.nbytes if qplt else 0)
        nsteps = max(1, min(args.steps, 3))
        split = {"inputs": 0.0, "generate+exchange": 0.0, "emit+d2h": 0.0}

        def e2e_step():
            t0 = time.perf_counter()
            ctx.set_power_spline(x, y, y2, pb.power.normalization, pb.power.Pk_smooth2)  # H2D + table kernel
            if qplt:
                ctx.set_eigenmodes(128, eig_tab)  # H2D
            t1 = time.perf_counter()
            if world > 1:
                pb.ws.begin()
            with torch.cuda.stream(stream):
                ctx.generate()
                if world > 1:
                    pb.ws.exchange()
            if world == 1:
                ctx.synchronize()
            t2 = time.perf_counter()
            for z0 in range(0, nloc, chunk):
                ctx.fetch_planes_ptr(z0, min(chunk, nloc - z0), pinned.data_ptr())  # D2H of every record
            t3 = time.perf_counter()
            split["inputs"] += t1 - t0
            split["generate+exchange"] += t2 - t1
            split["emit+d2h"] += t3 - t2

        e2e_step()
        pb.barrier()
        for k in split:
            split[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(nsteps):
            e2e_step()
        pb.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / nsteps] + [split[k] / nsteps for k in split], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dtl = [float(v) for v in dt.tolist()]
        e2e = {"value": N**3 / dtl[0], "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(N**3 * rb), "steps": nsteps, "seconds_per_step": dtl[0],
               "seconds_split_max_over_ranks": dict(zip(split.keys(), dtl[1:])),
               "d2h_gbs_per_gpu": nloc * plane / dtl[3] / 1e9 if dtl[3] > 0 else None, "numa": numa,
               "note": f"host spline+eigenmode tables -> device, every record -> pinned host ({chunk} planes per fetch call), wall clock; "
                       "bound by the PCIe device->host copy of the records"}
        del pinned, keep_alive

    ppd2048 = None
    e2e_files = None
    pb.close()
    del pb, ctx

    # ---- the drop-in call itself: zeldovich <param_file>, every ic_* file written ----------
    if world == 1 and rank == 0 and not args.no_files:
        need = N**3 * rb + (4 << 30)
        free = shutil.disk_usage(tmp).free
        if free < need or mem_available_gb() * 2**30 < need:
            e2e_files = {"value": None, "note": f"skipped: {need / 1e9:.0f} GB of ic files do not fit {tmp} ({free / 1e9:.0f} GB free)"}
        else:
            try:
                t0 = time.perf_counter()
                rep = pkg.run_param_file(write_inputs(tmp, N, qplt, args.icformat, synth, outdir="ic_files"), device=local_rank, write_files=True)
                wall = time.perf_counter() - t0
                e2e_files = {"value": N**3 / wall, "unit": UNIT, "seconds": wall, "seconds_preamble": rep.seconds_preamble,
                             "seconds_device_and_d2h": rep.seconds_device, "seconds_fwrite": rep.seconds_write,
                             "bytes_written": int(rep.bytes_written), "files": int(rep.files_written), "out_of_core_passes": int(rep.ooc_passes),
                             "note": "zplt_run_param_file = what bin/zeldovich <param_file> runs: parameter file, P(k) normalisation, "
                                     f"eigenmode file, generation, every ic_* file written under {tmp} (the reference arm's e2e writes its "
                                     "files the same way)"}
            except Exception as e:  # noqa: BLE001
                e2e_files = {"value": None, "note": f"failed: {e}"}
            shutil.rmtree(os.path.join(tmp, "ic_files"), ignore_errors=True)

    # ---- BASELINE configs[4]: PPD=2048 qPLT + rescale across 8 GPUs ---------------------------
    want2048 = world == 8 and args.exchange == "p2p" and not args.no_ppd2048 and N != 2048
    if want2048:
        # every rank must be able to hold its 128 GiB of slabs plus a few record planes, or none starts (a rank failing alone
        # would leave the others waiting in the exchange set-up)
        torch.cuda.empty_cache()
        ok = torch.tensor([1.0 if torch.cuda.mem_get_info(dev)[0] > (140 << 30) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0.0:
            want2048 = False
            ppd2048 = {"error": "not every GPU has 140 GiB free for the PPD=2048 slabs"}
    if want2048:
        try:
            pb2 = Problem(pkg, synth, zd, torch, dist, args, 2048, True, rank, world, local_rank, dev, tmp)
            ms2, a2a2, st2, _ = pb2.measure(3, 2)
            par2 = None if args.no_parity else pb2.parity(max_planes=3)
            rf2 = roofline_block(2048, 4, pb2.rb, world, st2, ms2, True)
            sent = 16 * 4 * 2048**3 // world * (world - 1) // world
            ppd2048 = {"workload": "PPD=2048 qPLT+rescale RVZel (BASELINE configs[4]) on 8 GPUs", "ms_per_step": ms2, "value": 2048**3 / (ms2 * 1e-3),
                       "unit": UNIT, "steps": 3, "warmup": 2, "stage_ms": {"stage 1 (generate+x-FFT || z-FFT+exchange)": st2[0] + st2[1], "y-FFT+emit": st2[2]},
                       "nvlink_gbs_per_gpu": sent / ((st2[0] + st2[1]) * 1e-3) / 1e9, "bytes_sent_per_gpu": sent,
                       "ypass_frac": rf2["kernels"]["y-FFT+emit"]["frac"], "step_frac": rf2["step_frac"], "parity": par2}
            pb2.close()
        except Exception as e:  # noqa: BLE001
            ppd2048 = {"error": str(e)[:300]}

    if rank == 0:
        p2p = world > 1 and args.exchange == "p2p"
        cfg = workload_config(N, qplt, args.icformat, world, args.exchange)
        cfg["record_bytes"] = rb
        line = {
            "metric": METRIC, "value": N**3 / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "stage_ms": dict(zip(KERNELS, stage)),
            "roofline": roofline_block(N, na, rb, world, stage, ms_per_step, p2p),
            "clocks": clocks, "gpu_launches": launches, "parity": parity,
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
        if e2e_files:
            line["e2e_files"] = e2e_files
        if ppd2048:
            line["ppd2048"] = ppd2048
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, synth, qplt)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    shutil.rmtree(tmp, ignore_errors=True)


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
