#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel of an .ncu-rep, grouped by source region.

    python tools/ncu_regions.py <report.ncu-rep> <kernel-name regex>

Reads `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo and --import-source on) and sums the per-line
"# Samples" / "Instructions Executed" columns over the regions of csrc/ that make up the three hot kernels.
"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur, hdr, per_line = None, None, collections.OrderedDict()
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Function Name":
        continue
    if hdr is None:
        hdr = r
        continue
    d = dict(zip(hdr, r))
    ln = d.get("Line No", "")
    if not ln.isdigit():
        continue
    try:
        s, n = int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0)
    except ValueError:
        continue
    a = per_line.setdefault((cur, int(ln)), [0, 0])
    a[0] += s
    a[1] += n


def src_line(fname, ln):
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zeldovich-plt_b200", "csrc", fname)
    try:
        return open(p).read().splitlines()[ln - 1].strip()
    except Exception:
        return ""


def region(fname, ln):
    t = src_line(fname, ln)
    if fname == "zplt_device.cuh":
        return "mode physics (zplt_device.cuh: RNG, Box-Muller, eigenmodes, growth)"
    if fname == "zplt_fft.cuh":
        if "S[" in t and "=" in t and t.split("=")[0].strip().startswith("S["):
            return "FFT: exchange stores (STS)"
        if "= S[" in t:
            return "FFT: exchange loads (LDS)"
        if "__syncthreads" in t or "__syncwarp" in t:
            return "FFT: barriers"
        return "FFT: butterflies and twiddles (FP64)"
    if "__syncthreads" in t:
        return "kernel: CTA barriers outside the FFT"
    if "st_stream" in t or "float4" in t or "put(" in t or "park_" in t or "ushort4" in t:
        return "kernel: global stores / parked fields"
    if "ld_stream" in t or "mbar_wait" in t or "= L[" in t:
        return "kernel: global loads / ring reads"
    return "kernel: other (index math, pencil build, statistics)"


tot_s = sum(a[0] for a in per_line.values()) or 1
tot_n = sum(a[1] for a in per_line.values()) or 1
agg = collections.Counter(), collections.Counter()
for (f, ln), (s, n) in per_line.items():
    r = region(f, ln)
    agg[0][r] += s
    agg[1][r] += n
print(f"kernel ~ /{kern}/ : {tot_s} stall samples, {tot_n} executed warp instructions (CUDA + SASS views summed)")
for r, s in agg[0].most_common():
    print(f"  {100 * s / tot_s:5.1f} % of samples  {100 * agg[1][r] / tot_n:5.1f} % of instructions   {r}")
