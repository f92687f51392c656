# memcheck + racecheck of the whole hot path at small sizes (the reference's CI uses ASan/UBSan)
cat > /tmp/san_case.py <<'PY'
import os, sys, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from __graft_entry__ import load_package, load_synth
import numpy as np
pkg, synth = load_package(), load_synth()
# N = 64 exercises the TMA-ring kernels (z pass and RVZel emission); the last case runs the ZD_f_NL potential pass
for N, qplt, fmt, G, fnl in ((32, 1, "RVZel", 1, 0), (64, 1, "RVZel", 1, 0), (64, 0, "RVdoubleZel", 1, 0), (32, 1, "RVZel", 2, 0), (64, 0, "RVZel", 1, 2000)):
    tmp = tempfile.mkdtemp()
    synth.write_power_table(os.path.join(tmp, "pk.pow"))
    over = dict(NP=N**3, ICFormat='"%s"' % fmt, ZD_Pk_filename='"%s"' % os.path.join(tmp, "pk.pow"))
    if fnl:
        over.update(ZD_f_NL=fnl, ZD_n_s="0.96", Omega_M="0.3")
    if qplt:
        synth.write_eigmodes(os.path.join(tmp, "eig"), 16)
        over.update(ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ZD_PLT_filename='"%s"' % os.path.join(tmp, "eig"))
    P = pkg.Parameters(synth.write_param(os.path.join(tmp, "c.par"), **over))
    pw = pkg.PowerSpectrum(P)
    for r in range(G):
        cfg = P.config(device=0); cfg.rank, cfg.nranks = r, G
        ctx = pkg.Context(cfg); pw.apply(ctx)
        if qplt: ctx.load_eigenmodes_file(P.PLT_filename)
        ctx.generate()
        if G == 1:
            rec = ctx.fetch_planes(0, N); print(N, fmt, "ok", rec["displ"].std() if "displ" in rec.dtype.names else "")
        else:
            ctx.synchronize(); print("slab stage 1 ok", r)
        ctx.close()
PY
compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py 2>&1 | tail -6
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_case.py 2>&1 | tail -6
