# 2-GPU slab CLI vs the single-GPU CLI: the ic_* files must be byte-identical (CPD < PPD: shared files across ranks)
set -e
T=$(mktemp -d); cd $T
python - <<PY
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from __graft_entry__ import load_synth
s = load_synth()
s.write_power_table("pk.pow"); s.write_eigmodes("eig", 32)
for name, out in (("one.par", "out1"), ("two.par", "out2")):
    s.write_param(name, NP=64**3, CPD=5, ZD_qPLT=1, ZD_qPLT_rescale=1, ZD_PLT_target_z="5.0", ICFormat='"RVZel"',
                  ZD_Pk_filename='"pk.pow"', ZD_PLT_filename='"eig"', InitialConditionsDirectory='"%s"' % out)
PY
R=${GRAFT_REPO_ROOT:-/root/repo}
$R/zeldovich-plt_b200/bin/zeldovich one.par 2>&1 | grep -E "rms|took"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 $R/zeldovich-plt_b200/run_distributed.py two.par 2>&1 | grep -E "rms|took"
ls out1 | sort > a.txt; ls out2 | sort > b.txt; diff a.txt b.txt && for f in $(cat a.txt); do cmp out1/$f out2/$f; done && echo "MGPU CLI: ic files identical ($(wc -l < a.txt) files)"
