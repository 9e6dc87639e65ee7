cat > dbg_modes.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from util import load_golden
from lehamoc_b200.engine import Engine
from lehamoc_b200.setup_host import pack_script
_, P = load_golden("test1")
P.update(grid_g_el=90., grid_nu=60., time_end=3.)
sets = [dict(P, B0=1.0 + 0.1 * i, p_el=1.8 + 0.02 * i) for i in range(5)]
b = pack_script(sets)
outs = []
for name, cfg in (("shared", b.cfg), ("perwalker", dict(b.cfg, shared_ic_grids=0)), ("notable", dict(b.cfg, shared_ic_grids=0, ic_pair_table=0))):
    eng = Engine(cfg, b.W)
    N, sed = eng.run_host(b, snapshots=3)
    outs.append((name, N.copy(), sed.copy()))
    eng.close()
ref = outs[0]
for name, N, sed in outs[1:]:
    bad = np.argwhere(sed != ref[2])
    print(name, "sed mismatches:", len(bad), bad[:8].tolist(), "N mismatches:", int((N != ref[1]).sum()))
PY
echo "== default"; python dbg_modes.py
echo "== LHM_CONST_B=0"; LHM_CONST_B=0 python dbg_modes.py
