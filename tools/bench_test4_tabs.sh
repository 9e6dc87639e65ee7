#!/bin/bash
# Test 4 at full size with and without the tabulated photo-hadronic operators (tools/config_benches.py config3)
python - <<'PY'
import json, os, sys
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import config_benches as cb
r = cb.config3(148, host_fit=False)
print(json.dumps({k: r[k] for k in ("ms_per_batch", "ms_per_model", "parity", "hadtab")}, indent=1))
os.environ["LHM_HAD_TABLES"] = "0"
r = cb.config3(148, host_fit=False)
print("per-step evaluation:", json.dumps({k: r[k] for k in ("ms_per_batch", "ms_per_model")}))
PY
