"""Time of the global correlation sampler (`corr`, a10) at BASELINE configs[1] size:
python tools/time_corr.py [N]   (prints one JSON line; GPU box)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from casmcode_clexmonte_b200 import _capi  # noqa: E402
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
systems = json.loads((ROOT / "tests/golden/systems.json").read_text())
sysd = systems["fcc"]
tab = _capi.Tables(ClexulatorTables.load(ROOT / "tests/golden/tables" / f"{sysd['tables']}.npz"))
st = _capi.State(tab, N)
eci = sysd["eci_sparse"]
st.set_eci(eci["index"], eci["value"])
st.randomize(3)
out = {}
for name, flags in (("streaming", 0), ("faithful", _capi.CMX_SWEEP_FORCE_GENERIC)):
    st.set_sweep_flags(flags)
    ref = st.global_corr()
    reps = 20 if name == "streaming" else 2
    t0 = time.perf_counter()
    for _ in range(reps):
        c = st.global_corr()
    dt = (time.perf_counter() - t0) / reps
    out[name] = {"us_per_call_host_clock": dt * 1e6, "corr_head": [float(x) for x in c[:4]]}
a = np.array(out["streaming"]["corr_head"]); b = np.array(out["faithful"]["corr_head"])
out["max_rel_diff_head"] = float(np.max(np.abs(a - b) / np.maximum(1e-300, np.abs(b))))
out["n_sites"] = st.n_sites
print(json.dumps(out))
