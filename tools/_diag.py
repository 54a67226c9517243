import json, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
from casmcode_clexmonte_b200.potential import semigrand_exchange_table
G = "tests/golden"
sysd = json.load(open(G + "/systems.json"))["fcc"]
tables = _capi.Tables(ClexulatorTables.load(G + "/tables/fcc_default.npz"))
ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.0, 0.0], 3)
N = tuple(int(x) for x in sys.argv[1:4]); nsw = int(sys.argv[4])
res = {}
for name, flags in {"generic": 2, "fused": 0, "nofusion": 8, "block": 4}.items():
    st = _capi.State(tables, N)
    eci = sysd["eci_sparse"]
    st.set_eci(eci["index"], eci["value"]); st.set_conditions(800.0, ex); st.randomize(2026)
    st.set_sweep_flags(flags)
    cnt = st.sgc_sweep(nsw, seed=7)
    res[name] = (st.download_occ(0, dtype=np.int8).reshape(N[2], N[1], N[0]), cnt[0].n_accept)
    st.close()
g = res["generic"][0]
for name in ("fused", "nofusion", "block"):
    d = res[name][0] != g
    print(name, "diff sites", int(d.sum()), "n_accept", res[name][1], "generic", res["generic"][1])
    if d.any():
        kk, jj, ii = np.nonzero(d)
        print("  k hist (parity):", np.bincount(kk % 2), " j parity:", np.bincount(jj % 2), "k range", kk.min(), kk.max(), "uniq k", len(np.unique(kk)))
        print("  first:", list(zip(kk[:8], jj[:8], ii[:8])))
