"""Timing probe: one colour pass on boxes of decreasing depth (is a half / eighth slab launch efficient?)"""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np, torch
from pathlib import Path
from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
from casmcode_clexmonte_b200.potential import semigrand_exchange_table
G = Path('tests/golden')
sysd = json.loads((G / 'systems.json').read_text())['fcc']
tables = _capi.Tables(ClexulatorTables.load(G / 'tables' / 'fcc_default.npz'))
ex = semigrand_exchange_table(sysd['occ_to_species'], sysd['axes']['Rt'], (0.0, 0.0), 3)
for N2 in (512, 256, 128, 64):
    for halo in (0, 1):
        st = _capi.State(tables, (512, 512, N2), 1, halo)
        st.set_eci(sysd['eci_sparse']['index'], sysd['eci_sparse']['value'])
        st.set_conditions(800.0, ex)
        st.randomize(1)
        stream = torch.cuda.ExternalStream(st.stream())
        def run(n):
            for w in range(n):
                for g in range(2):
                    st.sgc_sweep_kgroup(1, w, g)
        run(10); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); run(50); e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f'N2={N2} halo={halo}: {ms/200*1e3:.1f} us per launch, {50*512*512*N2/ms/1e-3:.3e} steps/s', flush=True)
        st.close()
