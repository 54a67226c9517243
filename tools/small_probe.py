"""Timing probe: sweeps on small single boxes, one launch per pass vs the cooperative whole-sweep kernel."""
import sys, json, os
sys.path.insert(0, '.')
import torch
from pathlib import Path
from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
from casmcode_clexmonte_b200.potential import semigrand_exchange_table
G = Path('tests/golden')
sysd = json.loads((G / 'systems.json').read_text())['fcc']
tables = _capi.Tables(ClexulatorTables.load(G / 'tables' / 'fcc_default.npz'))
ex = semigrand_exchange_table(sysd['occ_to_species'], sysd['axes']['Rt'], (0.0, 0.0), 3)
for N in (32, 64, 128, 256):
    st = _capi.State(tables, (N, N, N), 1)
    st.set_eci(sysd['eci_sparse']['index'], sysd['eci_sparse']['value'])
    st.set_conditions(800.0, ex)
    st.randomize(1)
    stream = torch.cuda.ExternalStream(st.stream())
    st.sgc_sweep(20, seed=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream); st.sgc_sweep(200, seed=1, first_sweep=20, counters=False); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'N={N}: {ms/200*1e3:.1f} us per sweep, {200*N**3/ms/1e-3:.3e} steps/s', flush=True)
    st.close()
