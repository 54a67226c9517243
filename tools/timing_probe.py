import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from pathlib import Path
from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
from casmcode_clexmonte_b200.potential import semigrand_exchange_table
G = Path('tests/golden')
sysd = json.loads((G / 'systems.json').read_text())['fcc']
tables = _capi.Tables(ClexulatorTables.load(G / 'tables' / 'fcc_default.npz'))
ex = semigrand_exchange_table(sysd['occ_to_species'], sysd['axes']['Rt'], (0.0, 0.0), 3)
for N in (128, 256, 512):
    st = _capi.State(tables, (N, N, N))
    st.set_eci(sysd['eci_sparse']['index'], sysd['eci_sparse']['value'])
    st.set_conditions(800.0, ex)
    st.randomize(1)
    st.energy(); st.composition(); st.global_corr()
    for name, f in (('energy', st.energy), ('composition', st.composition), ('global_corr', st.global_corr)):
        t0 = time.perf_counter()
        for _ in range(3): f()
        dt = (time.perf_counter() - t0) / 3
        print(f'N={N} {name}: {dt*1e3:.3f} ms  ({N**3/dt/1e9:.2f} Gsites/s)')
    t0 = time.perf_counter(); st.sgc_sweep(10, seed=1); dt = time.perf_counter() - t0
    print(f'N={N} 10 sweeps: {dt*1e3:.3f} ms ({10*N**3/dt:.3e} steps/s)')
    st.close()
# replica grid: 64 x 128^3 (config 2)
st = _capi.State(tables, (128, 128, 128), 64)
st.set_eci(sysd['eci_sparse']['index'], sysd['eci_sparse']['value'])
r = 0
for mu in np.linspace(-1, 1, 8):
    for T in np.arange(400.0, 1801.0, 200.0):
        st.set_conditions(T, semigrand_exchange_table(sysd['occ_to_species'], sysd['axes']['Rt'], (mu, 0.0), 3), r)
        r += 1
st.randomize(7)
st.sgc_sweep(10, seed=1)
t0 = time.perf_counter(); cnt = st.sgc_sweep(50, seed=1, first_sweep=10); dt = time.perf_counter() - t0
print(f'config2: 64 x 128^3, 50 sweeps: {dt*1e3:.2f} ms  {50*64*128**3/dt:.3e} steps/s; accept range',
      min(c.n_accept / c.n_attempt for c in cnt), max(c.n_accept / c.n_attempt for c in cnt))
st.close()
