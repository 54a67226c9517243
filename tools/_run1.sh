python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/s7_tests.log 2>&1; tail -15 gpurun_out/s7_tests.log
for v in "CMX_SWEEP_SLICE_MB=24" "CMX_SWEEP_SLICE_MB=0" "CMX_SWEEP_SLICE_MB=0 CMX_SWEEP_PDL=0" "CMX_SWEEP_SLICE_MB=48" "CMX_SWEEP_SLICE_MB=12" "CMX_SWEEP_SLICE_MB=24 CMX_SWEEP_BLOCKS_PER_SM=8"; do
  echo "== $v"; env $v python bench.py --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','accept_rate')}, d['roofline']['kernel_ms'], d['clocks'])
    else: print(l.rstrip()[:300])
"
done
