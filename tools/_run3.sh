timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s9_tests.log 2>&1; tail -5 gpurun_out/s9_tests.log
for v in "X=1" "CMX_SWEEP_SLICE_LAYERS=16" "CMX_SWEEP_SLICE_LAYERS=48" "CMX_SWEEP_SLICE_LAYERS=96" "CMX_SWEEP_SLICE_LAYERS=256" "CMX_SWEEP_BLOCKS_PER_SM=3"; do
  echo "== $v"; env $v timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','accept_rate')}, d['roofline']['kernel_ms'], d['clocks'])
    else: print(l.rstrip()[:300])
"
done
