#!/bin/bash
# round-end evidence on one B200: tests, smoke, bench (both arms), launch list, workloads, ncu summaries
out=gpurun_out/$1; mkdir -p $out
python -m pytest tests -m gpu -q > $out/tests.log 2>&1; tail -3 $out/tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench.json 2> $out/bench.err; tail -c 400 $out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 300 $out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
python tools/bench_workloads.py > $out/workloads.jsonl 2> $out/workloads.err; wc -l $out/workloads.jsonl
ncu --set full --clock-control none -k regex:k_sweep_generic_warp -s 2 -c 1 -o $out/generic_warp python tools/bench_workloads.py c4 > $out/ncu_c4.log 2>&1
ncu --set full --clock-control none -k regex:k_canonical_pairs -s 2 -c 1 -o $out/canon python tools/bench_workloads.py c1 > $out/ncu_c1.log 2>&1
ncu --set full --clock-control none -k regex:k_kmc_run -s 1 -c 1 -o $out/kmc python tools/bench_workloads.py c5 > $out/ncu_c5.log 2>&1
ncu --set full --clock-control none -k regex:k_sweep_generic -s 2 -c 1 -o $out/generic_dense python tools/bench_workloads.py c3full > $out/ncu_c3full.log 2>&1
ls -la $out
