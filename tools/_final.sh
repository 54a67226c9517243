#!/bin/bash
# round-end evidence on one B200: tests, smoke, bench (both arms), launch list, ncu capture of the
# headline kernel at the bench workload, secondary workloads
out=gpurun_out/$1; mkdir -p $out
python -m pytest tests -m gpu -q > $out/tests.log 2>&1; tail -3 $out/tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 200 $out/bench_ref.json
python bench.py --eci full > $out/bench_eci_full.json 2> $out/bench_eci_full.err; head -c 200 $out/bench_eci_full.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_pass16 -s 3 -c 1 -o $out/pass16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_pass16.log 2>&1
python tools/bench_workloads.py > $out/workloads.jsonl 2> $out/workloads.err; wc -l $out/workloads.jsonl
ls $out
