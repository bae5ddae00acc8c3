#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/j3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j3_pytest.log
tail -4 $O/j3_pytest.log
timeout 900 python bench.py > $O/j3_bench.json 2> $O/j3_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/j3_ncu_launches.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/j3_bench_ref.json 2> $O/j3_bench_ref.err; echo "ref rc=$?"; tail -c 400 $O/j3_bench_ref.json
