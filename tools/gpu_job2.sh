#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_psd.py -m gpu -x -q > $O/j2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j2_pytest.log
tail -3 $O/j2_pytest.log
for a in "32768 f64 4096 2" "32768 f32 4096 2" "65536 f32 2048 2" "16384 f32 4096 2"; do timeout 300 python tools/prof_csd.py $a; done > $O/j2_csd.log 2>&1; cat $O/j2_csd.log
timeout 300 python tools/psd_time.py > $O/j2_psd_plain.log 2>&1; cat $O/j2_psd_plain.log
