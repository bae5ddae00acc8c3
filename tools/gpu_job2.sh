#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_of1x1.py tests/test_gpu_api.py tests/test_gpu_layouts.py tests/test_gpu_properties.py tests/test_gpu_windows.py -m gpu -x -q > $O/j2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j2_pytest.log
tail -15 $O/j2_pytest.log
timeout 600 python tools/prof_of.py 32768 f64 2048 c2 > $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 32768 f64 8192 c2 >> $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 32768 f64 2048 c1 >> $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 16384 f64 4096 c2 >> $O/j2_prof_plain.log 2>&1
cat $O/j2_prof_plain.log
