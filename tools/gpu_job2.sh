#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/j2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/j2_pytest.log
tail -6 $O/j2_pytest.log
timeout 600 python tools/prof_of.py 65536 f32 2048 c2 > $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 65536 f64 2048 c2 >> $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 32768 f64 8192 c2 >> $O/j2_prof_plain.log 2>&1
timeout 600 python tools/prof_of.py 32768 f32 8192 c2 >> $O/j2_prof_plain.log 2>&1
cat $O/j2_prof_plain.log
timeout 300 python tools/psd_time.py > $O/j2_psd_plain.log 2>&1; cat $O/j2_psd_plain.log
timeout 300 python tools/trig_time.py > $O/j2_trig_plain.log 2>&1; cat $O/j2_trig_plain.log
