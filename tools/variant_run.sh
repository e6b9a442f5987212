#!/bin/bash
# usage: tools/variant_run.sh <label> <lib.so or "default"> [ncu]   -- one timing run (and optionally an ncu capture) of a kernel variant
label=$1; lib=$2; mode=$3
if [ "$lib" != "default" ]; then export BIOEM_B200_LIB=$PWD/$lib; fi
timeout 300 python tools/prof_lik.py cfg2 150 > gpurun_out/var_${label}.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/var_${label}.log
if [ "$mode" == "test" ] || [ "$mode" == "ncu+test" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cfg2_slice or headline or cfg5_slice" > gpurun_out/var_${label}_tests.log 2>&1
fi
if [ "$mode" == "ncu" ] || [ "$mode" == "ncu+test" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:likelihood_kernel -s 1 -c 1 -o gpurun_out/var_${label} -f python tools/prof_lik.py cfg2 150 > gpurun_out/var_${label}_ncu.log 2>&1
fi
tail -1 gpurun_out/var_${label}.log
