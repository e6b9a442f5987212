#!/bin/bash
# usage: tools/variant_run360.sh <label> <lib.so or "default">   -- N = 360 (cfg4 shape) timing of a kernel variant
label=$1; lib=$2
if [ "$lib" != "default" ]; then export BIOEM_B200_LIB=$PWD/$lib; fi
timeout 600 python tools/front_time_cfg4.py 8 592 > gpurun_out/var360_${label}.log 2>&1 || echo "FAILED rc=$?" >> gpurun_out/var360_${label}.log
tail -1 gpurun_out/var360_${label}.log
