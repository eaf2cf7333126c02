#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2o}
timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; echo "sweep rc=$?"
cat gpurun_out/${T}_sweep.txt; tail -3 gpurun_out/${T}_sweep.err
