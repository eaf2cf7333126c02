#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2q}
timeout 300 python -m pytest tests/test_gpu_spmv.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
SWEEP_MODEL=concrete timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep_concrete.txt 2> gpurun_out/${T}_sweep_concrete.err; echo "sweep concrete rc=$?"
cat gpurun_out/${T}_sweep_concrete.txt; tail -3 gpurun_out/${T}_sweep_concrete.err
timeout 300 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; cat gpurun_out/${T}_sweep.txt
