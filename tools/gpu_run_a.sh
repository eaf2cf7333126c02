#!/bin/bash
# round-2 GPU session A (1 GPU): smoke, full -m gpu suite, SpMV sweep, default bench line, gated coloured-EBE tests
mkdir -p gpurun_out
T=r2a
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${T}_pytest.log
timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; echo "sweep rc=$?"
cat gpurun_out/${T}_sweep.txt
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${T}_bench_n1.json
PCGB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_ebe_colored_experimental.py -q --timeout 200 -p no:cacheprovider > gpurun_out/${T}_ebe_colored.log 2>&1; echo "colored rc=$?"
tail -5 gpurun_out/${T}_ebe_colored.log
