#!/bin/bash
# round-2 GPU session K (1 GPU): in-place partials / 3-stage ring of the node-block kernel - tests, sweep, bench
mkdir -p gpurun_out
T=${1:-r2k}
timeout 900 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_pcg.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${T}_pytest.log
timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; echo "sweep rc=$?"
cat gpurun_out/${T}_sweep.txt; tail -3 gpurun_out/${T}_sweep.err
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n1.json").read().strip().splitlines()[-1])
print("ms/iter", d["ms_per_step"], "spmv", d["roofline"]["mean_launch_ms"], "frac", d["roofline"]["frac"], "streamed", d["roofline"]["streamed_GBps"], d["details"]["plan"], d["parity"]["max_rel_err"])
PY
