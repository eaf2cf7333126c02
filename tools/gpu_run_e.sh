#!/bin/bash
# round-2 GPU session E (1 GPU): node-block kernel v2 (uniform fast path, 8 lanes per row in phase 2) - tests, sweep, bench
mkdir -p gpurun_out
T=${1:-r2e}
timeout 900 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_concrete.py tests/test_gpu_pcg.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${T}_pytest.log
timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; echo "sweep rc=$?"
cat gpurun_out/${T}_sweep.txt; tail -3 gpurun_out/${T}_sweep.err
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n1.json").read().strip().splitlines()[-1])
print("ms/iter", d["ms_per_step"], "spmv", d["roofline"]["mean_launch_ms"], "frac", d["roofline"]["frac"], "streamed", d["roofline"]["streamed_GBps"], d["details"]["plan"], d["parity"]["max_rel_err"], d["roofline"]["phase_ms_per_iteration"])
PY
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --workload concrete > gpurun_out/${T}_bench_concrete_n1.json 2> gpurun_out/${T}_bench_concrete_n1.err; echo "bench concrete rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_concrete_n1.json").read().strip().splitlines()[-1])
print("concrete ms/iter", d["ms_per_step"], "spmv", d["roofline"]["mean_launch_ms"], "frac", d["roofline"]["frac"], d["details"]["plan"], d["full_solve"])
PY
