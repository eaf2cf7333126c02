#!/bin/bash
# round-2 GPU session T (1 GPU, final sanity): full -m gpu suite, smoke, the driver's bench command (both arms), one-CTA sweep
mkdir -p gpurun_out
T=${1:-r2t}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1_k20.json 2> gpurun_out/${T}_bench_n1_k20.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref_n1.json 2> gpurun_out/${T}_bench_ref_n1.err; echo "bench ref rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n1_k20.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/${T}_bench_ref_n1.json").read().strip().splitlines()[-1])
print("b200", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("reference", r["value"], r["cpu_baseline"]["cores"], "same config", d["config"] == r["config"], "ratio", d["value"]/r["value"], "e2e ratio", d["e2e"]["value"]/r["value"])
PY
timeout 300 python tools/spmv_sweep.py > gpurun_out/${T}_sweep.txt 2> gpurun_out/${T}_sweep.err; cat gpurun_out/${T}_sweep.txt; tail -2 gpurun_out/${T}_sweep.err
