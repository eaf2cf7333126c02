#!/bin/bash
# round-2 GPU session H (1 GPU): full -m gpu suite, concrete SpMV sweep, final bench lines, ncu capture + SASS evidence, EBE timing
mkdir -p gpurun_out
T=${1:-r2h}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${T}_pytest.log
if [ "$2" = "concrete" ]; then
SWEEP_MODEL=concrete timeout 400 python tools/spmv_sweep.py > gpurun_out/${T}_sweep_concrete.txt 2> gpurun_out/${T}_sweep_concrete.err; echo "sweep concrete rc=$?"
cat gpurun_out/${T}_sweep_concrete.txt; tail -3 gpurun_out/${T}_sweep_concrete.err
fi
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --operator ebe > gpurun_out/${T}_bench_n1_ebe.json 2> gpurun_out/${T}_bench_n1_ebe.err; echo "bench ebe rc=$?"
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --workload concrete > gpurun_out/${T}_bench_n1_concrete.json 2> gpurun_out/${T}_bench_n1_concrete.err; echo "bench concrete rc=$?"
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_n1_k20.json 2> gpurun_out/${T}_bench_n1_k20.err; echo "bench k20 rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --block 256 > gpurun_out/${T}_bench_n1_b256.json 2> gpurun_out/${T}_bench_n1_b256.err; echo "bench b256 rc=$?"
python - <<PY
import json
for f in ("bench_n1", "bench_n1_k20", "bench_n1_b256", "bench_n1_ebe", "bench_n1_concrete"):
    try:
        d=json.loads(open("gpurun_out/${T}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/iter", round(d["ms_per_step"],4), "it/s", round(d["iterations_per_s"],1), "e2e frac", round(d["e2e"]["fraction_of_value"],3), "spmv", round(d["roofline"]["mean_launch_ms"],4), "frac", round(d["roofline"]["frac"],3),
              "streamed frac", round(d["roofline"]["streamed_frac_of_peak"],3), "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d["clocks"], "full", d.get("full_solve"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 python tools/ebe_quick.py > gpurun_out/${T}_ebe_quick.json 2> gpurun_out/${T}_ebe_quick.err; echo "ebe rc=$?"; cat gpurun_out/${T}_ebe_quick.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_bsr3 -s 8 -c 1 -o gpurun_out/${T}_spmv_bsr3 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:k_ebe_t24 -s 3 -c 1 -o gpurun_out/${T}_ebe_t24 python tools/ebe_quick.py > gpurun_out/${T}_ncu_ebe.log 2>&1; echo "ncu ebe rc=$?"
