#!/bin/bash
# round-2 GPU session S (4 GPUs, short): final build - N=4 weak-scaling line and the 160^3 line with the oracle goldens
mkdir -p gpurun_out
T=${1:-r2s}
for cfg in "n4_peer --steps 200 --warmup 20 --no-cpu" "n4_b160 --steps 100 --warmup 10 --no-cpu --block 160"; do
  set -- $cfg; name=$1; shift
  PCGB_COMM=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus 4 "$@" > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms/iter", round(d["ms_per_step"],4), "it/s", round(d["iterations_per_s"],1), "e2e frac", round(d["e2e"]["fraction_of_value"],3), "spmv ms", round(d["roofline"]["mean_launch_ms"],4),
          "transport", d["details"]["transport"], "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], (d["parity"]["preflight"] or {}).get("ok"))
except Exception as e:
    print("${name} failed:", e)
PY
done
