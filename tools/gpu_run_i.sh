#!/bin/bash
# round-2 GPU session I (2 GPUs): multi-GPU parity with the final kernels, N=2 lines (fork on / off, NCCL)
mkdir -p gpurun_out
T=${1:-r2i}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
run() {
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus 2 "$@" > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms/iter", round(d["ms_per_step"],4), "it/s", round(d["iterations_per_s"],1), "e2e frac", round(d["e2e"]["fraction_of_value"],3), "spmv ms", round(d["roofline"]["mean_launch_ms"],4),
          "transport", d["details"]["transport"], "launches", d["gpu_launches"], "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], (d["parity"]["preflight"] or {}).get("ok"))
    print("   phases", {k: round(v, 4) for k, v in (d["roofline"].get("phase_ms_per_iteration") or {}).items()})
except Exception as e:
    print("${name} failed:", e)
PY
}
run peer PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu
run peer_nofork PCGB_COMM=peer PCGB_FORK=0 -- --steps 200 --warmup 20 --no-cpu
run nccl PCGB_COMM=nccl -- --steps 200 --warmup 20 --no-cpu
run concrete PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu --workload concrete
