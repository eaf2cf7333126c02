#!/bin/bash
# round-2 GPU session J (8 GPUs): final weak-scaling lines (node-block kernel, forked unpack), 160^3, C4, C3
mkdir -p gpurun_out
T=${1:-r2j4}
run() {
  name=$1; np=$2; shift; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29500 \
      bench.py --gpus $np "$@" > gpurun_out/${T}_${name}.json 2> gpurun_out/${T}_${name}.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms/iter", round(d["ms_per_step"],4), "it/s", round(d["iterations_per_s"],1), "e2e frac", round(d["e2e"]["fraction_of_value"],3), "spmv ms", round(d["roofline"]["mean_launch_ms"],4),
          "transport", d["details"]["transport"], "launches", d["gpu_launches"], "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], (d["parity"]["preflight"] or {}).get("ok"))
    print("   phases", {k: round(v, 4) for k, v in (d["roofline"].get("phase_ms_per_iteration") or {}).items()}, "full_solve", d.get("full_solve"))
except Exception as e:
    print("${name} failed:", e)
PY
}
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -p no:cacheprovider -k '4-1' > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
run n4_peer 4 PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu
run n4_peer_k20 4 PCGB_COMM=peer -- --steps 20 --warmup 5 --no-cpu
run n4_nccl 4 PCGB_COMM=nccl -- --steps 200 --warmup 20 --no-cpu
run n4_b160 4 PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu --block 160
run n4_concrete 4 PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu --workload concrete
run n4_hexmetis 4 PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu --workload hex_metis --block 64
run n4_b256 4 PCGB_COMM=peer -- --steps 100 --warmup 10 --no-cpu --block 256
