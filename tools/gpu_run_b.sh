#!/bin/bash
# round-2 GPU session B (2 GPUs): multi-GPU parity tests (peer + NCCL transport), device assembly tests, bench N=2 variants
mkdir -p gpurun_out
T=r2b
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_assemble.py tests/test_gpu_spmv.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${T}_pytest.log
run() {  # name, env..., -- bench args
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
          "transport", d["details"]["transport"], "iface tiles", d["details"]["plan"]["interface_tiles"], "launches", d["gpu_launches"], "parity", d["parity"]["max_rel_err"], d["parity"]["ok"], (d["parity"]["preflight"] or {}).get("ok"))
except Exception as e:
    print("${name} failed:", e)
PY
}
run peer PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu
run peer_noverlap PCGB_COMM=peer PCGB_OVERLAP=0 -- --steps 200 --warmup 20 --no-cpu
run nccl PCGB_COMM=nccl -- --steps 200 --warmup 20 --no-cpu
run peer_k20 PCGB_COMM=peer -- --steps 20 --warmup 5 --no-cpu
run concrete PCGB_COMM=peer -- --steps 200 --warmup 20 --no-cpu --workload concrete
run concrete_nccl PCGB_COMM=nccl -- --steps 200 --warmup 20 --no-cpu --workload concrete
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --workload concrete > gpurun_out/${T}_concrete_n1.json 2> gpurun_out/${T}_concrete_n1.err; echo "concrete n1 rc=$?"
tail -c 1500 gpurun_out/${T}_concrete_n1.json
