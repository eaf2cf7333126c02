#!/bin/bash
# bring-up bisect: every stage under its own timeout so a hung kernel cannot eat the gpurun budget
for st in spmv_tma0 spmv solve_direct solve_graph solve_default_stream_graph; do
  if [ "$st" = "spmv_tma0" ]; then export PCGB_SPMV_TMA=0; else unset PCGB_SPMV_TMA; fi
  timeout 60 python tools/gpu_bisect.py $st 2>&1 | tail -6
  echo "== stage $st rc=${PIPESTATUS[0]}"
done
