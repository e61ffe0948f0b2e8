#!/bin/bash
# Run the GPU test groups in separate processes (a trapped kernel poisons its CUDA context),
# each under its own timeout; logs go to gpurun_out/ which gpurun brings back.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python - <<'EOF' > gpurun_out/env.txt 2>&1
import torch
print(torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
EOF
run() {  # name timeout pytest-args...
  local name=$1; local to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "$name exit $?" >> gpurun_out/summary.txt
  tail -n 3 "gpurun_out/$name.log" >> gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
if [ "$1" == "tc" ] || [ -z "$1" ]; then
  run tc_gemm 300 tests/test_gpu_kernels.py -k "tc_gemm or truncates" -s
  run tc_lse 200 tests/test_gpu_kernels.py -k "lse"
fi
if [ "$1" == "rest" ] || [ -z "$1" ]; then
  run kernels 400 tests/test_gpu_kernels.py -k "not tc_gemm and not truncates and not lse"
  run api 600 tests/test_gpu_api.py tests/test_gpu_multi.py
  run parity_fixture 600 tests/test_gpu_parity.py -k "fixture"
  run parity_oracle 900 tests/test_gpu_parity.py -k "not fixture"
  run overlap 300 tests/test_gpu_overlap.py
  run cond 300 tests/test_gpu_cond.py
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/summary.txt
  tail -n 2 gpurun_out/smoke.log >> gpurun_out/summary.txt
fi
if [ "$1" == "bench" ] || [ -z "$1" ]; then
  timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench.json >> gpurun_out/summary.txt
  tail -n 5 gpurun_out/bench.err >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
