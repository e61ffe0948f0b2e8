#!/bin/bash
# chunked dW_d GEMM + Adam on the side stream (gradient kept in L2): correctness under the new schedule, then a sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/chunk_sweep.txt
B200VAE_WD_CHUNKS=5 timeout 600 python -m pytest tests/test_gpu_overlap.py tests/test_gpu_parity.py -q -p no:cacheprovider -x > gpurun_out/chunk_tests.log 2>&1
echo "tests under WD_CHUNKS=5: exit $?" >> gpurun_out/chunk_sweep.txt
tail -n 3 gpurun_out/chunk_tests.log >> gpurun_out/chunk_sweep.txt
run() {  # label env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - "$label" <<'PY' >> gpurun_out/chunk_sweep.txt
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
    print("%-24s %8.1f us/step  %9.0f users/s   e2e %9.0f users/s  launches/step %.1f" % (sys.argv[1], 1e3 * d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print(sys.argv[1], repr(e))
PY
}
run base B200VAE_WD_CHUNKS=0 B200VAE_HOST_OVERLAP=1
run c5 B200VAE_WD_CHUNKS=5 B200VAE_HOST_OVERLAP=1
run c5_nodiscard B200VAE_WD_CHUNKS=5 B200VAE_WD_DISCARD=0 B200VAE_HOST_OVERLAP=1
run c10 B200VAE_WD_CHUNKS=10 B200VAE_HOST_OVERLAP=1
run c3 B200VAE_WD_CHUNKS=3 B200VAE_HOST_OVERLAP=1
run c5_ctas4 B200VAE_WD_CHUNKS=5 B200VAE_WD_CHUNK_CTAS=4 B200VAE_HOST_OVERLAP=1
run c5_ctas2 B200VAE_WD_CHUNKS=5 B200VAE_WD_CHUNK_CTAS=2 B200VAE_HOST_OVERLAP=1
cat gpurun_out/chunk_sweep.txt
