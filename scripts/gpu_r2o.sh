#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "topk or metrics" > gpurun_out/topk_test.log 2>&1
echo "topk tests exit $?"; tail -n 2 gpurun_out/topk_test.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_eval.csv python scripts/profile_step.py --steps 3 --warmup 0 --eval > gpurun_out/ncu4.log 2>&1
grep -i "topk" gpurun_out/launches_eval.csv | tail -2 | awk -F'","' '{print $5, $NF}'
timeout 600 python bench.py --mode eval --steps 8 --no-cpu-baseline > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
python -c "import json; d=json.load(open('gpurun_out/bench_eval.json')); print('eval', d['value'], d['ms_per_step'], d['e2e']['value'])"
