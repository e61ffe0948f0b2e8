#!/bin/bash
# round-2 call F: EASE tests, ncu launch list + full-set captures of the step's kernels (shipped build), eval kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ease.py -q -p no:cacheprovider > gpurun_out/ease_test.log 2>&1
echo "ease exit $?" > gpurun_out/summary_f.txt
tail -n 12 gpurun_out/ease_test.log >> gpurun_out/summary_f.txt
# every launch of the last two of 6 steps (23 launches per step; ~12 set-up launches before the first step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 104 -c 46 --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary_f.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_gemm|k_adam|k_spmm|k_target_fixup|k_splitk" \
    -s 64 -c 17 -o gpurun_out/prof_step -f python scripts/profile_step.py --steps 2 --warmup 4 > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?" >> gpurun_out/summary_f.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_topk|k_tc_gemm|k_mask" \
    -s 6 -c 6 -o gpurun_out/prof_eval -f python scripts/profile_step.py --steps 4 --warmup 0 --eval > gpurun_out/ncu3.log 2>&1
echo "eval capture exit $?" >> gpurun_out/summary_f.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_eval.csv python scripts/profile_step.py --steps 4 --warmup 0 --eval > gpurun_out/ncu4.log 2>&1
timeout 600 python bench.py --config cfg3 --steps 50 --warmup 10 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 900 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 900 python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
cat gpurun_out/summary_f.txt
ls -la gpurun_out/*.ncu-rep
tail -n 3 gpurun_out/bench_cfg5.err gpurun_out/bench_cfg4.err
