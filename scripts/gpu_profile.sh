#!/bin/bash
# ncu passes (B200_PROFILING.md): (1) every launch of two training steps with its device time,
# (2) full-set captures of the tcgen05 kernels.  Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SKIP=${SKIP:-140}
COUNT=${COUNT:-80}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s "$SKIP" -c "$COUNT" --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
if [ "$1" == "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 8 -c 4 \
      -o gpurun_out/prof_tc -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
  echo "full capture exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_adam|k_spmm" -s 6 -c 4 \
      -o gpurun_out/prof_mem -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
  echo "mem capture exit $?"
fi
python - <<'EOF'
import csv, collections
rows = []
with open("gpurun_out/launches.csv") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    try:
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
    except Exception:
        pass
agg = collections.OrderedDict()
for n, v, u in rows:
    v_us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    k = n.split("(")[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v_us
tot = sum(a[1] for a in agg.values())
print("total %.1f us over %d launches" % (tot, len(rows)))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%%  x%-3d %s" % (t, 100 * t / tot, c, k))
EOF
