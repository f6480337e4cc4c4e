#!/bin/bash
# A/B bench of kernel variants (no tests): tools/gpu_ab.sh <tag> <ncu-variant|none> variants...
tag=$1; ncuv=$2; shift; shift
for v in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/bench_${tag}_v$v.log 2>&1
  python - <<P
import json
l=[x for x in open("gpurun_out/bench_${tag}_v$v.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("variant $v: %.2f Gvox/s  sweep %.3f ms  frac %.3f" % (j["value"], j["kernel_ms"]["sweep"], j["roofline"]["frac"]))
else:
    print("variant $v: no result"); print(open("gpurun_out/bench_${tag}_v$v.log").read()[-1500:])
P
done
if [ "$ncuv" != "none" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 2 -c 1 -f -o gpurun_out/prof_${tag}_v$ncuv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --variant $ncuv > gpurun_out/ncu_$tag.log 2>&1
fi
