#!/bin/bash
# One GPU measurement cycle: parity tests, bench per kernel variant, one ncu capture of the sweep kernel.
# usage: tools/gpu_cycle.sh <tag> [variants...]
tag=$1; shift; variants=${@:-0}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/tests_$tag.log
for v in $variants; do
  timeout 300 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_${tag}_v$v.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 2 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -4 gpurun_out/tests_$tag.log
for v in $variants; do python - <<P
import json
l=[x for x in open("gpurun_out/bench_${tag}_v$v.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("variant $v: %.2f Gvox/s  sweep %.3f ms  regress %.3f ms  frac %.3f  e2e %.2f ms  clocks %s" % (j["value"], j["kernel_ms"]["sweep"], j["kernel_ms"]["regress"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["clocks"]))
else:
    print("variant $v: no result"); print(open("gpurun_out/bench_${tag}_v$v.log").read()[-1500:])
P
done
