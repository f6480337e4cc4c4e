#!/bin/bash
# Round-2 evidence: launch list of the default bench command, ncu --set full of the cfg2 sweep kernel, bench lines.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_quad -s 2 -c 1 -f -o gpurun_out/prof_r2_cfg2 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_r2_cfg2.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 > gpurun_out/bench_r2_cfg3.json 2> gpurun_out/bench_r2_cfg3.err
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 > gpurun_out/bench_r2_cfg4.json 2> gpurun_out/bench_r2_cfg4.err
timeout 300 python bench.py --workload fuse --steps 10 --warmup 3 > gpurun_out/bench_r2_fuse.json 2> gpurun_out/bench_r2_fuse.err
for v in 8 9; do timeout 300 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_r2_variant$v.json 2>/dev/null; done
python tests/aten_gpu_baseline.py > gpurun_out/r2_aten_gpu_baseline.txt 2>&1
tail -n 2 gpurun_out/*.err | tail -20
