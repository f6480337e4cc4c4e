#!/bin/bash
# on the GPU box: pipeline tests + default bench + cfg5
timeout 300 python -m pytest tests/test_sweep_gpu.py -m gpu -q --timeout 100 -x -k "view_pipeline" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e1.json 2> gpurun_out/bench_e1.err; tail -c 1500 gpurun_out/bench_e1.json | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f e2e %s' % (j['value'], json.dumps(j['e2e'])[:600]))"
timeout 300 python bench.py --workload cfg5 > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; python -c "
import json
j=json.loads(open('gpurun_out/bench_cfg5_n1.json').read().strip().splitlines()[-1]); print('cfg5 value %.2f views/s %.1f ms/view %.3f h2d %d' % (j['value'], j['ref_views_per_s'], j['ms_per_step'], j['e2e']['h2d_bytes_per_step']))" || tail -5 gpurun_out/bench_cfg5_n1.err
