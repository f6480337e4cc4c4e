#!/bin/bash
# usage: tools/gpu_scale.sh N  -- default bench and the cfg5 scene block on N GPUs of one box
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg5 > gpurun_out/cfg5_n$N.json 2> gpurun_out/cfg5_n$N.err
python - <<P
import json
for f in ("gpurun_out/scale_n$N.json", "gpurun_out/cfg5_n$N.json"):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.2f ms/step %.3f e2e %.2f (%.3f ms/step, h2d %d) views/s %.1f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["e2e"]["h2d_bytes_per_step"], j["ref_views_per_s"]))
    except Exception as e:
        print(f, "no result", e)
P
tail -n 3 gpurun_out/scale_n$N.err; tail -n 3 gpurun_out/cfg5_n$N.err
