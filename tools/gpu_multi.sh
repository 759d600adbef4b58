#!/bin/bash
# usage: tools/gpu_multi.sh N  -- multi-GPU tests + the bench line at N GPUs (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/multi_n$N.gpus
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs 2>&1 | tail -12 > gpurun_out/multi_n$N.pytest.log; cat gpurun_out/multi_n$N.pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.err | tail -5; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"]); print("strong", d.get("strong")); print("lum", d.get("lum_link_sharded"))
except Exception as e: print("parse failed", e)
PY
