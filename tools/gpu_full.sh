#!/bin/bash
# GPU box: the whole gpu suite, then the bench line and the reference-arm line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 600 gpurun_out/bench_ours.err; cut -c1-400 gpurun_out/bench_ours.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
