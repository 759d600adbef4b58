#!/bin/bash
# GPU box: parity tests on the default library, then A/B of the named variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
PPC=4 tools/ab.sh "$@" 2>&1 | tee gpurun_out/ab.log
