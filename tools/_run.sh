timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== fused (default)"; PPC=4 bash tools/ab.sh libb200icp
echo "== split"; B200ICP_SPLIT=1 PPC=4 bash tools/ab.sh libb200icp
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01b.json; cat gpurun_out/bench_r01b.json
B200ICP_SPLIT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r01b_split.json; cut -c1-300 gpurun_out/bench_r01b_split.json
