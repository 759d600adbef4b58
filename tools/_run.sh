timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_last.json; cut -c1-200 gpurun_out/bench_last.json
