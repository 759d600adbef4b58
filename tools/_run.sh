set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_final.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_final_ref.json; cut -c1-300 gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_final.csv
