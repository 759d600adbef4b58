echo "== parity one"; B200ICP_LIB=$PWD/3dtk_b200/lib/one.so timeout 800 python -m pytest tests/test_gpu_parity.py tests/test_dat_config1.py tests/test_do_icp.py -m gpu -x -q 2>&1 | tail -3
PPC=4 bash tools/ab.sh base one base one
