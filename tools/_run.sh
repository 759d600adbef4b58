timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
PPC=4 bash tools/ab.sh base libb200icp base libb200icp
