PPC=4 bash tools/ab.sh base dyn1 hy4 hy6 hy7
