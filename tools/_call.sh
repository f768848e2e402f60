O=gpurun_out/r2_wide.txt; : > $O
W=$PWD/uemda_b200/libuem_b200_wide.so
UEM_B200_LIB=$W timeout 600 python -m pytest tests -m gpu -x -q -k "refine or golden or chain or mining_step" 2>&1 | tail -3 >> $O
echo "== default lib, cfg3" >> $O
timeout 300 python tools/kbench.py --workload cfg3_loveda_16x7x1024 --iters 24 --only label_refine,mine_chain >> $O 2>&1
echo "== wide (c=7: two columns per lane, 248 registers, 2 CTAs/SM), cfg3" >> $O
UEM_B200_LIB=$W timeout 300 python tools/kbench.py --workload cfg3_loveda_16x7x1024 --iters 24 --only label_refine,mine_chain >> $O 2>&1
cat $O
