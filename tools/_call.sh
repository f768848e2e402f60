K="--only label_refine,mine_chain --iters 120"
echo "== form1 default (MINB 3) cfg2" > gpurun_out/r2_kb_refine4.txt; python tools/kbench.py $K --refine-form 1 >> gpurun_out/r2_kb_refine4.txt 2>&1
echo "== variant a: form1 MINB=2 cfg2" >> gpurun_out/r2_kb_refine4.txt; UEM_B200_LIB=$PWD/uemda_b200/libuem_b200_a.so python tools/kbench.py $K --refine-form 1 >> gpurun_out/r2_kb_refine4.txt 2>&1
for t in b c d; do echo "== variant $t cfg2" >> gpurun_out/r2_kb_refine4.txt; UEM_B200_LIB=$PWD/uemda_b200/libuem_b200_$t.so python tools/kbench.py $K --refine-form 0 >> gpurun_out/r2_kb_refine4.txt 2>&1; done
export UEM_B200_LIB=$PWD/uemda_b200/libuem_b200_c.so
python tools/kbench.py --only label_refine --iters 24 --refine-form 0 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:refine_col2 -s 4 -c 2 -o gpurun_out/prof_col2_r02c python tools/kbench.py --only label_refine --iters 24 --refine-form 0 > gpurun_out/ncu4.log 2>&1
