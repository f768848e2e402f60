O=gpurun_out/r2_refine_ablation.txt; : > $O
for v in 0 1 2 4 8 16 3 18 19; do
  if [ $v = 0 ]; then L=uemda_b200/libuem_b200.so; else L=uemda_b200/libuem_b200_abl$v.so; fi
  echo "== UEM_ABL=$v" >> $O
  UEM_B200_LIB=$PWD/$L timeout 200 python tools/kbench.py --only label_refine >> $O 2>&1
done
echo "== cfg5 (batch 32)" >> $O
for v in 0 1 2 3 19; do
  if [ $v = 0 ]; then L=uemda_b200/libuem_b200.so; else L=uemda_b200/libuem_b200_abl$v.so; fi
  echo "== UEM_ABL=$v" >> $O
  UEM_B200_LIB=$PWD/$L timeout 200 python tools/kbench.py --workload cfg5_sweep_32x6x512 --iters 24 --only label_refine >> $O 2>&1
done
grep -v "^entry" $O
