python -m pytest tests -m gpu -x -q -k "refine or golden or chain or oracle_mining or full_size or workspace" 2>&1 | tail -4
K="--only label_refine,mine_chain --iters 120"
echo "== form1 + prologue overlap + tap prefetch cfg2" > gpurun_out/r2_kb_refine6.txt; python tools/kbench.py $K --refine-form 1 >> gpurun_out/r2_kb_refine6.txt 2>&1
echo "== cfg5" >> gpurun_out/r2_kb_refine6.txt; python tools/kbench.py $K --refine-form 1 --workload cfg5_sweep_32x6x512 >> gpurun_out/r2_kb_refine6.txt 2>&1
echo "== cfg3" >> gpurun_out/r2_kb_refine6.txt; python tools/kbench.py $K --refine-form 1 --workload cfg3_loveda_16x7x1024 >> gpurun_out/r2_kb_refine6.txt 2>&1
grep -v "^entry" gpurun_out/r2_kb_refine6.txt
python tools/refine_timing.py > gpurun_out/r2_refine_timeline_b.txt 2>&1; cat gpurun_out/r2_refine_timeline_b.txt
