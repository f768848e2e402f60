# Scratch script for `gpurun -- 'bash tools/_call.sh'` (rewritten per call during development).
# This version reproduces the committed one-GPU evidence: GPU tests, smoke, the bench line, the reference arm and the
# ncu launch list of the bench command (the latter only after the same command has exited 0 without ncu).
O=gpurun_out/evidence_1gpu.txt; : > $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?" >> $O
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-parity > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> $O
cat $O
