python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_a.json 2> gpurun_out/r2_bench_final_a.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_a.csv python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/ncu_launch_a.log 2>&1
tail -c 600 gpurun_out/r2_bench_final_a.err; tail -2 gpurun_out/ncu_launch_a.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:refine_col_kernel -s 30 -c 2 -o gpurun_out/prof_refine_r02 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_b.log 2>&1
tail -2 gpurun_out/ncu_full_b.log | cut -c1-300
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_final_a.json 2> gpurun_out/r2_ref_final_a.err
