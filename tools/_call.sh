O=gpurun_out/r2_l2hints2.txt; : > $O
for v in "l2_region=0" "l2_region=1" "l2_region=2" "l2_region=1 l2_keep=2" "l2_region=1 l2_last_use=0" "l2_stream=0 l2_last_use=0" "l2_region=0" "l2_region=1"; do
  o=""; for kv in $v; do o="$o --opt $kv"; done
  timeout 300 python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline --no-parity $o > gpurun_out/tmp.json 2>gpurun_out/tmp.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('bench [$v]', round(d['value']), round(d['ms_per_step']*1e3,1), 'roofline', round(d['roofline']['frac'],3), d['roofline'].get('kernel_ms'))" >> $O 2>&1
done
timeout 300 python tools/l2_probe.py >> $O 2>&1 || exit 1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:"pearson_tma|region_max_smem|refine_col|select_stats" --csv --log-file gpurun_out/r2_l2probe2.csv python tools/l2_probe.py >> $O 2>&1
grep -c "pass" gpurun_out/r2_l2probe2.csv >> $O
cat $O
