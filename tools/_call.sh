O=gpurun_out/r2_tests_g.txt; : > $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 >> $O
timeout 300 python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline --no-parity > gpurun_out/tmp.json 2>gpurun_out/tmp.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('1gpu', round(d['value']), round(d['ms_per_step']*1e3,1))" >> $O
cat $O
