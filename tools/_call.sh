O=gpurun_out/r2_selstream.txt; : > $O
run() { # label, env...
  lbl=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline --no-parity > gpurun_out/tmp.json 2>gpurun_out/tmp.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('bench [$lbl]', round(d['value']), round(d['ms_per_step']*1e3,1))" >> $O 2>&1 || tail -5 gpurun_out/tmp.err >> $O
}
run base A=1
run sel_low UEM_BENCH_SELECT_STREAM=1 UEM_BENCH_SEL_PRIO=l
run sel_high UEM_BENCH_SELECT_STREAM=1 UEM_BENCH_SEL_PRIO=h
run base A=1
run sel_low UEM_BENCH_SELECT_STREAM=1 UEM_BENCH_SEL_PRIO=l
UEM_BENCH_SELECT_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-e2e > gpurun_out/tmp2.json 2>gpurun_out/tmp2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp2.json') if l.startswith('{')][-1]); print('parity run', round(d['value']), round(d['ms_per_step']*1e3,1), json.dumps(d.get('parity'))[:400])" >> $O 2>&1 || tail -5 gpurun_out/tmp2.err >> $O
cat $O
