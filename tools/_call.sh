O=gpurun_out/r2_pdl_pearson.txt; : > $O
UEM_B200_OPTS=pdl_pearson=1 timeout 600 python -m pytest tests -m gpu -x -q -k "pearson or chain or mining_step or golden or staged or hint" 2>&1 | tail -3 >> $O
run() { lbl=$1; shift
  timeout 300 python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline --no-parity "$@" > gpurun_out/tmp.json 2>gpurun_out/tmp.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('bench [$lbl]', round(d['value']), round(d['ms_per_step']*1e3,1))" >> $O 2>&1 || tail -3 gpurun_out/tmp.err >> $O
}
run default
run pdl_pearson --opt pdl_pearson=1
run default
run pdl_pearson --opt pdl_pearson=1
echo "== kbench mine_chain default / pdl" >> $O
timeout 200 python tools/kbench.py --only mine_chain,pearson_nchw | grep -v "^entry" >> $O 2>&1
timeout 200 python tools/kbench.py --only mine_chain,pearson_nchw --opt pdl_pearson=1 | grep -v "^entry" >> $O 2>&1
cat $O
