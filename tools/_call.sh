export NCCL_DEBUG=WARN
O=gpurun_out/r2_scale5.txt; : > $O
runN() { # N tag env
  N=$1; tag=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2_bench_${N}gpu_$tag.json 2> gpurun_out/r2_bench_${N}gpu_$tag.err
  echo "${N}gpu $tag rc=$?" >> $O
  python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu_$tag.json') if l.startswith('{')][-1]); p=d.get('parity') or {}
v=p.get('vs_one_gpu_concatenated_batch') or {}
good = d['impl_detail']['exchange_status']==0 and p.get('prototypes_bit_identical_across_ranks') and v.get('prototypes_within_1e-5') and sum(v.get('hard_label_mismatch_per_rank_step0') or [1])==0
print('${N}gpu $tag', round(d['value']), round(d['ms_per_step']*1e3,1), d['impl_detail']['exchange'], d['impl_detail']['exchange_status'], d['kernels_per_step'], 'e2e', round(d['e2e']['value']), 'GOOD' if good else 'BAD', json.dumps(p)[:300])
sys.exit(0 if good else 1)" >> $O 2>&1
}
runN 8 h A=1 || runN 8 h_noregsend UEM_BENCH_REGION_SEND=0
runN 4 h A=1
runN 2 h A=1
python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/tmp.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('1gpu', round(d['value']), round(d['ms_per_step']*1e3,1))" >> $O
cat $O
