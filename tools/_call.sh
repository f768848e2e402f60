python -m pytest tests -m gpu -x -q -k "peer_exchange" 2>&1 | tail -3
export NCCL_DEBUG=WARN
B="--steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline"
for o in "" "--force-peer"; do
  python bench.py $B $o > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  echo "== [$o] $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step']*1e3,1), 'us; kernels', d['kernels_per_step'], d['impl_detail']['exchange_status'])" 2>&1 | tail -1) $(tail -c 300 gpurun_out/tmp.err | tr '\n' ' ')" >> gpurun_out/r2_split.txt
done
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 10 --no-e2e > gpurun_out/r2_bench_${N}gpu_d.json 2> gpurun_out/r2_bench_${N}gpu_d.err
echo "${N}gpu rc=$?" >> gpurun_out/r2_split.txt; tail -c 300 gpurun_out/r2_bench_${N}gpu_d.err | tr '\n' ' ' >> gpurun_out/r2_split.txt
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu_d.json') if l.startswith('{')][-1]); print('${N}gpu', round(d['value']), round(d['ms_per_step']*1e3,1), d['impl_detail']['exchange'], d['impl_detail']['exchange_status'], d['kernels_per_step'], json.dumps(d.get('parity'))[:330])" >> gpurun_out/r2_split.txt
cat gpurun_out/r2_split.txt
