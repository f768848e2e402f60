python -m pytest tests -m gpu -x -q -k "peer_exchange" 2>&1 | tail -3
export NCCL_DEBUG=WARN
B="--steps 60 --warmup 10 --no-extra --no-e2e --no-cpu-baseline"
python bench.py $B > gpurun_out/tmp.json 2> gpurun_out/tmp.err
echo "== [N=1] $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step']*1e3,1), 'us', d['kernels_per_step'])" 2>&1 | tail -1) $(tail -c 200 gpurun_out/tmp.err | tr '\n' ' ')" >> gpurun_out/r2_pdlid.txt
python bench.py $B --force-peer > gpurun_out/tmp.json 2> gpurun_out/tmp.err
echo "== [N=1 force-peer] $(python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step']*1e3,1), 'us', d['kernels_per_step'], d['impl_detail']['exchange_status'])" 2>&1 | tail -1) $(tail -c 200 gpurun_out/tmp.err | tr '\n' ' ')" >> gpurun_out/r2_pdlid.txt
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 60 --warmup 10 --no-e2e > gpurun_out/tmp2.json 2> gpurun_out/tmp2.err
echo "${N}gpu rc=$?" >> gpurun_out/r2_pdlid.txt; tail -c 200 gpurun_out/tmp2.err | tr '\n' ' ' >> gpurun_out/r2_pdlid.txt
python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp2.json') if l.startswith('{')][-1]); print('${N}gpu', round(d['value']), round(d['ms_per_step']*1e3,1), d['impl_detail']['exchange'], d['impl_detail']['exchange_status'], d['kernels_per_step'], json.dumps(d.get('parity'))[:330])" >> gpurun_out/r2_pdlid.txt
cat gpurun_out/r2_pdlid.txt
