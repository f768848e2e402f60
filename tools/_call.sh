export NCCL_DEBUG=WARN
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2_bench_${N}gpu_f.json 2> gpurun_out/r2_bench_${N}gpu_f.err
echo "${N}gpu rc=$?" >> gpurun_out/r2_scale3.txt; tail -c 200 gpurun_out/r2_bench_${N}gpu_f.err | tr '\n' ' ' >> gpurun_out/r2_scale3.txt
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu_f.json') if l.startswith('{')][-1]); print('${N}gpu', round(d['value']), round(d['ms_per_step']*1e3,1), d['impl_detail']['exchange'], d['impl_detail']['exchange_status'], d['kernels_per_step'], 'e2e', round(d['e2e']['value']), json.dumps(d.get('parity'))[:330])" >> gpurun_out/r2_scale3.txt
done
python bench.py --steps 40 --warmup 10 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/tmp.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/tmp.json') if l.startswith('{')][-1]); print('1gpu', round(d['value']), round(d['ms_per_step']*1e3,1))" >> gpurun_out/r2_scale3.txt
cat gpurun_out/r2_scale3.txt
