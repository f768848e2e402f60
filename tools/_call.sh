export NCCL_DEBUG=WARN
python tools/regen_bench.py --distinct 2 > gpurun_out/r02_regen_cfg4_1gpu.json 2> gpurun_out/regen1.err; tail -c 300 gpurun_out/regen1.err
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N tools/regen_bench.py --distinct 2 > gpurun_out/r02_regen_cfg4_${N}gpu.json 2> gpurun_out/regen$N.err
echo "N=$N rc=$?"; tail -c 300 gpurun_out/regen$N.err | tr '\n' ' '
done
for N in 1 2 4 8; do grep "^{" gpurun_out/r02_regen_cfg4_${N}gpu.json | cut -c1-600; done
