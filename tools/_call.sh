python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest6.log; tail -c 700 gpurun_out/r2_pytest6.log
python tools/lossbench.py > gpurun_out/r2_lossbench.txt 2>&1; cat gpurun_out/r2_lossbench.txt
