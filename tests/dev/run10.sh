cd /root/repo/tests
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
python gpu_flops.py > ../gpurun_out/flops_runs_plain.jsonl 2> ../gpurun_out/flops_plain.err
ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/flops_ncu.csv python gpu_flops.py > ../gpurun_out/flops_runs.jsonl 2> ../gpurun_out/flops.err
cd /root/repo
python profiles/flops_from_ncu.py gpurun_out/flops_ncu.csv gpurun_out/flops_runs.jsonl profiles/executed_flops.json
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k refill > gpurun_out/r02_t13.log 2>&1; tail -4 gpurun_out/r02_t13.log
python bench.py > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; tail -c 400 gpurun_out/r02_bench3.json
cp profiles/executed_flops.json gpurun_out/executed_flops.json
