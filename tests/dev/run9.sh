cd /root/repo
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k refill > gpurun_out/r02_t12.log 2>&1; tail -4 gpurun_out/r02_t12.log
cd tests
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum
python gpu_flops.py > ../gpurun_out/flops_runs_plain.jsonl 2> ../gpurun_out/flops_plain.err; tail -3 ../gpurun_out/flops_runs_plain.jsonl
ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/flops_ncu.csv python gpu_flops.py > ../gpurun_out/flops_runs.jsonl 2> ../gpurun_out/flops.err
export TSB_AUTOTUNE=0
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 2 --launch-count 1 -f -o ../gpurun_out/r02_rlc_v5 python gpu_one.py rlc 1048576 share_time_grid=1 > ../gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_diode2_v5 python gpu_one.py diode2 1048576 > ../gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_bjt2_v5 python gpu_one.py bjt2 1048576 > ../gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_mosfet1_v5 python gpu_one.py mosfet1 1048576 > ../gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 2 --launch-count 1 -f -o ../gpurun_out/r02_transformer2_v5 python gpu_one.py transformer2 262144 share_time_grid=1 > ../gpurun_out/ncu5.log 2>&1
unset TSB_AUTOTUNE
cd /root/repo
python bench.py > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; tail -c 600 gpurun_out/r02_bench2.json
