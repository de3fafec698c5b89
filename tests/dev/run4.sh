cd /root/repo/tests
python gpu_ab.py rlc 1048576 "share_time_grid=0" "share_time_grid=1" "share_time_grid=1||TSB_TGRID_REUSE=0" "share_time_grid=1,min_blocks=5||TSB_TGRID_REUSE=1" "share_time_grid=0,min_blocks=5" "share_time_grid=1|TSB_TG_CACHED=0;TSB_TG_PREFETCH=0" "share_time_grid=1|TSB_TG_PREFETCH=0" > ../gpurun_out/r02_ab3.log 2>&1
python gpu_ab.py rlc 256 "share_time_grid=0" "share_time_grid=1||TSB_TGRID_REUSE=0" >> ../gpurun_out/r02_ab3.log 2>&1
export TSB_TGRID_REUSE=1
python gpu_ab.py rl 1048576 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab3.log 2>&1
python gpu_ab.py transformer1 1048576 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab3.log 2>&1
python gpu_ab.py rc 1048576 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab3.log 2>&1
cat ../gpurun_out/r02_ab3.log
cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shared_time_grid or strict_build_is_bit or skipping" > gpurun_out/r02_t4.log 2>&1; tail -5 gpurun_out/r02_t4.log
python -m pytest tests/test_gpu_job.py tests/test_gpu_extra.py -m gpu -q > gpurun_out/r02_t5.log 2>&1; tail -12 gpurun_out/r02_t5.log
export TSB_AUTOTUNE=0
cd tests
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 2 --launch-count 1 -f -o ../gpurun_out/r02_rlc_v3_share1 python gpu_one.py rlc 1048576 share_time_grid=1 > ../gpurun_out/ncu1.log 2>&1
