cd /root/repo/tests
python gpu_ab.py rlc 1048576 "share_time_grid=0" "share_time_grid=1" "share_time_grid=1|TSB_TG_CACHED=0;TSB_TG_PREFETCH=0" "share_time_grid=1|TSB_TG_PREFETCH=0" "share_time_grid=1|TSB_TG_PUBLISH_EVERY=32" "share_time_grid=1,min_blocks=5" > ../gpurun_out/r02_ab2.log 2>&1
python gpu_ab.py rlc 131072 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab2.log 2>&1
python gpu_ab.py rlc 4194304 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab2.log 2>&1
cat ../gpurun_out/r02_ab2.log
export TSB_AUTOTUNE=0
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 3 --launch-count 1 -f -o ../gpurun_out/r02_rlc_share1 python gpu_one.py rlc 1048576 share_time_grid=1 > ../gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_rlc_share0 python gpu_one.py rlc 1048576 share_time_grid=0 > ../gpurun_out/ncu0.log 2>&1
tail -3 ../gpurun_out/ncu1.log ../gpurun_out/ncu0.log
ls -la ../gpurun_out/*.ncu-rep
