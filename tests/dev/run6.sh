cd /root/repo/tests
python gpu_ab.py rlc 1048576 "share_time_grid=0" "share_time_grid=1" > ../gpurun_out/r02_ab5.log 2>&1
for d in diode2 diode4 mosfet1 bjt2; do python gpu_ab.py $d 4194304 "||TSB_TRANFAST=0" "||TSB_TRANFAST=1" >> ../gpurun_out/r02_ab5.log 2>&1; done
unset TSB_TRANFAST
for d in rl transformer1 transformer2 transformer3; do python gpu_ab.py $d 1048576 "" >> ../gpurun_out/r02_ab5.log 2>&1; done
cat ../gpurun_out/r02_ab5.log
export TSB_AUTOTUNE=0
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 2 --launch-count 1 -f -o ../gpurun_out/r02_rlc_v4 python gpu_one.py rlc 1048576 share_time_grid=1 > ../gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_diode2_v4 python gpu_one.py diode2 1048576 > ../gpurun_out/ncu2.log 2>&1
unset TSB_AUTOTUNE
cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r02_t8.log 2>&1; tail -8 gpurun_out/r02_t8.log
