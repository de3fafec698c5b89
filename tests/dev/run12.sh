cd /root/repo/tests
python gpu_flips.py > ../gpurun_out/r02_flips.txt 2> ../gpurun_out/r02_flips.err; tail -5 ../gpurun_out/r02_flips.txt
export TSB_AUTOTUNE=0
ncu --set full --clock-control none --import-source on -k regex:tsb_optran --launch-skip 1 --launch-count 1 -f -o ../gpurun_out/r02_diode2_v6 python gpu_one.py diode2 1048576 > ../gpurun_out/ncu2.log 2>&1
unset TSB_AUTOTUNE
cd /root/repo
python -m pytest tests -m gpu -q > gpurun_out/r02_t15.log 2>&1; tail -4 gpurun_out/r02_t15.log
