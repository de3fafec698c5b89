cd /root/repo/tests
python gpu_ab.py rlc 1048576 "share_time_grid=0||TSB_TRANFAST=0" "share_time_grid=0||TSB_TRANFAST=1" "share_time_grid=1||TSB_TRANFAST=1" "share_time_grid=1|TSB_TG_CACHED=0;TSB_TG_PREFETCH=0|TSB_TRANFAST=1" "share_time_grid=1,min_blocks=5||TSB_TRANFAST=1" "share_time_grid=0,min_blocks=5||TSB_TRANFAST=1" > ../gpurun_out/r02_ab4.log 2>&1
export TSB_TRANFAST=1
python gpu_ab.py rl 1048576 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab4.log 2>&1
python gpu_ab.py rc 1048576 "share_time_grid=0" "share_time_grid=1" >> ../gpurun_out/r02_ab4.log 2>&1
for d in diode2 diode4 mosfet1; do
python gpu_ab.py $d 4194304 "||TSB_TRANFAST=0" "||TSB_TRANFAST=1" >> ../gpurun_out/r02_ab4.log 2>&1
done
python gpu_ab.py transformer1 1048576 "strict_fp=0||TSB_TRANFAST=0" "strict_fp=0||TSB_TRANFAST=1" "strict_fp=1" >> ../gpurun_out/r02_ab4.log 2>&1
cat ../gpurun_out/r02_ab4.log
unset TSB_TRANFAST
cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or auto or shared_time_grid or golden" > gpurun_out/r02_t6.log 2>&1; tail -12 gpurun_out/r02_t6.log
python -m pytest tests/test_gpu_extra.py tests/test_random_decks.py tests/test_gpu_job.py -m gpu -q > gpurun_out/r02_t7.log 2>&1; tail -12 gpurun_out/r02_t7.log
