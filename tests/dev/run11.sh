cd /root/repo/tests
for d in diode2 diode4; do python gpu_ab.py $d 4194304 "|TSB_X_EXPC=0;TSB_X_CONVSEL=0" "|TSB_X_EXPC=1;TSB_X_CONVSEL=0" "|TSB_X_EXPC=0;TSB_X_CONVSEL=1" "" >> ../gpurun_out/r02_ab6.log 2>&1; done
for d in mosfet1 bjt2; do python gpu_ab.py $d 4194304 "|TSB_X_CONVSEL=0" "" >> ../gpurun_out/r02_ab6.log 2>&1; done
cat ../gpurun_out/r02_ab6.log
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_extra.py -m gpu -q -x -k "diode or dio or mos or bjt or pnp" > gpurun_out/r02_t14.log 2>&1; tail -4 gpurun_out/r02_t14.log
