cd /root/repo
python -m pytest tests/test_ac.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02_t10.log 2>&1; tail -8 gpurun_out/r02_t10.log
python -m pytest tests -m gpu -q > gpurun_out/r02_t11.log 2>&1; tail -8 gpurun_out/r02_t11.log
