cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r02_t9.log 2>&1; tail -8 gpurun_out/r02_t9.log
python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -c 3000 gpurun_out/r02_bench1.json
