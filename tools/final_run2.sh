set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/r2c_k2_launches.csv python tools/k2_once.py 250 > gpurun_out/r2c_k2_once.log 2>&1
cat gpurun_out/r2c_tests.log; head -c 300 gpurun_out/r2c_bench_n1.json; tail -n 3 gpurun_out/r2c_k2_once.log
