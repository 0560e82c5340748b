set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_n1_steps3.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ber_tconv2 -s 3 -c 1 -f -o gpurun_out/prof_r2b_k1 python tools/quick_bench.py wtx 256 27 > gpurun_out/r2b_p1.log 2>&1
K1AB_CASES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ber_tconv2 -s 3 -c 1 -f -o gpurun_out/prof_r2b_k1_n1024 python tools/k1_ab.py 1024 3 2 > gpurun_out/r2b_p2.log 2>&1
tail -3 gpurun_out/r2b_tests.log; head -c 600 gpurun_out/r2b_bench_n1.json; tail -2 gpurun_out/r2b_p1.log gpurun_out/r2b_p2.log
