set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2i_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo smoke rc=$?
timeout 600 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo bench rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_gemm_f16 -s 2 -c 1 -f -o gpurun_out/prof_r2i_mask_gemm python tools/mask_once.py 4 > gpurun_out/r2i_p1.log 2>&1
tail -3 gpurun_out/r2i_tests.log; head -c 400 gpurun_out/r2i_bench_n1.json; tail -2 gpurun_out/r2i_p1.log
