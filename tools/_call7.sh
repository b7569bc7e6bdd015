set -x
(timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log); tail -5 gpurun_out/r2_t7.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.log 2> gpurun_out/r2_bench_full.err; echo rc=$?; tail -c 4500 gpurun_out/r2_bench_full.log; tail -5 gpurun_out/r2_bench_full.err
