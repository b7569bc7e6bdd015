set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log); tail -5 gpurun_out/r2_t4.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c2.log 2> gpurun_out/r2_bench_c2.err; tail -c 2500 gpurun_out/r2_bench_c2.log; tail -3 gpurun_out/r2_bench_c2.err
timeout 900 python bench.py --workload config3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c3.log 2> gpurun_out/r2_bench_c3.err; tail -c 2500 gpurun_out/r2_bench_c3.log; tail -3 gpurun_out/r2_bench_c3.err
KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_mzasync.so timeout 900 python bench.py --workload config3 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-oracle > gpurun_out/r2_bench_c3_mzasync.log 2> gpurun_out/r2_bench_c3_mzasync.err; tail -c 1500 gpurun_out/r2_bench_c3_mzasync.log; tail -3 gpurun_out/r2_bench_c3_mzasync.err
