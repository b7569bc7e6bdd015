set -x
nvidia-smi -L | head -3
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t5_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5_multi.log); tail -15 gpurun_out/r2_t5_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_n2.log 2> gpurun_out/r2_bench_c2_n2.err; echo rc=$?; tail -c 3000 gpurun_out/r2_bench_c2_n2.log; tail -5 gpurun_out/r2_bench_c2_n2.err
timeout 900 $TR bench.py --gpus 2 --workload config3 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_bench_c3_n2.log 2> gpurun_out/r2_bench_c3_n2.err; echo rc=$?; tail -c 2500 gpurun_out/r2_bench_c3_n2.log; tail -5 gpurun_out/r2_bench_c3_n2.err
