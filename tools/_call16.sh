set -x
P=29511
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_v8_c2_n2.log 2> gpurun_out/r2_v8_c2_n2.err; tail -1 gpurun_out/r2_v8_c2_n2.log | cut -c1-1800; tail -3 gpurun_out/r2_v8_c2_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --workload config3 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_v8_c3_n2.log 2> gpurun_out/r2_v8_c3_n2.err; tail -1 gpurun_out/r2_v8_c3_n2.log | cut -c1-600; tail -3 gpurun_out/r2_v8_c3_n2.err
(timeout 600 python -m pytest tests -m gpu -x -q tests/test_gpu_multi.py > gpurun_out/r2_t16_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t16_multi.log); tail -3 gpurun_out/r2_t16_multi.log
