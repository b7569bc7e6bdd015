# usage: bash tools/records_ngpu.sh N V   -- multi-GPU records (torchrun, one rank per GPU)
N=$1
V=$2
P=29711
if [ $N = 2 ]; then (timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_${V}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${V}_pytest_multi.log); tail -2 gpurun_out/r2_${V}_pytest_multi.log; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --no-files > gpurun_out/r2_${V}_c2_n$N.log 2> gpurun_out/r2_${V}_c2_n$N.err; tail -1 gpurun_out/r2_${V}_c2_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2 N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['host_transport'][:50], d['checks'])"; tail -2 gpurun_out/r2_${V}_c2_n$N.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --workload config3 --steps 3 --warmup 3 --no-e2e --no-files > gpurun_out/r2_${V}_c3_n$N.log 2> gpurun_out/r2_${V}_c3_n$N.err; tail -1 gpurun_out/r2_${V}_c3_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c3 N=$N', d['value'], d['ms_per_step'], d['roofline']['frac_step'], d['checks'])"; tail -2 gpurun_out/r2_${V}_c3_n$N.err
