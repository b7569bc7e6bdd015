set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t12.log); tail -5 gpurun_out/r2_t12.log
timeout 400 python tools/sweep.py --grid r2 --steps 3 > gpurun_out/r2_sweep_prefetch.jsonl 2>gpurun_out/r2_sweep_prefetch.err; cut -c1-330 gpurun_out/r2_sweep_prefetch.jsonl
KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_noprefetch.so timeout 400 python tools/sweep.py --grid r2 --steps 3 > gpurun_out/r2_sweep_noprefetch.jsonl 2>gpurun_out/r2_sweep_noprefetch.err; cut -c1-330 gpurun_out/r2_sweep_noprefetch.jsonl
