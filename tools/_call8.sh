set -x
timeout 600 python tools/text_route_profile.py 8000000 > gpurun_out/r2_text_profile.jsonl 2> gpurun_out/r2_text_profile.err; cat gpurun_out/r2_text_profile.jsonl; tail -3 gpurun_out/r2_text_profile.err
timeout 400 python tools/sweep.py --grid r2 --steps 3 > gpurun_out/r2_sweep_t128.jsonl 2>gpurun_out/r2_sweep_t128.err; cut -c1-300 gpurun_out/r2_sweep_t128.jsonl
KMB_LIB_PATH=$PWD/kmer_mapper_b200/libkmer_mapper_b200_t256.so timeout 400 python tools/sweep.py --grid r2 --steps 3 > gpurun_out/r2_sweep_t256.jsonl 2>gpurun_out/r2_sweep_t256.err; cut -c1-300 gpurun_out/r2_sweep_t256.jsonl
