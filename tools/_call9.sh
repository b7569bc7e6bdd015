set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log); tail -6 gpurun_out/r2_t9.log
timeout 600 python tools/text_route_profile.py 8000000 > gpurun_out/r2_text_profile2.jsonl 2> gpurun_out/r2_text_profile2.err; cat gpurun_out/r2_text_profile2.jsonl; tail -3 gpurun_out/r2_text_profile2.err
