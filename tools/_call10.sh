set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t10.log); tail -15 gpurun_out/r2_t10.log
timeout 600 python tools/gz_device_profile.py 8000000 > gpurun_out/r2_gz_device.jsonl 2> gpurun_out/r2_gz_device.err; cat gpurun_out/r2_gz_device.jsonl; tail -5 gpurun_out/r2_gz_device.err
