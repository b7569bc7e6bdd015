(timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r2_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t23.log); tail -2 gpurun_out/r2_t23.log
timeout 600 python tools/text_route_profile.py 8000000 2>gpurun_out/r2_text_profile3.err | grep '"rep": 1' > gpurun_out/r2_text_profile3.jsonl; cat gpurun_out/r2_text_profile3.jsonl
timeout 600 python tools/gz_device_profile.py 8000000 2> gpurun_out/r2_gz_device2.err | grep '"rep": 2' > gpurun_out/r2_gz_device2.jsonl; cat gpurun_out/r2_gz_device2.jsonl; tail -3 gpurun_out/r2_gz_device2.err
