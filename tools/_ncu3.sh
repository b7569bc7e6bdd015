ncu --set full --clock-control none --import-source on -k regex:"kmb_gz_inflate" -s 1 -c 1 -o gpurun_out/r2_gz_inflate python tools/gz_device_profile.py 2000000 > gpurun_out/r2_ncu3.log 2>&1
tail -3 gpurun_out/r2_ncu3.log
