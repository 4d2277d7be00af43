python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests3.log 2>&1; tail -4 gpurun_out/s3_tests3.log
python tools/probe_frame.py --config C4 --vrls 3000 --reps 2 > gpurun_out/s3_c4c.log 2>&1; tail -1 gpurun_out/s3_c4c.log
ALVRL_NO_BVH4=1 python tools/probe_frame.py --config C4 --vrls 3000 --reps 2 > gpurun_out/s3_c4c_bin.log 2>&1; tail -1 gpurun_out/s3_c4c_bin.log
ncu --set full --clock-control none --import-source on -k regex:k_build_R -c 1 -o gpurun_out/r2_c4_bvh_v2 -f python tools/probe_transport.py --config C4 --width 1920 --height 1080 --vrls 1000 --reps 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
ncu --set full --clock-control none --import-source on -k regex:k_build_R -c 1 -o gpurun_out/r2_c3_march_v1 -f python tools/probe_transport.py --config C3 --vrls 400 --reps 1 > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
