python tools/probe_transport.py --config C4 --width 1920 --height 1080 --vrls 500 --reps 2 > gpurun_out/s3_pt_c4.log 2>&1; tail -2 gpurun_out/s3_pt_c4.log
python tools/probe_transport.py --config C3 --vrls 200 --reps 2 > gpurun_out/s3_pt_c3.log 2>&1; tail -2 gpurun_out/s3_pt_c3.log
ncu --set full --clock-control none --import-source on -k regex:k_build_R -c 1 -o gpurun_out/r2_c4_bvh_v0 -f python tools/probe_transport.py --config C4 --width 1920 --height 1080 --vrls 500 --reps 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
ncu --set full --clock-control none --import-source on -k regex:k_build_R -c 1 -o gpurun_out/r2_c3_march_v0 -f python tools/probe_transport.py --config C3 --vrls 200 --reps 1 > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
