python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; tail -c 400 gpurun_out/r2_final_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/r2_final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 1 --warmup 3 --no-strict --no-parity --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1; tail -1 gpurun_out/ncu_launch2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"k_build_R_fast|k_render_fast" -c 2 -o gpurun_out/r2_transport_v12 -f python tools/probe_frame.py --reps 1 > gpurun_out/ncu_v12.log 2>&1; tail -2 gpurun_out/ncu_v12.log
timeout 300 python bench.py --config C3 --slice-range 0 4 --steps 1 --warmup 1 --no-strict --parity-seconds 3 --cpu-seconds 6 > gpurun_out/r2_bench_C3_slices0_4.json 2> gpurun_out/r2_bench_C3.err; tail -c 300 gpurun_out/r2_bench_C3_slices0_4.json
