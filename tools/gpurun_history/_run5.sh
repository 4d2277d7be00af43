python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 > gpurun_out/s3_c3c.log 2>&1; tail -1 gpurun_out/s3_c3c.log
python bench.py --config C4 --slice-range 0 4 --steps 2 --warmup 1 --no-strict --parity-seconds 4 --cpu-seconds 6 > gpurun_out/r2_bench_C4_slices0_4.json 2> gpurun_out/r2_bench_C4.err; tail -c 1500 gpurun_out/r2_bench_C4_slices0_4.json; tail -3 gpurun_out/r2_bench_C4.err
