python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
python tools/probe_frame.py --config C4 --vrls 3000 --reps 2 > gpurun_out/s3_c4.log 2>&1; tail -3 gpurun_out/s3_c4.log
python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 > gpurun_out/s3_c3.log 2>&1; tail -3 gpurun_out/s3_c3.log
ALVRL_PROFILE=1 python tools/probe_frame.py --slice-range 0 12 --reps 2 > gpurun_out/s3_c2_12.log 2>&1; tail -30 gpurun_out/s3_c2_12.log
