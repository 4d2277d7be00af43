python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests2.log 2>&1; tail -8 gpurun_out/s3_tests2.log
python -m pytest tests/test_c2_parity_gpu.py -m gpu -q -s -k "other_config" 2>&1 | grep -E "^C[345] " | cut -c1-400
python tools/probe_frame.py --config C4 --vrls 3000 --reps 2 > gpurun_out/s3_c4b.log 2>&1; tail -1 gpurun_out/s3_c4b.log
python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 > gpurun_out/s3_c3b.log 2>&1; tail -1 gpurun_out/s3_c3b.log
