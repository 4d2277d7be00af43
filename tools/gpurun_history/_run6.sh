python -m pytest tests/test_chain.py tests/test_film.py -m gpu -x -q -s > gpurun_out/s3_tests_new.log 2>&1; tail -25 gpurun_out/s3_tests_new.log
python -m pytest tests -m gpu -x -q --deselect tests/test_chain.py --deselect tests/test_film.py > gpurun_out/s3_tests4.log 2>&1; tail -4 gpurun_out/s3_tests4.log
python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 > gpurun_out/s3_c3c.log 2>&1; tail -1 gpurun_out/s3_c3c.log
python tools/probe_frame.py --reps 2 > gpurun_out/s3_c2.log 2>&1; tail -1 gpurun_out/s3_c2.log
