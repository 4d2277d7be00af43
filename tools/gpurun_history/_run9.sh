for v in gold ga gold; do echo "variant $v"; ALVRL_LIB=build/libalvrl_$v.so python tools/probe_frame.py --config C3 --vrls 2000 --reps 3 | tail -2; done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,clocks_throttle_reasons.active,temperature.gpu,power.draw --format=csv
