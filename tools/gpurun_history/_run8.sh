for v in ga gb gc gd; do echo "variant $v"; ALVRL_LIB=build/libalvrl_$v.so python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 | tail -1; done
echo "default"; python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 | tail -1
