python tools/_diag_c3.py 2>&1 | grep grid
echo default; python tools/probe_frame.py --config C3 --vrls 2000 --reps 2 | tail -1
