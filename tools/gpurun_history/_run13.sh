for v in gold cur; do
  lib=build/libalvrl_$v.so; [ $v = cur ] && lib=mitsuba-alvrl_b200/libalvrl.so
  ALVRL_LIB=$lib ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none -k regex:k_build_R -c 1 --csv --page raw --log-file gpurun_out/ncu_c3_$v.csv python tools/probe_transport.py --config C3 --vrls 400 --reps 1 > gpurun_out/ncu_c3_$v.log 2>&1
  tail -1 gpurun_out/ncu_c3_$v.log
done
