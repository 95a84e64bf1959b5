set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2r_gputests.txt; cat gpurun_out/r2r_gputests.txt
# launch list of the bench command (cold-cache, serialised: only the kernel's SHARE of the step is meaningful)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/r2r_b_under_ncu.log 2>&1
tail -2 gpurun_out/r2r_b_under_ncu.log | cut -c1-300
# full capture of the dominant kernel on the bench configuration
WL=raft N=4096 timeout 600 ncu --set full --import-source on --clock-control none -k regex:msim_step_kernel -c 1 -f -o gpurun_out/r2r_raft_bench python tools/prof_ml.py > gpurun_out/r2r_ncu_raft.log 2>&1; tail -2 gpurun_out/r2r_ncu_raft.log
WL=tonic N=32768 timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy --metrics sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:msim_step_kernel -c 1 -f -o gpurun_out/r2r_tonic python tools/prof_ml.py > gpurun_out/r2r_ncu_tonic.log 2>&1; tail -2 gpurun_out/r2r_ncu_tonic.log
