set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_gputests.txt 2>&1; tail -4 gpurun_out/r2b_gputests.txt
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 3000 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launch_list.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2b_under_ncu.log 2>&1
WL=raft N=4096 S=28 L=1 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2b_raft_bench python tools/prof_ml.py > gpurun_out/r2b_ncu_raft.log 2>&1; tail -1 gpurun_out/r2b_ncu_raft.log
TAG=r2b_small CASES=stream:16384,chaos:16384,signal:16384 GEOS=28:1,56:2,56:2k,112:4,112:4k python tools/gpu_ml_sweep.py > gpurun_out/r2b_sweep_small.txt 2>&1; cat gpurun_out/r2b_sweep_small.txt
TAG=r2b_etcd CASES=etcd:16384 GEOS=28:1,50:2,50:2k python tools/gpu_ml_sweep.py > gpurun_out/r2b_sweep_etcd.txt 2>&1; cat gpurun_out/r2b_sweep_etcd.txt
WL=stream N=16384 S=112 L=4 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2b_stream_free4 python tools/prof_ml.py > gpurun_out/r2b_ncu_stream.log 2>&1; tail -1 gpurun_out/r2b_ncu_stream.log
TAG=r2b_raftstd CASES=raft:8192,raft:16384 GEOS=28:1 python tools/gpu_ml_sweep.py > gpurun_out/r2b_sweep_raft_std.txt 2>&1; cat gpurun_out/r2b_sweep_raft_std.txt
MSIM_B200_LIB=$PWD/build_exp/libmsim_raft_compact.so TAG=r2b_raftc CASES=raft:4096,raft:8192,raft:16384 GEOS=28:1,56:1,56:2,56:2k CHECK=128 python tools/gpu_ml_sweep.py > gpurun_out/r2b_sweep_raft_compact.txt 2>&1; cat gpurun_out/r2b_sweep_raft_compact.txt
MSIM_B200_LIB=$PWD/build_exp/libmsim_raft_compact.so WL=raft N=16384 S=56 L=2 ncu --set full --clock-control none -k regex:msim_step -c 1 -o gpurun_out/r2b_raftc_free2 python tools/prof_ml.py > gpurun_out/r2b_ncu_raftc.log 2>&1; tail -1 gpurun_out/r2b_ncu_raftc.log
ls -la gpurun_out/*.ncu-rep
