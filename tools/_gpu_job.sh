set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_gputests.txt 2>&1; tail -3 gpurun_out/r2a_gputests.txt
TAG=r2a_small CASES=stream:16384,chaos:16384,signal:16384 GEOS=28:1,28:2,56:2,112:4,140:5,196:7,224:8 python tools/gpu_ml_sweep.py > gpurun_out/r2a_ml_sweep_small.txt 2>&1; cat gpurun_out/r2a_ml_sweep_small.txt
TAG=r2a_etcd CASES=etcd:16384,etcd30:16384 GEOS=28:1,50:2,42:2,28:2 python tools/gpu_ml_sweep.py > gpurun_out/r2a_ml_sweep_etcd.txt 2>&1; cat gpurun_out/r2a_ml_sweep_etcd.txt
TAG=r2a_raft CASES=raft2s:4096,raftf:1024 GEOS=28:1,28:2,28:4,28:7 python tools/gpu_ml_sweep.py > gpurun_out/r2a_ml_sweep_raft.txt 2>&1; cat gpurun_out/r2a_ml_sweep_raft.txt
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
WL=stream N=16384 S=112 L=4 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2a_stream_ml4 python tools/prof_ml.py > gpurun_out/r2a_ncu_stream.log 2>&1; tail -2 gpurun_out/r2a_ncu_stream.log
WL=etcd N=16384 S=50 L=2 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2a_etcd_ml2 python tools/prof_ml.py > gpurun_out/r2a_ncu_etcd.log 2>&1; tail -2 gpurun_out/r2a_ncu_etcd.log
WL=stream N=16384 S=28 L=1 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2a_stream_l1 python tools/prof_ml.py > gpurun_out/r2a_ncu_stream1.log 2>&1; tail -2 gpurun_out/r2a_ncu_stream1.log
ls -la gpurun_out/*.ncu-rep
