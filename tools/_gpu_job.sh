set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2c_gputests.txt 2>&1; tail -4 gpurun_out/r2c_gputests.txt
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 2500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
WL=tonic N=8192 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2c_tonic_unary python tools/prof_ml.py > gpurun_out/r2c_ncu_tonic.log 2>&1; tail -1 gpurun_out/r2c_ncu_tonic.log
WL=etcd N=16384 ncu --set full --clock-control none --import-source on -k regex:msim_step -c 1 -o gpurun_out/r2c_etcd python tools/prof_ml.py > gpurun_out/r2c_ncu_etcd.log 2>&1; tail -1 gpurun_out/r2c_ncu_etcd.log
TAG=r2c_tonic CASES=tonic:32768 GEOS=0:0,25:1,50:2,28:1 CHECK=64 python tools/gpu_ml_sweep.py > gpurun_out/r2c_sweep_tonic.txt 2>&1; cat gpurun_out/r2c_sweep_tonic.txt
python tools/gpu_fuzz.py 5 60 > gpurun_out/r2c_gpu_fuzz.txt 2>&1; tail -3 gpurun_out/r2c_gpu_fuzz.txt
