set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -c 1500 gpurun_out/r2f_bench_n2.json; tail -5 gpurun_out/r2f_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --ref-seeds 1024 > gpurun_out/r2f_ref_n2.json 2> gpurun_out/r2f_ref_n2.err; tail -c 600 gpurun_out/r2f_ref_n2.json; tail -3 gpurun_out/r2f_ref_n2.err
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "jobs or cpp_host or sharded" 2>&1 | tail -3
