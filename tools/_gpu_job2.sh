set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
tail -3 gpurun_out/r2q_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_n2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'])
PY
