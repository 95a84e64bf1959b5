set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2y_gputests.txt; cat gpurun_out/r2y_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'])
PY
timeout 600 python tools/gpu_fuzz.py 31 60 > gpurun_out/r2y_gpu_fuzz.txt 2>&1; tail -1 gpurun_out/r2y_gpu_fuzz.txt | cut -c1-200
