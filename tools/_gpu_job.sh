set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2v_gputests.txt; cat gpurun_out/r2v_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
