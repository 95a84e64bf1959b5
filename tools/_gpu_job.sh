set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2ab_gputests.txt; cat gpurun_out/r2ab_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ab_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ab_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-configs > gpurun_out/r2ab_b_under_ncu.log 2>&1
WL=raft N=4096 timeout 600 ncu --set full --import-source on --clock-control none -k regex:msim_step_kernel -c 1 -f -o gpurun_out/r2ab_raft_bench python tools/prof_ml.py > gpurun_out/r2ab_ncu_raft.log 2>&1; tail -2 gpurun_out/r2ab_ncu_raft.log
