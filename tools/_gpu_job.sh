set -x
mkdir -p gpurun_out
for lib in madsim_b200/libmsim_b200.so build_exp/lib_etcd_1024.so; do echo "== $lib"; MSIM_B200_LIB=$PWD/$lib TAG=r2h CASES=etcd:16384,etcd30:16384 GEOS=0:0,28:1,32:1 CHECK=128 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/r2h_etcd_threads.txt 2>&1; grep -E "==|True|False|error" gpurun_out/r2h_etcd_threads.txt | grep -v echo
for lib in madsim_b200/libmsim_b200.so build_exp/lib_election_1024.so; do echo "== $lib"; MSIM_B200_LIB=$PWD/$lib TAG=r2h CASES=election:16384 GEOS=0:0,32:1,64:2 CHECK=128 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/r2h_election.txt 2>&1; grep -E "==|True|False|error" gpurun_out/r2h_election.txt | grep -v echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'launches',d['gpu_launches'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'])
print(d['cpu_baseline'])
PY
