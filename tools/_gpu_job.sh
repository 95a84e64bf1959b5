set -x
mkdir -p gpurun_out
for lib in madsim_b200/libmsim_b200.so build_exp/lib_etcdkv_coldcheck.so build_exp/lib_etcdkv_inlinecheck.so build_exp/lib_etcdfull_inlinecheck.so; do echo "== $lib"; MSIM_B200_LIB=$PWD/$lib TAG=r2e CASES=etcd:16384,etcd30:16384 GEOS=28:1 CHECK=128 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/r2e_etcd_variants.txt 2>&1; grep -E "==|True|False" gpurun_out/r2e_etcd_variants.txt
python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'])
for c in d.get('configs',[]): print(c['config'],round(c['ms'],3),'ms',round(c['value']/1e9,3),'Gev/s',c['parity_vs_oracle'],c['geometry'])
PY
