set -x
mkdir -p gpurun_out
T=r2l
run() { # wl cases geos libs...
  wl=$1; cases=$2; geos=$3; shift 3
  for lib in "$@"; do echo "== $lib"; MSIM_B200_LIB=$PWD/build_exp/lib_$lib.so TAG=$T CASES=$cases GEOS=$geos CHECK=128 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/${T}_$wl.txt 2>&1
  grep -E "==|True|False|error" gpurun_out/${T}_$wl.txt | grep -v echo
}
run raft raft:4096 0:0 raft_h0 raft_rs
run etcd etcd:16384,etcd30:16384 0:0 etcd_h16 etcd_rs
run tonic tonic:32768 0:0 tonic_h0 tonic_rs
run election election:16384 0:0 election_h16 election_rs
run stream stream:16384 0:0 stream_h16 stream_rs
for v in etcd_h16 etcd_rs tonic_rs; do
  w=${v%%_*}; n=16384; [ $w = tonic ] && n=32768
  MSIM_B200_LIB=$PWD/build_exp/lib_$v.so WL=$w N=$n timeout 300 ncu --metrics smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__inst_executed.avg.per_cycle_elapsed -k regex:msim_step_kernel -c 1 python tools/prof_ml.py 2>&1 | grep -E "inst_executed|icc_request" | sed "s/^/$v /"
done > gpurun_out/${T}_inst.txt 2>&1
cat gpurun_out/${T}_inst.txt
