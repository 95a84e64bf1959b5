set -x
mkdir -p gpurun_out
T=r2x
run() { # wl cases geos libs...
  wl=$1; cases=$2; geos=$3; shift 3
  for lib in "$@"; do echo "== $lib"; MSIM_B200_LIB=$PWD/build_exp/lib_$lib.so TAG=$T CASES=$cases GEOS=$geos CHECK=64 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/${T}_$wl.txt 2>&1
  grep -E "==|True|False|error" gpurun_out/${T}_$wl.txt | grep -v echo
}
run etcd etcd:16384,etcd30:16384 0:0 etcd_base etcd_spawn0 etcd_finish1 etcd_ch1 etcd_h0
run tonic tonic:32768 0:0 tonic_base tonic_spawn1 tonic_finish1 tonic_ch1 tonic_now1 tonic_h1
run election election:16384 0:0 election_base election_spawn1 election_finish1 election_ch1 election_now1 election_h0
run raft raft:4096 0:0 raft_base raft_spawn1 raft_finish1
