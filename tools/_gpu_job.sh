set -x
mkdir -p gpurun_out
T=r2x2
run() { # wl cases geos libs...
  wl=$1; cases=$2; geos=$3; shift 3
  for lib in "$@"; do echo "== $lib"; MSIM_B200_LIB=$PWD/build_exp/lib_$lib.so TAG=$T CASES=$cases GEOS=$geos CHECK=64 python tools/gpu_ml_sweep.py 2>&1 | grep -v "^#"; done > gpurun_out/${T}_$wl.txt 2>&1
  grep -E "==|True|False|error" gpurun_out/${T}_$wl.txt | grep -v echo
}
run tonic tonic:32768 0:0 tonic_base tonic_h1 tonic_now1 tonic_full tonic_bind0
