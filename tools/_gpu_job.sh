set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "golden or escalat or large or capacity" 2>&1 | tail -6 | tee gpurun_out/r2aa_gputests.txt
