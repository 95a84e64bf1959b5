set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2af_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
