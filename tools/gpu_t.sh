# usage: bash tools/gpu_t.sh <logtag> <pytest args...> -- run the given GPU tests, log to gpurun_out/<logtag>.log
TAG=$1; shift
mkdir -p gpurun_out
timeout 1400 python -m pytest "$@" -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/$TAG.log
