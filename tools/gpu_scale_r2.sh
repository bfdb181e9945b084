# usage: bash tools/gpu_scale_r2.sh <N> <tag> [peer|nccl] [steps] [warmup] -- the driver's multi-GPU launch of bench.py (torchrun, one rank per GPU)
N=${1:-2}; TAG=${2:-x}; EX=${3:-peer}; K=${4:-20}; W=${5:-5}
mkdir -p gpurun_out
GSB_EXCHANGE=$EX timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $K --warmup $W --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
tail -c 800 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N $EX fps %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("quality", {k: d["config"]["quality"][k] for k in ("psnr_db", "psnr_tsdf_only_db")})
    print("full_run", {k: d["config"]["full_run"][k] for k in ("fps", "ms_per_step_p50", "ms_per_step_last", "gaussians_final")})
    r = d["roofline"]
    for k, v in r.get("kernels_us", {}).items(): print("  %-40s %8.1f us" % (k, v))
except Exception as e:
    print("no bench line:", e)
PY
