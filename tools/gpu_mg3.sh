# usage: bash tools/gpu_mg3.sh <N> <tag> <modes...> -- on an N-GPU box: shard + parallel tests (N >= 2), then bench with GSB_TSDF_SHARD_MODE in <modes>
# (mode "r" = replicated TSDF, GSB_TSDF_SHARD=0)
N=${1:-2}; TAG=${2:-x}; shift; shift
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_tsdf_shard_gpu.py tests/test_parallel_gpu.py -m gpu -q -x > gpurun_out/mg3_tests_$TAG.log 2>&1; tail -5 gpurun_out/mg3_tests_$TAG.log
fi
for MODE in "$@"; do
if [ "$MODE" = "r" ]; then export GSB_TSDF_SHARD=0; else export GSB_TSDF_SHARD=1; export GSB_TSDF_SHARD_MODE=$MODE; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.json 2> gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.err
tail -c 400 gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.err | grep -v "^\*\|OMP_NUM" 
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.json").read().strip().splitlines()[-1])
    print("N=$N mode=$MODE fps %.1f e2e %.1f ms/step %.2f  full_run fps %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["full_run"]["fps"]))
    print("  psnr", d["config"]["quality"]["psnr_db"], "breakdown", {k: round(v, 2) for k, v in d["config"]["breakdown"].items()})
    r = d["roofline"]
    print("  " + ", ".join("%s %.0f" % (k.split("(")[0], v) for k, v in r.get("kernels_us", {}).items()))
    print("  fresh", {k: round(v, 1) for k, v in r["tsdf_integrate"]["fresh_frame_stages_us"].items()})
except Exception as e:
    print("no bench line:", e)
PY
done
