# usage: bash tools/gpu_mg.sh <N> <tag> -- multi-GPU checks on an N-GPU box: sharded-TSDF / parallel tests, then bench at N with the voxel
# hash sharded and (A/B) replicated
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tsdf_shard_gpu.py tests/test_parallel_gpu.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/mg_tests_$TAG.log
for SH in 1 0; do
GSB_TSDF_SHARD=$SH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}_shard$SH.json 2> gpurun_out/bench_${N}gpu_${TAG}_shard$SH.err
tail -c 600 gpurun_out/bench_${N}gpu_${TAG}_shard$SH.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_${TAG}_shard$SH.json").read().strip().splitlines()[-1])
    print("N=$N shard=$SH fps %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("quality", {k: d["config"]["quality"][k] for k in ("psnr_db", "psnr_tsdf_only_db")})
    print("breakdown", d["config"]["breakdown"])
    r = d["roofline"]
    for k, v in r.get("kernels_us", {}).items(): print("  %-40s %8.1f us" % (k, v))
    print("fresh", r["tsdf_integrate"]["fresh_frame_stages_us"])
except Exception as e:
    print("no bench line:", e)
PY
done
