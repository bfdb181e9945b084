# usage: bash tools/gpu_mg2.sh <N> <tag> -- shard-mode A/B on an N-GPU box
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tsdf_shard_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/mg2_tests_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 tests/mp_slam_worker.py /tmp/w2t.json 21 0.5 1 > gpurun_out/track2_$TAG.log 2>&1; tail -c 3000 gpurun_out/track2_$TAG.log | grep -v "^\s*$" | tail -30
for MODE in 1; do
GSB_TSDF_SHARD_MODE=$MODE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.json 2> gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.err
tail -c 600 gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_${TAG}_mode$MODE.json").read().strip().splitlines()[-1])
    print("N=$N mode=$MODE fps %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("quality", {k: d["config"]["quality"][k] for k in ("psnr_db", "psnr_tsdf_only_db")})
    print("breakdown", d["config"]["breakdown"])
    r = d["roofline"]
    for k, v in r.get("kernels_us", {}).items(): print("  %-40s %8.1f us" % (k, v))
    print("fresh", r["tsdf_integrate"]["fresh_frame_stages_us"])
except Exception as e:
    print("no bench line:", e)
PY
done
