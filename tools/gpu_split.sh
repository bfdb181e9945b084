# usage: bash tools/gpu_split.sh <N> <tag> [variants: 1 0] -- functional-split test (N >= 4) + bench at N with GSB_SPLIT in the given variants
N=${1:-4}; TAG=${2:-x}; shift; shift
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 600 python -m pytest tests/test_parallel_gpu.py -k functional_split -m gpu -q -x > gpurun_out/split_tests_$TAG.log 2>&1; tail -5 gpurun_out/split_tests_$TAG.log; grep -n "Error\|error\|assert" gpurun_out/split_tests_$TAG.log | head -20
fi
for SP in "$@"; do
GSB_SPLIT=$SP timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}_split$SP.json 2> gpurun_out/bench_${N}gpu_${TAG}_split$SP.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_${N}gpu_${TAG}_split$SP.err | tail -12
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_${TAG}_split$SP.json").read().strip().splitlines()[-1])
    print("N=$N split=$SP fps %.1f e2e %.1f ms/step %.2f  full_run fps %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["full_run"]["fps"]))
    print("  psnr", d["config"]["quality"]["psnr_db"], d["config"]["quality"]["psnr_tsdf_only_db"], "gaussians", d["config"]["gaussians"], "overflow", d["config"]["overflow_flags"], "mbox", d["config"].get("mailbox_errors"))
    print("  breakdown", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["config"]["breakdown"].items() if k != "note"})
    r = d["roofline"]
    print("  " + ", ".join("%s %.0f" % (k.split("(")[0], v) for k, v in r.get("kernels_us", {}).items()))
except Exception as e:
    print("no bench line:", e)
PY
done
