# usage: bash tools/gpu_exp.sh <tag> -- TSDF/ICP/golden parity + bench with and without the experiment switch given in $EXP (e.g. EXP="GSB_RAYCAST_MINB=8")
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tsdf_parity_gpu.py tests/test_icp_parity_gpu.py tests/test_golden_gpu.py tests/test_checkpoint_gpu.py -m gpu -q -x 2>&1 | tail -4
for V in base exp; do
  if [ $V = exp ]; then [ -z "$EXP" ] && break; export $EXP; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$V.json 2> gpurun_out/bench_${TAG}_$V.err; tail -c 300 gpurun_out/bench_${TAG}_$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$V.json"))
print("$V fps %.1f e2e %.1f psnr %.3f launches %d" % (d["value"], d["e2e"]["value"], d["config"]["quality"]["psnr_db"], d["gpu_launches"]), d["config"]["ms_per_step_each"])
for k,v in d["roofline"]["kernels_us"].items(): print("  %-40s %8.1f us" % (k, v))
PY
done
