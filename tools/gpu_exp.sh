# usage: EXP="VAR=val" bash tools/gpu_exp.sh <tag> -- quick TSDF parity (hard timeout: a hung kernel must not hold the box), then bench base / exp
TAG=${1:-x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tsdf_parity_gpu.py ${TESTS} -m gpu -q -x > gpurun_out/exp2_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/exp2_tests_$TAG.log
for V in base exp; do
  if [ $V = exp ]; then [ -z "$EXP" ] && break; export $EXP; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$V.json 2> gpurun_out/bench_${TAG}_$V.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$V.json"))
print("$V fps %.1f e2e %.1f psnr %.3f" % (d["value"], d["e2e"]["value"], d["config"]["quality"]["psnr_db"]), d["config"]["ms_per_step_each"])
for k,v in d["roofline"]["kernels_us"].items(): print("  %-40s %8.1f us" % (k, v))
PY
done
