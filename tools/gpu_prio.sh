for V in "0 0" "-1 0" "0 -1"; do
  set -- $V
  GSB_TSDF_STREAM_PRIORITY=$1 GSB_MAIN_STREAM_PRIORITY=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/prio.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/prio.json"))
print("tsdf prio $1 main prio $2: fps %.1f e2e %.1f clocks %s" % (d["value"], d["e2e"]["value"], d["clocks"]))
PY
done
