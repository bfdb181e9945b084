# builds gs_raster.cu variants on the box and times the train-step kernels for each (nvcc is present in the image)
cd gps_slam_b200
for V in "0 1" "0 3" "1 1" "1 3"; do
  set -- $V
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off -I ../include -DBWD_PREFETCH_CURSOR=$1 -DBWD_GRID_MULT=$2 -c csrc/gs_raster.cu -o build/gs_raster.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libgpsslam_b200.so build/tsdf_kernels.o build/tsdf_engine.o build/icp_kernels.o build/gs_project.o build/gs_raster.o build/gs_spawn.o build/gs_staged.o build/gs_engine.o -lcudart
  cd ..
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/var.json"))
k=d["roofline"]["kernels_us"]
print("prefetch=$1 gridmult=$2: fps %.1f bwd %.1f fwd %.1f step %.1f" % (d["value"], k["gs_raster_bwd"], k["gs_raster_fwd_train"], k["gs_train_step(7 kernels, no flush)"]))
PY
  cd gps_slam_b200
done
