# usage: gpurun --gpus N -- bash tools/gpu_scale.sh N <tag> -- the bench exactly as the driver launches it for N > 1
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
echo rc=$?; tail -c 1500 gpurun_out/bench_${N}gpu_$TAG.err; cut -c 1-400 gpurun_out/bench_${N}gpu_$TAG.json
