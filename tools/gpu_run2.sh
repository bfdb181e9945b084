python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1500 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | cut -c 1-1500
python bench.py --steps 10 --warmup 3 --track 1 --no-cpu-baseline > gpurun_out/bench_track.json 2> gpurun_out/bench_track.err
tail -c 800 gpurun_out/bench_track.err; cat gpurun_out/bench_track.json | cut -c 1-1200
