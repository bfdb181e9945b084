python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -c 2000 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
