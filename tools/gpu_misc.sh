# tracking-mode bench (config 3), reference arm, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.err; cut -c 1-600 gpurun_out/bench_reference.json
python bench.py --steps 10 --warmup 3 --track 1 --no-cpu-baseline > gpurun_out/bench_track.json 2> gpurun_out/bench_track.err; tail -c 600 gpurun_out/bench_track.err; cut -c 1-900 gpurun_out/bench_track.json
