set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload ecoli --steps 5 --warmup 3 > gpurun_out/bench_ecoli_v5.json 2> gpurun_out/bench_ecoli_v5.err; echo rc=$?
python bench.py --workload yeast --steps 3 --warmup 3 > gpurun_out/bench_yeast_v5.json 2> gpurun_out/bench_yeast_v5.err; echo rc=$?
tail -3 gpurun_out/bench_yeast_v5.err
