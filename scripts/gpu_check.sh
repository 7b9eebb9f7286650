# GPU check used during development: parity tests, then the two single-GPU benches.  Usage: bash scripts/gpu_check.sh TAG
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload ecoli --steps 5 --warmup 3 > gpurun_out/bench_ecoli_$TAG.json 2> gpurun_out/bench_ecoli_$TAG.err; echo rc=$?
python bench.py --workload yeast --steps 3 --warmup 3 > gpurun_out/bench_yeast_$TAG.json 2> gpurun_out/bench_yeast_$TAG.err; echo rc=$?
tail -3 gpurun_out/bench_yeast_$TAG.err
