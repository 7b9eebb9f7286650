# round-end measurement on one GPU: parity tests, the four single-GPU bench lines, the ncu launch list of the bench
# command and one full ncu capture of the two dominant kernels.  Usage: bash scripts/final_run.sh TAG
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; tail -2 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_yeast.json 2> $O/${TAG}_yeast.err; echo yeast rc=$?
python bench.py --workload ecoli > $O/${TAG}_ecoli.json 2> $O/${TAG}_ecoli.err; echo ecoli rc=$?
python bench.py --workload thaliana --steps 3 --warmup 3 > $O/${TAG}_thaliana.json 2> $O/${TAG}_thaliana.err; echo thaliana rc=$?
python bench.py --workload human --genome-scale 0.015625 --chunk-tiles 4096 --steps 2 --warmup 3 > $O/${TAG}_human64.json 2> $O/${TAG}_human64.err; echo human64 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-sequence > $O/${TAG}_ncu_bench.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on --kernel-name 'regex:pair_tiles_merge_kernel|pair_emit_rows_bulk_kernel' \
  --launch-skip 4 --launch-count 2 -f -o $O/${TAG}_pair python scripts/profile_step.py yeast 4 > $O/${TAG}_ncu_pair.log 2>&1; echo capture rc=$?
