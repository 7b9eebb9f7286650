# development A/B on the GPU box: programmatic dependent launches on / off (MSGPU_PDL), wall time per construction
for w in ${WORKLOADS:-yeast ecoli}; do
  for v in 1 0; do
    echo "== MSGPU_PDL=$v $w"; MSGPU_PDL=$v python scripts/quick_step.py $w 2>&1 | tail -4 | cut -c1-330
  done
done
