# development A/B on the GPU box: per-kernel times of one step for the product library, for other builds of it
# (muchsalsa_b200/libmsgpu_<tag>.so, see muchsalsa_b200/build.py --variant) and for the run-time variants
for w in ${WORKLOADS:-yeast}; do
  echo "== product $w"; python scripts/quick_step.py $w 2>&1 | tail -4
  for lib in muchsalsa_b200/libmsgpu_*.so; do
    [ -e "$lib" ] || continue
    echo "== $lib $w"; MSGPU_LIB=$PWD/$lib python scripts/quick_step.py $w 2>&1 | tail -4
  done
  echo "== MSGPU_EMIT=scalar $w"; MSGPU_EMIT=scalar python scripts/quick_step.py $w 2>&1 | tail -4
done
