# development A/B on the GPU box: the pipelined emission (pair_fused.cu, pairing_pipelined) by number of chunks
for w in ${WORKLOADS:-yeast}; do
  for pc in ${CHUNKS:-0 2 4 6 8 12 16}; do
    echo "== MSGPU_PIPE_CHUNKS=$pc $w"; MSGPU_PIPE_CHUNKS=$pc python scripts/quick_step.py $w 2>&1 | tail -4 | cut -c1-400
  done
  for lib in muchsalsa_b200/libmsgpu_*.so; do
    [ -e "$lib" ] || continue
    echo "== $lib $w"; MSGPU_LIB=$PWD/$lib python scripts/quick_step.py $w 2>&1 | tail -4 | cut -c1-400
  done
done
