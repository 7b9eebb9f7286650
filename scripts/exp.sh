# quick A/B of kernel variants on the yeast workload (development only)
mkdir -p gpurun_out
for v in "MSGPU_EXP_RANK=v MSGPU_EXP_EMIT=0" "MSGPU_EXP_RANK=m MSGPU_EXP_EMIT=1" "MSGPU_EXP_RANK=v MSGPU_EXP_EMIT=2" "MSGPU_EXP_RANK=v MSGPU_EXP_EMIT=3"; do
  echo "== $v"
  env $v python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from muchsalsa_b200 import synth
from muchsalsa_b200.construct import GraphBuilder
paf = synth.generate_paf("yeast")
dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
with GraphBuilder(0) as gb:
    for _ in range(3): gb.build_from_device(dev.data_ptr(), len(paf))
    gb.profile(True)
    gb.build_from_device(dev.data_ptr(), len(paf))
    ks = gb.kernel_stats()
    print({k: round(v[1], 3) for k, v in ks.items() if k.startswith("pair_")}, gb.counts()["n_edge_matches"])
PY
done
