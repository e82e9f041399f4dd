#!/bin/bash
# Round 2, call M: PDL modes, whole-sampler graph on/off, batch sizes (fused prep on).
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-torch-reference --config4 0 $EXTRA > gpurun_out/m_$label.json 2> gpurun_out/m_$label.err
  python - <<PY
import json
d=json.load(open("gpurun_out/m_$label.json"))
print("$label: value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"launches",d["gpu_launches"])
PY
}
EXTRA=""
run pdl2 FLOWSE_PDL=2
run pdl3 FLOWSE_PDL=3
run pdl0 FLOWSE_PDL=0
run wg0 FLOWSE_WHOLE_GRAPH=0
run wg1 FLOWSE_WHOLE_GRAPH=1
EXTRA="--batch 4"; run b4 FLOWSE_PDL=2
EXTRA="--batch 8"; run b8 FLOWSE_PDL=2
EXTRA="--batch 8"; run b8_unfused FLOWSE_FUSE_PREP=0
