#!/bin/bash
# retention x tile-bytes sweep (device-side only); output: gpurun_out/sweep_tile.log
mkdir -p gpurun_out
: > gpurun_out/sweep_tile.log
for ret in ${RETS:-0.1 0.2 0.3 0.4}; do
  for tile in ${TILES:-16384 24576 32768 49152}; do
    line=$(python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --retention $ret --tile-bytes $tile 2>/dev/null | tail -1)
    python - "$ret" "$tile" "$line" >> gpurun_out/sweep_tile.log <<'PY'
import json, sys
d = json.loads(sys.argv[3])
print(f"retention {sys.argv[1]} tile {sys.argv[2]:>6}: value {d['value']:9.1f} Gbp/s  ms/step {d['ms_per_step']:.3f}  k_emit_ms {d['roofline'].get('kernel_ms', 0):.3f}  plan_ms {d['roofline'].get('plan_ms', 0):.3f}  frac {d['roofline']['frac']:.3f}  verify {d.get('verify')}")
PY
  done
done
cat gpurun_out/sweep_tile.log
