#!/bin/bash
# retention x flat-run-bytes sweep (device-side only); output: gpurun_out/sweep_flat.log
mkdir -p gpurun_out
: > gpurun_out/sweep_flat.log
for ret in 0.1 0.2 0.3 0.5 0.7 0.9; do
  for flat in 0 384 640 1024 1048576; do
    line=$(python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --retention $ret --flat-run-bytes $flat 2>/dev/null | tail -1)
    python - "$ret" "$flat" "$line" >> gpurun_out/sweep_flat.log <<'PY'
import json, sys
d = json.loads(sys.argv[3])
print(f"retention {sys.argv[1]} flat {sys.argv[2]:>8}: value {d['value']:9.1f} Gbp/s  ms/step {d['ms_per_step']:.3f}  k_emit_ms {d['roofline'].get('kernel_ms', 0):.3f}  frac {d['roofline']['frac']:.3f}  verify {d.get('verify')}")
PY
  done
done
cat gpurun_out/sweep_flat.log
