#!/bin/bash
# retention x (tile bytes, run-table entries) sweep (device-side only); output: gpurun_out/sweep_cfg.log
# CFGS is a list of tile:runtable pairs.
mkdir -p gpurun_out
: > gpurun_out/sweep_cfg.log
for ret in ${RETS:-0.1 0.3 0.5 0.9}; do
  for cfg in ${CFGS:-49152:64 49152:32 45056:48 40960:64}; do
    tile=${cfg%%:*}; rt=${cfg##*:}
    line=$(python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --retention $ret --tile-bytes $tile --run-table $rt ${EXTRA:-} 2>/dev/null | tail -1)
    python - "$ret" "$cfg" "$line" >> gpurun_out/sweep_cfg.log <<'PY'
import json, sys
d = json.loads(sys.argv[3])
print(f"retention {sys.argv[1]} tile:rt {sys.argv[2]:>9}: value {d['value']:9.1f} Gbp/s  ms/step {d['ms_per_step']:.3f}  k_emit_ms {d['roofline'].get('kernel_ms', 0):.3f}  plan_ms {d['roofline'].get('plan_ms', 0):.3f}  frac {d['roofline']['frac']:.3f}  verify {d.get('verify')}")
PY
  done
done
cat gpurun_out/sweep_cfg.log
