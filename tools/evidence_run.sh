#!/bin/bash
# One-GPU evidence run (under gpurun): GPU tests, the default bench line, the reference arm, the ncu
# launch list of the bench command, and full ncu captures of the emit and plan kernels.  Everything
# lands in gpurun_out/ev2/ ; the summaries worth keeping are copied into profiles/ by hand.
mkdir -p gpurun_out/ev2
cd "$(dirname "$0")/.."
O=gpurun_out/ev2
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>&1; echo "reference rc=$?"
LIGHT="--no-e2e --no-cpu-baseline --no-dropin --no-sharded --no-sweep --no-c5"
CMD="python bench.py --steps 2 --warmup 3 $LIGHT"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
# kernels before the timed region: 2 sizing plans x 3 + 4 warm-up steps x 4 + 3 solo passes x 4 = 34
CMD2="python bench.py --steps 1 --warmup 3 $LIGHT --verify 0"
$CMD2 > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_plan|k_keep|k_scan" -s 34 -c 4 -o $O/prof_all -f $CMD2 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
