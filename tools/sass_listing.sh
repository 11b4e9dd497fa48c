#!/bin/bash
# profiles/r02_sass_<kernel>.txt: per kernel of libgm2.so, the ptxas resource line, counts of the SASS mnemonics
# that matter on this path (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, STG.E.*128 / LDS.128 = vector stores /
# shared loads, SHF = funnel shifts, REDUX / SHFL / VOTE = warp primitives, ATOMS = shared atomics) and the first
# 60 SASS lines.  Run from the repo root after a build:  bash tools/sass_listing.sh
set -e
LIB=genome-minimizer-2_b200/libgm2.so
OUT=profiles
cuobjdump -sass $LIB > /tmp/gm2_all.sass
cuobjdump -res-usage $LIB 2>/dev/null > /tmp/gm2_res.txt || true
for pat in '_Z6k_emitILi1ELi3ELi1ELi2E:k_emit_72reg_flat2' '_Z6k_emitILi1ELi4ELi1ELi2E:k_emit_64reg_flat2' '_Z13k_emit_packed:k_emit_packed' '_Z6k_plan:k_plan' '_Z15k_keep_from_ids:k_keep_from_ids' '_Z17k_keep_from_probs:k_keep_from_probs' '_Z14k_scan_records:k_scan_records'; do
  sym=${pat%%:*}; name=${pat##*:}
  awk -v s="$sym" '/Function :/ {on = index($0, s) > 0} on {print}' /tmp/gm2_all.sass > /tmp/gm2_one.sass
  f=$OUT/r02_sass_$name.txt
  {
    echo "# $name — $(grep -m1 'Function :' /tmp/gm2_one.sass | sed 's/^\s*//')"
    echo "# cuobjdump -sass $LIB (sm_100a, CUDA 12.9); resource usage:"
    grep -A1 "Function $sym" /tmp/gm2_res.txt | tail -1 | sed 's/^/#   /' || true
    echo "# SASS lines: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/gm2_one.sass)"
    for m in UBLKCP SYNCS 'STG\.E[A-Z.]*\.128' 'STG\.E' 'LDS\.128' 'LDS' 'LDG' 'SHF\.' REDUX SHFL VOTE 'ATOMS' POPC 'BAR\.SYNC' 'HMMA|IMMA|UTCMMA|UTMALDG'; do
      echo "#   $(printf '%-28s' "$m") $(grep -cE "\b($m)" /tmp/gm2_one.sass)"
    done
    echo "# ---- first 60 instructions"
    grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/gm2_one.sass | head -60 | sed 's#/\* 0x[0-9a-f]\{16\} \*/##; s/ *$//'
  } > $f
  echo "$f: $(wc -l < $f) lines"
done
