#!/bin/bash
# SASS evidence for the search kernel (k = 8, Gaussian f64-over-f32 epilogue): instruction mix and the lines around
# the bulk copies / mbarrier waits and the shared-memory candidate loop.   usage: tools/sass_excerpt.sh > profiles/rN_sass_query_kernel.txt
O=pyresample_b200/build/prs_query_float_8_3.o
S=/tmp/sass_q83.txt
cuobjdump -sass $O > $S
echo "# cuobjdump -sass $O  (sm_100a; query_kernel<float, 8, GAUSS_DF>)"
grep -m1 "Function :" $S
echo
echo "## instruction mix (count of mnemonic)"
for m in UBLKCP SYNCS.ARRIVE SYNCS.PHASECHK "SYNCS" FENCE "LDS.128" "LDS.64" "LDS " "LDG.E.128" "LDG" "STG" "BAR.SYNC" "SHFL" "REDUX" "VOTE" "ATOMS" "MUFU" "DADD\|DMUL\|DFMA" "FADD\|FMUL\|FFMA" "CALL" "BRA"; do
  printf "%-22s %6d\n" "$m" "$(grep -c "$m" $S)"
done
echo
echo "## bulk copy issue (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes) and mbarrier wait"
grep -n -B6 -A4 "UBLKCP" $S | cut -c1-150 | head -60
echo
grep -n -B3 -A6 "SYNCS.PHASECHK" $S | cut -c1-150 | head -40
echo
echo "## candidate loop from shared memory (scan_home_staged): 16-byte record reads"
grep -n -A14 -m2 "LDS.128" $S | cut -c1-150 | head -60
