#!/bin/bash
# SASS evidence that the tensor-core / TMA kernels of libsd_b200.so really use tcgen05 + TMEM + TMA on sm_100a
# (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk,
# SYNCS = mbarrier, UTCBAR = tcgen05.commit).   usage: tools/sass_evidence.sh > profiles/r02_sass_evidence.md
so=${1:-soccerdiffusion_b200/libsd_b200.so}
echo "# SASS evidence — tcgen05 / TMEM / TMA instructions per kernel of \`$so\` (cuobjdump -sass, sm_100a)"
echo
echo "| kernel | UTCHMMA (tcgen05.mma) | UTMALDG (TMA tensor load) | UBLKCP (TMA bulk copy) | LDTM (tcgen05.ld) | UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | R2UR.BROADCAST (per-instruction issue loop; 0 = operands in uniform registers) |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|"
cuobjdump -sass "$so" 2>/dev/null | awk '
/Function :/ {fn=$3}
/UTCHMMA|UTCQMMA|UTCOMMA/ {mma[fn]++; seen[fn]=1}
/UTMALDG/ {tma[fn]++; seen[fn]=1}
/UBLKCP/ {blk[fn]++; seen[fn]=1}
/LDTM/ {ldtm[fn]++; seen[fn]=1}
/UTCBAR/ {bar[fn]++}
/SYNCS/ {syncs[fn]++}
/R2UR.BROADCAST/ {bc[fn]++}
END {for (f in seen) printf "%s %d %d %d %d %d %d %d\n", f, mma[f]+0, tma[f]+0, blk[f]+0, ldtm[f]+0, bar[f]+0, syncs[f]+0, bc[f]+0}' |
while read -r f a b c d e g h; do
  name=$(echo "$f" | c++filt | sed 's/(anonymous namespace):://g; s/(.*//' | cut -c1-80)
  echo "| \`$name\` | $a | $b | $c | $d | $e | $g | $h |"
done | sort | awk '!/gemm_tc_kernel/ {print} /gemm_tc_kernel/ {if (!g++) print; n++} END {if (n > 1) print "| … " n - 1 " more `gemm_tc_kernel<A_KM, B_KN, EPI, MNMAJ>` instantiations with the same counts | | | | | | | |"}'
