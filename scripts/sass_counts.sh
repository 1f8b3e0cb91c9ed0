#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
# UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF (TMA), HMMA (legacy mma.sync, expected 0).
# Usage: scripts/sass_counts.sh > profiles/r2_sass_counts.txt
LIB=${1:-neural-audio-fp_b200/csrc/libnafp.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
{ for (i = 1; i <= NF; i++) if ($i ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCBAR|HMMA|FFMA2|SYNCS)/) { split($i, a, "."); c[fn" "a[1]]++ } }
END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{ n=$NF; m=$(NF-1); $NF=""; $(NF-1)=""; printf "%-8s %5d  %s\n", m, n, $0 }' | sort -k3,3 -k1,1
