#!/bin/bash
# SASS evidence of the shipped library: per-kernel counts of the tcgen05 / TMA / TMEM / mbarrier opcodes (cuobjdump -sass).
set -e
cd "$(dirname "$0")/.."
OUT=profiles/r02_sass_summary.txt
SO=alpha_zero_b200/libaz_b200.so
{
  echo "# cuobjdump -sass $SO ($(date -u +%FT%TZ), $(sha1sum $SO | cut -c1-12))"
  echo "# opcode counts per kernel: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load, LDTM = tcgen05.ld,"
  echo "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTMAPF = TMA prefetch, REDUX = redux.sync"
  cuobjdump -sass $SO | awk '
    /Function :/ { fn=$3 }
    /UTCHMMA|UTMALDG|LDTM|UTCBAR|UTMAPF|SYNCS|UTCATOMSWS|REDUX|HMMA|UTMASTG/ {
      op=$0; sub(/^[^A-Z@]*/, "", op); sub(/^@!?U?P[0-9T]+ /, "", op); split(op, a, /[ ;]/); key=fn " " a[1]; c[key]++ }
    END { for (k in c) print c[k], k }' | sort -k2,2 -k1,1nr | awk '{printf "%-90s %-40s %6d\n", $2, $3, $1}'
} > $OUT
wc -l $OUT
