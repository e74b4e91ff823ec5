#!/bin/bash
# Which tensor / TMA / TMEM instructions each object of librg_b200.so holds (cuobjdump -sass, sm_100a) and the register /
# spill figures of every kernel (cuobjdump -res-usage): the static evidence beside the ncu captures under profiles/.
#   tools/sass_summary.sh > profiles/sass_r2_final.txt
cd "$(dirname "$0")/.."
echo "# mnemonic counts per object (UTCQMMA = tcgen05.mma kind::f8f6f4, UTCIMMA = kind::i8, UTCHMMA = kind::tf32 / f16, UTMALDG = TMA load,"
echo "# UTMASTG = TMA store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, DMMA = FP64 mma, SYNCS = mbarrier)"
for o in regenie_b200/_obj/*.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  line=""
  for m in UTCQMMA UTCIMMA UTCHMMA UTMALDG UTMASTG UTMAPF LDTM STTM UTCBAR DMMA SYNCS; do
    n=$(grep -c "$m" <<<"$s")
    [ "$n" -gt 0 ] && line="$line $m=$n"
  done
  printf "%-24s%s\n" "$(basename "$o")" "$line"
done
echo
echo "# registers, shared memory, spills per kernel (cuobjdump -res-usage)"
for o in regenie_b200/_obj/*.o; do
  cuobjdump -res-usage "$o" 2>/dev/null | awk -v f="$(basename "$o")" '
    /Function/ { name=$2; sub(/:$/, "", name) }
    /REG:/ { reg=""; sh=""; st=""; loc="";
             for (i = 1; i <= NF; ++i) { if ($i ~ /^REG:/) reg=$i; if ($i ~ /^SHARED:/) sh=$i; if ($i ~ /^STACK:/) st=$i; if ($i ~ /^LOCAL:/) loc=$i }
             printf "%-22s %-90s %s %s %s %s\n", f, substr(name, 1, 90), reg, sh, st, loc }'
done | c++filt 2>/dev/null
