#!/bin/bash
# Per-kernel SASS evidence of the Blackwell-native paths (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
# UTMALDG/UTMASTG = TMA loads / stores, LDGSTS = cp.async).  Usage: bash scripts/sass_counts.sh > profiles/rNN_sass_counts.txt
LIB=stove_b200/csrc/libstove_b200.so
echo "# cuobjdump -sass $LIB : instruction counts per kernel (only kernels with at least one of the mnemonics)"
cuobjdump -sass $LIB | awk '
/Function :/ { name=$3; next }
/UTC[A-Z]*MMA/ { mma[name]++ }
/LDTM/ { ldtm[name]++ }
/UTMALDG/ { tmal[name]++ }
/UTMASTG/ { tmas[name]++ }
/UTCBAR|UTCCP/ { utc[name]++ }
/LDGSTS/ { cpa[name]++ }
/SYNCS/ { syncs[name]++ }
END {
  printf "%-90s %8s %6s %8s %8s %7s %7s\n", "kernel", "UTC*MMA", "LDTM", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS"
  for (n in mma) seen[n]=1; for (n in tmal) seen[n]=1; for (n in cpa) seen[n]=1
  for (n in seen) printf "%-90s %8d %6d %8d %8d %7d %7d\n", substr(n,1,90), mma[n], ldtm[n], tmal[n], tmas[n], cpa[n], syncs[n]
}' | sort
