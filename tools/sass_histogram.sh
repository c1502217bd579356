#!/bin/bash
# SASS opcode histogram of the shipped library: proof that the hot kernels are tcgen05 / TMEM / TMA code
# (B200_PROFILING.md mnemonics).  usage: tools/sass_histogram.sh > profiles/rNN_sass_opcodes.txt
LIB=${1:-nind_denoise_b200/libnind_b200.so}
S=$(mktemp); cuobjdump -sass $LIB > $S
echo "library: $LIB  kernel sources: $(python -c 'from nind_denoise_b200 import _build; print(_build.build_key()[:16])')"
echo "kernels: $(grep -c 'Function :' $S)"
for op in UTCHMMA 'UTCHMMA.2CTA' LDTM UTMALDG UTMASTG UTCBAR UTCATOMSWS 'SYNCS.ARRIVE' 'SYNCS.PHASECHK' UBLKCP UTMACCTL UTMACMDFLUSH 'FENCE.VIEW.ASYNC' '[^C]HMMA' '[^C]IMMA'; do
  printf "%-18s %6d\n" "$op" "$(grep -c "$op" $S)"
done
echo "per kernel (UTCHMMA / LDTM / UTMALDG / UTMASTG):"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} END {for (n in a) printf "  %-90s %4d %4d %4d %4d\n", n, a[n], b[n], c[n], d[n]}' $S | sort
rm -f $S
