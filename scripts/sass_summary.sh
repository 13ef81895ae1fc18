#!/bin/bash
# SASS evidence of the Blackwell-native paths in the built library -> profiles/r2_sass_summary.txt
cd "$(dirname "$0")/.."
SO=yoloret_b200/libyoloret_b200.so
cuobjdump -sass $SO > /tmp/yr_all.sass 2>/dev/null
{
echo "# SASS evidence of the Blackwell-native paths in $SO (round 2, final build)"
echo "# cuobjdump -sass $SO | grep -c <mnemonic>   (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)"
for m in UTCHMMA "UTCHMMA.2CTA" UTCBAR LDTM STTM UTMALDG UBLKCP FFMA2 SYNCS UCGABAR; do printf "%-14s %s\n" "$m" "$(grep -c "$m" /tmp/yr_all.sass)"; done
printf "%-14s %s\n" "HMMA (legacy)" "$(grep -c "[^C]HMMA\." /tmp/yr_all.sass)"
echo "#   UTCHMMA = tcgen05.mma (.2CTA = cta_group::2, the CTA-pair kernel), LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG = cp.async.bulk.tensor (TMA),"
echo "#   UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UCGABAR = cluster barrier (SE gate kernel, CTA-pair kernel), FFMA2 = packed fma.rn.f32x2; HMMA (legacy mma.sync) = 0"
echo
echo "# per kernel: UTCHMMA LDTM STTM UTMALDG UBLKCP FFMA2  function"
awk '/Function : /{f=$3} /UTCHMMA/{a[f]++} / LDTM/{b[f]++} / STTM/{c[f]++} /UTMALDG/{d[f]++} /UBLKCP/{e[f]++} /FFMA2/{g[f]++} /Function : /{n[f]=1} END{for (k in n) printf "%4d %4d %4d %4d %4d %5d  %s\n", a[k],b[k],c[k],d[k],e[k],g[k],k}' /tmp/yr_all.sass | sort -k7 | while read a b c d e g f; do [ $((a+b+c+d+e+g)) -gt 0 ] && printf "%4d %4d %4d %4d %4d %5d  %s\n" $a $b $c $d $e $g "$(echo $f | c++filt | cut -c1-110)"; done
} > profiles/r2_sass_summary.txt
head -16 profiles/r2_sass_summary.txt; grep -c . profiles/r2_sass_summary.txt
