#!/bin/bash
# Per-kernel count of the Blackwell-native SASS mnemonics in libtu_b200.so (tcgen05 MMA / TMEM / TMA), written to profiles/.
# usage: tools/sass_census.sh <out file>   (runs here: cuobjdump needs no GPU)
OUT=${1:-profiles/sass_census.txt}
SO=transformerupscaler_b200/libtu_b200.so
{
echo "# cuobjdump -sass $SO (sm_100a): instructions per kernel"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA load/store, HMMA = mma.sync"
printf "%8s %6s %6s %6s %8s %8s %6s  %s\n" UTCHMMA 2CTA LDTM STTM UTMALDG UTMASTG HMMA kernel
cuobjdump -sass $SO 2>/dev/null | awk '
/Function : /{ if (name!="") printf "%8d %6d %6d %6d %8d %8d %6d  %s\n", u, c2, l, s, tl, ts, h, name; name=$3; u=c2=l=s=tl=ts=h=0 }
/UTCHMMA/{u++} /UTCHMMA.*2CTA/{c2++} /LDTM/{l++} /STTM/{s++} /UTMALDG/{tl++} /UTMASTG/{ts++} / HMMA/{h++}
END{ printf "%8d %6d %6d %6d %8d %8d %6d  %s\n", u, c2, l, s, tl, ts, h, name }' | c++filt | sed 's/(anonymous namespace):://; s/(CUtensorMap_st.*//; s/(.*//' | sort -k1,1nr -k7,7nr
} > $OUT
head -48 $OUT | cut -c1-160
