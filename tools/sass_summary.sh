#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA in the shipped library.
# usage: tools/sass_summary.sh > profiles/sass_summary.txt
LIB=${1:-compression_b200/libtfcb200.so}
echo "# cuobjdump -sass $LIB : per-kernel mnemonic counts (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,"
echo "# UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA tensor maps), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,"
echo "# SYNCS = mbarrier, CREDUX/REDUX = warp reductions, RED = red.global)"
cuobjdump -sass "$LIB" | c++filt | sed -e 's/tfcb::(anonymous namespace):://g' -e 's/(anonymous namespace):://g' | awk '
/Function :/ { fn=$0; sub(/^.*Function : /, "", fn); sub(/\(.*$/, "", fn); gsub(/^void /, "", fn); names[++n]=fn; cur=fn }
{ for (i=1;i<=NF;i++) { t=$i; sub(/\..*/, "", t);
    if (t=="UTCHMMA"||t=="LDTM"||t=="STTM"||t=="UTMALDG"||t=="UTMASTG"||t=="UBLKCP"||t=="UTCBAR"||t=="SYNCS"||t=="CREDUX"||t=="REDUX"||t=="UTMAPF"||t=="UBLKPF"||t=="RED"||t=="IMAD"||t=="HMMA") c[cur,t]++ } }
END { for (k=1;k<=n;k++) { f=names[k]; printf "%-60s", substr(f,1,60);
        split("UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP UBLKPF UTCBAR SYNCS CREDUX REDUX RED", m, " ");
        for (j=1;j<=12;j++) if (c[f,m[j]]>0) printf " %s=%d", m[j], c[f,m[j]]; printf "\n" } }'
