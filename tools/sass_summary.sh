#!/bin/bash
# SASS evidence for profiles/: per kernel of libwun.so, counts of the Blackwell tensor-core / bulk-copy / TMEM / mbarrier instructions.
cd "$(dirname "$0")/.."
cuobjdump -sass wave-u-net_b200/libwun.so | awk '
/Function :/ { fn=$3; next }
/UTCHMMA|UTCBAR|LDTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|UTCATOM|ELECT|REDG|RED\./ {
  n=split("UTCHMMA UTCBAR LDTM UBLKCP UTMALDG UTMASTG SYNCS ELECT RED", K, " ");
  for (i=1;i<=n;i++) if (index($0, K[i])) c[fn,K[i]]++; seen[fn]=1 }
END { for (f in seen) { printf "%s:", f; n=split("UTCHMMA UTCBAR LDTM UBLKCP UTMALDG UTMASTG SYNCS ELECT RED", K, " "); for (i=1;i<=n;i++) if (c[f,K[i]]) printf " %s=%d", K[i], c[f,K[i]]; printf "\n" } }' | c++filt | sort
