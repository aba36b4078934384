#!/bin/bash
# SASS excerpts for profiles/: for every tcgen05 kernel of libwun.so, the first occurrence of each Blackwell instruction family
# (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, SYNCS = mbarrier, ELECT) with two
# lines of context, plus the longest run of consecutive UTCHMMA instructions (the unrolled issue loop).
cd "$(dirname "$0")/.."
cuobjdump -sass wave-u-net_b200/libwun.so | c++filt | awk '
function flush() {
  if (fn == "" || !has_mma) return;
  printf "==== %s\n", fn;
  printf "     longest run of consecutive UTCHMMA: %d\n", best;
  n = split("UTCHMMA UTCBAR LDTM UBLKCP SYNCS.ARRIVE SYNCS.PHASECHK ELECT", K, " ");
  for (i = 1; i <= n; i++) if (first[K[i]] > 0) {
    printf "  -- first %s\n", K[i];
    for (j = first[K[i]] - 2; j <= first[K[i]] + 2; j++) if (j >= 1 && j <= nl) printf "     %s\n", line[j];
  }
}
/Function :/ { flush(); fn = $0; sub(/.*Function : /, "", fn); nl = 0; has_mma = 0; run = 0; best = 0; delete first; delete line; next }
/^[ \t]*\/\*[0-9a-f]+\*\// {
  txt = $0; sub(/^[ \t]*/, "", txt); sub(/[ \t]*\/\* 0x[0-9a-f]+ \*\/[ \t]*$/, "", txt);
  nl++; line[nl] = txt;
  if (index(txt, "UTCHMMA")) { has_mma = 1; run++; if (run > best) best = run } else run = 0;
  n = split("UTCHMMA UTCBAR LDTM UBLKCP SYNCS.ARRIVE SYNCS.PHASECHK ELECT", K, " ");
  for (i = 1; i <= n; i++) if (!(K[i] in first) && index(txt, K[i])) first[K[i]] = nl;
}
END { flush() }'
