#!/bin/sh
# SASS evidence per kernel of the built library: TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), warp match/vote,
# dp4a, warp reductions, shared-memory atomics.   usage: tools/sass_markers.sh > profiles/r02_sass_markers.txt
LIB=${1:-zlib.es_b200/libzles.so}
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $LIB  git $(git rev-parse --short HEAD)  ($(cuobjdump -lelf $LIB | head -1))"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3 }
  /UBLKCP/ {c[fn,"UBLKCP"]++} /SYNCS/ {c[fn,"SYNCS"]++} /MATCH\.ANY/ {c[fn,"MATCH.ANY"]++} /VOTE/ {c[fn,"VOTE"]++}
  /IDP\.4A/ {c[fn,"IDP.4A"]++} /REDUX/ {c[fn,"REDUX"]++} /ATOMS/ {c[fn,"ATOMS"]++} /SHFL/ {c[fn,"SHFL"]++} /LOP3/ {c[fn,"LOP3"]++}
  /[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z]/ {n[fn]++}
  END { printf "%-28s %8s %7s %6s %10s %6s %7s %6s %6s %6s %6s\n","kernel","instr","UBLKCP","SYNCS","MATCH.ANY","VOTE","IDP.4A","REDUX","ATOMS","SHFL","LOP3";
        for (f in n) { g=f; sub(/^_ZN4zles[0-9]+/,"",g); sub(/E.*$/,"",g);
          printf "%-28s %8d %7d %6d %10d %6d %7d %6d %6d %6d %6d\n", g, n[f], c[f,"UBLKCP"], c[f,"SYNCS"], c[f,"MATCH.ANY"], c[f,"VOTE"], c[f,"IDP.4A"], c[f,"REDUX"], c[f,"ATOMS"], c[f,"SHFL"], c[f,"LOP3"] } }' | sort
