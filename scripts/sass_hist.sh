#!/bin/bash
# opcode histogram of one kernel of a model library: scripts/sass_hist.sh <lib.so> <substring of the mangled kernel name>
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); c[op]++} END {for (k in c) printf "%6d %s\n", c[k], k}' | sort -rn
