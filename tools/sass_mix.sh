#!/bin/bash
# static SASS instruction mix of one kernel of the built library: tools/sass_mix.sh <mangled-name-substring> [lib]
LIB=${2:-/root/repo/clifford-vae_b200/clifford_b200/libclifford_b200.so}
FUN=$(cuobjdump -sass "$LIB" | grep "Function :" | grep "$1" | head -1 | awk '{print $3}')
echo "kernel: $FUN"
cuobjdump -sass -fun "$FUN" "$LIB" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-24}
cuobjdump -sass -fun "$FUN" "$LIB" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | wc -l
