#!/bin/bash
# usage: tools/sass.sh <substring of the mangled kernel name> [lib]  -> instruction list (address opcode operands) on stdout
LIB=${2:-/root/repo/3dfacerecon_b200/lib3dfacerecon_b200.so}
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function : /{f=(index($0,pat)>0)} f' | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s+\/\*.*//'
