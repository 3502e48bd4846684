#!/bin/bash
# sass.sh <object> <function substring> : one instruction per line "addr opcode operands"
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{f=index($0,pat)>0} f' | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//'
