#!/bin/bash
# usage: tools/sass_dump.sh <lib.so> <function-name-substring> : compact SASS listing (address, opcode, operands)
cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f=index($0,pat)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##; s/^\s*//; s/\s\+;\s*$//; s/  \+/ /g'
