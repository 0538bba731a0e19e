#!/bin/bash
# usage: sass_kernel.sh <object> <mangled-substring>   prints the SASS of the first matching kernel, one instruction per line
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{p = index($0, pat) > 0} p' | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
