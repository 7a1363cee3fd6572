#!/bin/bash
# tools/sass_of.sh <mangled-name-substring> [librhj.so]: SASS of one kernel of the built library, encodings stripped
# (used to read the hot loops before spending GPU time: instruction counts, spills, back-to-back LDS / ATOMS).
so=${2:-$(dirname "$0")/../radixhashjoin_b200/librhj.so}
cuobjdump -sass "$so" | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f' | grep -v '^\s*/\* 0x' |
    sed 's#/\* 0x[0-9a-f]* \*/##; s#/\*[0-9a-f]*\*/##; s/^ *//; s/ *$//'
