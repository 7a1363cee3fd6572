#!/bin/bash
# tools/sass_same_as.sh <commit>: rebuilds librhj.so of <commit> in a scratch directory and compares ALL device code (SASS, encodings
# stripped) with the library built in the tree.  Used when only host code or uninstantiated template code changed after the last
# GPU run: identical SASS means the kernels the GPU tests validated are the kernels that ship.  Optional sed expression in $2 maps
# renamed kernels (e.g. 's/k_join_posILi3ELb1ELi/k_join_posILi3ELi/').
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$here" archive "$1" radixhashjoin_b200/csrc include | tar -x -C "$tmp"
make -C "$tmp/radixhashjoin_b200/csrc" > /dev/null 2>&1
strip_enc() { cuobjdump -sass "$1" | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##; s#/\*[0-9a-f]*\*/##'; }
strip_enc "$tmp/radixhashjoin_b200/librhj.so" | sed "${2:-s/^//}" > "$tmp/a.sass"
strip_enc "$here/radixhashjoin_b200/librhj.so" > "$tmp/b.sass"
if cmp -s "$tmp/a.sass" "$tmp/b.sass"; then echo "device code identical to $1 ($(wc -l < "$tmp/b.sass") SASS lines)"; else echo "device code DIFFERS from $1"; diff "$tmp/a.sass" "$tmp/b.sass" | head -20; exit 1; fi
rm -rf "$tmp"
