#!/bin/bash
# Are the GPU-verified kernels still the same machine code?  Builds revision $1 (default: the last revision that ran on a
# GPU, 836631a) in a temporary worktree and compares the SASS of every kernel that exists in both builds.  No GPU needed.
set -e
cd "$(dirname "$0")/.."
REV=${1:-836631a}
T=$(mktemp -d)
git worktree add -q "$T/old" "$REV"
(cd "$T/old/stark-verifier_b200" && bash build.sh > /dev/null 2>&1)
old="$T/old/stark-verifier_b200/libsvb200.so"; new=stark-verifier_b200/libsvb200.so
for k in $(cuobjdump -sass "$old" | grep -oE "Function : \S+" | awk '{print $3}' | sort -u); do
  a=$(cuobjdump -sass -fun "$k" "$old" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | md5sum | cut -c1-12)
  b=$(cuobjdump -sass -fun "$k" "$new" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | md5sum | cut -c1-12)
  [ "$a" == "$b" ] && echo "same     $k" || echo "CHANGED  $k"
done
git worktree remove --force "$T/old"
