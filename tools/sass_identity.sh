#!/bin/bash
# Is the DEVICE code of the working tree the same as that of an earlier commit?  Compiles cudacollision.cu of both to a cubin for
# sm_100a and compares the complete SASS listings.  Used after host-only changes made without GPU time: identical SASS means the
# kernels are the ones the GPU runs of that commit verified.  usage: tools/sass_identity.sh <commit>
set -e
commit=${1:?usage: tools/sass_identity.sh <commit>}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$commit" openrave_b200/csrc include | tar -x -C "$tmp"
flags="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -cubin"
(cd "$tmp" && nvcc $flags -o "$tmp/old.cubin" openrave_b200/csrc/cudacollision.cu)
nvcc $flags -o "$tmp/head.cubin" "$root/openrave_b200/csrc/cudacollision.cu"
cuobjdump -sass "$tmp/old.cubin" > "$tmp/old.sass"
cuobjdump -sass "$tmp/head.cubin" > "$tmp/head.sass"
echo "commit $commit: $(wc -l < "$tmp/old.sass") SASS lines, sha256 $(sha256sum < "$tmp/old.sass" | cut -c1-16)"
echo "working tree ($(git -C "$root" rev-parse --short HEAD)+): $(wc -l < "$tmp/head.sass") SASS lines, sha256 $(sha256sum < "$tmp/head.sass" | cut -c1-16)"
if cmp -s "$tmp/old.sass" "$tmp/head.sass"; then echo "device code: IDENTICAL"; rc=0; else echo "device code: DIFFERS"; diff "$tmp/old.sass" "$tmp/head.sass" | head -20; rc=1; fi
rm -rf "$tmp"
exit $rc
