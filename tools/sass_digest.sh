#!/bin/bash
# Digest of the device code of every object of the product build, ignoring instruction encodings
# and the per-build hash nvcc puts into the names of internal-linkage kernels.  Two builds whose
# digests agree run byte-identical SASS: host-only changes can be checked against a GPU-validated
# build without a GPU (this is how the host-path changes at the end of round 1 were vetted).
#   tools/sass_digest.sh [build-dir]      (default bellman_b200/csrc/build)
dir=${1:-bellman_b200/csrc/build}
for o in "$dir"/*.o; do
    d=$(cuobjdump -sass "$o" | grep -v '^\s*/\* 0x' | sed 's/\/\*[0-9a-f]*\*\///' | grep -v 'Function :' | md5sum | cut -d' ' -f1)
    echo "$(basename "$o") $d"
done
