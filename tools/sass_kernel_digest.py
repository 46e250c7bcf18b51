#!/usr/bin/env python
"""Per-kernel digest of the SASS of an object file (instruction text without encodings and without the per-build
hash nvcc puts into internal-linkage names).  Two builds agree on a kernel's digest iff they run the same
instructions for it: with no GPU at hand this shows which kernels a change did NOT touch relative to a build that
was validated on hardware.

    python tools/sass_kernel_digest.py build/msm.o                  # list
    python tools/sass_kernel_digest.py old/msm.o new/msm.o          # diff
"""
import hashlib
import re
import subprocess
import sys


def digests(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    res, name, lines = {}, None, []

    def flush():
        if name is not None:
            res[name] = hashlib.md5("\n".join(lines).encode()).hexdigest()[:12]

    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", m.group(1))
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            lines = []
            continue
        if re.match(r"\s*/\* 0x", line):
            continue
        line = re.sub(r"/\*[0-9a-f]+\*/", "", line)
        line = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", line)
        lines.append(line.strip())
    flush()
    return res


if __name__ == "__main__":
    if len(sys.argv) == 2:
        for k, v in sorted(digests(sys.argv[1]).items()):
            print(v, k)
    else:
        a, b = digests(sys.argv[1]), digests(sys.argv[2])
        same = sorted(k for k in a if b.get(k) == a[k])
        print(f"{len(same)} kernels identical, {len(set(a) | set(b)) - len(same)} differ or exist on one side only")
        for k in sorted(set(a) | set(b)):
            if a.get(k) != b.get(k):
                print(f"  {a.get(k, '-'):12s} {b.get(k, '-'):12s} {k}")
