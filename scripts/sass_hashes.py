#!/usr/bin/env python
"""Per-kernel SASS fingerprints of libmfvidip.so: proof that a change left the GPU-verified kernels untouched.

    python scripts/sass_hashes.py --write profiles/r01_sass_hashes.json     # fingerprint the current build
    python scripts/sass_hashes.py --check profiles/r01_sass_hashes.json     # every recorded kernel body must still be present

A fingerprint is the md5 of a kernel's `cuobjdump -sass` text (instructions and encodings) with mangled names blanked, so that
adding a template parameter with a default — which renames the symbol but not the code — still matches.  Kernels that only
exist in the current build (new code) are listed, not failed.
"""
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mfvi_dip_mia_b200", "csrc", "libmfvidip.so")


def fingerprints(lib=LIB):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if line.startswith("Fatbin elf code"):
            cur = None
        if cur and line.strip():
            out[cur].append(re.sub(r"_ZN\S+", "SYM", line).rstrip())
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in out.items()}


def main():
    mode, path = sys.argv[1], sys.argv[2]
    now = fingerprints()
    if mode == "--write":
        with open(path, "w") as f:
            json.dump(now, f, indent=0, sort_keys=True)
        print(f"{len(now)} kernels -> {path}")
        return 0
    with open(path) as f:
        old = json.load(f)
    have = collections.Counter(now.values())
    missing = [k for k, h in old.items() if have[h] == 0]
    new = [k for k, h in now.items() if h not in set(old.values())]
    print(f"{len(old)} recorded kernels, {len(now)} in the current build, {len(new)} new, {len(missing)} changed or missing")
    for k in missing:
        print("  CHANGED:", k)
    return 1 if missing else 0


if __name__ == "__main__":
    sys.exit(main())
