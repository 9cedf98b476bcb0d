#!/usr/bin/env python
"""Which kernels of the current build differ from the build of a given commit?  Used to show that a change which adds
opt-in kernels leaves every kernel of a hardware-verified commit bit-identical (compares SASS text without addresses).

    python scripts/sass_diff.py <commit>          # exit status 1 if a kernel of <commit> changed or disappeared
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import build  # noqa: E402


def kernels(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            out[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in out.items()}


def main():
    commit = sys.argv[1]
    cur = kernels(build.build_library())
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "gbp_b200/csrc", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        lib = os.path.join(tmp, "lib.so")
        subprocess.run([build.find_nvcc()] + build.NVCC_FLAGS + ["-o", lib, os.path.join(tmp, "gbp_b200", "csrc", "gbp_ba.cu")], check=True)
        old = kernels(lib)
    # a new template parameter renames every instantiation: match by code, not by name
    bodies = set(cur.values())
    changed = sorted(k for k in old if k in cur and old[k] != cur[k] and old[k] not in bodies)
    gone = sorted(k for k in old if k not in cur and old[k] not in bodies)
    renamed = sorted(k for k in old if k not in cur and old[k] in bodies)
    old_bodies = set(old.values())
    new = sorted(k for k in cur if cur[k] not in old_bodies)
    print(f"{len(old)} kernels at {commit}: {len(old) - len(changed) - len(gone)} with identical code ({len(renamed)} of them under a new name), "
          f"{len(changed)} changed, {len(gone)} removed; {len(new)} new")
    try:
        for tag, names in (("CHANGED", changed), ("REMOVED", gone), ("NEW", new)):
            for k in names:
                print(tag, k)
    except BrokenPipeError:
        pass
    sys.exit(1 if changed or gone else 0)


if __name__ == "__main__":
    main()
