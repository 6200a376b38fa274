#!/usr/bin/env python
"""Opcode counts per kernel of the built library -> profiles/rNN_sass_summary.md.

  python scripts/sass_summary.py vpint_b200/libvpint_b200.so profiles/r02_sass_summary.md "Round 2"
"""
import collections
import re
import subprocess
import sys

COLS = ["DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "SYNCS", "STAS", "UBLKCP", "UTMALDG", "UCGABAR",
        "ATOM", "RED", "BRA", "F2F", "I2F", "MEMBAR", "LDL", "STL"]


def main():
    lib, out = sys.argv[1:3]
    title = sys.argv[3] if len(sys.argv) > 3 else "SASS"
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur, k = [], None, -1
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            k += 1
            name = names[k] if k < len(names) and names[k] else m.group(1)
            name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name).replace("vpint::(anonymous namespace)::", "").replace("vpint::<unnamed>::", "")
            name = re.sub(r"^void ", "", name)
            cur = [name, collections.Counter(), 0]
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
            cur[2] += 1
    kernels.sort(key=lambda c: -c[2])
    with open(out, "w") as o:
        o.write("# %s: SASS instruction summary per kernel (%s)\n\n" % (title, ", ".join(sorted(archs)) or "sm_100a"))
        o.write("Source: `cuobjdump -sass %s` (the library `__graft_entry__.build()` produces), opcode counts per kernel (%d kernels); "
                "written by `scripts/sass_summary.py`.\n" % (lib, len(kernels)))
        o.write("`STAS` = st.async into distributed shared memory, `SYNCS` = mbarrier operations, `UBLKCP` = cp.async.bulk (copy engine, "
                "global -> shared), `UTMALDG` = TMA tensor load, `UCGABAR` = cluster barrier.\n"
                "No `UTC*MMA` / `HMMA` / TMEM instructions anywhere: the path is a stencil, not a contraction.\n\n")
        o.write("| kernel | instructions | " + " | ".join(COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
        for name, cnt, n in kernels:
            o.write("| `%s` | %d | %s |\n" % (name, n, " | ".join(str(cnt.get(c, 0)) for c in COLS)))


if __name__ == "__main__":
    main()
