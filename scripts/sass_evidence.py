#!/usr/bin/env python
"""Static SASS evidence (no GPU needed): per kernel of lib/*.o, how many tcgen05 MMA (UTC*MMA), TMEM load/store (LDTM / STTM),
TMA (UTMALDG / UTMASTG / UBLKCP / UBLKRED) and legacy mma.sync (HMMA) instructions the sm_100a binary contains.
python scripts/sass_evidence.py > profiles/sass_evidence_r01.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "patchrefinerv2_b200", "lib")
PATTERNS = [("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
            ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UBLKCP / UBLKRED (bulk copy / reduce)", r"\bUBLK(CP|RED)"),
            ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (legacy mma.sync)", r"\bHMMA"), ("MUFU.EX2", r"MUFU\.EX2"), ("total", r"^\s+/\*[0-9a-f]{4}\*/")]


def main():
    print("# SASS evidence, sm_100a (`cuobjdump -sass` of `patchrefinerv2_b200/lib/*.o`, built by `__graft_entry__.build()`)\n")
    print("Instruction counts per kernel; only kernels that use the tensor core, TMEM, TMA / bulk copies or mbarriers are listed in the")
    print("first table. No kernel contains `HMMA` (legacy `mma.sync`): every matrix product goes through `tcgen05.mma`.\n")
    rows = []
    for obj in sorted(os.listdir(LIB)):
        if not obj.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
        name, counts = None, None
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if name:
                    rows.append((obj, name, counts))
                name, counts = m.group(1), collections.Counter()
                continue
            if name:
                for label, pat in PATTERNS:
                    if re.search(pat, line):
                        counts[label] += 1
        if name:
            rows.append((obj, name, counts))
    def demangle(n):
        full = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", full)
        return m.group(1) if m else full[:60]
    labels = [p[0] for p in PATTERNS]
    print("| object | kernel | " + " | ".join(labels) + " |")
    print("|---|---|" + "---|" * len(labels))
    rest = []
    for obj, name, c in rows:
        if any(c[l] for l in labels[:7]):
            print(f"| {obj} | `{demangle(name)}` | " + " | ".join(str(c[l]) for l in labels) + " |")
        else:
            rest.append((obj, demangle(name), c["total"]))
    assert not any(c[labels[7]] for _, _, c in rows), "legacy HMMA found"
    print("\nOther kernels (plain SIMT, HBM-bound): " + ", ".join(f"`{n.split('::')[-1]}` ({t})" for _, n, t in rest))


if __name__ == "__main__":
    sys.exit(main())
