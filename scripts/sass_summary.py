"""Per-kernel SASS evidence of the built library: counts of the Blackwell-native opcodes (tcgen05.mma = UTC*MMA,
tcgen05.ld/st = LDTM/STTM, bulk TMA = UBLKCP, tcgen05.commit = UTCBAR, mbarrier = SYNCS, cluster ops), the legacy
tensor path (HMMA: must be 0), plus registers / spills from the ptxas logs of the in-tree build.

    python scripts/sass_summary.py > profiles/r2_sass_summary.json     (no GPU needed)
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "plenoctree_b200", "libplenoctree_b200.so")
OPS = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UTMALDG", "UTMASTG",
       "HMMA", "UCGABAR", "STG", "LDG", "ATOMG", "REDG", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {"instructions": 0, **{o: 0 for o in OPS}}
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        k = kernels[cur]
        k["instructions"] += 1
        base = op.split(".")[0]
        if base in k:
            k[base] += 1
        if op.startswith("UTCHMMA.2CTA"):
            k["UTCHMMA.2CTA"] += 1
    regs = {}
    for log in sorted(os.listdir(os.path.join(ROOT, "plenoctree_b200", "build"))):
        if not log.endswith(".o.log"):
            continue
        txt = open(os.path.join(ROOT, "plenoctree_b200", "build", log)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
            regs[m.group(1)] = {"registers": int(m.group(5)), "stack_bytes": int(m.group(2)),
                                "spill_store_bytes": int(m.group(3)), "spill_load_bytes": int(m.group(4))}
    names = demangle(list(kernels))
    out = {}
    for mangled, k in kernels.items():
        rec = {o: c for o, c in k.items() if c}
        rec.update(regs.get(mangled, {}))
        out[names.get(mangled, mangled)[:120]] = rec
    total = {o: sum(k.get(o, 0) for k in kernels.values()) for o in OPS}
    json.dump({"library": os.path.relpath(LIB, ROOT), "arch": "sm_100a", "totals": {o: c for o, c in total.items() if c},
               "legacy_tensor_path_HMMA": total["HMMA"], "kernels": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
