"""Per-kernel SASS instruction histogram of the built extension (cuobjdump -sass on the per-size objects).
usage: python scripts/sass_histogram.py [L ...] > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "bounty_matrix_inversion_b200", "csrc", "build")
sizes = [int(a) for a in sys.argv[1:]] or [11]
files = [os.path.join(OBJ, f"launch_L{L}.o") for L in sizes] + [os.path.join(OBJ, "engine.o")]
NOTE = {"UBLKCP": "TMA bulk copy (cp.async.bulk)", "STAS": "st.async (DSMEM store with mbarrier completion)", "SYNCS": "mbarrier",
        "UCGABAR_ARV": "cluster barrier", "UCGABAR_WAIT": "cluster barrier", "SHFL": "warp shuffle", "IMAD.WIDE.U32": "32x32->64 multiply-add",
        "LDG.E.64.CONSTANT": "ld.global.nc 64-bit", "LDG.E.128.CONSTANT": "ld.global.nc 128-bit"}
print("# SASS instruction histogram per kernel (`cuobjdump -sass`, sm_100a)\n")
print("Counts are static instructions in the kernel body; the evidence columns name the Blackwell/Hopper-class features present.\n")
for path in files:
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    print(f"## {os.path.basename(path)}\n")
    for name, h in hist.items():
        total = sum(h.values())
        if total < 40 or "NOP" in name:
            continue
        groups = collections.Counter()
        for op, n in h.items():
            groups[op.split(".")[0] if not op.startswith(("IMAD.WIDE", "LDG", "STAS", "SHFL", "SYNCS", "UBLKCP")) else ".".join(op.split(".")[:4])] += n
        top = ", ".join(f"{op} {n}" for op, n in groups.most_common(9) if op != "NOP")
        feats = sorted({NOTE[k] for k in NOTE for op in h if op.startswith(k)})
        short = re.sub(r"\(PbsArgs\)|void ", "", name)
        print(f"* `{short}` — {total} instructions: {top}.  Features: {'; '.join(feats) if feats else '-'}")
    print()
