"""Per-kernel SASS evidence from an .ncu-rep (source page): opcode histogram of the instructions that prove which units
a kernel uses (DMMA / UTCIMMA / LDTM / UTMALDG / UBLKCP / SYNCS / LDG.E.128 ...), and the hottest instructions by stall
samples.   usage: ncu_kernel_sass.py report.ncu-rep kernel-name-substring [topn]"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# sections: a "Kernel Name" row, then a header row starting with "Address", then instruction rows
secs, i = [], 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name" and i + 1 < len(rows) and rows[i + 1] and rows[i + 1][0] == "Address":
        name, hdr, j = r[1], rows[i + 1], i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        secs.append((name, hdr, body))
        i = j
    else:
        i += 1
KEY = re.compile(r"^(DMMA|UTCIMMA|UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|HMMA|IMMA|LDG\.E\.128|LDG\.E\.64|"
                 r"LDG|STG|LDS|STS|LDSM|BAR|SHFL|DFMA|DADD|DMUL|FRND|F2I|I2F|REDUX|ATOMG|RED)")
seen = set()
for name, hdr, body in secs:
    if pat not in name or name in seen:
        continue
    seen.add(name)
    isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(c, k) for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    hist, dyn = Counter(), Counter()
    for r in body:
        op = r[isrc].split()
        op = [t for t in op if not t.startswith("@")]
        if not op:
            continue
        m = KEY.match(op[0])
        if m:
            key = op[0] if op[0].startswith(("LDG.E.128", "LDG.E.64", "DMMA", "UTC", "LDTM", "UTMA", "UBLKCP")) else m.group(1)
            key = re.sub(r"\.(SYNC|ALIGN|SYS|STRONG|CONSTANT|GPU|CTA)", "", key)
            hist[key] += 1
            dyn[key] += int(r[iex] or 0)
    tot = sum(int(r[ismp] or 0) for r in body)
    print(f"=== {name[:110]}")
    print(f"    {len(body)} SASS instructions, {tot} stall samples")
    print("    static count / warp-level executions of the unit-identifying opcodes:")
    for k, v in sorted(hist.items(), key=lambda kv: -dyn[kv[0]]):
        print(f"      {k:28s} {v:5d}  {dyn[k]:>14d}")
    idx = sorted(range(len(body)), key=lambda k: -int(body[k][ismp] or 0))[:topn]
    print(f"    hottest {topn} instructions by stall samples (index, samples, share, executions, SASS, top stall reasons):")
    for k in sorted(idx):
        r = body[k]
        st = sorted(((int(r[j] or 0), c) for c, j in stall_cols), reverse=True)[:2]
        print(f"      {k:5d} {int(r[ismp] or 0):7d} {100 * int(r[ismp] or 0) / max(tot, 1):5.1f}% ex={r[iex]:>10s}  {r[isrc][:84]:84s} {st}")
