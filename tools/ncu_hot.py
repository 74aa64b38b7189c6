"""Hot SASS instructions of an .ncu-rep source page: top-N by stall samples, plus samples by region."""
import csv, subprocess, sys
rep, launch, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
s = heads[launch]; e = heads[launch + 1] if launch + 1 < len(heads) else len(rows)
h = rows[s]
ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(c, i) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[s + 1:e] if len(r) == len(h)]
tot = sum(int(r[ismp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ismp] or 0))[:topn]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[j] or 0), c) for c, j in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ismp] or 0):6d} {100*int(r[ismp] or 0)/max(tot,1):5.1f}% ex={r[iex]:>8s} {r[isrc][:70]:70s} {st}")
