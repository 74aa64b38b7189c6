"""Per-panel event timeline of one damped QR factorisation (LSO_QR_TIMELINE=1): summarises time per phase."""
import os, subprocess, sys, re, collections
if os.environ.get("LSO_QR_TIMELINE") != "1":
    env = dict(os.environ, LSO_QR_TIMELINE="1")
    out = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    lines = [l for l in out.stderr.splitlines() if l.startswith("TL")]
    # keep only the last solve's lines
    starts = [i for i, l in enumerate(lines) if "leaf chain end" in l and "panel   0" in l]
    lines = lines[starts[-1]:] if starts else lines
    prev = None; agg = collections.defaultdict(float); per = []
    t_prev = None
    for l in lines:
        mm = re.match(r"TL\s+([\d.]+) ms\s+(\w)\s+panel\s+(\d+)\s+(.*)", l)
        t, what = float(mm.group(1)), mm.group(4)
        if t_prev is not None: agg[what] += t - t_prev
        per.append((int(mm.group(3)), what, t))
        t_prev = t
    print(out.stdout.strip())
    for k, v in agg.items(): print(f"{k:20s} {v:8.3f} ms")
    print("first panels:")
    for p in per[:12]: print("  ", p)
    sys.exit(0)
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n = 100000, 1000
ctx = L.Context.default(0)
if len(sys.argv) > 1: ctx.set_option('qr_apply', int(sys.argv[1]))
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
for _ in range(2):
    ws.ldiv(x, A, y, dtd)
ctx.sync()
print("ok", x.norm())
