"""SM clock / power / throttle reasons sampled every 20 ms while a loop of damped QR solves runs (is the fp64 path power-capped?)."""
import subprocess, sys, threading, time
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
m, n, reps = 100000, 1000, 60
ctx = L.Context.default(0)
if len(sys.argv) > 1: ctx.set_option('qr_apply', int(sys.argv[1]))
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
rows = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                         "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
def rd():
    for l in proc.stdout: rows.append((time.perf_counter(), l.strip()))
th = threading.Thread(target=rd, daemon=True); th.start()
time.sleep(0.5)
ws.ldiv(x, A, y, dtd); ctx.sync()
t0 = time.perf_counter()
for _ in range(reps): ws.ldiv(x, A, y, dtd)
ctx.sync()
t1 = time.perf_counter()
time.sleep(0.2); proc.terminate()
print("ms per solve", (t1 - t0) / reps * 1e3)
busy = [r for t, r in rows if t0 + 0.1 < t < t1]
idle = [r for t, r in rows if t < t0 - 0.1]
print("idle sample:", idle[-1] if idle else None)
print("samples under load:", len(busy))
for r in busy[:: max(1, len(busy) // 12)]: print("  ", r)
