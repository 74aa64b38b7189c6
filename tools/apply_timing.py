import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
from lsob200._lib import check, lib
ctx = L.Context.default(0)
if len(sys.argv) > 1: ctx.set_option('qr_apply', int(sys.argv[1]))
fn = lib().lso_debug_apply_timing
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
buf = np.zeros(32, dtype=np.int64)
fn(ctx.handle, buf.ctypes.data)           # arm
m, n = 100000, 1000
A = L.DenseMatrix(ctx, m, n)
check(lib().lso_synth_dense_matrix(ctx.handle, m, n, 0, 20240608, A.ptr, A.ld), ctx.handle)
y = L.DeviceVector(ctx, m); check(lib().lso_synth_vector(ctx.handle, m, 0, 77, 1.0, y.ptr), ctx.handle)
dtd, x = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
A.colsumabs2(dtd); L.api._lm_damping(ctx, dtd, 0.1)
ws = L.DenseQRAllocatedSolver(ctx, m, n, True)
for _ in range(2):
    ws.ldiv(x, A, y, dtd)
fn(ctx.handle, buf.ctypes.data)           # read + re-arm
ws.ldiv(x, A, y, dtd)
fn(ctx.handle, buf.ctypes.data)
names = ["wait full", "GEMM1+store partial", "sync1", "reduce", "sync2", "T-mult", "sync3", "C-init loads", "GEMM2", "stores + loop top (V switch)"]
nj = max(int(buf[10]), 1)
tot = sum(buf[:10])
print("tiles", nj, "cycles/tile", tot / nj)
for i in range(10):
    print(f"  {names[i]:22s} {buf[i] / nj:8.0f} cycles  {100 * buf[i] / tot:5.1f}%")
print("CTA 0 total cycles in kernel", int(buf[11]), "=> per tile", buf[11] / nj)
print("CTA 0 wall ns", int(buf[12]), "=> effective SM clock %.0f MHz" % (buf[11] / max(buf[12], 1) * 1e3))
print("producer of the middle CTA: ns since kernel start at the first job of each level, then at exit:", [int(v) for v in buf[16:22]])
t0 = int(buf[24]); print("all CTAs: first job at 0, slowest CTA enters level l at (us):", [round((int(v) - t0) / 1e3, 1) if v else None for v in buf[26:30]], "last exit", round((int(buf[25]) - t0) / 1e3, 1))
n = max(int(buf[31]), 1); print("producer exit times over the %d CTAs (us after the first job): first %.1f, mean %.1f, last %.1f" % (n, (int(buf[23]) - t0) / 1e3, (int(buf[30]) / n - t0) / 1e3, (int(buf[25]) - t0) / 1e3))
