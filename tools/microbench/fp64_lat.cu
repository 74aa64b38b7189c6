// fp64 latency / issue-interval probe for one SM (sm_100a): a single CTA of W warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int NCHAIN>
__global__ void dep_kernel(int iters, double* out, long long* cyc) {
    double c[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lds_fma_kernel(int iters, double* out, long long* cyc) {
    __shared__ __align__(16) double buf[64];
    if (threadIdx.x < 64) buf[threadIdx.x] = threadIdx.x * 0.001;
    __syncthreads();
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    double coef = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double* cb = buf + (it & 1) * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const double2 v = *reinterpret_cast<const double2*>(cb + i);
            x[i] = fma(-coef, v.x, x[i]);
            x[i + 1] = fma(-coef, v.y, x[i + 1]);
        }
        coef = x[3] * 1e-9;      // serialise iterations through one value (like coef of the next step)
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void sqrt_kernel(int iters, double* out, long long* cyc, int mode) {
    double q = 2.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) q = sqrt(q) + 3.0;
        else if (mode == 1) q = 1.0 / q + 3.0;
        else if (mode == 2) q = rsqrt(q) + 3.0;
        else { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q)); q = y + 3.0; }
    }
    long long t1 = clock64();
    if (q == 12345.678) out[0] = q;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    long long h;
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16, 32}) {
        dep_kernel<1><<<1, warps * 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d  dependent DFMA chain: %.1f cycles/op\n", warps, (double)h / iters);
        dep_kernel<4><<<1, warps * 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d  4 chains: %.1f cycles/DFMA (per warp)\n", warps, (double)h / iters / 4);
        dep_kernel<16><<<1, warps * 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d  16 chains: %.1f cycles/DFMA (per warp)\n", warps, (double)h / iters / 16);
        lds_fma_kernel<<<1, warps * 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d  8 LDS.128 + 16 DFMA + dependent mul: %.1f cycles/iteration\n", warps, (double)h / iters);
    }
    for (int mode = 0; mode < 4; ++mode) {
        sqrt_kernel<<<1, 32>>>(iters, out, cyc, mode); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d (0 sqrt,1 div,2 rsqrt,3 rsqrt.approx) + add: %.1f cycles\n", mode, (double)h / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
