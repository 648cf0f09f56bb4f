// Dependent-issue latencies on the target GPU (cycles per operation in a dependent chain), one warp / one CTA:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/latency_probe scripts/latency_probe.cu && /tmp/latency_probe
// Used to reason about the local-BA reduced solve (pivot chains) and the Schur phase.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 512;
#define CHAIN(name, body)                                                                  \
    __global__ void name(double* out, long long* cyc, double x0, double y0) {              \
        __shared__ double sh[64];                                                          \
        sh[threadIdx.x & 63] = y0 + threadIdx.x;                                           \
        __syncthreads();                                                                   \
        double x = x0 + threadIdx.x * 1e-9, y = y0;                                        \
        int idx = threadIdx.x & 31;                                                        \
        (void)idx;                                                                         \
        const long long t0 = clock64();                                                    \
        _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; }                         \
        const long long t1 = clock64();                                                    \
        out[blockIdx.x * blockDim.x + threadIdx.x] = x;                                    \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                   \
    }

CHAIN(k_dfma, x = fma(x, y, 1e-3))
CHAIN(k_dadd, x = x + y)
CHAIN(k_dmul, x = x * y)
CHAIN(k_ffma, { float f = (float)x; _Pragma("unroll") for (int k = 0; k < 1; ++k) f = fmaf(f, 1.0001f, 0.5f); x = f; })
CHAIN(k_rcp64, { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.0; })
CHAIN(k_rsqrt64, { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 2.0; })
CHAIN(k_div64, x = 1.0 / x + 1.5)
CHAIN(k_sqrt64, x = sqrt(x) + 1.5)
CHAIN(k_shfl64, x = __shfl_sync(0xffffffffu, x, (idx + 1) & 31) + 1e-3)
CHAIN(k_lds64, { x = sh[((int)x) & 63] ; })
CHAIN(k_lds_dfma_sts, { sh[idx] = fma(sh[idx], y, x); __syncwarp(); })
CHAIN(k_bar, { __syncthreads(); x += 1e-3; })
CHAIN(k_warpsum, { double v = x; _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); x = v * 1e-3; })

// fp32 chain measured directly
__global__ void k_ffma32(float* out, long long* cyc, float x0) {
    float x = x0 + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fmaf(x, 1.0001f, 0.5f);
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename K>
void run(const char* name, K k, int threads, double x0, double y0, double* d_out, long long* d_cyc) {
    k<<<1, threads>>>(d_out, d_cyc, x0, y0);
    k<<<1, threads>>>(d_out, d_cyc, x0, y0);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-28s threads %4d : %7.1f cycles per iteration\n", name, threads, (double)c / N);
}

int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 1024 * sizeof(double)); cudaMalloc(&d_cyc, 64);
    for (int threads : {32, 256}) {
        run("DFMA dependent", k_dfma, threads, 1.0, 0.999, d_out, d_cyc);
        run("DADD dependent", k_dadd, threads, 1.0, 1e-3, d_out, d_cyc);
        run("DMUL dependent", k_dmul, threads, 1.0, 0.9999, d_out, d_cyc);
        run("F2F+FFMA+F2F round trip", k_ffma, threads, 1.0, 0.0, d_out, d_cyc);
        run("rcp.approx.f64 + DADD", k_rcp64, threads, 1.5, 0.0, d_out, d_cyc);
        run("rsqrt.approx.f64 + DADD", k_rsqrt64, threads, 1.5, 0.0, d_out, d_cyc);
        run("1.0 / x (IEEE) + DADD", k_div64, threads, 1.5, 0.0, d_out, d_cyc);
        run("sqrt(x) (IEEE) + DADD", k_sqrt64, threads, 1.5, 0.0, d_out, d_cyc);
        run("SHFL.64 + DADD", k_shfl64, threads, 1.0, 0.0, d_out, d_cyc);
        run("LDS.64 dependent address", k_lds64, threads, 1.0, 3.0, d_out, d_cyc);
        run("LDS+DFMA+STS+syncwarp", k_lds_dfma_sts, threads, 1e-3, 0.5, d_out, d_cyc);
        run("__syncthreads + DADD", k_bar, threads, 1.0, 0.0, d_out, d_cyc);
        run("warp_sum fp64 (5 steps)", k_warpsum, threads, 1.0, 0.0, d_out, d_cyc);
    }
    float* f_out; cudaMalloc(&f_out, 4096);
    k_ffma32<<<1, 32>>>(f_out, d_cyc, 1.0f); k_ffma32<<<1, 32>>>(f_out, d_cyc, 1.0f);
    cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-28s threads   32 : %7.1f cycles per iteration\n", "FFMA dependent", (double)c / N);
    return 0;
}
