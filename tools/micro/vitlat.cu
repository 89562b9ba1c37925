// Microbenchmark: how fast does the SEQUENTIAL Viterbi recursion (fast_vit_step of hb_norm3_fast.cuh, the
// 12-DADD difference form) run as a function of the chains resident on one SM?  Inputs come from shared
// memory (a resident chain), nothing touches DRAM: this is the recursion's own dependent-instruction latency.
// It answers whether a chain-resident design (x kept on chip between the sweeps) can feed Viterbi:
// 227 KB of shared memory hold 31 chains of the workload's mean length (909 positions x 8 B) = 1 warp per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I honeybadger_b200/csrc tools/micro/vitlat.cu -o tools/micro/vitlat
#include <cstdio>
#include <cuda_runtime.h>
#include "hb_norm3_fast.cuh"
using namespace hb;

__global__ void vit(const double *uin, int T, Norm3Const c, long long *cyc, unsigned *sink)
{
    extern __shared__ double us[];
    for (int i = threadIdx.x; i < T; i += blockDim.x) us[i] = uin[i];
    __syncthreads();
    Fast3State s;
    s.lnK = -0.07 - 1e-3 * (threadIdx.x & 31);
    s.d0 = s.d2 = 0;
    s.mn = 0x7fffffffu;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int t = 0; t < T; t++) acc ^= fast_vit_step(c, s, us[t] * (1.0 + 1e-3 * (threadIdx.x & 7)), t == 0) << (t & 7);
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ s.mn;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    const int T = 4096;
    double *h = new double[T];
    unsigned long long z = 88172645463325252ull;
    for (int i = 0; i < T; i++) {
        z ^= z << 13; z ^= z >> 7; z ^= z << 17;
        h[i] = ((double)(z >> 11) / 9007199254740992.0 - 0.5) * 1.6; // u ~ U(-0.8, 0.8)
    }
    double *d;
    long long *cyc;
    unsigned *sink;
    cudaMalloc(&d, T * 8);
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 148 * 1024 * 4);
    cudaMemcpy(d, h, T * 8, cudaMemcpyHostToDevice);
    Norm3Const c;
    memset(&c, 0, sizeof c);
    c.tau = log(1e-6) - log(1 - 2e-6);
    c.dl0 = c.dl2 = -INFINITY;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("# warps/SM  chains/SM  cycles/step(one chain)  steps/clk/SM  G steps/s at %.0f MHz x 148 SMs (roofline of V+FB: 198 G)\n", clk / 1e3);
    for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) {
        vit<<<148, 32 * w, T * 8>>>(d, T, c, cyc, sink);
        vit<<<148, 32 * w, T * 8>>>(d, T, c, cyc, sink);
        cudaDeviceSynchronize();
        long long hc[148];
        cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < 148; i++) m += hc[i];
        m /= 148;
        const double per = m / T, rate = 32.0 * w / per;
        printf("%9d %10d %22.1f %13.3f %10.1f\n", w, 32 * w, per, rate, rate * 148 * clk * 1e3 / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
