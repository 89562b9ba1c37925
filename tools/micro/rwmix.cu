// Microbenchmark: plain streaming kernels with different read:write mixes (is the measured copy
// bandwidth reachable when 60 % of the bytes are writes, as in the chain kernels?).  Development aid.
#include <cstdio>
#include <cuda_runtime.h>
template <int NR, int NW>
__global__ void __launch_bounds__(256) mix(const double *const *r, double *const *w, long n)
{
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
        double v = 0;
#pragma unroll
        for (int k = 0; k < NR; k++) v += __ldcs(r[k] + i);
#pragma unroll
        for (int k = 0; k < NW; k++) __stcs(w[k] + i, v + k);
    }
}
template <int NR, int NW>
void run(const double *const *dr, double *const *dw, long n, int blocks)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) mix<NR, NW><<<blocks, 256>>>(dr, dw, n);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) mix<NR, NW><<<blocks, 256>>>(dr, dw, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("reads %d writes %d blocks %5d: %.3f ms  %.0f GB/s  [%s]\n", NR, NW, blocks, ms, (NR + NW) * 8.0 * n / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const long n = 200000000L; // 1.6 GB per array
    double *a[8]; for (int i = 0; i < 8; i++) { cudaMalloc(&a[i], n * 8); cudaMemset(a[i], 0, n * 8); }
    const double **dr; double **dw;
    cudaMalloc(&dr, 64); cudaMalloc(&dw, 64);
    const double *hr[4] = {a[0], a[1], a[2], a[3]}; double *hw[4] = {a[4], a[5], a[6], a[7]};
    cudaMemcpy(dr, hr, 32, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw, 32, cudaMemcpyHostToDevice);
    for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
        run<1, 1>(dr, dw, n, blocks);
        run<2, 3>(dr, dw, n, blocks);
        run<2, 2>(dr, dw, n, blocks);
        run<3, 1>(dr, dw, n, blocks);
        run<1, 3>(dr, dw, n, blocks);
        run<4, 0>(dr, dw, n, blocks);
    }
    return 0;
}
