// Microbenchmark: fp64 dependent-issue latency and per-SMSP throughput on sm_100a (development aid).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int iters)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) v[i] = __fma_rn(v[i], b, a);
                if (OP == 1) v[i] = __dadd_rn(v[i], b);
                if (OP == 2) v[i] = __dmul_rn(v[i], b);
                if (OP == 3) { double q = __dadd_rn(v[i], -b); v[i] = (__double2hiint(q) < 0) ? b : v[i]; v[i] = __dadd_rn(v[i], a); }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(const char *name, int threads, int blocks)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<ILP, OP><<<blocks, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
    k<ILP, OP><<<blocks, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 16.0 * ILP * (OP == 3 ? 2 : 1));
    printf("%-6s ILP=%d threads/block=%4d blocks=%4d : %.2f cycles per fp64 instr per warp (chain step %.1f cycles)\n", name, ILP, threads, blocks, per, per * ILP * (OP == 3 ? 2 : 1));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1, 0>("DFMA", 32, 1); run<2, 0>("DFMA", 32, 1); run<4, 0>("DFMA", 32, 1); run<8, 0>("DFMA", 32, 1);
    run<1, 1>("DADD", 32, 1); run<1, 2>("DMUL", 32, 1); run<1, 3>("ADDSEL", 32, 1);
    // throughput: 4 warps (1/SMSP), 16 warps (4/SMSP), 32 warps (8/SMSP) per SM
    run<1, 0>("DFMA", 128, 148); run<1, 0>("DFMA", 512, 148); run<1, 0>("DFMA", 1024, 148);
    run<4, 0>("DFMA", 128, 148); run<4, 0>("DFMA", 512, 148); run<2, 0>("DFMA", 512, 148); run<8,0>("DFMA", 512, 148);
    return 0;
}
