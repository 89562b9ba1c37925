// Microbenchmark: the memory access structure of the chain kernels with (almost) no arithmetic.
// thread per column, block = 128 columns of one segment; forward sweep reads x (+ writes psi/ckpt
// like scratch), backward sweep re-reads x and writes 3 fp64 planes + 1 byte plane.  Development aid.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) pat(const double *x, long ld, long T, long B, const long *seg, const int *order, int nblk,
                    double *g, unsigned char *y, double *ck, int mode, int tiled)
{
    extern __shared__ double sm[];
    int s = order[blockIdx.x / nblk];
    long b = (long)(blockIdx.x % nblk) * 128 + threadIdx.x;
    if (b >= B) return;
    long t0 = seg[s], t1 = seg[s + 1];
    double acc = 0;
    const long tb = (b >> 5) * (T * 32) + (b & 31), tcb = (b >> 5) * ((T / 8 + 64) * 4 * 32) + (b & 31);
#define XI(t) (tiled ? tb + (t) * 32 : (t) * ld + b)
#define CI(r) (tiled ? tcb + (r) * 32 : (r) * B + b)
    if (mode & 1) { // forward sweep: read x, write 6 B/step of scratch every 8 steps
        for (long t = t0; t < t1; t += 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (t + j < t1) ? x[XI(t + j)] : 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) acc = acc * 0.999 + v[j];
            long c = t / 8;
            ck[CI(c * 4 + 0)] = acc; ck[CI(c * 4 + 1)] = acc + 1; ck[CI(c * 4 + 2)] = acc + 2; ck[CI(c * 4 + 3)] = acc + 3;
        }
    }
    if (mode & 2) { // backward sweep
        long plane = T * ld;
        for (long t = ((t1 - 1 - t0) / 8) * 8 + t0; t >= t0; t -= 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (t + j < t1) ? x[XI(t + j)] : 0.0;
            long c = t / 8;
            acc += ck[CI(c * 4 + 0)] + ck[CI(c * 4 + 1)] + ck[CI(c * 4 + 2)] + ck[CI(c * 4 + 3)];
#pragma unroll
            for (int j = 7; j >= 0; j--) {
                if (t + j < t1) {
                    acc = acc * 0.999 + v[j];
                    if (tiled == 3) {
                        // pair layout: steps (2k, 2k+1) of a lane adjacent -> one 16-byte store per plane and pair
                        if (j & 1) {
                            const long o2 = (b >> 5) * (T * 32) + ((t + j) >> 1) * 64 + (b & 31) * 2;
                            double2 v0 = make_double2(acc, acc + 3), v1 = make_double2(acc + 1, acc + 4), v2 = make_double2(acc + 2, acc + 5);
                            *reinterpret_cast<double2 *>(g + o2) = v0;
                            *reinterpret_cast<double2 *>(g + o2 + plane) = v1;
                            *reinterpret_cast<double2 *>(g + o2 + 2 * plane) = v2;
                            *reinterpret_cast<unsigned short *>(y + o2) = (unsigned short)(int)acc;
                        }
                    } else if (tiled == 2) {
                        // one record per (tile, chunk): 3 planes x 8 steps x 32 lanes doubles, then 8 x 32 bytes
                        const long rec = ((b >> 5) * (T / 8 + 24) + c + s) * (3 * 8 * 32 + 32); // in doubles (y: 256 B = 32 doubles)
                        double *gr = g + rec;
                        gr[(0 * 8 + j) * 32 + (b & 31)] = acc; gr[(1 * 8 + j) * 32 + (b & 31)] = acc + 1; gr[(2 * 8 + j) * 32 + (b & 31)] = acc + 2;
                        ((unsigned char *)(gr + 3 * 8 * 32))[j * 32 + (b & 31)] = (unsigned char)(int)acc;
                    } else {
                    long o = XI(t + j);
                    g[o] = acc; g[o + plane] = acc + 1; g[o + 2 * plane] = acc + 2;
                    y[o] = (unsigned char)(int)acc;
                    }
                }
            }
        }
    }
    if (acc == 1.2345) sm[threadIdx.x] = acc;
}
int main(int argc, char **argv)
{
    long T = 20000, B = 10000, ld = B;
    int cnt[22] = {704,392,379,268,279,431,388,194,223,94,328,331,35,82,165,218,311,81,338,150,77,119};
    std::vector<long> seg(23); long tot = 0; for (int c : cnt) tot += c;
    long acc = 0; for (int i = 0; i < 22; i++) { seg[i] = acc * T / tot; acc += cnt[i]; } seg[22] = T;
    std::vector<int> order(22); std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg[a + 1] - seg[a] > seg[b + 1] - seg[b]; });
    double *x, *g, *ck; unsigned char *y; long *dseg; int *dord;
    cudaMalloc(&x, T * ld * 8); cudaMalloc(&g, 3 * T * ld * 8 + T * ld * 2 + (1 << 24)); cudaMalloc(&y, T * ld); cudaMalloc(&ck, (T / 8 + 64) * 4 * B * 8);
    cudaMemset(x, 0, T * ld * 8);
    cudaMalloc(&dseg, 23 * 8); cudaMalloc(&dord, 22 * 4);
    cudaMemcpy(dseg, seg.data(), 23 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dord, order.data(), 22 * 4, cudaMemcpyHostToDevice);
    int nblk = (B + 127) / 128;
    B = 9984; ld = B; nblk = (B + 127) / 128; // multiple of 32 for the tiled variant
    for (int tiled : {1, 2, 3}) {
        size_t smem = 53248;
        cudaFuncSetAttribute(pat, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
        for (int mode : {1, 2, 3}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int i = 0; i < 2; i++) pat<<<22 * nblk, 128, smem>>>(x, ld, T, B, dseg, dord, nblk, g, y, ck, mode, tiled);
            cudaEventRecord(e0);
            for (int i = 0; i < 5; i++) pat<<<22 * nblk, 128, smem>>>(x, ld, T, B, dseg, dord, nblk, g, y, ck, mode, tiled);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            double bytes = 0;
            if (mode & 1) bytes += T * B * (8.0 + 4.0);
            if (mode & 2) bytes += T * B * (8.0 + 4.0 + 24.0 + 1.0);
            printf("tiled %d smem/block %6zu mode %d (%s): %.3f ms, %.2f GB -> %.0f GB/s  [%s]\n", tiled, smem, mode, mode == 1 ? "fwd" : mode == 2 ? "bwd" : "fwd+bwd", ms, bytes / 1e9, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
