// Micro-benchmark: issue rate of IMAD vs IDP (dp2a / dp4a) vs PRMT on sm_100a -- decides whether the packed 16-bit x 8-bit dot
// product is worth using in the warp / sub-pel filters of K2.   nvcc -arch=sm_100a -O3 -o idp_bench idp_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int* out, int a0, int b0, int iters) {
    int a = a0 + threadIdx.x, b = b0, c0 = 0, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) { c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b; c4 = c4 * a + b; c5 = c5 * a + b; c6 = c6 * a + b; c7 = c7 * a + b; }
            if (MODE == 1) { c0 = __dp2a_lo(a, b, c0); c1 = __dp2a_hi(a, b, c1); c2 = __dp2a_lo(a, b, c2); c3 = __dp2a_hi(a, b, c3); c4 = __dp2a_lo(a, b, c4); c5 = __dp2a_hi(a, b, c5); c6 = __dp2a_lo(a, b, c6); c7 = __dp2a_hi(a, b, c7); }
            if (MODE == 2) { c0 = __dp4a(a, b, c0); c1 = __dp4a(a, b, c1); c2 = __dp4a(a, b, c2); c3 = __dp4a(a, b, c3); c4 = __dp4a(a, b, c4); c5 = __dp4a(a, b, c5); c6 = __dp4a(a, b, c6); c7 = __dp4a(a, b, c7); }
            if (MODE == 3) { c0 = __byte_perm(c0, a, 0x5410 + (b & 1)); c1 = __byte_perm(c1, a, 0x5410 + (b & 1)); c2 = __byte_perm(c2, a, 0x5410 + (b & 1)); c3 = __byte_perm(c3, a, 0x5410 + (b & 1)); c4 = __byte_perm(c4, a, 0x5410 + (b & 1)); c5 = __byte_perm(c5, a, 0x5410 + (b & 1)); c6 = __byte_perm(c6, a, 0x5410 + (b & 1)); c7 = __byte_perm(c7, a, 0x5410 + (b & 1)); }
            if (MODE == 4) {  // 4 IMAD + 4 PRMT interleaved (two pipes)
                c0 = c0 * a + b; c1 = __byte_perm(c1, a, 0x5410 + (b & 1)); c2 = c2 * a + b; c3 = __byte_perm(c3, a, 0x5410 + (b & 1));
                c4 = c4 * a + b; c5 = __byte_perm(c5, a, 0x5410 + (b & 1)); c6 = c6 * a + b; c7 = __byte_perm(c7, a, 0x5410 + (b & 1)); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
template <int MODE>
static void run(const char* name, int* d) {
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 3, 0x01020304, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 3, 0x01020304, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 64;
    printf("%-28s %.3f ms  %.1f Gop/s  (%.1f thread-ops/clk/SM at 1.965 GHz)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    int* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("IMAD", d); run<1>("IDP.2A (dp2a)", d); run<2>("IDP.4A (dp4a)", d); run<3>("PRMT", d); run<4>("IMAD+PRMT interleaved", d);
    return 0;
}
