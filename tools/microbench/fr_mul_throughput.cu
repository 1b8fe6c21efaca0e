// Microbenchmark: Montgomery product throughput of fr.cuh on one GPU, as a function of resident warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I arithmetic-circuits_b200/csrc \
//        tools/microbench/fr_mul_throughput.cu -o gpurun_out/fr_mul_throughput && gpurun_out/fr_mul_throughput
// Each thread runs a dependent chain x <- x * y (ITERS products, operands in registers, no memory traffic).
// Prints products per cycle per SM and the implied floor for 2^20 constraints at 2.5 products per constraint.
#include <cstdio>
#include <cuda_runtime.h>
#include "fr.cuh"
#include "fr_variants.cuh"
using namespace acg;

// VAR: 0 the shipped product (8 x 32-bit split accumulator), 1 Karatsuba, 2 the 4 x 64-bit-limb CIOS
template <class P, int CHAINS, int VAR = 0>
__global__ void k_chain(fr_t* out, int iters, unsigned long long* cycles) {
    fr_t x[CHAINS], y;
    for (int i = 0; i < 8; ++i) y.l[i] = 0x1234567u * (threadIdx.x + 3) + i;
    y.l[7] &= 0x0fffffffu;
    for (int c = 0; c < CHAINS; ++c) {
        for (int i = 0; i < 8; ++i) x[c].l[i] = 0x9e3779b9u * (blockIdx.x + c + 1) + i * threadIdx.x;
        x[c].l[7] &= 0x0fffffffu;
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (CHAINS == 2) {
            fr_mul2<P>(x[0], x[1], x[0], y, x[1], y);
        } else {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c)
                x[c] = VAR == 1 ? fr_mul_karatsuba<P>(x[c], y) : (VAR == 2 ? fr_mul_u64<P>(x[c], y) : fr_mul<P>(x[c], y));
        }
    }
    const long long t1 = clock64();
    fr_t acc = x[0];
    for (int c = 1; c < CHAINS; ++c) acc = fr_add<P>(acc, x[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
}

template <int CHAINS, int VAR = 0>
void run(int warps_per_sm, int sms) {
    const int iters = 2000;
    const int threads = 32 * warps_per_sm;
    fr_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, sizeof(fr_t) * threads * sms);
    cudaMalloc(&cyc, 8);
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_chain<Bn254Fr, CHAINS, VAR><<<sms, threads>>>(out, 10, cyc);
    cudaMemset(cyc, 0, 8);
    cudaEventRecord(e0);
    k_chain<Bn254Fr, CHAINS, VAR><<<sms, threads>>>(out, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double prods_per_sm = (double)iters * CHAINS * threads;
    const double ppc = prods_per_sm / (double)c;
    printf("%schains=%d warps/SM=%2d: %.1f cycles per product per warp, %.3f products/cycle/SM, %.2f Gprod/s (chip), "
           "floor for 2.5*2^20 products: %.1f us  [%s]\n",
           VAR == 1 ? "karatsuba " : (VAR == 2 ? "4xu64 " : ""), CHAINS, warps_per_sm, (double)c / (iters * CHAINS), ppc, prods_per_sm * sms / (ms * 1e6),
           2.5 * 1048576.0 / (prods_per_sm * sms / (ms * 1e-3)) * 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

__global__ void k_check_u64(unsigned int* bad) {
    fr_t x, y;
    for (int i = 0; i < 8; ++i) {
        x.l[i] = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1) + 0x85ebca6bu * i;
        y.l[i] = 0xc2b2ae35u * (threadIdx.x + 7) + 0x27d4eb2fu * (i + blockIdx.x);
    }
    x.l[7] &= 0x0fffffffu;
    y.l[7] &= 0x0fffffffu;
    for (int it = 0; it < 64; ++it) {
        const fr_t a = fr_mul<Bn254Fr>(x, y), b = fr_mul_u64<Bn254Fr>(x, y);
        if (!fr_eq(a, b)) atomicAdd(bad, 1u);
        x = y;
        y = a;
    }
}
void check_u64() {
    unsigned int* bad;
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    k_check_u64<<<64, 128>>>(bad);
    unsigned int h = 1;
    cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
    printf("4xu64 vs shipped product on %d products: %u mismatches [%s]\n", 64 * 128 * 64, h, cudaGetErrorString(cudaGetLastError()));
    cudaFree(bad);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {1, 2, 4, 8, 12, 16, 20, 24, 32}) run<1>(w, p.multiProcessorCount);
    for (int w : {4, 8, 12, 16, 20}) run<2>(w, p.multiProcessorCount);
    for (int w : {4, 8, 16}) run<4>(w, p.multiProcessorCount);
    for (int w : {1, 4, 8, 12, 16, 20, 32}) run<1, 1>(w, p.multiProcessorCount);
    for (int w : {1, 4, 8, 12, 16, 20, 32}) run<1, 2>(w, p.multiProcessorCount);
    // correctness of the 4 x 64-bit form against the shipped product on random operands
    check_u64();
    return 0;
}
