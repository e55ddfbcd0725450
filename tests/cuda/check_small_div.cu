// Exhaustive check of lattice.cuh's small_div() with the hardware's approximate reciprocal (small_rcp):
// every divisor 1..8192 against every dividend below 2^20 (the range small_div takes the float path for).
// Built and run by tests/test_gpu_properties.py::test_small_div_exhaustive on the GPU box.
#include <cstdio>
#include "cuda_compat.cuh"
#include "lattice.cuh"

__global__ void check(unsigned long long* bad) {
    const int d = blockIdx.x + 1;
    const float inv = lt::small_rcp(d);
    unsigned long long mine = 0;
    for (int q = threadIdx.x; q < (1 << 20); q += blockDim.x)
        if (lt::small_div(q, d, inv) != q / d) ++mine;
    // the triangular decode: span from t with the one-step correction
    if (blockIdx.x == 0)
        for (int t = threadIdx.x; t < (1 << 20); t += blockDim.x) {
            int span = (int)((1.0f + lt::approx_sqrt(8.0f * (float)t + 1.0f)) * 0.5f);
            span -= (span * (span - 1) / 2 > t) ? 1 : 0;
            span += (span * (span + 1) / 2 <= t) ? 1 : 0;
            if (!(span * (span - 1) / 2 <= t && t < span * (span + 1) / 2)) ++mine;
        }
    if (mine) atomicAdd(bad, mine);
}

int main() {
    unsigned long long* d_bad;
    unsigned long long bad = 0;
    cudaMalloc(&d_bad, 8);
    cudaMemcpy(d_bad, &bad, 8, cudaMemcpyHostToDevice);
    check<<<8192, 256>>>(d_bad);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 2; }
    cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
    printf("mismatches %llu\n", bad);
    return bad ? 1 : 0;
}
