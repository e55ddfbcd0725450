// Self-test of the emulator's LT_SIMT_MEMCHECK mode (tests/test_emulated_kernels.py::test_memcheck_self_test):
// argv[1] = "global" | "shared" | "ok" — a kernel that stores one element past a device buffer, one element
// past its dynamic shared memory, or stays inside both.
#include "simt.h"

__global__ void poke(int* out, int n_out, int n_smem, int over_global, int over_shared) {
    LT_DYN_SMEM(raw);
    int* sm = reinterpret_cast<int*>(raw);
    const int t = threadIdx.x;
    if (t < n_smem + over_shared) sm[t] = t;
    __syncthreads();
    if (t < n_out + over_global) out[t] = sm[t < n_smem ? t : 0];
}

int main(int argc, char** argv) {
    const std::string what = argc > 1 ? argv[1] : "ok";
    const int n = 64;                                   // 256 bytes: the buffers end exactly at their guard pages
    int* d = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d), n * sizeof(int)) != cudaSuccess) return 3;
    LT_LAUNCH(poke, 1, 96, n * sizeof(int), nullptr, d, n, n, what == "global" ? 1 : 0, what == "shared" ? 1 : 0);
    int h[n];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) if (h[i] != i) return 4;
    cudaFree(d);
    printf("clean\n");
    return 0;
}
