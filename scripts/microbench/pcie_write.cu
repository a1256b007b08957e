// Micro-benchmark: how fast can 1.21 MB (g + dense Jacobian of config[1]) reach pinned host memory?
//   (a) cudaMemcpyAsync D2H from a device buffer (copy engine)
//   (b) SM stores into mapped host memory (zero-copy), 8-byte and 16-byte stores, different grid shapes
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_write pcie_write.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <chrono>
__global__ void w8(double* dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)i;
}
__global__ void w16(double2* dst, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = make_double2((double)i, 1.0);
}
// block-contiguous chunks of `chunk` doubles (like one (t, link) block writing its 140 + 20 rows)
__global__ void wchunk(double* dst, size_t n, int chunk) {
    const size_t base = (size_t)blockIdx.x * chunk;
    for (int e = threadIdx.x; e < chunk && base + e < n; e += blockDim.x) dst[base + e] = (double)e;
}
int main() {
    const size_t n = 18844 * 8;   // doubles: g (m) + jac (7 m)
    double *h, *d, *hd;
    cudaHostAlloc(&h, n * 8, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hd, h, 0);
    cudaMalloc(&d, n * 8);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn) {
        std::vector<float> ev; std::vector<double> wall;
        for (int it = 0; it < 60; it++) {
            auto t0 = std::chrono::high_resolution_clock::now();
            cudaEventRecord(e0, s); fn(); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
            auto t1 = std::chrono::high_resolution_clock::now();
            float ms; cudaEventElapsedTime(&ms, e0, e1); ev.push_back(ms * 1e3f); wall.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
        }
        std::sort(ev.begin(), ev.end()); std::sort(wall.begin(), wall.end());
        printf("%-44s events p50 %7.1f us min %7.1f | wall p50 %7.1f us  => %.1f GB/s (events p50)\n", name, ev[30], ev[0], wall[30], n * 8 / (ev[30] * 1e-6) / 1e9);
    };
    timeit("cudaMemcpyAsync D2H 1.21 MB", [&] { cudaMemcpyAsync(h, d, n * 8, cudaMemcpyDeviceToHost, s); });
    timeit("device write only (w8, 148x256)", [&] { w8<<<148, 256, 0, s>>>(d, n); });
    for (int grid : {37, 74, 148, 296, 592, 897}) {
        char nm[64]; snprintf(nm, 64, "zero-copy w8  grid %d x 128", grid);
        timeit(nm, [&] { w8<<<grid, 128, 0, s>>>(hd, n); });
        snprintf(nm, 64, "zero-copy w16 grid %d x 128", grid);
        timeit(nm, [&] { w16<<<grid, 128, 0, s>>>((double2*)hd, n / 2); });
    }
    timeit("zero-copy chunks of 160 doubles, 943 blocks", [&] { wchunk<<<(unsigned)((n + 159) / 160), 128, 0, s>>>(hd, n, 160); });
    timeit("zero-copy chunks of 1280 doubles", [&] { wchunk<<<(unsigned)((n + 1279) / 1280), 128, 0, s>>>(hd, n, 1280); });
    timeit("empty launch (w8, n = 0)", [&] { w8<<<1, 32, 0, s>>>(d, 0); });
    return 0;
}
