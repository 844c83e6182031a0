// variant_bench.cu -- times K1 (cdf_tables) built with different FLIC_*_XU switches (development aid).
//   nvcc ... -DFLIC_ARG_XU=1 -DFLIC_V_XU=1 -I finalproject-losslessimagecompression_b200/csrc tools/variant_bench.cu finalproject-losslessimagecompression_b200/csrc/cdf_tables.cu
#include "flic_kernels.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
int main(int argc, char** argv) {
    const int64_t n = 1ll << 27;
    std::vector<float> x(n), m(n), s(n);
    uint64_t st = 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) / 9007199254740992.0; };
    for (int64_t i = 0; i < n; ++i) {
        m[i] = (float)(((int)(rnd() * 513) - 256) / 256.0);
        s[i] = (float)(exp(10 * rnd() - 5) / 256);
        x[i] = (float)(round(((double)m[i] + (double)s[i] * (10 * rnd() - 5)) * 256) / 256);
    }
    float *dx, *dm, *ds; uint32_t *a, *b; int32_t* w;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dm, n * 4); cudaMalloc(&ds, n * 4); cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&w, 4);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dm, m.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(ds, s.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemset(w, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        flic::launch_cdf_tables(dx, dm, ds, n, a, b, w, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<uint32_t> ha(n), hb(n);
    cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost);
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < n; ++i) { h = (h ^ ha[i]) * 1099511628211ull; h = (h ^ hb[i]) * 1099511628211ull; }
    int32_t flags; cudaMemcpy(&flags, w, 4, cudaMemcpyDeviceToHost);
    printf("%s: %.3f ms  %.1f Gsym/s  %.1f cycles/warp-row  checksum %016llx flags %d %s\n", argc > 1 ? argv[1] : "", best, n / best / 1e6,
           best * 1e-3 * 1.965e9 * 592 / (n / 32.0), (unsigned long long)h, flags, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
