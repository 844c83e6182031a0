// Latency of the primitives a lone in-order warp can build a serial decode chain from (development
// aid; run on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_mb tools/chain_microbench.cu
// Each test is a dependent chain of N steps executed by ONE warp; the figure is cycles per step.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 4096;
constexpr unsigned kFull = 0xffffffffu;

template <int V>
__global__ void chain(uint32_t seed, uint32_t* out, long long* cycles) {
    __shared__ uint32_t s_mem[1024];
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 1024; i += 32) s_mem[i] = (uint32_t)((i * 37 + 11) & 1023);
    __syncwarp();
    uint32_t x = seed + out[lane];   // uniform at run time (out is zeroed), not provably so at compile time
    float f = (float)seed + 1.5f;
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (V == 0) {          // REDUX.OR of a value only one lane contributes
            x = __reduce_or_sync(kFull, lane == (int)(x & 31u) ? x + 1u : 0u);
        } else if (V == 1) {   // ballot -> index of the set bit -> shuffle
            const unsigned b = __ballot_sync(kFull, lane == (int)(x & 31u));
            x = __shfl_sync(kFull, x + 1u, __ffs((int)b) - 1);
        } else if (V == 2) {   // shared-memory broadcast: the owner stores, everybody loads
            if (lane == (int)(x & 31u)) s_mem[0] = x + 1u;
            __syncwarp();
            x = *(volatile uint32_t*)&s_mem[0];
            __syncwarp();
        } else if (V == 3) {   // shuffle alone
            x = __shfl_sync(kFull, x + 1u, (int)(x & 31u));
        } else if (V == 4) {   // shared-memory load alone (pointer chase)
            x = *(volatile uint32_t*)&s_mem[x & 1023u];
        } else if (V == 5) {   // lg2.approx
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
        } else if (V == 6) {   // int -> float -> int
            f = (float)x; x = (uint32_t)(int)f + 1u;
        } else if (V == 7) {   // IMAD.WIDE chain
            const uint64_t p = (uint64_t)x * 2654435761u + x;
            x = (uint32_t)(p >> 32) ^ (uint32_t)p;
        } else if (V == 8) {   // plain ALU op
            x = (x ^ 0x9e3779b9u) + (x >> 3);
        } else if (V == 9) {   // two REDUX.OR in flight, the pop between them (the decoder's singles step)
            const uint32_t hi = x, lo = x * 3u;
            const bool mine = lane == (int)(x & 31u);
            const uint64_t pr = (uint64_t)((hi << 8) | (lo >> 24)) * 12345u + (lo & 0xffffffu);
            const uint32_t a = __reduce_or_sync(kFull, mine ? (uint32_t)(pr >> 32) : 0u);
            const uint32_t b = __reduce_or_sync(kFull, mine ? (uint32_t)pr : 0u);
            x = a ^ b;
        } else if (V == 10) {  // match on a ballot, then two shuffles from the owner
            const uint32_t hi = x, lo = x * 3u;
            const bool mine = lane == (int)(x & 31u);
            const uint64_t pr = (uint64_t)((hi << 8) | (lo >> 24)) * 12345u + (lo & 0xffffffu);
            const unsigned bal = __ballot_sync(kFull, mine);
            const int src = __ffs((int)bal) - 1;
            const uint32_t a = __shfl_sync(kFull, (uint32_t)(pr >> 32), src);
            const uint32_t b = __shfl_sync(kFull, (uint32_t)pr, src);
            x = a ^ b;
        } else if (V == 11) {  // ballot alone (predicate -> mask -> predicate)
            const unsigned b = __ballot_sync(kFull, lane == (int)(x & 31u));
            x = b + x;
        } else if (V == 12) {  // ffs alone
            x = (uint32_t)__ffs((int)(x | 0x80000000u)) + x;
        } else if (V == 13) {  // popc alone
            x = (uint32_t)__popc(x) + x;
        } else if (V == 14) {  // float add (FADD latency)
            f = f + 1.25f;
        } else if (V == 15) {  // select chain
            x = (x & 1u) ? x + 3u : x + 5u;
        }
    }
    const long long t1 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; }
    out[lane] = x + (uint32_t)__float_as_uint(f);
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    cudaMemset(out, 0, 128);
    const char* names[] = {"REDUX.OR (one contributor) + select", "ballot + ffs + shfl", "STS by owner + LDS by all", "SHFL.IDX", "LDS pointer chase",
                           "MUFU.LG2", "I2F + F2I", "IMAD.WIDE + xor", "2 ALU ops", "singles step: pop + 2 REDUX", "singles step: ballot + ffs + 2 SHFL",
                           "ballot + add", "ffs + add", "popc + add", "FADD", "select"};
#define RUN(V) { cudaMemset(out, 0, 128); chain<V><<<1, 32>>>(12345u, out, cyc); cudaMemset(out, 0, 128); chain<V><<<1, 32>>>(12345u, out, cyc); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
                 printf("%-45s %7.1f cycles/step\n", names[V], (double)c / N); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
