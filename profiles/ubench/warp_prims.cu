// Microbenchmark (B200): throughput per SM of the warp primitives the ordered commit can be built from.
// Prints cycles per warp-instruction per SM for match.any (vs number of distinct keys), shfl, vote, redux,
// with 1..16 resident warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_prims warp_prims.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096;
template <int OP>
__global__ void k(unsigned *out, int distinct, long long *cyc) {
    const unsigned lane = threadIdx.x & 31;
    unsigned key = (lane % distinct) * 2654435761u + blockIdx.x;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < kIters; i++) {
        if (OP == 0) acc += __match_any_sync(0xffffffffu, key + acc * 0);          // independent matches
        if (OP == 1) acc += __shfl_sync(0xffffffffu, key + i, (lane + i) & 31);
        if (OP == 2) acc += __ballot_sync(0xffffffffu, (key >> (i & 15)) & 1);
        if (OP == 3) acc += __reduce_or_sync(0xffffffffu, key + i);
        if (OP == 4) {                                                           // dependent matches (latency)
            acc = __match_any_sync(0xffffffffu, key ^ (acc & 0x80000000u));
        }
        if (OP == 5) acc += __reduce_min_sync(0xffffffffu, key + i);
        key += (OP == 0) ? 32u * 2654435761u * 0u : 0u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP>
void run(const char *name, int distinct, int warps_per_sm) {
    const int sms = 148;
    unsigned *out; long long *cyc;
    int threads = 32 * warps_per_sm;
    cudaMalloc(&out, sms * threads * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, threads>>>(out, distinct, cyc);
    k<OP><<<sms, threads>>>(out, distinct, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    printf("%-14s distinct=%2d warps/SM=%2d : %7.2f cycles per warp-instr per warp, %6.2f cycles per warp-instr per SM\n", name, distinct,
           warps_per_sm, avg / kIters, avg / kIters / warps_per_sm);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16}) {
        for (int d : {1, 2, 4, 8, 16, 32}) run<0>("match.any", d, w);
        run<4>("match.any dep", 32, w);
        run<1>("shfl", 32, w);
        run<2>("vote.ballot", 32, w);
        run<3>("redux.or", 32, w);
        run<5>("redux.min", 32, w);
    }
    return 0;
}
