// tools/ubench_fp32x2.cu -- micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2
// (fma.rn.f32x2 / add.rn.f32x2, sm_100+) issue rate per SM, alone and mixed with
// shared-memory loads.  Decides whether the FFT butterflies should use packed math.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp32x2 ubench_fp32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, float seed)
{
    __shared__ float2 sm[1024];
    sm[threadIdx.x] = make_float2(seed, seed * 0.5f);
    __syncthreads();
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(seed + i, seed - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(seed * 1e-9f, seed * 1e-9f);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) {  // scalar FFMA x2
                acc[i].x = fmaf(acc[i].x, m.x, c.x);
                acc[i].y = fmaf(acc[i].y, m.y, c.y);
            } else if (MODE == 1) {  // packed FFMA2
                acc[i] = __ffma2_rn(acc[i], m, c);
            } else if (MODE == 2) {  // scalar FADD x2
                acc[i].x = acc[i].x + c.x;
                acc[i].y = acc[i].y + c.y;
            } else if (MODE == 3) {  // packed FADD2
                acc[i] = __fadd2_rn(acc[i], c);
            } else if (MODE == 4) {  // scalar FFMA x2 + one LDS.64 per 4 complex ops
                acc[i].x = fmaf(acc[i].x, m.x, c.x);
                acc[i].y = fmaf(acc[i].y, m.y, c.y);
                if ((i & 3) == 0) { float2 v = sm[(idx + i) & 1023]; acc[i].x += v.x; acc[i].y += v.y; idx += 33; }
            } else if (MODE == 5) {  // packed FFMA2 + one LDS.64 per 4 complex ops
                acc[i] = __ffma2_rn(acc[i], m, c);
                if ((i & 3) == 0) { float2 v = sm[(idx + i) & 1023]; acc[i] = __fadd2_rn(acc[i], v); idx += 33; }
            } else if (MODE == 6) {  // packed FMUL2
                acc[i] = __fmul2_rn(acc[i], m);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm, double flop_per_iter_per_thread, int nsm, float* d)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<MODE><<<nsm * blocks_per_sm, threads>>>(d, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    bench<MODE><<<nsm * blocks_per_sm, threads>>>(d, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double thr = (double)nsm * blocks_per_sm * threads;
    double lane_ops = thr * ITERS * NACC * 2.0;  // scalar-lane FP ops (each complex op = 2 lanes)
    printf("%-28s thr/blk=%4d blk/SM=%d  %8.3f ms  %7.2f T lane-ops/s  %6.1f lane-ops/clk/SM@1.9GHz  (%s)\n", name, threads,
           blocks_per_sm, ms, lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / nsm / 1.9e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    float* d;
    cudaMalloc(&d, 4);
    int nsm = p.multiProcessorCount;
    for (int threads : {256, 1024}) {
        int bps = 2048 / threads;
        if (bps > 2) bps = 2;
        run<0>("scalar FFMA", threads, bps, 0, nsm, d);
        run<1>("packed FFMA2", threads, bps, 0, nsm, d);
        run<2>("scalar FADD", threads, bps, 0, nsm, d);
        run<3>("packed FADD2", threads, bps, 0, nsm, d);
        run<6>("packed FMUL2", threads, bps, 0, nsm, d);
        run<4>("scalar FFMA + LDS.64/4", threads, bps, 0, nsm, d);
        run<5>("packed FFMA2 + LDS.64/4", threads, bps, 0, nsm, d);
    }
    return 0;
}
