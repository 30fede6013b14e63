// tools/ubench_ffma2_lat.cu -- dependent-chain latency / throughput of FFMA2 vs ILP and warps per SM sub-partition
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 8192
template <int ILP, bool PACKED>
__global__ void k(float* out, float s, long long* cyc)
{
    float2 a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = make_float2(s + i, s - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(s * 1e-9f, s * 2e-9f);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
            else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i].x + a[i].y;
    if (r == 1.2345f) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP, bool PACKED>
void run(int warps_per_smsp, float* d, long long* dc)
{
    int threads = 128 * warps_per_smsp;  // 4 SMSPs
    k<ILP, PACKED><<<148, threads>>>(d, 1.f, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (ITERS * ILP);
    printf("%s ILP=%d warps/SMSP=%d : %.2f cycles per %s per warp  -> pipe busy %.0f%%\n", PACKED ? "FFMA2 " : "2xFFMA", ILP, warps_per_smsp,
           per, PACKED ? "FFMA2" : "FFMA pair", 100.0 * 2.0 * warps_per_smsp / per);
}
int main()
{
    float* d; long long* dc;
    cudaMalloc(&d, 4); cudaMalloc(&dc, 8);
    for (int w : {1, 2, 3, 4}) {
        run<1, true>(w, d, dc); run<2, true>(w, d, dc); run<4, true>(w, d, dc); run<8, true>(w, d, dc);
    }
    for (int w : {1, 3}) { run<1, false>(w, d, dc); run<4, false>(w, d, dc); }
    return 0;
}
