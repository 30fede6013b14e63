// tools/ubench_tmem.cu -- can tensor memory serve as a per-thread constant store (FFT twiddles) next to a kernel that
// already saturates the shared-memory port?  Each warp loops over "rows": 64 packed FMAs that consume 16 complex
// per-thread constants, plus a background of 32 conflict-free 64-bit shared loads and 32 stores (the exchange traffic
// of the row FFT).  The constants come from
//   mode 0: shared memory, 16 x LDS.64 (lane-private addresses)          -- what a twiddle table costs
//   mode 1: tensor memory, 2 x tcgen05.ld.32x32b.x16 + one wait::ld       -- per-thread columns, written once
//   mode 2: nowhere (kept in registers)                                   -- lower bound
// Reported: clocks per row per SM at 12 warps/SM (3 CTAs x 128 threads).
// Build on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/bin/ubench_tmem tools/ubench_tmem.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                                        \
        }                                                                                        \
    } while (0)

constexpr int kThreads = 128;
constexpr int kRows = 4096;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE, int BG>
__global__ void __launch_bounds__(kThreads, 3) k_rows(float* out)
{
    __shared__ __align__(16) float2 buf[32 * kThreads];   // background exchange tile, 32 KB
    __shared__ __align__(16) float2 tws[16 * kThreads];   // mode 0: per-thread constants, 16 KB
    __shared__ unsigned s_taddr;
    const int t = threadIdx.x;
    for (int i = t; i < 32 * kThreads; i += kThreads) buf[i] = make_float2(1.f / (1 + i), 0.5f);
    for (int i = t; i < 16 * kThreads; i += kThreads) tws[i] = make_float2(0.999f, 0.001f * (i & 7));
    unsigned taddr = 0;
    if (MODE == 1) {
        if (t < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_taddr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        taddr = s_taddr + (((unsigned)(t / 32) * 32u) << 16);
        for (int c = 0; c < 32; c += 8) {
            unsigned z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = __float_as_uint((j & 1) ? 0.001f * ((c + j) & 7) : 0.999f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c), "r"(z[0]), "r"(z[1]),
                         "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    float2 v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = make_float2((float)t, (float)j);
    float2 creg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) creg[j] = make_float2(0.999f, 0.001f * j);
    for (int row = 0; row < kRows; ++row) {
        if (BG) {  // background: the exchange (32 STS.64 + 32 LDS.64 per thread, conflict-free)
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(smem_u32(buf + j * kThreads + t)), "f"(v[j].x), "f"(v[j].y) : "memory");
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[j].x), "=f"(v[j].y) : "r"(smem_u32(buf + j * kThreads + (t ^ 1))) : "memory");
        }
        float2 c[16];
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(c[j].x), "=f"(c[j].y) : "r"(smem_u32(tws + j * kThreads + t)) : "memory");
        } else if (MODE == 1) {
            unsigned r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr + 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = creg[j];
        }
        // 64 packed FMAs: each constant used by 2 complex multiplies (2 FFMA2-class ops each)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float2 w = c[j & 15];
            const float2 a = v[j];
            float2 m = __fmul2_rn(a, make_float2(w.x, w.x));
            v[j] = __ffma2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y), m);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j].x + v[j].y;
    out[blockIdx.x * kThreads + t] = s;
    if (MODE == 1) {
        __syncthreads();
        if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(s_taddr) : "memory");
    }
}

template <int MODE, int BG>
void run(const char* name, int n_sms, float* d_out)
{
    const int grid = 3 * n_sms;
    k_rows<MODE, BG><<<grid, kThreads>>>(d_out);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_rows<MODE, BG><<<grid, kThreads>>>(d_out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    std::printf("%-44s %.3f ms  -> %.0f ns per row per CTA (3 CTAs/SM)\n", name, ms, 1e6 * ms / kRows);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    std::printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    float* d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * 3 * prop.multiProcessorCount * kThreads));
    for (int rep = 0; rep < 2; ++rep) {
        run<2, 1>("exchange + FMAs, constants in registers", prop.multiProcessorCount, d_out);
        run<0, 1>("exchange + FMAs, constants by 16 LDS.64", prop.multiProcessorCount, d_out);
        run<1, 1>("exchange + FMAs, constants by 2 LDTM.x16", prop.multiProcessorCount, d_out);
        run<2, 0>("FMAs only, constants in registers", prop.multiProcessorCount, d_out);
        run<0, 0>("FMAs only, constants by 16 LDS.64", prop.multiProcessorCount, d_out);
        run<1, 0>("FMAs only, constants by 2 LDTM.x16", prop.multiProcessorCount, d_out);
    }
    return 0;
}
