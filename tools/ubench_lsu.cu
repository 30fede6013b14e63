// tools/ubench_lsu.cu -- which on-chip data paths can the row FFT lean on next to shared memory?
//
// The fused receiver kernels are bound by the L1TEX/LSU data pipe (profiles/r01_ncu_full_summary.txt: 82 % of the
// shared-memory wavefront peak at N = 1024; the 2048/4096-point plans add a second exchange pass).  This
// micro-benchmark measures, per SM and per clock, on the machine it runs on:
//   lds64 / sts64      conflict-free 64-bit shared loads / stores                (warp-instructions, bytes)
//   shfl               32-bit butterfly shuffles                                   (do they cost LSU wavefronts?)
//   lds64+shfl         both interleaved: additive (same pipe) or overlapped?
//   ldtm / sttm        tcgen05.ld / tcgen05.st 32x32b.x8 (tensor memory as a per-thread scratch pad)
//   lds64+ldtm         both interleaved
//   sel                predicated selects (the conditional swaps a shuffle transpose needs)
// Build on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/bin/ubench_lsu tools/ubench_lsu.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

constexpr int kThreads = 128;
constexpr int kIters = 2048;
constexpr int kUnroll = 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

enum { T_LDS = 0, T_STS, T_SHFL, T_LDS_SHFL, T_LDTM, T_STTM, T_LDS_LDTM, T_SEL, T_FFMA2, T_FFMA2_LDS, T_FFMA2_LDTM, T_COUNT };
const char* kNames[T_COUNT] = {"lds64", "sts64", "shfl", "lds64+shfl", "ldtm.x8", "sttm.x8", "lds64+ldtm.x8", "sel", "ffma2", "ffma2+lds64", "ffma2+ldtm.x8"};

template <int TEST>
__global__ void __launch_bounds__(kThreads) k_bench(float* out, long long* cycles)
{
    __shared__ __align__(16) float2 buf[kThreads * kUnroll];
    __shared__ unsigned s_taddr;
    const int t = threadIdx.x;
    for (int i = t; i < kThreads * kUnroll; i += kThreads) buf[i] = make_float2((float)i, 1.f);
    unsigned taddr = 0;
    constexpr bool TM = (TEST == T_LDTM || TEST == T_STTM || TEST == T_LDS_LDTM || TEST == T_FFMA2_LDTM);
    if (TM) {
        if (t < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_taddr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        taddr = s_taddr + (((unsigned)(t / 32) * 32u) << 16);
        unsigned z[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = (unsigned)t + j;
        for (int c = 0; c < 64; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c), "r"(z[0]), "r"(z[1]),
                         "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7])
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    float2 acc[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) acc[j] = make_float2((float)t, (float)j);
    unsigned r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        if (TEST == T_LDS || TEST == T_LDS_SHFL || TEST == T_LDS_LDTM || TEST == T_FFMA2_LDS) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(buf + j * kThreads + t)));
                acc[j].x += v.x;
                acc[j].y += v.y;
            }
        }
        if (TEST == T_STS) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j)
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(smem_u32(buf + j * kThreads + t)), "f"(acc[j].x), "f"(acc[j].y) : "memory");
        }
        if (TEST == T_SHFL || TEST == T_LDS_SHFL) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, 1);
                acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, 2);
            }
        }
        if (TEST == T_LDTM || TEST == T_LDS_LDTM || TEST == T_FFMA2_LDTM) {
#pragma unroll
            for (int j = 0; j < kUnroll / 4; ++j) {  // 2 x (x8) = 16 words = the bytes of 8 shuffles-pairs / 8 lds64
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                             : "r"(taddr + (unsigned)((j * 8 + it) & 56)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q % kUnroll].x += __uint_as_float(r[q]);
            }
        }
        if (TEST == T_STTM) {
#pragma unroll
            for (int j = 0; j < kUnroll / 4; ++j) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + (unsigned)((j * 8 + it) & 56)),
                             "r"(__float_as_uint(acc[0].x)), "r"(__float_as_uint(acc[1].x)), "r"(__float_as_uint(acc[2].x)),
                             "r"(__float_as_uint(acc[3].x)), "r"(__float_as_uint(acc[4].x)), "r"(__float_as_uint(acc[5].x)),
                             "r"(__float_as_uint(acc[6].x)), "r"(__float_as_uint(acc[7].x))
                             : "memory");
                acc[j].x += 1.f;
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (TEST == T_SEL) {
            const bool odd = (t & 1) != 0;
#pragma unroll
            for (int j = 0; j < kUnroll; j += 2) {
                const float a = acc[j].x, b = acc[j + 1].x;
                acc[j].x = odd ? b : a;
                acc[j + 1].x = odd ? a : b;
                const float c = acc[j].y, d = acc[j + 1].y;
                acc[j].y = (it & 1) ? d : c;
                acc[j + 1].y = (it & 1) ? c : d;
                asm volatile("" : "+f"(acc[j].x), "+f"(acc[j + 1].x), "+f"(acc[j].y), "+f"(acc[j + 1].y));
            }
        }
        if (TEST == T_FFMA2 || TEST == T_FFMA2_LDS || TEST == T_FFMA2_LDTM) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < kUnroll; ++j) acc[j] = __ffma2_rn(acc[j], make_float2(0.999f, 1.001f), make_float2(0.5f, 0.25f));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * kThreads + t] = s + (float)r[0];
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
    if (TM) {
        __syncthreads();
        if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(s_taddr) : "memory");
    }
}

template <int TEST>
void run(int ctas_per_sm, int n_sms, float* d_out, long long* d_cyc)
{
    const int grid = ctas_per_sm * n_sms;
    k_bench<TEST><<<grid, kThreads>>>(d_out, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_bench<TEST><<<grid, kThreads>>>(d_out, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)std::malloc(sizeof(long long) * grid);
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    std::free(h);
    // warp-instructions of the op under test per SM per clock
    const double warps = (double)ctas_per_sm * kThreads / 32;
    double per_iter = kUnroll;  // lds / sts / shfl-pairs / sel-groups
    if (TEST == T_SHFL) per_iter = 2 * kUnroll;
    if (TEST == T_LDTM || TEST == T_STTM) per_iter = kUnroll / 4;
    if (TEST == T_FFMA2) per_iter = 4 * kUnroll;
    std::printf("%-16s ctas/SM %d: %8.0f clk per CTA, %.3f ms -> %.3f primary warp-instr/clk/SM (%.1f clk per iteration of %d)\n", kNames[TEST],
                ctas_per_sm, mean, ms, warps * per_iter * kIters / mean, mean / kIters, kUnroll);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    std::printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    float* d_out;
    long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(float) * 8 * prop.multiProcessorCount * kThreads));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * 8 * prop.multiProcessorCount));
    for (int c : {1, 3, 6}) {
        run<T_LDS>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_STS>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_SHFL>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_LDS_SHFL>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_LDTM>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_STTM>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_LDS_LDTM>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_SEL>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_FFMA2>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_FFMA2_LDS>(c, prop.multiProcessorCount, d_out, d_cyc);
        run<T_FFMA2_LDTM>(c, prop.multiProcessorCount, d_out, d_cyc);
    }
    return 0;
}
