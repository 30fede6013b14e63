// fft_radix.cuh -- in-register radix-R DFT butterflies (R = 2..32) for sm_100a.
//
// The receive path needs an unnormalised forward DFT per antenna row
// (reference: cpuLS.hpp:165-174 fftOneRow / gpuLS.cu:377-380 cuFFT C2C forward).
// A row transform is decomposed (lsmrc_kernels.cu) into register-resident
// sub-transforms of R points per thread, joined by shared-memory exchanges.
// This header holds the register part: a fully unrolled decimation-in-frequency
// radix-2 recursion whose twiddles are compile-time constants, so every
// multiply folds to an FFMA/FMUL with an immediate operand and nothing spills.
//
// fft_dif<R>(v): in place; X[k] ends up at v[brev<R>(k)] (bit-reversed slot).
#pragma once
#include <cuda_runtime.h>

namespace lsmrc {

__host__ __device__ constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }

template <int R>
__host__ __device__ constexpr int brev(int k)
{
    int r = 0;
    for (int b = 0; b < ilog2c(R); ++b)
        if (k & (1 << b)) r |= 1 << (ilog2c(R) - 1 - b);
    return r;
}

// cos/sin(2*pi*j/32), j = 0..15, rounded once from double
__device__ constexpr float kCos32[16] = {
    1.f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f, 0.382683426f,
    0.195090324f, 0.f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f,
    -0.831469595f, -0.923879504f, -0.980785251f};
__device__ constexpr float kSin32[16] = {
    0.f, 0.195090324f, 0.382683426f, 0.555570245f, 0.707106769f, 0.831469595f, 0.923879504f,
    0.980785251f, 1.f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f,
    0.382683426f, 0.195090324f};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

// v * exp(-2*pi*i*j32/32) for a compile-time-foldable j32 in [0,16)
__device__ __forceinline__ float2 mul_w32(float2 v, int j32)
{
    if (j32 == 0) return v;
    if (j32 == 8) return make_float2(v.y, -v.x);  // * (-i)
    if (j32 == 4) {
        const float c = 0.707106769f;  // (1 - i)/sqrt(2)
        return make_float2((v.x + v.y) * c, (v.y - v.x) * c);
    }
    if (j32 == 12) {
        const float c = 0.707106769f;  // (-1 - i)/sqrt(2)
        return make_float2((v.y - v.x) * c, -(v.x + v.y) * c);
    }
    const float wr = kCos32[j32], wi = -kSin32[j32];
    return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
}

template <int R>
__device__ __forceinline__ void fft_dif(float2* v)
{
    if constexpr (R >= 2) {
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            const float2 a = v[j], b = v[j + R / 2];
            v[j] = cadd(a, b);
            v[j + R / 2] = mul_w32(csub(a, b), j * (32 / R));
        }
        fft_dif<R / 2>(v);
        fft_dif<R / 2>(v + R / 2);
    }
}

}  // namespace lsmrc
