// fft_radix.cuh -- in-register radix-R DFT butterflies (R = 2..32) for sm_100a.
//
// The receive path needs an unnormalised forward DFT per antenna row
// (reference: cpuLS.hpp:165-174 fftOneRow / gpuLS.cu:377-380 cuFFT C2C forward).
// A row transform is decomposed (lsmrc_kernels.cuh) into register-resident
// sub-transforms of R points per thread, joined by shared-memory exchanges.
// This header holds the register part.
//
// Blackwell-specific design: a complex value lives in one aligned 64-bit register
// pair (re, im) and all arithmetic is issued as packed fp32x2 instructions
// (FADD2 / FMUL2 / FFMA2, sm_100+).  Their operands take free modifiers in SASS --
// half swap (.LO_HI), per-half negate (.NP/.PN), scalar broadcast (.F32) and
// immediates -- so a complex add is ONE instruction and a complex multiply TWO,
// with no shuffling of register halves.  Packed ops run at the same lane rate as
// scalar ones (measured: tools/ubench_fp32x2.cu, 128 lane-ops/clk/SM either way),
// so the gain is issue slots: ~2.3x fewer instructions per transform, which is
// what bounds this kernel next to HBM.
//
// The transform itself is a decimation-in-time radix-2 network whose twiddled
// butterflies use the multiply-add factorisation  a +- w*b = a +- c*(b + i*tau*b)
// (tau = Im w / Re w, or the mirrored form when |Im w| > |Re w|): three packed FMAs
// per butterfly instead of a complex multiply plus two adds.  All twiddles are
// compile-time constants of W_32.
//
// fft_reg<R>(v): in place; input v[n] natural order, X[k] ends up at v[brev<R>(k)].
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

// W_32^j = exp(-2*pi*i*j/32) = kWr[j] + i*kWi[j], j = 0..15, rounded once from double;
// kTau = Wi/Wr where |Wr| >= |Wi|, kSig = Wr/Wi elsewhere (the unused entries are 0)
__device__ constexpr float kWr[16] = {1.f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f,
                                      0.382683426f, 0.195090324f, 0.f, -0.195090324f, -0.382683426f, -0.555570245f,
                                      -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f};
__device__ constexpr float kWi[16] = {0.f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f, -0.831469595f,
                                      -0.923879504f, -0.980785251f, -1.f, -0.980785251f, -0.923879504f, -0.831469595f,
                                      -0.707106769f, -0.555570245f, -0.382683426f, -0.195090324f};
__device__ constexpr float kTau[16] = {0.f, -0.198912367f, -0.414213568f, -0.668178618f, -1.f, 0.f, 0.f, 0.f,
                                       0.f, 0.f, 0.f, 0.f, 0.f, 0.668178618f, 0.414213568f, 0.198912367f};
__device__ constexpr float kSig[16] = {0.f, 0.f, 0.f, 0.f, 0.f, -0.668178618f, -0.414213568f, -0.198912367f,
                                       0.f, 0.198912367f, 0.414213568f, 0.668178618f, 1.f, 0.f, 0.f, 0.f};

__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * w, both runtime values: 2 packed instructions
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    const float2 t = __fmul2_rn(a, make_float2(w.x, w.x));
    return __ffma2_rn(swp(a), make_float2(-w.y, w.y), t);
}
// acc + a * w
__device__ __forceinline__ float2 cmac(float2 acc, float2 a, float2 w)
{
    acc = __ffma2_rn(a, make_float2(w.x, w.x), acc);
    return __ffma2_rn(swp(a), make_float2(-w.y, w.y), acc);
}

// (a, b) <- (a + w*b, a - w*b), w = W_32^j32 with j32 a compile-time-foldable value in [0,16)
__device__ __forceinline__ void bfly_dit(float2& a, float2& b, int j32)
{
    if (j32 == 0) {
        const float2 s = cadd(a, b);
        b = csub(a, b);
        a = s;
    } else if (j32 == 8) {  // w = -i : w*b = (b.y, -b.x)
        const float2 sb = swp(b);
        const float2 s = __ffma2_rn(sb, make_float2(1.f, -1.f), a);
        b = __ffma2_rn(sb, make_float2(-1.f, 1.f), a);
        a = s;
    } else if (kTau[j32] != 0.f) {  // w = c*(1 + i*tau)
        const float tau = kTau[j32], c = kWr[j32];
        const float2 u = __ffma2_rn(swp(b), make_float2(-tau, tau), b);
        const float2 s = __ffma2_rn(u, make_float2(c, c), a);
        b = __ffma2_rn(u, make_float2(-c, -c), a);
        a = s;
    } else {  // w = i*wi*(1 - i*sig)
        const float sig = kSig[j32], wi = kWi[j32];
        const float2 u = __ffma2_rn(swp(b), make_float2(sig, -sig), b);
        const float2 su = swp(u);
        const float2 s = __ffma2_rn(su, make_float2(-wi, wi), a);
        b = __ffma2_rn(su, make_float2(wi, -wi), a);
        a = s;
    }
}

template <int R>
__device__ __forceinline__ void fft_reg(float2* v)
{
    // decimation in time on the bit-reversed view w[p] = v[brev(p)]
#pragma unroll
    for (int len = 2; len <= R; len <<= 1) {
#pragma unroll
        for (int s = 0; s < R; s += len) {
#pragma unroll
            for (int k = 0; k < len / 2; ++k)
                bfly_dit(v[brev<R>(s + k)], v[brev<R>(s + k + len / 2)], k * (32 / len));
        }
    }
}

}  // namespace lsmrc
