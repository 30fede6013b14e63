// lsmrc_kernels.cuh -- fused uplink-receiver kernels for sm_100a.
//
// One kernel template, two modes, replaces the reference's whole per-symbol
// chain of library calls and tiny kernels (gpuLS.cu):
//
//   MODE_PILOT : CP strip -> N-pt FFT -> drop DC -> LS divide by the pilot ->
//                conj -> store Hconj[a][k] and sum_a |H|^2
//                (dropPrefix :143 + cuFFT :377 + findHs :158 + findDistSqrd :185;
//                 CPU: cpuLS.hpp:247-317)
//   MODE_DATA  : CP strip -> N-pt FFT -> drop DC -> x conj(H) -> sum over
//                antennas in registers -> / sum|H|^2 -> ascending-frequency
//                store -> hard QAM demap -> packed bits
//                (cuFFT :441 + multiplyWithChannelConj :212 + combineForMRC :236
//                 + shiftOneRow :109; CPU: cpuLS.hpp:319-389; demap is new)
//
// Decomposition of one N-point row transform (Cooley-Tukey, decimation in
// frequency, N = P * R2 * R3):
//   a "team" of T = N/P threads owns a row; every thread keeps P points in
//   registers.  Stage 1 loads x[n1*T + t] straight from global memory (64-bit
//   coalesced loads that skip the cyclic prefix), runs a P-point register DFT
//   over n1, applies W_N^(t*k1) and writes row k1 of a [P][T+1] shared tile.
//   Stage 2 reads the tile transposed (lanes run over k1, stride T+1 complex:
//   conflict free), runs R2-point DFTs; for three-stage plans the result goes
//   back in place with W_(T)^(m2*k2) applied and stage 3 runs R3-point DFTs.
//   After the last stage lane order equals bin order (bin = c + (N/R_last)*j),
//   so the Hconj reads, the combined-symbol stores and the MRC accumulators
//   are all coalesced / register resident: the antenna reduction never leaves
//   the register file (no shared or global intermediate, unlike
//   gpuLS.cu:212-259 which writes and re-reads the full [S][A][K] tensor).
//
// N = 1024 uses P = 32, R2 = 32: one warp per row, one __syncwarp per row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_radix.cuh"

namespace lsmrc {

enum { MODE_PILOT = 0, MODE_DATA = 1 };

struct KernelParams {
    // input: antenna-samples, complex64.  element (f, s, a, n) at
    //   rx + f*frame_stride + s*sym_stride + a*ant_stride + n   (n includes the CP)
    const float2* rx;
    long long frame_stride;
    long long sym_stride;
    int ant_stride;
    int cp;
    int first_sym;  // symbol index of work item 0 inside a frame (0 pilot, 1 first data)
    int n_ant;      // A
    int n_sym_work; // symbols handled per frame by this launch (1 for pilot, S-1 for data)
    int n_frames;
    int qam_bits;
    // per-frame state: Hconj [F][A][K] (conj of the LS estimate, bin order) and sum|H|^2 [F][K]
    float2* hconj;
    float* hsqrd;
    const float2* pilot_bin;  // X in FFT-bin order, K entries (bin k+1 at index k)
    // outputs of MODE_DATA
    float2* combined;  // [F][n_sym_work][K], ascending frequency
    uint8_t* bits;     // [F][n_sym_work][row_bytes] or nullptr
    int bits_row_bytes;
    const float2* twiddles;  // plan table: tw1 [(P-1)][T] then tw2 [(R2-1)][R3]
};

template <int N_, int P_, int R2_, int R3_, int TEAMS_, int NBUF_ = 2>
struct Plan {
    static constexpr int N = N_, P = P_, R2 = R2_, R3 = R3_, TEAMS = TEAMS_, NBUF = NBUF_;
    static constexpr int T = N / P;        // threads per team == M1 (points per row of the tile)
    static constexpr int ROW = T + 1;      // padded tile row (complex elements)
    static constexpr int NB2 = P / R2;     // stage-2 butterflies per thread
    static constexpr int NB3 = P / R3;     // stage-3 butterflies per thread (R3 > 1)
    static constexpr int RL = (R3 > 1) ? R3 : R2;  // radix of the last stage
    static constexpr int NBL = P / RL;             // last-stage butterflies per thread
    static constexpr int THREADS = T * TEAMS;
    static constexpr int TW1 = (P - 1) * T;
    static constexpr int TW2 = (R3 > 1) ? (R2 - 1) * R3 : 0;
    static constexpr int TWN = TW1 + TW2;
    static constexpr int TILE = P * ROW;   // complex elements per tile
    static constexpr size_t SMEM_BYTES = sizeof(float2) * (size_t)(TWN + TEAMS * NBUF * TILE);
    static_assert(P * R2 * R3 == N, "plan must factor N");
    static_assert(P >= R2 && P >= R3, "thread must own whole butterflies");
    static_assert(THREADS <= 1024, "block too large");
    static_assert(T <= 32 || TEAMS <= 15, "named barriers 1..15");
};

// streaming 64-bit load of one antenna-sample: read once, keep it out of L1
__device__ __forceinline__ float2 ld_stream(const float2* p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

template <class PL>
__device__ __forceinline__ void team_sync(int team)
{
    if constexpr (PL::T <= 32) {
        __syncwarp();
    } else if constexpr (PL::TEAMS == 1) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(PL::T) : "memory");
    }
}

// Hard decision, Gray-mapped square QAM (3GPP TS 38.211 5.1.3-5.1.5, unit average
// power); identical constants and comparisons to oracle/cpuls_oracle.c demap_one().
__device__ __forceinline__ unsigned demap_symbol(float re, float im, int qam_bits)
{
    const float t16 = (float)0.6324555320336759;
    const float t4 = (float)0.6172133998483676;
    const float t2 = (float)0.3086066999241838;
    unsigned v = 0;
    const float are = fabsf(re), aim = fabsf(im);
    if (re < 0.0f) v |= 1u;
    if (im < 0.0f) v |= 2u;
    if (qam_bits == 4) {
        if (are > t16) v |= 4u;
        if (aim > t16) v |= 8u;
    } else if (qam_bits == 6) {
        if (are > t4) v |= 4u;
        if (aim > t4) v |= 8u;
        if (fabsf(__fsub_rn(are, t4)) > t2) v |= 16u;
        if (fabsf(__fsub_rn(aim, t4)) > t2) v |= 32u;
    }
    return v;
}

// One N-point forward DFT of the row starting at `x` (CP already skipped), the
// team's P*T outputs handed to `sink(slot, bin, value)` where slot in [0,P) is
// the thread-local accumulator index and bin = c + (N/RL)*j is the FFT bin.
template <class PL, class Sink>
__device__ __forceinline__ void row_fft(const float2* __restrict__ x, float2* __restrict__ tile,
                                        const float2* __restrict__ s_tw1,
                                        const float2* __restrict__ s_tw2, int t, int team, Sink&& sink)
{
    constexpr int P = PL::P, T = PL::T, ROW = PL::ROW, R2 = PL::R2, R3 = PL::R3;
    {
        float2 v[P];
#pragma unroll
        for (int n1 = 0; n1 < P; ++n1) v[n1] = ld_stream(x + n1 * T + t);
        fft_dif<P>(v);
#pragma unroll
        for (int k1 = 0; k1 < P; ++k1) {
            float2 val = v[brev<P>(k1)];
            if (k1 > 0) val = cmul(val, s_tw1[(k1 - 1) * T + t]);
            tile[k1 * ROW + t] = val;
        }
    }
    team_sync<PL>(team);
#pragma unroll
    for (int i = 0; i < PL::NB2; ++i) {
        const int b = t + T * i;
        const int k1 = b % P, m2 = b / P;
        float2 u[R2];
        float2* col = tile + k1 * ROW + m2;
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) u[n2] = col[n2 * R3];
        fft_dif<R2>(u);
        if constexpr (R3 == 1) {
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) sink(i * R2 + k2, b + (PL::N / R2) * k2, u[brev<R2>(k2)]);
        } else {
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                float2 val = u[brev<R2>(k2)];
                if (k2 > 0) val = cmul(val, s_tw2[(k2 - 1) * R3 + m2]);
                col[k2 * R3] = val;
            }
        }
    }
    if constexpr (R3 > 1) {
        team_sync<PL>(team);
#pragma unroll
        for (int i = 0; i < PL::NB3; ++i) {
            const int c = t + T * i;
            const int k1 = c % P, k2 = c / P;
            float2 w[R3];
            const float2* src = tile + k1 * ROW + k2 * R3;
#pragma unroll
            for (int m2 = 0; m2 < R3; ++m2) w[m2] = src[m2];
            fft_dif<R3>(w);
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) sink(i * R3 + k3, c + (PL::N / R3) * k3, w[brev<R3>(k3)]);
        }
    }
    if constexpr (PL::NBUF == 1) team_sync<PL>(team);
}

template <class PL, int MODE, int MINB>
__global__ void __launch_bounds__(PL::THREADS, MINB) lsmrc_kernel(const KernelParams p)
{
    constexpr int N = PL::N, P = PL::P, T = PL::T, K = N - 1;
    extern __shared__ float2 smem[];
    float2* s_tw1 = smem;
    float2* s_tw2 = smem + PL::TW1;
    float2* s_tiles = smem + PL::TWN;

    for (int i = threadIdx.x; i < PL::TWN; i += PL::THREADS) smem[i] = p.twiddles[i];
    __syncthreads();

    const int team = threadIdx.x / T;
    const int t = threadIdx.x % T;
    float2* my_tiles = s_tiles + team * (PL::NBUF * PL::TILE);

    if constexpr (MODE == MODE_PILOT) {
        // one CTA per frame; antennas dealt round-robin to the teams
        const int f = blockIdx.x;
        const float2* x0 = p.rx + (long long)f * p.frame_stride + (long long)p.first_sym * p.sym_stride + p.cp;
        float2* hc_frame = p.hconj + (long long)f * p.n_ant * K;
        float e[P];
        float2 xp[P];   // pilot value per owned bin
        float xden[P];  // |X|^2, hoisted: cpuLS.hpp:240-241 recomputes it per element
#pragma unroll
        for (int sl = 0; sl < P; ++sl) {
            const int i = sl / PL::RL, j = sl % PL::RL;
            const int bin = t + T * i + (N / PL::RL) * j;
            e[sl] = 0.f;
            xp[sl] = p.pilot_bin[bin > 0 ? bin - 1 : 0];
            xden[sl] = xp[sl].x * xp[sl].x + xp[sl].y * xp[sl].y;
        }
        const int n_iter = (p.n_ant + PL::TEAMS - 1) / PL::TEAMS;
        for (int it = 0; it < n_iter; ++it) {
            const int a_raw = it * PL::TEAMS + team;
            const bool a_ok = a_raw < p.n_ant;
            const int a = a_ok ? a_raw : p.n_ant - 1;
            float2* tile = my_tiles + (PL::NBUF == 2 ? (it & 1) * PL::TILE : 0);
            float2* hc_row = hc_frame + (long long)a * K;
            row_fft<PL>(x0 + (long long)a * p.ant_stride, tile, s_tw1, s_tw2, t, team,
                        [&](int sl, int bin, float2 z) {
                            // LS estimate, naive complex division of cpuLS.hpp:233-244, then conj (:303-307)
                            const float2 X = xp[sl];
                            const float re = (z.x * X.x + z.y * X.y) / xden[sl];
                            const float im = (z.y * X.x - z.x * X.y) / xden[sl];
                            if (a_ok && bin > 0) {
                                hc_row[bin - 1] = make_float2(re, -im);
                                e[sl] += re * re + im * im;  // cpuLS.hpp:211-228
                            }
                        });
        }
        // deterministic cross-team sum of the energy partials
        __syncthreads();
        float* s_e = reinterpret_cast<float*>(s_tiles);  // [TEAMS][N]
#pragma unroll
        for (int sl = 0; sl < P; ++sl) {
            const int i = sl / PL::RL, j = sl % PL::RL;
            const int bin = t + T * i + (N / PL::RL) * j;
            s_e[team * N + bin] = e[sl];
        }
        __syncthreads();
        const int n_live = p.n_ant < PL::TEAMS ? p.n_ant : PL::TEAMS;
        for (int bin = 1 + threadIdx.x; bin < N; bin += PL::THREADS) {
            float acc = s_e[bin];
            for (int tm = 1; tm < n_live; ++tm) acc += s_e[tm * N + bin];
            p.hsqrd[(long long)f * K + bin - 1] = acc;
        }
    } else {
        // one team per (frame, data symbol); loop over all antennas, accumulate in registers
        const long long n_work = (long long)p.n_frames * p.n_sym_work;
        long long work = (long long)blockIdx.x * PL::TEAMS + team;
        const bool valid = work < n_work;
        if (!valid) work = n_work - 1;
        const int f = (int)(work / p.n_sym_work);
        const int s = (int)(work % p.n_sym_work);
        const float2* x0 = p.rx + (long long)f * p.frame_stride + (long long)(p.first_sym + s) * p.sym_stride + p.cp;
        const float2* hc_frame = p.hconj + (long long)f * p.n_ant * K;
        float2 acc[P];
#pragma unroll
        for (int sl = 0; sl < P; ++sl) acc[sl] = make_float2(0.f, 0.f);

        for (int a = 0; a < p.n_ant; ++a) {
            float2* tile = my_tiles + (PL::NBUF == 2 ? (a & 1) * PL::TILE : 0);
            const float2* hc_row = hc_frame + (long long)a * K;
            row_fft<PL>(x0 + (long long)a * p.ant_stride, tile, s_tw1, s_tw2, t, team,
                        [&](int sl, int bin, float2 y) {
                            // cpuLS.hpp:187-208: acc += Y * Hconj
                            const float2 h = __ldg(hc_row + (bin > 0 ? bin - 1 : 0));
                            acc[sl].x += y.x * h.x - y.y * h.y;
                            acc[sl].y += y.x * h.y + y.y * h.x;
                        });
        }

        // epilogue: normalise (cpuLS.hpp:364-367), reorder (cpuLS.hpp:135-149), demap, pack
        const float* e_row = p.hsqrd + (long long)f * K;
        float2* out_row = p.combined + ((long long)f * p.n_sym_work + s) * K;
        uint8_t* s_idx = reinterpret_cast<uint8_t*>(my_tiles);
        team_sync<PL>(team);  // everyone is done reading the last tile before it is reused
#pragma unroll
        for (int sl = 0; sl < P; ++sl) {
            const int i = sl / PL::RL, j = sl % PL::RL;
            const int bin = t + T * i + (N / PL::RL) * j;
            if (bin > 0) {
                const float e = e_row[bin - 1];
                const float2 o = make_float2(acc[sl].x / e, acc[sl].y / e);
                const int pos = (bin < N / 2) ? (bin - 1 + N / 2) : (bin - N / 2);
                if (valid) out_row[pos] = o;
                s_idx[pos] = (uint8_t)demap_symbol(o.x, o.y, p.qam_bits);
            }
        }
        if (p.bits != nullptr) {
            team_sync<PL>(team);
            uint8_t* bits_row = p.bits + ((long long)f * p.n_sym_work + s) * p.bits_row_bytes;
            const int b = p.qam_bits;
            for (int byte = t; byte < p.bits_row_bytes; byte += T) {
                unsigned v = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int pos = byte * 8 + q;
                    const int sym = pos / b, bit = pos - sym * b;
                    if (sym < K) v |= ((s_idx[sym] >> bit) & 1u) << q;
                }
                if (valid) bits_row[byte] = (uint8_t)v;
            }
        }
    }
}

}  // namespace lsmrc
