// gpuLS.hpp -- host-side facade with the reference's GPU function surface: `class gpuLS`
// with the member names and argument lists of gpuLS.cuh:72-113, plus the free-function
// spellings that gpuLS_main.cu:104-141 actually calls (copyPilotToGPU, firstVector,
// demodOneSymbol, printTimes, storeTimes).  Written from scratch: every method is a thin
// forward into the C ABI (include/ofdm_lsmrc.h); no CUDA header is needed to compile a caller.
//
// Argument meaning follows the reference (rows = antennas, cols = FFT size, it = symbol index).
// Buffers: the reference makes the caller allocate dY/Y/dH/dX/Hsqrd (gpuLS_main.cu:73-91) and
// re-plans / re-allocates scratch inside every call.  Here the handle owns all device state;
// the caller's pointers are honoured where they carry results:
//   firstVector      -> dH (device, [rows][cols-1] conj(H)) and Hsqrd (device, [cols-1]) are filled;
//   demodOneSymbol   -> dY[0 .. cols-1) (host) receives the combined symbols, ascending frequency
//                       (gpuLS.cu:461-464);
//   demodOneFrame*   -> dY[0 .. (S-1)*(cols-1)) (host) receives all combined symbols (gpuLS.cu:560).
// `Y` (the reference's in-place FFT scratch) is not written: the fused kernels never materialise
// the transformed antenna rows.  rows/cols are checked against the handle; a mismatch aborts
// with a message (the reference would silently corrupt memory).
#ifndef LSMRC_HOST_GPULS_HPP_
#define LSMRC_HOST_GPULS_HPP_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifndef cudaEn
#define cudaEn  // compile the ring's GPU attachment (gpuLS.cuh:30-32 does the same)
#endif
#include "../../include/ofdm_lsmrc.h"
#include "ShMemSymBuff.hpp"

// Stand-ins so callers written against the CUDA types compile without the CUDA headers.
#if !defined(CU_COMPLEX_H_)
struct lsmrc_float2 {
    float x, y;
};
typedef lsmrc_float2 cuFloatComplex;
#endif
#if !defined(__DRIVER_TYPES_H__)
typedef void* cudaStream_t;
#endif
#if !defined(__VECTOR_TYPES_H__)
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
#endif

#ifndef fileNameForX
#define fileNameForX "Pilots.dat"  // gpuLS.cuh:56
#endif
#ifndef LSMRC_QAM_BITS
#define LSMRC_QAM_BITS 2  // demapper order; the reference has no demapper
#endif
// 0 = slave: attach to a ring somebody else created.  The reference makes this a macro called `mode`
// (gpuLS.cuh:57); a macro of that name breaks every later header with a parameter called mode (cublas_v2.h,
// which gpuLS_main.cu:37 includes next), so it is honoured when the build defines it but never defined here.
#ifdef mode
constexpr int kLsmrcRingMode = (mode);
#else
constexpr int kLsmrcRingMode = 0;
#endif

class gpuLS {
   public:
    ShMemSymBuff* buffPtr = nullptr;  // gpuLS.cuh:74
    lsmrc_handle handle = nullptr;

    // Reference constructor (gpuLS.cu:43-47): opens the ring named shmemID as slave, macro dims.
    gpuLS() : gpuLS(numOfRows, dimension, prefix, lenOfBuffer, LSMRC_QAM_BITS, lenOfBuffer, std::string(shmemID), kLsmrcRingMode, 0) {}

    // Runtime dimensions.  ring_slots == 0 -> no ring (device / host tensors only).
    gpuLS(int rows, int cols, int cp, int n_sym, int qam_bits, int ring_slots, const std::string& shm_uid,
          int is_master, int device, int n_lanes = 3, int max_frames = 1)
        : rows_(rows), cols_(cols), cp_(cp), n_sym_(n_sym), qam_bits_(qam_bits)
    {
        lsmrc_config c;
        c.n_ant = rows;
        c.fft_size = cols;
        c.cp_len = cp;
        c.n_sym = n_sym;
        c.qam_bits = qam_bits;
        c.max_frames = max_frames;
        c.device = device;
        c.n_lanes = n_lanes;
        check(lsmrc_create(&c, &handle), "lsmrc_create");
        bits_.resize(lsmrc_bits_row_bytes(cols, qam_bits) * (size_t)(n_sym > 1 ? n_sym - 1 : 1));
        if (ring_slots > 0) {
            buffPtr = new ShMemSymBuff(shm_uid, is_master, rows, cols, cp, ring_slots);
            buffPtr->attach(handle);
        }
    }
    ~gpuLS()
    {
        delete buffPtr;
        if (handle) lsmrc_destroy(handle);
    }
    gpuLS(const gpuLS&) = delete;
    gpuLS& operator=(const gpuLS&) = delete;

    // Pilots.dat -> X in FFT-bin order (gpuLS.cu:53-86).  Missing file: the CPU reference's
    // fallback 0.707+0.707i (cpuLS.hpp:85-88); the GPU reference's 1+1i disagrees with it.
    void matrix_readX(cuFloatComplex* X, int cols)
    {
        std::vector<cuFloatComplex> p((size_t)cols);
        FILE* f = std::fopen(fileNameForX, "rb");
        if (!f || std::fread(p.data(), sizeof(cuFloatComplex), p.size(), f) != p.size()) {
            fprintf(stderr, "Unable to open file %s, filling in 0.707+0.707i for x\n", fileNameForX);
            for (auto& v : p) v.x = v.y = 0.707f;
        }
        if (f) std::fclose(f);
        for (int k = 0; k < cols; ++k) X[k] = p[(size_t)((k + (cols + 1) / 2) % cols)];
    }

    // gpuLS.cu:88-106 replicated X rows*(cols-1) times on the device; the kernels broadcast one
    // copy instead.  dX (device, >= cols-1 elements) still receives X for callers that read it.
    void copyPilotToGPU(cuFloatComplex* dX, int rows, int cols)
    {
        checkDims(rows, cols);
        const int rc = lsmrc_set_pilot_file(handle, fileNameForX);
        if (rc < 0) check(rc, "lsmrc_set_pilot_file");
        if (dX) {
            std::vector<cuFloatComplex> x((size_t)(cols - 1));
            matrix_readX(x.data(), cols - 1);
            for (int r = 0; r < rows; ++r)
                check(lsmrc_copy_to_device(handle, dX + (size_t)r * (cols - 1), x.data(), x.size() * sizeof(cuFloatComplex)),
                      "copy pilot");
        }
    }
    void setPilot(const cuFloatComplex* pilot_asc, int K) { check(lsmrc_set_pilot(handle, &pilot_asc->x, K), "lsmrc_set_pilot"); }

    // odd-length roll to ascending frequency, on the host (gpuLS.cu:127-141)
    void shiftOneRowCPU(cuFloatComplex* Y, int cols, int row)
    {
        cuFloatComplex* r = Y + (size_t)row * cols;
        std::vector<cuFloatComplex> t(r, r + cols);
        for (int i = 0; i < cols; ++i) r[i] = t[(size_t)((i + (cols - 1) / 2) % cols)];
    }

    // The per-step kernel wrappers of gpuLS.cuh:87-99, for callers that drive the chain step by step
    // on their own device buffers.  (firstVector / demod* do NOT go through these: there the steps
    // are fused into two kernels.)  The reference derives the amount of work from the launch
    // geometry the caller passes; the same convention is honoured here: rows of work =
    // grid.y * block.y where the reference indexes that way, and the stream argument is accepted
    // but the work runs on the handle's stream.
    // gpuLS.cu:263-266 / kernel :109-125 -- roll every length-`cols` row to ascending frequency, in place
    void ShiftOneRow(cuFloatComplex* Y, int cols, int rows, dim3 block, dim3 grid, cudaStream_t*)
    {
        (void)rows;
        needK(cols);
        const long long n_rows = (long long)grid.y * block.y;
        void* tmp = scratch((size_t)n_rows * cols * sizeof(cuFloatComplex));
        check(lsmrc_stage_shift_rows(handle, Y, tmp, n_rows), "lsmrc_stage_shift_rows");
        check(lsmrc_copy_device(handle, Y, tmp, (size_t)n_rows * cols * sizeof(cuFloatComplex)), "lsmrc_copy_device");
    }
    // gpuLS.cu:268-271 / :143-156 -- Y[r][n] = dY[r][n + prefix]
    void DropPrefix(cuFloatComplex* Y, cuFloatComplex* dY, int rows, int cols, dim3, dim3, cudaStream_t*)
    {
        needN(cols);
        check(lsmrc_stage_drop_prefix(handle, Y, dY, rows), "lsmrc_stage_drop_prefix");
    }
    // gpuLS.cu:273-276 / :158-182 -- dH[a][k] = conj(dY[a][k+1] / dX[a][k]); dY holds the FFT of the pilot symbol
    void FindLeastSquaresGPU(cuFloatComplex* dY, cuFloatComplex* dH, cuFloatComplex* dX, int rows, int cols, dim3, dim3, cudaStream_t*)
    {
        checkDims(rows, cols);
        check(lsmrc_stage_find_hs(handle, dY, dH, dX), "lsmrc_stage_find_hs");
    }
    // gpuLS.cu:278-282 / :185-209 -- Hsqrd[k] = sum_a |H[a][k]|^2   (cols here is K = N-1, as in the reference)
    void FindHsqrdforMRC(cuFloatComplex* H, float* Hsqrd, int rows, int cols, dim3, dim3, cudaStream_t*)
    {
        checkDims(rows, cols + 1);
        check(lsmrc_stage_find_hsqrd(handle, H, Hsqrd), "lsmrc_stage_find_hsqrd");
    }
    // gpuLS.cu:284-287 / :212-233 -- Yf[s][a][k] = Y[s][a][k+1] * Hconj[a][k] for `syms` symbols
    void MultiplyWithChannelConj(cuFloatComplex* Y, cuFloatComplex* Hconj, cuFloatComplex* Yf, int rows, int cols, int syms, dim3, dim3, cudaStream_t*)
    {
        checkDims(rows, cols);
        check(lsmrc_stage_mult_conj(handle, Y, Hconj, Yf, syms), "lsmrc_stage_mult_conj");
    }
    // gpuLS.cu:289-293 / :236-259 -- Y[s*K + k] = sum_a Y[s][a][k] / Hsqrd[k], symbols = grid.y, written back into
    // the head of Y as the reference does (through a scratch buffer: the reference's in-place form races)
    void CombineForMRC(cuFloatComplex* Y, float* Hsqrd, int rows, int cols, dim3, dim3 grid, cudaStream_t*)
    {
        checkDims(rows, cols + 1);
        const int syms = (int)grid.y;
        void* tmp = scratch((size_t)syms * cols * sizeof(cuFloatComplex));
        check(lsmrc_stage_combine(handle, Y, Hsqrd, tmp, syms), "lsmrc_stage_combine");
        check(lsmrc_copy_device(handle, Y, tmp, (size_t)syms * cols * sizeof(cuFloatComplex)), "lsmrc_copy_device");
    }
    // gpuLS.cu:343-349 -- in-place forward FFT of `rows` rows of `cols` points (hand-written kernel, no cuFFT plan)
    void batchedFFT(cuFloatComplex* Y, int rows, int cols, cudaStream_t*)
    {
        needN(cols);
        check(lsmrc_stage_fft(handle, Y, rows), "lsmrc_stage_fft");
    }

    // Pilot symbol: next ring slot -> H.  gpuLS.cu:351-408.
    void firstVector(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dH, cuFloatComplex* dX, float* Hsqrd,
                     int rows, int cols, int it)
    {
        (void)dY; (void)Y; (void)dX;
        checkDims(rows, cols);
        needRing("firstVector");
        const clock_t t0 = clock();
        const complexF* slot = buffPtr->peekSlot();
        check(lsmrc_first_vector(handle, slot, 0), "lsmrc_first_vector");
        if (dH || Hsqrd) check(lsmrc_get_channel_device(handle, dH, Hsqrd), "lsmrc_get_channel_device");
        else check(lsmrc_sync(handle), "lsmrc_sync");
        buffPtr->releaseSlots(1);
        if (timerEn) buffPtr->setDecode((float)(clock() - t0) / (float)CLOCKS_PER_SEC, it);
    }

    // One data symbol: next ring slot -> combined symbols in dY[0..cols-1).  gpuLS.cu:410-473.
    void demodOneSymbol(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* Hconj, float* Hsqrd, int rows, int cols, int it)
    {
        (void)Y; (void)Hconj; (void)Hsqrd;
        checkDims(rows, cols);
        needRing("demodOneSymbol");
        const clock_t t0 = clock();
        const complexF* slot = buffPtr->peekSlot();
        check(lsmrc_demod_one_symbol(handle, slot, 0, dY, bits_.data()), "lsmrc_demod_one_symbol");
        buffPtr->releaseSlots(1);
        if (timerEn) buffPtr->setDecode((float)(clock() - t0) / (float)CLOCKS_PER_SEC, it);
    }

    // Whole frame out of the ring: S slots -> combined symbols in dY[0..(S-1)*(cols-1)).  gpuLS.cu:475-573.
    void demodOneFrame(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dX, cuFloatComplex* Hconj, float* Hsqrd,
                       int rows, int cols)
    {
        (void)Y; (void)dX; (void)Hconj; (void)Hsqrd;
        checkDims(rows, cols);
        needRing("demodOneFrame");
        const complexF *first = nullptr, *second = nullptr;
        int n_first = 0;
        buffPtr->waitFrame(n_sym_, &first, &n_first, &second);
        check(lsmrc_ring_submit_split(handle, 0, first, n_first, second), "lsmrc_ring_submit_split");
        check(lsmrc_ring_copy_done(handle, 0), "lsmrc_ring_copy_done");
        buffPtr->releaseSlots(n_sym_);  // slots are reusable as soon as the H2D copy has drained them
        const void *comb = nullptr, *bits = nullptr;
        check(lsmrc_ring_wait(handle, 0, &comb, &bits, nullptr), "lsmrc_ring_wait");
        std::memcpy(dY, comb, (size_t)(n_sym_ - 1) * (size_t)(cols - 1) * sizeof(cuFloatComplex));
        std::memcpy(bits_.data(), bits, bits_.size());
    }

    // Frame already on the device in Y as [S][rows][cols+prefix] -> dY (host).  gpuLS.cu:575-675.
    void demodOneFrameCUDA(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dX, cuFloatComplex* Hconj, float* Hsqrd,
                           int rows, int cols)
    {
        (void)dX;
        checkDims(rows, cols);
        const size_t n_out = (size_t)(n_sym_ - 1) * (size_t)(cols - 1);
        if (!d_comb_) check(lsmrc_dev_alloc(handle, n_out * sizeof(cuFloatComplex), &d_comb_), "lsmrc_dev_alloc");
        check(lsmrc_demod_frames_device(handle, Y, 1, Hconj, Hsqrd, d_comb_, nullptr), "lsmrc_demod_frames_device");
        check(lsmrc_copy_to_host(handle, dY, d_comb_, n_out * sizeof(cuFloatComplex)), "lsmrc_copy_to_host");
    }
    // gpuLS.cu:677-769 and :771-857 were alternative launch shapes of the same computation
    void demodOptimized(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dX, cuFloatComplex* Hconj, float* Hsqrd, int rows, int cols)
    {
        demodOneFrameCUDA(dY, Y, dX, Hconj, Hsqrd, rows, cols);
    }
    void demodCuBlas(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dX, cuFloatComplex* Hconj, float* Hsqrd, int rows, int cols)
    {
        demodOneFrameCUDA(dY, Y, dX, Hconj, Hsqrd, rows, cols);
    }

    // Soft output (no reference counterpart): as demodOneFrameCUDA plus max-log LLRs, llr (host) is
    // [S-1][cols-1][qamBits()] floats in ascending-frequency order, LLR > 0 <=> bit 0.
    void demodOneFrameSoftCUDA(cuFloatComplex* dY, float* llr, cuFloatComplex* Y, float noiseVar, int rows, int cols)
    {
        checkDims(rows, cols);
        const size_t n_out = (size_t)(n_sym_ - 1) * (size_t)(cols - 1);
        if (!d_comb_) check(lsmrc_dev_alloc(handle, n_out * sizeof(cuFloatComplex), &d_comb_), "lsmrc_dev_alloc");
        if (!d_llr_) check(lsmrc_dev_alloc(handle, n_out * qam_bits_ * sizeof(float), &d_llr_), "lsmrc_dev_alloc");
        check(lsmrc_demod_frames_device_soft(handle, Y, 1, d_comb_, nullptr, d_llr_, noiseVar), "lsmrc_demod_frames_device_soft");
        check(lsmrc_copy_to_host(handle, dY, d_comb_, n_out * sizeof(cuFloatComplex)), "lsmrc_copy_to_host");
        check(lsmrc_copy_to_host(handle, llr, d_llr_, n_out * qam_bits_ * sizeof(float)), "lsmrc_copy_to_host");
    }

    // demapped bits of the most recent demodOneSymbol (one row) / demodOneFrame (S-1 rows)
    const uint8_t* lastBits() const { return bits_.data(); }
    size_t bitsRowBytes() const { return lsmrc_bits_row_bytes(cols_, qamBits()); }
    int qamBits() const { return qam_bits_; }

   private:
    void check(int rc, const char* what)
    {
        if (rc < 0) {
            fprintf(stderr, "gpuLS: %s failed: %s (%s)\n", what, lsmrc_error_name(rc), lsmrc_last_error(handle));
            exit(EXIT_FAILURE);
        }
    }
    void checkDims(int rows, int cols)
    {
        if (rows != rows_ || cols != cols_) {
            fprintf(stderr, "gpuLS: called with %dx%d but the handle was created for %dx%d\n", rows, cols, rows_, cols_);
            exit(EXIT_FAILURE);
        }
    }
    void needRing(const char* who)
    {
        if (!buffPtr) {
            fprintf(stderr, "gpuLS::%s needs the symbol ring (constructed with ring_slots == 0)\n", who);
            exit(EXIT_FAILURE);
        }
    }
    void needN(int cols)
    {
        if (cols != cols_) {
            fprintf(stderr, "gpuLS: called with %d columns but the handle was created for FFT size %d\n", cols, cols_);
            exit(EXIT_FAILURE);
        }
    }
    void needK(int cols)
    {
        if (cols != cols_ - 1) {
            fprintf(stderr, "gpuLS: called with rows of %d but the handle has %d used subcarriers\n", cols, cols_ - 1);
            exit(EXIT_FAILURE);
        }
    }
    void* scratch(size_t bytes)
    {
        if (bytes > scratch_bytes_) {
            if (d_scratch_) lsmrc_dev_free(handle, d_scratch_);
            d_scratch_ = nullptr;
            check(lsmrc_dev_alloc(handle, bytes, &d_scratch_), "lsmrc_dev_alloc");
            scratch_bytes_ = bytes;
        }
        return d_scratch_;
    }
    int rows_, cols_, cp_, n_sym_, qam_bits_;
    std::vector<uint8_t> bits_;
    void* d_comb_ = nullptr;
    void* d_llr_ = nullptr;
    void* d_scratch_ = nullptr;
    size_t scratch_bytes_ = 0;
};

// ---- free-function spellings used by gpuLS_main.cu:104-141 ------------------------------------------
inline gpuLS*& lsmrc_default_gpuLS()
{
    static gpuLS* g = nullptr;
    return g;
}
inline gpuLS& lsmrc_the_gpuLS()
{
    if (!lsmrc_default_gpuLS()) lsmrc_default_gpuLS() = new gpuLS();
    return *lsmrc_default_gpuLS();
}
inline void copyPilotToGPU(cuFloatComplex* dX, int rows, int cols) { lsmrc_the_gpuLS().copyPilotToGPU(dX, rows, cols); }
inline void firstVector(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* dH, cuFloatComplex* dX, float* Hsqrd, int rows, int cols, int it)
{
    lsmrc_the_gpuLS().firstVector(dY, Y, dH, dX, Hsqrd, rows, cols, it);
}
inline void demodOneSymbol(cuFloatComplex* dY, cuFloatComplex* Y, cuFloatComplex* Hconj, float* Hsqrd, int rows, int cols, int it)
{
    lsmrc_the_gpuLS().demodOneSymbol(dY, Y, Hconj, Hsqrd, rows, cols, it);
}
inline void printTimes(bool cpu) { if (lsmrc_the_gpuLS().buffPtr) lsmrc_the_gpuLS().buffPtr->printTimes(cpu); }
inline void storeTimes(bool cpu) { if (lsmrc_the_gpuLS().buffPtr) lsmrc_the_gpuLS().buffPtr->storeTimes(cpu); }

#endif
