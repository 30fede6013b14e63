// ShMemSymBuff.hpp -- the shared-memory symbol ring that feeds the receiver, same class name,
// constructor and method names as the reference's three ShMemSymBuff headers
// (ShMemSymBuff.hpp:193-484, ShMemSymBuff_cucomplex.hpp, ShMemSymBuff_gpu.hpp:123-503) and a
// byte-identical segment layout, so the reference's producer (rx_and_corr.cpp:64-87) can feed it:
//
//     struct symbolBuffer { int size; int readPtr; int writePtr; symbol symbols[len]; }
//     symbol = complexF[rows * (dimension + prefix)]                (ShMemSymBuff.hpp:92-106)
//
// i.e. a 12-byte header followed by `len` slots of rows*(dimension+prefix) complex64.
// Written from scratch; what changes, and why:
//   * dimensions are runtime constructor arguments (the -D macros numOfRows / dimension /
//     prefix / lenOfBuffer of ShMemSymBuff.hpp:42-67 remain as defaults), because one binary
//     serves all five benchmark configs;
//   * the three indices are accessed with acquire/release atomics (the reference spins on
//     plain ints shared between processes: ShMemSymBuff.hpp:215,242,248 -- undefined behaviour);
//   * flow control is a real single-producer/single-consumer ring: empty <=> writePtr==readPtr
//     (or -1 before the first write), full <=> next(writePtr)==readPtr, one slot kept free.
//     The reference cannot tell full from empty and its WithWait writer spins forever at the
//     wrap (ShMemSymBuff_gpu.hpp:471).  writePtr still means "next slot to write" and starts at
//     -1, readPtr "next slot to read" and starts at 0, so a reference NoWait producer works;
//   * frames can be consumed whole: waitFrame() hands out pointers to S consecutive slots
//     (at most two pieces when the frame wraps) for zero-copy ingest into the GPU;
//   * attach(handle) pins the mapping (cudaHostRegister through the C ABI) so that the
//     H2D copies are truly asynchronous; the reference copies from pageable shm, which makes
//     cudaMemcpyAsync synchronous (ShMemSymBuff_gpu.hpp:386-387).
#ifndef LSMRC_HOST_SHMEMSYMBUFF_HPP_
#define LSMRC_HOST_SHMEMSYMBUFF_HPP_

#include <sched.h>
#include <time.h>

#include <complex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#ifdef cudaEn  // GPU attachment compiled in only for consumers, as in the reference (gpuLS.cuh:30-32)
#include "../../include/ofdm_lsmrc.h"
#endif
#include "CSharedMemSimple.hpp"

#ifndef numOfRows
#define numOfRows 16
#endif
#ifndef dimension
#define dimension 1024
#endif
#ifndef prefix
#define prefix 0
#endif
#ifndef lenOfBuffer
#define lenOfBuffer 10
#endif
#ifndef timerEnabled
#define timerEnabled true
#endif
#ifndef testEnabled
#define testEnabled true
#endif
#define numberOfSymbolsToTest lenOfBuffer
#ifndef shmemID
#define shmemID "/blah"
#endif
#define timerEn timerEnabled
#define testEn testEnabled

struct complexF {  // == ShMemSymBuff.hpp:86-89
    float real;
    float imag;
};

class ShMemSymBuff {
   public:
    // reference signature (ShMemSymBuff.hpp:202): dimensions from the macros
    ShMemSymBuff(std::string shm_uid, int isMaster) : ShMemSymBuff(shm_uid, isMaster, numOfRows, dimension, prefix, lenOfBuffer) {}

    // runtime dimensions: rows = antennas, dim = FFT size, cp = prefix, len = ring slots
    ShMemSymBuff(std::string shm_uid, int isMaster, int rows, int dim, int cp, int len)
        : rows_(rows), dim_(dim), cp_(cp), len_(len), master_(isMaster == 1)
    {
        slot_elems_ = (size_t)rows_ * (size_t)(dim_ + cp_);
        shm_ = new CSharedMemSimple(shm_uid, segmentBytes(rows, dim, cp, len));
        hdr_ = static_cast<int*>(shm_->ptr());
        slots_ = reinterpret_cast<complexF*>(hdr_ + 3);
        times_.assign(4, std::vector<float>((size_t)len_, 0.f));
        if (master_) {
            shm_->set_master_mode();
            store(kRead, 0);
            store(kWrite, -1);
            store(kSize, len_);
        } else {
            while (load(kSize) <= 0) relax();  // wait for the master (ShMemSymBuff.hpp:215)
        }
    }
    ShMemSymBuff(const ShMemSymBuff&) = delete;
    ShMemSymBuff& operator=(const ShMemSymBuff&) = delete;

    ~ShMemSymBuff()
    {
#ifdef cudaEn
        detach();
#endif
        if (!master_) store(kSize, -1);  // tell the master the reader is gone (ShMemSymBuff.hpp:227-229)
        delete shm_;
    }

    static size_t segmentBytes(int rows, int dim, int cp, int len)
    {
        return 3 * sizeof(int) + (size_t)len * (size_t)rows * (size_t)(dim + cp) * sizeof(complexF);
    }

    void info() { shm_->info(); }
    void setBuffLen(int n)  // ShMemSymBuff_cucomplex.hpp:251
    {
        if (master_ && n > 0 && n <= len_) store(kSize, n);
    }
    int rows() const { return rows_; }
    int fftSize() const { return dim_; }
    int prefixLen() const { return cp_; }
    int slots() const { return len_; }
    size_t slotElems() const { return slot_elems_; }
    size_t slotBytes() const { return slot_elems_ * sizeof(complexF); }
    bool readerGone() { return load(kSize) == -1; }

    // ---- producer side ---------------------------------------------------------------------
    // Blocks while the ring is full (ShMemSymBuff.hpp:437-462, minus the wrap deadlock).
    template <typename T>
    void writeNextSymbolWithWait(T* Yf)
    {
        int w = load(kWrite);
        if (w < 0) w = 0;
        const int nxt = next(w);
        while (nxt == load(kRead)) {
            if (load(kSize) == -1) return;  // the reader has gone away: drop the symbol instead of spinning
            relax();
        }
        std::memcpy(slot(w), Yf, slotBytes());
        store(kWrite, nxt);
    }
    // Two-step form of writeNextSymbolWithWait for producers that fill the slot themselves (e.g. with
    // several threads, or straight from a DMA engine): wait for a free slot, fill it, publish it.
    complexF* acquireWriteSlot()
    {
        int w = load(kWrite);
        if (w < 0) w = 0;
        const int nxt = next(w);
        while (nxt == load(kRead)) {
            if (load(kSize) == -1) return nullptr;
            relax();
        }
        return slot(w);
    }
    // Producers that keep several slots in flight (ring_feeder's PipelinedWriter) address slots directly and check
    // the read index themselves; slots are still published one by one, in order, with commitWriteSlot().
    complexF* slotAt(int i) { return slot(i); }
    int readIndex() const { return load(kRead); }
    void commitWriteSlot()
    {
        int w = load(kWrite);
        if (w < 0) w = 0;
        store(kWrite, next(w));
    }
    // Never blocks; overruns the reader if it is slow (ShMemSymBuff.hpp:464-482) -- the
    // behaviour rx_and_corr.cpp:83 relies on.
    template <typename T>
    void writeNextSymbolNoWait(T* Yf)
    {
        int w = load(kWrite);
        if (w < 0) w = 0;
        std::memcpy(slot(w), Yf, slotBytes());
        store(kWrite, next(w));
    }

    // ---- consumer side, one symbol at a time --------------------------------------------------
    // Copies the next slot into Y and strips the cyclic prefix: Y[a][n] = slot[a][n+prefix]
    // (ShMemSymBuff.hpp:237-295).  `it` indexes the timing arrays only.
    template <typename T>
    void readNextSymbol(T* Y, int it)
    {
        const complexF* src = waitSlot();
        const clock_t t0 = clock();
        stripInto(reinterpret_cast<complexF*>(Y), src);
        addTime(kDropT, it, t0);
        release(1);
    }
    template <typename T>
    void readLastSymbol(T* Y)  // same thing: the fixed protocol needs no special last read (:297-331)
    {
        readNextSymbol(Y, len_ - 1);
    }
#ifdef cudaEn
    // Next slot, CP included, onto the device (ShMemSymBuff_gpu.hpp:373-406).  Needs attach().
    template <typename T>
    void readNextSymbolCUDA(T* dY, int it)
    {
        const complexF* src = waitSlot();
        const clock_t t0 = clock();
        if (handle_) lsmrc_copy_to_device(handle_, dY, src, slotBytes());
        addTime(kReadT, it, t0);
        release(1);
    }
    template <typename T>
    void readLastSymbolCUDA(T* dY)
    {
        readNextSymbolCUDA(dY, len_ - 1);
    }
#endif

    // ---- consumer side, whole frames (new) ------------------------------------------------------
    // Blocks until n_sym consecutive slots are readable; first/n_first/second describe them
    // (second is null unless the frame wraps past the end of the ring).  Nothing is copied.
    void waitFrame(int n_sym, const complexF** first, int* n_first, const complexF** second)
    {
        while (available() < n_sym) relax();
        const int r = load(kRead);
        const int until_end = len_ - r;
        *first = slot(r);
        *n_first = n_sym < until_end ? n_sym : until_end;
        *second = (*n_first < n_sym) ? slot(0) : nullptr;
    }
    // The same for a frame that sits `skip` slots behind the read pointer, i.e. behind frames that have been handed
    // out but not released yet (several frames in flight on the GPU at once).
    void waitFrameAt(int skip, int n_sym, const complexF** first, int* n_first, const complexF** second)
    {
        while (available() < skip + n_sym) relax();
        const int r = (load(kRead) + skip) % len_;
        const int until_end = len_ - r;
        *first = slot(r);
        *n_first = n_sym < until_end ? n_sym : until_end;
        *second = (*n_first < n_sym) ? slot(0) : nullptr;
    }
    bool frameReady(int n_sym) { return available() >= n_sym; }
    void releaseSlots(int n) { release(n); }
    int available()
    {
        const int w = load(kWrite);
        if (w < 0) return 0;
        const int r = load(kRead);
        return (w - r + len_) % len_;
    }
    // direct view of the next readable slot (blocks until there is one)
    const complexF* peekSlot() { return waitSlot(); }

#ifdef cudaEn
    // ---- GPU attachment ---------------------------------------------------------------------------
    // Pins the whole mapping so slots can be DMA'd straight out of shared memory.
    int attach(lsmrc_handle h)
    {
        detach();
        const int rc = lsmrc_host_register(h, shm_->ptr(), shm_->nBytes());
        if (rc == LSMRC_OK) {
            handle_ = h;
            pinned_ = true;
        } else {
            handle_ = h;  // still usable: copies fall back to staged transfers
        }
        return rc;
    }
    void detach()
    {
        if (pinned_ && handle_) lsmrc_host_unregister(handle_, shm_->ptr());
        pinned_ = false;
        handle_ = nullptr;
    }
#endif
    // stream helpers of ShMemSymBuff_gpu.hpp:364-371: streams live inside the lsmrc handle now
    void createStream(int) {}
    void destroyStream(int) {}

    // ---- timing, same arrays and file format as ShMemSymBuff_gpu.hpp:113-119,157-257 -------------
    void setReadT(float t, int it) { bump(kReadT, it, t); }
    void setDecode(float t, int it) { bump(kDecodeT, it, t); }
    void setDrop(float t, int it) { bump(kDropT, it, t); }
    void setFft(float t, int it) { bump(kFftT, it, t); }
    void setNumTimes(int n) { num_times_ = n > 0 ? n : 1; }

    void printTimes(bool cpu)
    {
        float avg[4], var[4];
        for (int k = 0; k < 4; ++k) avgVar(times_[(size_t)k], k == kDecodeT ? 1 : 0, &avg[k], &var[k]);
        printf("\t \t Avg Time(s) \t Variance (s^2) \n");
        printf("Read: \t \t %e \t %e \n", avg[kReadT] / num_times_, var[kReadT] / num_times_);
        printf("ChanEst: \t %e \n", times_[kDecodeT][0] / num_times_);
        printf("Decode: \t %e \t %e \n", avg[kDecodeT] / num_times_, var[kDecodeT] / num_times_);
        printf("FFT: \t \t %e \t %e \n", avg[kFftT] / num_times_, var[kFftT] / num_times_);
        if (cpu) printf("Drop: \t \t %e \t %e \n", avg[kDropT] / num_times_, var[kDropT] / num_times_);
    }
    // five raw floats: read, chanEst, decode, FFT, drop  (ShMemSymBuff.hpp:183-188)
    void storeTimes(bool cpu)
    {
        float avg[4], var[4];
        for (int k = 0; k < 4; ++k) avgVar(times_[(size_t)k], k == kDecodeT ? 1 : 0, &avg[k], &var[k]);
        const float rec[5] = {avg[kReadT] / num_times_, times_[kDecodeT][0] / num_times_, avg[kDecodeT] / num_times_,
                              avg[kFftT] / num_times_, avg[kDropT] / num_times_};
        std::ofstream f(cpu ? "time_cpu.dat" : "time_gpu.dat", std::ofstream::binary);
        f.write(reinterpret_cast<const char*>(rec), sizeof(rec));
    }

   private:
    enum { kSize = 0, kRead = 1, kWrite = 2 };
    enum { kReadT = 0, kDecodeT = 1, kDropT = 2, kFftT = 3 };

    int load(int i) const { return __atomic_load_n(hdr_ + i, __ATOMIC_ACQUIRE); }
    void store(int i, int v) { __atomic_store_n(hdr_ + i, v, __ATOMIC_RELEASE); }
    static void relax() { sched_yield(); }
    int next(int i) const { return (i + 1 >= len_) ? 0 : i + 1; }
    complexF* slot(int i) { return slots_ + (size_t)i * slot_elems_; }

    const complexF* waitSlot()
    {
        for (;;) {
            const int w = load(kWrite);
            if (w >= 0 && w != load(kRead)) break;
            relax();
        }
        return slot(load(kRead));
    }
    void release(int n) { store(kRead, (load(kRead) + n) % len_); }
    void stripInto(complexF* Y, const complexF* src)
    {
        for (int a = 0; a < rows_; ++a)
            std::memcpy(Y + (size_t)a * dim_, src + (size_t)a * (dim_ + cp_) + cp_, (size_t)dim_ * sizeof(complexF));
    }
    void bump(int which, int it, float t)
    {
        if (it >= 0 && it < len_) times_[(size_t)which][(size_t)it] += t;
    }
    void addTime(int which, int it, clock_t t0)
    {
        if (timerEn) bump(which, it, (float)(clock() - t0) / (float)CLOCKS_PER_SEC);
    }
    static void avgVar(const std::vector<float>& v, int first, float* avg, float* var)
    {
        const int n = (int)v.size() - first;
        float a = 0.f, q = 0.f;
        for (int i = first; i < (int)v.size(); ++i) a += v[(size_t)i];
        a = n > 0 ? a / n : 0.f;
        for (int i = first; i < (int)v.size(); ++i) q += (v[(size_t)i] - a) * (v[(size_t)i] - a);
        *avg = a;
        *var = n > 0 ? q / n : 0.f;
    }

    int rows_, dim_, cp_, len_;
    bool master_;
    size_t slot_elems_ = 0;
    CSharedMemSimple* shm_ = nullptr;
    int* hdr_ = nullptr;
    complexF* slots_ = nullptr;
#ifdef cudaEn
    lsmrc_handle handle_ = nullptr;
    bool pinned_ = false;
#endif
    int num_times_ = 1;
    std::vector<std::vector<float>> times_;
};

#endif
