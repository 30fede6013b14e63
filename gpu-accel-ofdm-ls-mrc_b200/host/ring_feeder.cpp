// ring_feeder.cpp -- synthetic producer for the symbol ring, standing in for the radio
// front end: it reproduces the write pattern of rx_and_corr.cpp:64-87 (one slot per OFDM
// symbol, antenna-major rows, a frame = S consecutive slots starting at the pilot) but leaves
// the cyclic prefix IN the slot (prefix > 0 layout), so the CP strip happens on the GPU.
// Creates the segment (master), waits for a consumer, streams the frames of a file.
//
//   ring_feeder --file rx.bin --rows A --cols N --prefix C --syms S --ring L [--frames F]
//               [--shm /blah] [--repeat R] [--nowait] [--threads T] [--stream-stores]
// Slots are filled with non-temporal stores when --threads T > 1 or --stream-stores is given (see stream_copy).
// --threads T > 1 copies every symbol into its slot with T helper threads (a radio front end delivers
// the antennas in parallel; one memcpy thread tops out near 9 GB/s and would hide what the consumer can do).
// rx.bin holds [F][S][A][N+C] complex64.  --nowait uses writeNextSymbolNoWait like the
// reference producer (can overrun a slow reader); the default blocks on a full ring.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "ShMemSymBuff.hpp"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// Copy into a ring slot with non-temporal stores.  The slot is written once and next read by the GPU's DMA engine,
// never by this CPU: ordinary stores would first read every destination line into the cache (a third of the memory
// traffic of the copy) and then evict what the producer actually reuses.  Slots sit behind the ring's 12-byte header,
// i.e. only 4-byte aligned, so the head and tail of the range are copied normally.
static void stream_copy(void* dst, const void* src, size_t bytes)
{
#if defined(__SSE2__)
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (bytes < 256 || head > bytes) {
        std::memcpy(d, s, bytes);
        return;
    }
    std::memcpy(d, s, head);
    d += head, s += head, bytes -= head;
    const size_t n16 = bytes / 16;
    for (size_t i = 0; i < n16; ++i)
        _mm_stream_si128(reinterpret_cast<__m128i*>(d) + i, _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + i));
    std::memcpy(d + n16 * 16, s + n16 * 16, bytes - n16 * 16);
    _mm_sfence();  // the stores are globally visible before the slot is committed to the consumer
#else
    std::memcpy(dst, src, bytes);
#endif
}

// Splits one symbol copy over a few persistent helper threads (spin-synchronised: copies are ~100 us apart).
class ParallelCopy {
   public:
    explicit ParallelCopy(int n) : n_(n > 1 ? n : 1)
    {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~ParallelCopy()
    {
        stop_.store(true, std::memory_order_release);
        gen_.fetch_add(1, std::memory_order_release);
        for (auto& t : workers_) t.join();
    }
    void copy(void* dst, const void* src, size_t bytes)
    {
        if (n_ == 1) {
            stream_copy(dst, src, bytes);
            return;
        }
        dst_ = static_cast<char*>(dst);
        src_ = static_cast<const char*>(src);
        bytes_ = bytes;
        done_.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        part(0);
        while (done_.load(std::memory_order_acquire) != n_ - 1) sched_yield();
    }

   private:
    void part(int i)
    {
        const size_t chunk = ((bytes_ + n_ - 1) / n_ + 63) & ~(size_t)63;
        const size_t b = chunk * (size_t)i, e = b + chunk < bytes_ ? b + chunk : bytes_;
        if (b < e) stream_copy(dst_ + b, src_ + b, e - b);
    }
    void loop(int i)
    {
        unsigned seen = 0;
        for (;;) {
            unsigned g;
            while ((g = gen_.load(std::memory_order_acquire)) == seen) sched_yield();
            seen = g;
            if (stop_.load(std::memory_order_acquire)) return;
            part(i);
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> done_{0};
    std::atomic<bool> stop_{false};
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
};

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, ring = 0, frames = -1, repeat = 1, threads = 1;
    bool nowait = false, stream_stores = false;
    std::string shm = shmemID, file;
    for (int i = 1; i < argc; ++i) {
        auto val = [&](const char* name) -> const char* {
            if (std::strcmp(argv[i], name) == 0 && i + 1 < argc) return argv[++i];
            return nullptr;
        };
        const char* v;
        if ((v = val("--rows"))) rows = atoi(v);
        else if ((v = val("--cols"))) cols = atoi(v);
        else if ((v = val("--prefix"))) cp = atoi(v);
        else if ((v = val("--syms"))) syms = atoi(v);
        else if ((v = val("--ring"))) ring = atoi(v);
        else if ((v = val("--frames"))) frames = atoi(v);
        else if ((v = val("--repeat"))) repeat = atoi(v);
        else if ((v = val("--threads"))) threads = atoi(v);
        else if ((v = val("--shm"))) shm = v;
        else if ((v = val("--file"))) file = v;
        else if (std::strcmp(argv[i], "--nowait") == 0) nowait = true;
        else if (std::strcmp(argv[i], "--stream-stores") == 0) stream_stores = true;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    if (ring <= 0) ring = syms + 1;
    std::ifstream in(file.c_str(), std::ifstream::binary);
    if (!in) {
        fprintf(stderr, "ring_feeder: cannot open %s\n", file.c_str());
        return 2;
    }
    const size_t slot = (size_t)rows * (size_t)(cols + cp);
    in.seekg(0, in.end);
    const size_t n_elems = (size_t)in.tellg() / sizeof(complexF);
    in.seekg(0, in.beg);
    const int in_file = (int)(n_elems / (slot * (size_t)syms));
    if (frames < 0 || frames > in_file) frames = in_file;
    std::vector<complexF> buf((size_t)frames * (size_t)syms * slot);
    in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(complexF)));

    ShMemSymBuff ringbuf(shm, /*isMaster=*/1, rows, cols, cp, ring);
    ParallelCopy pc(threads);
    for (int r = 0; r < repeat; ++r)
        for (int f = 0; f < frames; ++f)
            for (int s = 0; s < syms; ++s) {
                complexF* sym = buf.data() + ((size_t)f * syms + (size_t)s) * slot;
                if (threads > 1 || stream_stores) {
                    // same protocol as writeNextSymbolWithWait, with the slot filled by the helper threads
                    complexF* dst = ringbuf.acquireWriteSlot();
                    if (!dst) break;  // reader gone
                    pc.copy(dst, sym, ringbuf.slotBytes());
                    ringbuf.commitWriteSlot();
                } else if (nowait) {
                    ringbuf.writeNextSymbolNoWait(sym);
                } else {
                    ringbuf.writeNextSymbolWithWait(sym);
                }
            }
    // keep the segment alive until the reader has drained it and gone away
    while (ringbuf.available() > 0 && !ringbuf.readerGone()) sched_yield();
    return 0;
}
