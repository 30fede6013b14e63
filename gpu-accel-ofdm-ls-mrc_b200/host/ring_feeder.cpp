// ring_feeder.cpp -- synthetic producer for the symbol ring, standing in for the radio
// front end: it reproduces the write pattern of rx_and_corr.cpp:64-87 (one slot per OFDM
// symbol, antenna-major rows, a frame = S consecutive slots starting at the pilot) but leaves
// the cyclic prefix IN the slot (prefix > 0 layout), so the CP strip happens on the GPU.
// Creates the segment (master), waits for a consumer, streams the frames of a file.
//
//   ring_feeder --file rx.bin --rows A --cols N --prefix C --syms S --ring L [--frames F]
//               [--shm /blah] [--repeat R] [--nowait] [--threads T] [--stream-stores] [--first-lap-only]
// Slots are filled with non-temporal stores when --threads T > 1 or --stream-stores is given (see stream_copy).
// --threads T > 1 keeps T slots in flight, one producer thread each, published in order (PipelinedWriter; a radio
// front end delivers in parallel; one memcpy thread tops out near 9 GB/s and would hide what the consumer can do).
// --first-lap-only (with --threads) writes every slot once and afterwards only publishes it again: the consumer's
// ceiling without any producer traffic in host memory (diagnostic).
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

// Several producer threads, one whole slot each: thread i takes the next unwritten symbol, waits until its slot has been
// released by the reader, fills it, and flags it; the main thread publishes slots strictly in order.  (Round 2 first
// split every symbol over the threads with a join per symbol -- a 45 us copy paid ~10 us of fork/join; whole slots in
// flight need no join at all, as the antennas of a radio front end arrive in parallel anyway.)
class PipelinedWriter {
   public:
    PipelinedWriter(ShMemSymBuff& ring, int threads, bool copy_every_lap)
        : ring_(ring), n_(threads > 1 ? threads : 1), len_(ring.slots()), every_lap_(copy_every_lap), done_((size_t)ring.slots())
    {
        for (auto& d : done_) d.store(0, std::memory_order_relaxed);
    }
    // symbol j comes from src(j); returns false when the reader went away
    template <class Src>
    bool run(long long total, Src&& src)
    {
        std::vector<std::thread> workers;
        for (int i = 0; i < n_; ++i)
            workers.emplace_back([this, total, &src] {
                for (;;) {
                    const long long j = next_.fetch_add(1, std::memory_order_relaxed);
                    if (j >= total) return;
                    if (!wait_free(j)) return;
                    if (every_lap_ || j < len_) stream_copy(ring_.slotAt((int)(j % len_)), src(j), ring_.slotBytes());
                    done_[(size_t)(j % len_)].store(j + 1, std::memory_order_release);
                }
            });
        bool ok = true;
        for (long long j = 0; j < total && ok; ++j) {
            while (done_[(size_t)(j % len_)].load(std::memory_order_acquire) != j + 1) {
                if (ring_.readerGone()) {
                    ok = false;
                    break;
                }
                sched_yield();
            }
            if (ok) {
                ring_.commitWriteSlot();
                committed_.store(j + 1, std::memory_order_release);
            }
        }
        if (!ok) {
            abort_.store(true, std::memory_order_release);
            next_.store(total, std::memory_order_relaxed);
        }
        for (auto& t : workers) t.join();
        return ok;
    }

   private:
    // slot j % len is free once the reader has released symbol j - len; one slot of the ring always stays empty
    bool wait_free(long long j)
    {
        for (;;) {
            const long long c = committed_.load(std::memory_order_acquire);
            const int r = ring_.readIndex();
            const long long in_ring = ((c % len_) - r + len_) % len_;  // == available(); read after c, so never too small
            if (j - (c - in_ring) <= len_ - 2) return true;
            if (abort_.load(std::memory_order_acquire) || ring_.readerGone()) return false;
            sched_yield();
        }
    }
    ShMemSymBuff& ring_;
    int n_;
    long long len_;
    bool every_lap_;
    std::vector<std::atomic<long long>> done_;
    std::atomic<long long> next_{0}, committed_{0};
    std::atomic<bool> abort_{false};
};

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, ring = 0, frames = -1, repeat = 1, threads = 1;
    bool nowait = false, stream_stores = false, first_lap_only = false;
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
        else if (std::strcmp(argv[i], "--first-lap-only") == 0) first_lap_only = true;
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
    if (threads > 1) {
        const long long per_lap = (long long)frames * syms;
        PipelinedWriter pw(ringbuf, threads, !first_lap_only);
        pw.run(per_lap * repeat, [&](long long j) { return buf.data() + (size_t)(j % per_lap) * slot; });
    } else {
        for (int r = 0; r < repeat; ++r)
            for (int f = 0; f < frames; ++f)
                for (int s = 0; s < syms; ++s) {
                    complexF* sym = buf.data() + ((size_t)f * syms + (size_t)s) * slot;
                    if (stream_stores) {
                        // same protocol as writeNextSymbolWithWait, with the slot filled by non-temporal stores
                        complexF* dst = ringbuf.acquireWriteSlot();
                        if (!dst) break;  // reader gone
                        stream_copy(dst, sym, ringbuf.slotBytes());
                        ringbuf.commitWriteSlot();
                    } else if (nowait) {
                        ringbuf.writeNextSymbolNoWait(sym);
                    } else {
                        ringbuf.writeNextSymbolWithWait(sym);
                    }
                }
    }
    // keep the segment alive until the reader has drained it and gone away
    while (ringbuf.available() > 0 && !ringbuf.readerGone()) sched_yield();
    return 0;
}
