// stream_main.cpp -- streaming consumer for ring-fed slots (BASELINE config 3): frames are
// taken whole out of the pinned ring and rotated over the handle's lanes, so the H2D copy of
// frame i+1 runs while frame i is in the kernels and frame i-1 is copied back.  The reference
// has no counterpart: it copies one symbol at a time from pageable memory on the default stream
// with a device-wide sync after every step (ShMemSymBuff_gpu.hpp:386-387, gpuLS.cu:365-401).
//
//   stream_main --rows A --cols N --prefix C --syms S --qam b --ring L --frames F [--shm /blah]
//               [--bits-ring /name [--bits-slots n]]
// Writes Output_gpu.dat / Bits_gpu.dat and prints frames/s, antenna-samples/s and H2D GB/s.  With
// --bits-ring the packed bits of every frame also go out on a return ring (ShMemBitsBuff) for a
// downstream process.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "gpuLS.hpp"  // defines cudaEn before the ring header is seen
#include "ShMemBitsBuff.hpp"

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, qam = LSMRC_QAM_BITS, ring = 0, frames = 1;
    std::string shm = shmemID, pilots = fileNameForX, bits_ring;
    int bits_slots = 8;
    bool write_out = true;
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
        else if ((v = val("--qam"))) qam = atoi(v);
        else if ((v = val("--ring"))) ring = atoi(v);
        else if ((v = val("--frames"))) frames = atoi(v);
        else if ((v = val("--shm"))) shm = v;
        else if ((v = val("--pilots"))) pilots = v;
        else if ((v = val("--bits-ring"))) bits_ring = v;
        else if ((v = val("--bits-slots"))) bits_slots = atoi(v);
        else if (std::strcmp(argv[i], "--no-output") == 0) write_out = false;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    const int n_lanes = 3;
    if (ring <= 0) ring = (n_lanes + 1) * syms + 1;
    gpuLS ls(rows, cols, cp, syms, qam, ring, shm, 0, 0);
    if (lsmrc_set_pilot_file(ls.handle, pilots.c_str()) < 0) {
        fprintf(stderr, "pilot: %s\n", lsmrc_last_error(ls.handle));
        return 1;
    }
    const int K = cols - 1;
    const size_t comb_bytes = (size_t)(syms - 1) * K * sizeof(cuFloatComplex);
    const size_t bits_bytes = (size_t)(syms - 1) * lsmrc_bits_row_bytes(cols, qam);
    std::ofstream out, outb;
    if (write_out) {
        out.open("Output_gpu.dat", std::ofstream::binary | std::ofstream::trunc);
        outb.open("Bits_gpu.dat", std::ofstream::binary | std::ofstream::trunc);
    }
    ShMemBitsBuff* ret = bits_ring.empty() ? nullptr : new ShMemBitsBuff(bits_ring, 1, bits_bytes, bits_slots);
    auto collect = [&](int lane) {
        const void *comb = nullptr, *bits = nullptr;
        if (lsmrc_ring_wait(ls.handle, lane, &comb, &bits, nullptr) < 0) {
            fprintf(stderr, "ring_wait: %s\n", lsmrc_last_error(ls.handle));
            exit(1);
        }
        if (write_out) {
            out.write(static_cast<const char*>(comb), (std::streamsize)comb_bytes);
            outb.write(static_cast<const char*>(bits), (std::streamsize)bits_bytes);
        }
        if (ret && !ret->writeFrame(static_cast<const uint8_t*>(bits))) {
            fprintf(stderr, "return ring: the reader has gone away\n");
            exit(1);
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    // slots of frame f may only be released once its H2D copy has finished; do that lazily,
    // one frame behind, so the copy engine always has the next frame queued
    int pending_release = -1;
    for (int f = 0; f < frames; ++f) {
        const int lane = f % n_lanes;
        if (f >= n_lanes) collect(lane);
        const complexF *first = nullptr, *second = nullptr;
        int n_first = 0;
        // frame f sits behind the not-yet-released frame f-1 in the ring
        while (ls.buffPtr->available() < (pending_release >= 0 ? 2 : 1) * syms) sched_yield();
        if (pending_release >= 0) {
            lsmrc_ring_copy_done(ls.handle, pending_release);
            ls.buffPtr->releaseSlots(syms);
        }
        ls.buffPtr->waitFrame(syms, &first, &n_first, &second);
        if (lsmrc_ring_submit_split(ls.handle, lane, first, n_first, second) < 0) {
            fprintf(stderr, "ring_submit: %s\n", lsmrc_last_error(ls.handle));
            return 1;
        }
        pending_release = lane;
    }
    if (pending_release >= 0) {
        lsmrc_ring_copy_done(ls.handle, pending_release);
        ls.buffPtr->releaseSlots(syms);
    }
    for (int f = (frames > n_lanes ? frames - n_lanes : 0); f < frames; ++f) collect(f % n_lanes);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double samples = (double)frames * syms * rows * (cols + cp);
    delete ret;  // unlinks the name; a reader that is still draining keeps its mapping
    char plan[256] = "";
    lsmrc_describe_plan(ls.handle, plan, sizeof plan);
    printf("{\"frames\": %d, \"seconds\": %.6f, \"frames_per_s\": %.2f, \"antenna_samples_per_s\": %.4e, \"h2d_gbs\": %.3f, \"plan\": \"%s\"}\n",
           frames, dt, frames / dt, samples / dt, samples * 8.0 / dt / 1e9, plan);
    return 0;
}
