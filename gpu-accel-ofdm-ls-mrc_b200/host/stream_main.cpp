// stream_main.cpp -- streaming consumer for ring-fed slots (BASELINE config 3): frames are
// taken whole out of the pinned ring and rotated over the handle's lanes, so the H2D copy of
// frame i+1 runs while frame i is in the kernels and frame i-1 is copied back.  The reference
// has no counterpart: it copies one symbol at a time from pageable memory on the default stream
// with a device-wide sync after every step (ShMemSymBuff_gpu.hpp:386-387, gpuLS.cu:365-401).
//
//   stream_main --rows A --cols N --prefix C --syms S --qam b --ring L --frames F [--shm /blah]
//               [--lanes n] [--bits-ring /name [--bits-slots n]]
// --lanes: frames in flight on the GPU at once (default 3, or 8 for frames below 1 MB, which are launch-latency bound).
// The ring must hold at least lanes frames + 1 slot (default (lanes+1)*S + 1).
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
    int bits_slots = 8, n_lanes = 0;
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
        else if ((v = val("--lanes"))) n_lanes = atoi(v);
        else if (std::strcmp(argv[i], "--no-output") == 0) write_out = false;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    // default: 3 frames in flight for large frames (copy-bound), 8 for frames below 1 MB (launch-latency bound)
    if (n_lanes < 1 || n_lanes > 64) n_lanes = ((size_t)syms * rows * (cols + cp) * sizeof(complexF) < (1u << 20)) ? 8 : 3;
    if (ring <= 0) ring = (n_lanes + 1) * syms + 1;
    if (ring < n_lanes * syms + 1) {
        fprintf(stderr, "the ring (%d slots) must hold %d frames of %d slots plus one\n", ring, n_lanes, syms);
        return 2;
    }
    gpuLS ls(rows, cols, cp, syms, qam, ring, shm, 0, 0, n_lanes);
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
    // Up to n_lanes frames are in flight.  Two things trail the submissions, both in frame order: the slots of a
    // frame go back to the producer as soon as the GPU no longer reads them (its H2D copy, or the in-place kernel,
    // has finished -- polled, so the producer refills while the kernels and D2H of later stages still run), and a
    // lane's results are collected just before the lane is reused.
    int unreleased = 0, busy = 0;   // frames [f - unreleased, f) still own their slots; [f - busy, f) not collected
    auto release_oldest = [&](int f, bool block) -> bool {
        const int lane = (f - unreleased) % n_lanes;
        if (block) {
            if (lsmrc_ring_copy_done(ls.handle, lane) < 0) exit(1);
        } else if (lsmrc_ring_copy_query(ls.handle, lane) != 1) {
            return false;
        }
        ls.buffPtr->releaseSlots(syms);
        --unreleased;
        return true;
    };
    for (int f = 0; f < frames; ++f) {
        const int lane = f % n_lanes;
        if (busy == n_lanes) {
            while (unreleased == n_lanes) release_oldest(f, true);  // the lane's events are about to be re-recorded
            collect(lane);
            --busy;
        }
        while (unreleased > 0 && release_oldest(f, false)) {
        }
        const complexF *first = nullptr, *second = nullptr;
        int n_first = 0;
        while (!ls.buffPtr->frameReady((unreleased + 1) * syms)) {
            if (unreleased > 0 && release_oldest(f, false)) continue;  // a free slot may be what the producer waits for
            sched_yield();
        }
        ls.buffPtr->waitFrameAt(unreleased * syms, syms, &first, &n_first, &second);
        if (lsmrc_ring_submit_split(ls.handle, lane, first, n_first, second) < 0) {
            fprintf(stderr, "ring_submit: %s\n", lsmrc_last_error(ls.handle));
            return 1;
        }
        ++unreleased;
        ++busy;
    }
    while (unreleased > 0) release_oldest(frames, true);
    for (int f = frames - busy; f < frames; ++f) collect(f % n_lanes);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double samples = (double)frames * syms * rows * (cols + cp);
    delete ret;  // unlinks the name; a reader that is still draining keeps its mapping
    char plan[256] = "";
    lsmrc_describe_plan(ls.handle, plan, sizeof plan);
    printf("{\"frames\": %d, \"seconds\": %.6f, \"frames_per_s\": %.2f, \"antenna_samples_per_s\": %.4e, \"h2d_gbs\": %.3f, \"plan\": \"%s\"}\n",
           frames, dt, frames / dt, samples / dt, samples * 8.0 / dt / 1e9, plan);
    return 0;
}
