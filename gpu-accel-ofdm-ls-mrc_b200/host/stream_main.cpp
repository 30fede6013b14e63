// stream_main.cpp -- streaming consumer for ring-fed slots (BASELINE config 3): frames are
// taken whole out of the pinned ring and rotated over the handle's lanes, so the H2D copy of
// frame i+1 runs while frame i is in the kernels and frame i-1 is copied back.  The reference
// has no counterpart: it copies one symbol at a time from pageable memory on the default stream
// with a device-wide sync after every step (ShMemSymBuff_gpu.hpp:386-387, gpuLS.cu:365-401).
//
//   stream_main --rows A --cols N --prefix C --syms S --qam b --ring L --frames F [--shm /blah]
//               [--lanes n] [--batch k] [--bits-ring /name [--bits-slots n]]
// --lanes: submissions in flight on the GPU at once (default 3, or 4 for frames below 1 MB).
// --batch: frames per submission when that many are already waiting in the ring (default 1, or 16 for frames below
// 1 MB).  A ring shorter than lanes*batch frames + 1 slot (the default is (lanes*batch + 1)*S + 1) simply limits how
// much of that is used.
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
    int bits_slots = 8, n_lanes = 0, batch = 0;
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
        else if ((v = val("--batch"))) batch = atoi(v);
        else if (std::strcmp(argv[i], "--no-output") == 0) write_out = false;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    // default: 3 submissions in flight for large frames (copy-bound), 4 for frames below 1 MB (launch-latency bound)
    if (n_lanes < 1 || n_lanes > 64) n_lanes = ((size_t)syms * rows * (cols + cp) * sizeof(complexF) < (1u << 20)) ? 4 : 3;
    const bool small = (size_t)syms * rows * (cols + cp) * sizeof(complexF) < (1u << 20);
    if (batch < 1 || batch > 64) batch = small ? 16 : 1;
    if (ring <= 0) ring = (n_lanes * batch + 1) * syms + 1;
    if (ring < syms + 1) {
        fprintf(stderr, "the ring (%d slots) must hold at least one frame of %d slots plus one\n", ring, syms);
        return 2;
    }
    gpuLS ls(rows, cols, cp, syms, qam, ring, shm, 0, 0, n_lanes, batch);
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
    std::vector<int> lane_n((size_t)n_lanes, 0);  // frames of the submission in flight on each lane
    auto collect = [&](int lane) {
        const void *comb = nullptr, *bits = nullptr;
        if (lsmrc_ring_wait(ls.handle, lane, &comb, &bits, nullptr) < 0) {
            fprintf(stderr, "ring_wait: %s\n", lsmrc_last_error(ls.handle));
            exit(1);
        }
        const int n = lane_n[(size_t)lane];
        if (write_out) {
            out.write(static_cast<const char*>(comb), (std::streamsize)(comb_bytes * n));
            outb.write(static_cast<const char*>(bits), (std::streamsize)(bits_bytes * n));
        }
        for (int i = 0; ret && i < n; ++i) {
            if (!ret->writeFrame(static_cast<const uint8_t*>(bits) + (size_t)i * bits_bytes)) {
                fprintf(stderr, "return ring: the reader has gone away\n");
                exit(1);
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    // Up to n_lanes submissions (of up to `batch` frames each) are in flight.  Two things trail them, both in order:
    // the slots of a submission go back to the producer as soon as the GPU no longer reads them (its H2D copy, or
    // the in-place kernel, has finished -- polled, so the producer refills while later stages still run), and a
    // lane's results are collected just before the lane is reused.
    const int ring_frames = (ring - 1) / syms;  // whole frames the ring can hold with one slot to spare
    int sub = 0, unreleased_subs = 0, busy = 0, unreleased_frames = 0;
    auto release_oldest = [&](bool block) -> bool {
        const int lane = (sub - unreleased_subs) % n_lanes;
        if (block) {
            if (lsmrc_ring_copy_done(ls.handle, lane) < 0) exit(1);
        } else if (lsmrc_ring_copy_query(ls.handle, lane) != 1) {
            return false;
        }
        ls.buffPtr->releaseSlots(lane_n[(size_t)lane] * syms);
        unreleased_frames -= lane_n[(size_t)lane];
        --unreleased_subs;
        return true;
    };
    for (int f = 0; f < frames;) {
        const int lane = sub % n_lanes;
        if (busy == n_lanes) {
            while (unreleased_subs == n_lanes) release_oldest(true);  // the lane's events are about to be re-recorded
            collect(lane);
            --busy;
        }
        while (unreleased_subs > 0 && release_oldest(false)) {
        }
        while (unreleased_frames + 1 > ring_frames) release_oldest(true);  // the ring is full of our own frames
        while (!ls.buffPtr->frameReady((unreleased_frames + 1) * syms)) {
            if (unreleased_subs > 0 && release_oldest(false)) continue;  // a free slot may be what the producer waits for
            sched_yield();
        }
        // as many whole frames as are already waiting, up to the batch size
        int nb = ls.buffPtr->available() / syms - unreleased_frames;
        if (nb > batch) nb = batch;
        if (nb > frames - f) nb = frames - f;
        if (nb < 1) nb = 1;
        const complexF *first = nullptr, *second = nullptr;
        int n_first = 0;
        ls.buffPtr->waitFrameAt(unreleased_frames * syms, nb * syms, &first, &n_first, &second);
        if (lsmrc_ring_submit_frames(ls.handle, lane, first, n_first, second, nb) < 0) {
            fprintf(stderr, "ring_submit: %s\n", lsmrc_last_error(ls.handle));
            return 1;
        }
        lane_n[(size_t)lane] = nb;
        unreleased_frames += nb;
        ++unreleased_subs;
        ++busy;
        ++sub;
        f += nb;
    }
    while (unreleased_subs > 0) release_oldest(true);
    for (int i = sub - busy; i < sub; ++i) collect(i % n_lanes);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double samples = (double)frames * syms * rows * (cols + cp);
    delete ret;  // unlinks the name; a reader that is still draining keeps its mapping
    char plan[256] = "";
    lsmrc_describe_plan(ls.handle, plan, sizeof plan);
    printf("{\"frames\": %d, \"seconds\": %.6f, \"frames_per_s\": %.2f, \"antenna_samples_per_s\": %.4e, \"h2d_gbs\": %.3f, \"plan\": \"%s\"}\n",
           frames, dt, frames / dt, samples / dt, samples * 8.0 / dt / 1e9, plan);
    return 0;
}
