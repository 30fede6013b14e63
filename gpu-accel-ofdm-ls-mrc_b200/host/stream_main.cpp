// stream_main.cpp -- streaming consumer for ring-fed slots (BASELINE config 3): frames are
// taken whole out of the pinned ring and rotated over the handle's lanes, so the H2D copy of
// frame i+1 runs while frame i is in the kernels and frame i-1 is copied back.  The reference
// has no counterpart: it copies one symbol at a time from pageable memory on the default stream
// with a device-wide sync after every step (ShMemSymBuff_gpu.hpp:386-387, gpuLS.cu:365-401).
//
//   stream_main --rows A --cols N --prefix C --syms S --qam b --ring L --frames F [--shm /blah]
//               [--lanes n] [--batch k] [--bits-ring /name [--bits-slots n]] [--gpus G] [--trace file]
// --gpus G: one ring (<shm>_g), one worker thread and one receiver per GPU inside this process (the
// reference is pinned to device 0, gpuLS_main.cu:69); --frames counts per GPU.
// --trace file: per submission the CUDA-event times of its H2D, kernels and D2H (lsmrc_ring_trace), as CSV.
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
#include <thread>
#include <vector>

#include "gpuLS.hpp"  // defines cudaEn before the ring header is seen
#include "ShMemBitsBuff.hpp"

struct Args {
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, qam = LSMRC_QAM_BITS, ring = 0, frames = 1;
    std::string shm = shmemID, pilots = fileNameForX, bits_ring, trace;
    int bits_slots = 8, n_lanes = 0, batch = 0, gpus = 1;
    bool write_out = true;
};
struct Result {
    int rc = 0;
    double seconds = 0;         // from the start of the consume loop (includes waiting for the producer's first frame
                                // and the first-use allocation of the lanes)
    double steady_seconds = 0;  // from the first submission
    std::string plan;
};

// run this thread on the cores next to GPU `gpu` (sysfs local_cpulist of its PCI device); best effort
static void pin_thread_near_gpu(int gpu)
{
    char bdf[64] = "";
    if (lsmrc_device_pci_bus_id(gpu, bdf, sizeof bdf) != LSMRC_OK) return;
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    std::ifstream in(std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist");
    std::string list;
    if (!in || !std::getline(in, list)) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    const char* p = list.c_str();
    while (*p) {
        char* e;
        long lo = strtol(p, &e, 10), hi = lo;
        if (e == p) break;
        if (*e == '-') hi = strtol(e + 1, &e, 10);
        for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c, ++n) CPU_SET((int)c, &set);
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',' ) break;
    }
    if (n > 0) sched_setaffinity(0, sizeof set, &set);
}

// One ring -> one GPU: frames are taken whole out of the pinned ring `shm` and rotated over the handle's lanes.
static void run_stream(const Args& a, int gpu, const std::string& shm, const std::string& suffix, Result* res)
{
    const int rows = a.rows, cols = a.cols, cp = a.cp, syms = a.syms, qam = a.qam, n_lanes = a.n_lanes, batch = a.batch, ring = a.ring,
              frames = a.frames;
    if (a.gpus > 1) pin_thread_near_gpu(gpu);
    gpuLS ls(rows, cols, cp, syms, qam, ring, shm, 0, gpu, n_lanes, batch);
    if (lsmrc_set_pilot_file(ls.handle, a.pilots.c_str()) < 0) {
        fprintf(stderr, "pilot: %s\n", lsmrc_last_error(ls.handle));
        res->rc = 1;
        return;
    }
    const int K = cols - 1;
    const size_t comb_bytes = (size_t)(syms - 1) * K * sizeof(cuFloatComplex);
    const size_t bits_bytes = (size_t)(syms - 1) * lsmrc_bits_row_bytes(cols, qam);
    std::ofstream out, outb, trace;
    if (a.write_out) {
        out.open(("Output_gpu" + suffix + ".dat").c_str(), std::ofstream::binary | std::ofstream::trunc);
        outb.open(("Bits_gpu" + suffix + ".dat").c_str(), std::ofstream::binary | std::ofstream::trunc);
    }
    if (!a.trace.empty()) {
        trace.open((a.trace + suffix).c_str(), std::ofstream::trunc);
        trace << "submission,lane,frames,t_enqueued_ms,t_h2d_done_ms,t_kernels_done_ms,t_results_on_host_ms\n";
    }
    ShMemBitsBuff* ret = a.bits_ring.empty() ? nullptr : new ShMemBitsBuff(a.bits_ring + suffix, 1, bits_bytes, a.bits_slots);
    std::vector<int> lane_n((size_t)n_lanes, 0);    // frames of the submission in flight on each lane
    std::vector<int> lane_sub((size_t)n_lanes, 0);  // ... and its running number
    auto collect = [&](int lane) {
        const void *comb = nullptr, *bits = nullptr;
        if (lsmrc_ring_wait(ls.handle, lane, &comb, &bits, nullptr) < 0) {
            fprintf(stderr, "ring_wait: %s\n", lsmrc_last_error(ls.handle));
            exit(1);
        }
        const int n = lane_n[(size_t)lane];
        if (trace.is_open()) {
            float t[4];
            if (lsmrc_ring_trace(ls.handle, lane, t) == LSMRC_OK)
                trace << lane_sub[(size_t)lane] << ',' << lane << ',' << n << ',' << t[0] << ',' << t[1] << ',' << t[2] << ',' << t[3] << '\n';
        }
        if (a.write_out) {
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
    if (lsmrc_ring_prepare(ls.handle) < 0) {  // lane buffers exist before the clock starts
        fprintf(stderr, "ring_prepare: %s\n", lsmrc_last_error(ls.handle));
        res->rc = 1;
        return;
    }
    const auto t0 = std::chrono::steady_clock::now();
    auto t_first = t0;
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
        if (sub == 0) t_first = std::chrono::steady_clock::now();
        if (lsmrc_ring_submit_frames(ls.handle, lane, first, n_first, second, nb) < 0) {
            fprintf(stderr, "ring_submit: %s\n", lsmrc_last_error(ls.handle));
            res->rc = 1;
            return;
        }
        lane_n[(size_t)lane] = nb;
        lane_sub[(size_t)lane] = sub;
        unreleased_frames += nb;
        ++unreleased_subs;
        ++busy;
        ++sub;
        f += nb;
    }
    while (unreleased_subs > 0) release_oldest(true);
    for (int i = sub - busy; i < sub; ++i) collect(i % n_lanes);
    res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res->steady_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_first).count();
    delete ret;  // unlinks the name; a reader that is still draining keeps its mapping
    char plan[256] = "";
    lsmrc_describe_plan(ls.handle, plan, sizeof plan);
    res->plan = plan;
}

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        auto val = [&](const char* name) -> const char* {
            if (std::strcmp(argv[i], name) == 0 && i + 1 < argc) return argv[++i];
            return nullptr;
        };
        const char* v;
        if ((v = val("--rows"))) a.rows = atoi(v);
        else if ((v = val("--cols"))) a.cols = atoi(v);
        else if ((v = val("--prefix"))) a.cp = atoi(v);
        else if ((v = val("--syms"))) a.syms = atoi(v);
        else if ((v = val("--qam"))) a.qam = atoi(v);
        else if ((v = val("--ring"))) a.ring = atoi(v);
        else if ((v = val("--frames"))) a.frames = atoi(v);
        else if ((v = val("--shm"))) a.shm = v;
        else if ((v = val("--pilots"))) a.pilots = v;
        else if ((v = val("--bits-ring"))) a.bits_ring = v;
        else if ((v = val("--bits-slots"))) a.bits_slots = atoi(v);
        else if ((v = val("--lanes"))) a.n_lanes = atoi(v);
        else if ((v = val("--batch"))) a.batch = atoi(v);
        else if ((v = val("--gpus"))) a.gpus = atoi(v);
        else if ((v = val("--trace"))) a.trace = v;
        else if (std::strcmp(argv[i], "--no-output") == 0) a.write_out = false;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    // default: 3 submissions in flight for large frames (copy-bound), 4 for frames below 1 MB (launch-latency bound)
    const bool small = (size_t)a.syms * a.rows * (a.cols + a.cp) * sizeof(complexF) < (1u << 20);
    if (a.n_lanes < 1 || a.n_lanes > 64) a.n_lanes = small ? 4 : 3;
    if (a.batch < 1 || a.batch > 64) a.batch = small ? 16 : 1;
    if (a.ring <= 0) a.ring = (a.n_lanes * a.batch + 1) * a.syms + 1;
    if (a.ring < a.syms + 1) {
        fprintf(stderr, "the ring (%d slots) must hold at least one frame of %d slots plus one\n", a.ring, a.syms);
        return 2;
    }
    if (a.gpus < 1 || a.gpus > 64) {
        fprintf(stderr, "--gpus must be in 1..64\n");
        return 2;
    }
    // --gpus G: G rings (<shm>_0 .. <shm>_{G-1}), one worker thread and one receiver handle per GPU, every worker on
    // the cores next to its GPU; `frames` frames per GPU.  Output files and the trace get the suffix _<g>.
    std::vector<Result> res((size_t)a.gpus);
    if (a.gpus == 1) {
        run_stream(a, 0, a.shm, "", &res[0]);
    } else {
        std::vector<std::thread> workers;
        for (int g = 0; g < a.gpus; ++g)
            workers.emplace_back([&a, &res, g] { run_stream(a, g, a.shm + "_" + std::to_string(g), "_" + std::to_string(g), &res[(size_t)g]); });
        for (auto& w : workers) w.join();
    }
    double dt = 0, dts = 0;
    for (const Result& r : res) {
        if (r.rc != 0) return r.rc;
        if (r.seconds > dt) dt = r.seconds;  // the slowest GPU
        if (r.steady_seconds > dts) dts = r.steady_seconds;
    }
    const double total_frames = (double)a.frames * a.gpus;
    const double samples = total_frames * a.syms * a.rows * (a.cols + a.cp);
    printf("{\"frames\": %.0f, \"gpus\": %d, \"seconds\": %.6f, \"frames_per_s\": %.2f, \"antenna_samples_per_s\": %.4e, \"h2d_gbs\": %.3f, "
           "\"seconds_from_first_submission\": %.6f, \"h2d_gbs_from_first_submission\": %.3f, \"plan\": \"%s\"}\n",
           total_frames, a.gpus, dt, total_frames / dt, samples / dt, samples * 8.0 / dt / 1e9, dts, samples * 8.0 / dts / 1e9, res[0].plan.c_str());
    return 0;
}
