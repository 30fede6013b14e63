// gpuLS_main.cpp -- ring consumer, the B200 counterpart of gpuLS_main.cu:66-141: attach to
// the symbol ring as slave, then per frame firstVector (pilot symbol) and demodOneSymbol for
// every data symbol, appending the combined symbols to Output_gpu.dat (gpuLS_main.cu:114-126)
// and the demapped bits to Bits_gpu.dat.  Written from scratch against host/gpuLS.hpp.
//
//   gpuLS_main [--rows A] [--cols N] [--prefix C] [--syms S] [--qam b] [--ring L]
//              [--frames F] [--shm /blah] [--pilots Pilots.dat] [--frame-mode]
// Defaults are the compile-time macros (ShMemSymBuff.hpp).  --frame-mode consumes whole
// frames with demodOneFrame (one overlapped H2D + two kernels per frame) instead of symbols.
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "gpuLS.hpp"

static volatile std::sig_atomic_t stop_signal_called = 0;
static void sig_int_handler(int) { stop_signal_called = 1; }

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, qam = LSMRC_QAM_BITS, ring = 0, frames = 1;
    std::string shm = shmemID, pilots = fileNameForX;
    bool frame_mode = false;
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
        else if (std::strcmp(argv[i], "--frame-mode") == 0) frame_mode = true;
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    if (ring <= 0) ring = frame_mode ? 3 * syms + 1 : syms + 1;
    std::signal(SIGINT, &sig_int_handler);

    gpuLS ls(rows, cols, cp, syms, qam, ring, shm, /*is_master=*/0, /*device=*/0);
    if (lsmrc_set_pilot_file(ls.handle, pilots.c_str()) < 0) {
        fprintf(stderr, "pilot: %s\n", lsmrc_last_error(ls.handle));
        return 1;
    }
    const int K = cols - 1;
    std::vector<cuFloatComplex> dY((size_t)(syms - 1) * K);
    std::ofstream out("Output_gpu.dat", std::ofstream::binary | std::ofstream::trunc);
    std::ofstream outb("Bits_gpu.dat", std::ofstream::binary | std::ofstream::trunc);
    const size_t row_bytes = lsmrc_bits_row_bytes(cols, qam);

    for (int f = 0; f < frames && !stop_signal_called; ++f) {
        if (frame_mode) {
            ls.demodOneFrame(dY.data(), nullptr, nullptr, nullptr, nullptr, rows, cols);
            out.write(reinterpret_cast<const char*>(dY.data()), (std::streamsize)(dY.size() * sizeof(cuFloatComplex)));
            outb.write(reinterpret_cast<const char*>(ls.lastBits()), (std::streamsize)(row_bytes * (size_t)(syms - 1)));
        } else {
            ls.firstVector(nullptr, nullptr, nullptr, nullptr, nullptr, rows, cols, 0);
            for (int i = 1; i < syms; ++i) {
                ls.demodOneSymbol(dY.data(), nullptr, nullptr, nullptr, rows, cols, i);
                out.write(reinterpret_cast<const char*>(dY.data()), (std::streamsize)((size_t)K * sizeof(cuFloatComplex)));
                outb.write(reinterpret_cast<const char*>(ls.lastBits()), (std::streamsize)row_bytes);
            }
        }
    }
    if (timerEn && ls.buffPtr) {
        ls.buffPtr->printTimes(false);
        ls.buffPtr->storeTimes(false);
    }
    return 0;
}
