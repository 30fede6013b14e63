// ring_feeder.cpp -- synthetic producer for the symbol ring, standing in for the radio
// front end: it reproduces the write pattern of rx_and_corr.cpp:64-87 (one slot per OFDM
// symbol, antenna-major rows, a frame = S consecutive slots starting at the pilot) but leaves
// the cyclic prefix IN the slot (prefix > 0 layout), so the CP strip happens on the GPU.
// Creates the segment (master), waits for a consumer, streams the frames of a file.
//
//   ring_feeder --file rx.bin --rows A --cols N --prefix C --syms S --ring L [--frames F]
//               [--shm /blah] [--repeat R] [--nowait]
// rx.bin holds [F][S][A][N+C] complex64.  --nowait uses writeNextSymbolNoWait like the
// reference producer (can overrun a slow reader); the default blocks on a full ring.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ShMemSymBuff.hpp"

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, ring = 0, frames = -1, repeat = 1;
    bool nowait = false;
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
        else if ((v = val("--shm"))) shm = v;
        else if ((v = val("--file"))) file = v;
        else if (std::strcmp(argv[i], "--nowait") == 0) nowait = true;
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
    for (int r = 0; r < repeat; ++r)
        for (int f = 0; f < frames; ++f)
            for (int s = 0; s < syms; ++s) {
                complexF* sym = buf.data() + ((size_t)f * syms + (size_t)s) * slot;
                if (nowait) ringbuf.writeNextSymbolNoWait(sym);
                else ringbuf.writeNextSymbolWithWait(sym);
            }
    // keep the segment alive until the reader has drained it and gone away
    while (ringbuf.available() > 0 && !ringbuf.readerGone()) sched_yield();
    return 0;
}
