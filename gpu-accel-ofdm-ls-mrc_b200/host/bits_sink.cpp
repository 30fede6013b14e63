// bits_sink.cpp -- the process at the far end of the return ring: attaches to the segment the
// receiver created (ShMemBitsBuff, slave), reads F frames of packed bits and writes them to a
// file.  Stands in for a channel decoder in tests.  No GPU, no CUDA.
//   bits_sink --shm /name --frame-bytes B --slots L --frames F --out file
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ShMemBitsBuff.hpp"

int main(int argc, char** argv)
{
    std::string shm = "/lsmrc_bits", out = "Bits_ring.dat";
    long frame_bytes = 0;
    int slots = 8, frames = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        if (a == "--shm") shm = argv[i + 1];
        else if (a == "--frame-bytes") frame_bytes = atol(argv[i + 1]);
        else if (a == "--slots") slots = atoi(argv[i + 1]);
        else if (a == "--frames") frames = atoi(argv[i + 1]);
        else if (a == "--out") out = argv[i + 1];
        else {
            fprintf(stderr, "usage: %s --shm /name --frame-bytes B --slots L --frames F --out file\n", argv[0]);
            return 2;
        }
    }
    if (frame_bytes <= 0 || slots < 2 || frames < 1) {
        fprintf(stderr, "bits_sink: need --frame-bytes > 0, --slots >= 2, --frames >= 1\n");
        return 2;
    }
    ShMemBitsBuff ring(shm, 0, (size_t)frame_bytes, slots);  // waits for the writer to initialise the segment
    std::ofstream f(out.c_str(), std::ofstream::binary | std::ofstream::trunc);
    std::vector<uint8_t> buf((size_t)frame_bytes);
    for (int i = 0; i < frames; ++i) {
        ring.readFrame(buf.data());
        f.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)buf.size());
    }
    f.flush();
    printf("{\"frames\": %d, \"bytes\": %ld}\n", frames, (long)frames * frame_bytes);
    return 0;
}
