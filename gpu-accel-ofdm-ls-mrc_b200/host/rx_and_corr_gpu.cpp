// rx_and_corr_gpu.cpp -- the receive front end of rx_and_corr.cpp (PN frame sync :332-360, frame stitching
// :372-393, symbol gather :64-87) moved onto the GPU and chained straight into the fused receiver: a capture
// goes host -> device once and comes back as combined symbols + bits; the CPU correlator, the shared-memory ring
// and the per-symbol copies are bypassed.  File-driven stand-in for the USRP receive loop (no UHD here).
//
//   rx_and_corr_gpu --buf1 b1.bin --buf2 b2.bin --pn pn.bin --samps M --rows A --cols N --prefix C --syms S
//                   [--qam b] [--thres 0.5] [--pilots Pilots.dat]
// b1/b2: [A][M] complex64 consecutive capture buffers per channel; pn: L complex64.
// Writes Output_gpu.dat / Bits_gpu.dat and prints the detected offset.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "gpuLS.hpp"

static std::vector<char> slurp(const std::string& path)
{
    std::ifstream f(path.c_str(), std::ifstream::binary);
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path.c_str());
        exit(2);
    }
    f.seekg(0, f.end);
    std::vector<char> v((size_t)f.tellg());
    f.seekg(0, f.beg);
    f.read(v.data(), (std::streamsize)v.size());
    return v;
}

int main(int argc, char** argv)
{
    int rows = numOfRows, cols = dimension, cp = prefix, syms = lenOfBuffer, qam = LSMRC_QAM_BITS, samps = 0;
    float thres = 0.5f;
    std::string b1, b2, pnf, pilots = fileNameForX;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const char* v = argv[i + 1];
        if (k == "--rows") rows = atoi(v);
        else if (k == "--cols") cols = atoi(v);
        else if (k == "--prefix") cp = atoi(v);
        else if (k == "--syms") syms = atoi(v);
        else if (k == "--qam") qam = atoi(v);
        else if (k == "--samps") samps = atoi(v);
        else if (k == "--thres") thres = (float)atof(v);
        else if (k == "--buf1") b1 = v;
        else if (k == "--buf2") b2 = v;
        else if (k == "--pn") pnf = v;
        else if (k == "--pilots") pilots = v;
        else {
            fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 2;
        }
    }
    std::vector<char> h1 = slurp(b1), h2 = slurp(b2), hp = slurp(pnf);
    const int L = (int)(hp.size() / sizeof(cuFloatComplex));
    if (samps <= 0) samps = (int)(h1.size() / sizeof(cuFloatComplex) / (size_t)rows);
    gpuLS ls(rows, cols, cp, syms, qam, /*ring_slots=*/0, "", 0, 0);
    lsmrc_handle h = ls.handle;
    if (lsmrc_set_pilot_file(h, pilots.c_str()) < 0) return 1;
    void *d1 = nullptr, *d2 = nullptr, *dp = nullptr, *drx = nullptr, *dcomb = nullptr, *dbits = nullptr;
    const size_t K = (size_t)cols - 1, nd = (size_t)syms - 1, row_bytes = lsmrc_bits_row_bytes(cols, qam);
    if (lsmrc_dev_alloc(h, h1.size(), &d1) || lsmrc_dev_alloc(h, h2.size(), &d2) || lsmrc_dev_alloc(h, hp.size(), &dp) ||
        lsmrc_dev_alloc(h, (size_t)syms * rows * (cols + cp) * sizeof(cuFloatComplex), &drx) ||
        lsmrc_dev_alloc(h, nd * K * sizeof(cuFloatComplex), &dcomb) || lsmrc_dev_alloc(h, nd * row_bytes, &dbits)) {
        fprintf(stderr, "alloc: %s\n", lsmrc_last_error(h));
        return 1;
    }
    lsmrc_copy_to_device(h, d1, h1.data(), h1.size());
    lsmrc_copy_to_device(h, d2, h2.data(), h2.size());
    lsmrc_copy_to_device(h, dp, hp.data(), hp.size());
    int off = -1, ch = -1;
    float metric = 0.f;
    if (lsmrc_sync_correlate(h, d1, rows, samps, dp, L, thres, &off, &ch, &metric, nullptr) < 0) {
        fprintf(stderr, "correlate: %s\n", lsmrc_last_error(h));
        return 1;
    }
    printf("{\"offset\": %d, \"channel\": %d, \"metric\": %.6f}\n", off, ch, metric);
    if (off < 0) return 3;  // no frame in this capture (rx_and_corr.cpp:362-364 `continue`)
    if (lsmrc_sync_assemble(h, d1, d2, samps, off, L, drx) < 0 ||
        lsmrc_demod_frames_device(h, drx, 1, nullptr, nullptr, dcomb, dbits) < 0) {
        fprintf(stderr, "demod: %s\n", lsmrc_last_error(h));
        return 1;
    }
    std::vector<char> comb(nd * K * sizeof(cuFloatComplex)), bits(nd * row_bytes);
    lsmrc_copy_to_host(h, comb.data(), dcomb, comb.size());
    lsmrc_copy_to_host(h, bits.data(), dbits, bits.size());
    std::ofstream("Output_gpu.dat", std::ofstream::binary).write(comb.data(), (std::streamsize)comb.size());
    std::ofstream("Bits_gpu.dat", std::ofstream::binary).write(bits.data(), (std::streamsize)bits.size());
    return 0;
}
