/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the reference's OWN CPU implementation, compiled from the sources
 * where they lie under /root/reference (nothing is copied into this repo):
 * ShMemSymBuff.hpp, CSharedMemSimple.hpp and cpuLS.hpp are #included exactly as
 * cpuLS_main.cpp:27-35 includes them.  <fftw3.h>/<cblas.h> resolve to the shims
 * in oracle/ref_shim/ because FFTW3/CBLAS are not installed here.
 *
 * What runs verbatim from the reference: the ring (ShMemSymBuff /
 * CSharedMemSimple, master mode), matrix_readX, fftOneRow, divideOneRow,
 * findDistSqrd and the whole of doOneSymbol (ring read, CP strip, FFT, drop
 * DC, matrixMultThenSum, normalise, shiftOneRow, Output_cpu.dat append).
 * What this driver restores: firstVector() as shipped never reads the pilot
 * symbol (cpuLS.hpp:266-272 is commented out) so it yields H == 0 / NaN; here
 * the pilot is read with readNextSymbol and the body of firstVector
 * (cpuLS.hpp:278-311) is replayed with the reference's own sub-functions.
 *
 * Dimensions are compile-time in the reference (-DnumOfRows -Ddimension
 * -Dprefix -DlenOfBuffer), so oracle/Makefile builds one binary per test case.
 *
 * usage: cpuls_ref_<case> <workdir> <rx.bin> <n_frames> <out_prefix> [pilots.dat]
 *   rx.bin      [F][S][A][N+C] complex64
 *   pilots.dat  K complex64 in ascending-frequency order (cpuLS.hpp:93); when
 *               absent the reference's own fallback 0.707+0.707i is exercised
 *   outputs     <out_prefix>.hconj [F][A][K] c64, .hsqrd [F][K] f32,
 *               .comb [F][S-1][K] c64
 */
#include <fftw3.h>
#include "CSharedMemSimple.hpp"
#include "ShMemSymBuff.hpp"
#include "cpuLS.hpp"

#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <vector>

static std::vector<char> slurp(const std::string &path)
{
    std::ifstream f(path.c_str(), std::ifstream::binary);
    if (!f) {
        fprintf(stderr, "ref_driver: cannot open %s\n", path.c_str());
        exit(2);
    }
    f.seekg(0, f.end);
    size_t n = (size_t)f.tellg();
    f.seekg(0, f.beg);
    std::vector<char> buf(n);
    f.read(buf.data(), (std::streamsize)n);
    return buf;
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s workdir rx.bin n_frames out_prefix [pilots.dat]\n", argv[0]);
        return 2;
    }
    const std::string workdir = argv[1], rx_path = argv[2], out_prefix = argv[4];
    const int F = atoi(argv[3]);
    const int rows = numOfRows, cols = dimension, S = lenOfBuffer, K = cols - 1;
    const size_t slot = (size_t)rows * (cols + prefix);

    std::vector<char> rx = slurp(rx_path);
    if (rx.size() != (size_t)F * S * slot * sizeof(complexF)) {
        fprintf(stderr, "ref_driver: %s has %zu bytes, expected %zu\n", rx_path.c_str(), rx.size(),
                (size_t)F * S * slot * sizeof(complexF));
        return 2;
    }
    std::vector<char> pil;
    if (argc >= 6) pil = slurp(argv[5]);
    if (chdir(workdir.c_str()) != 0) {
        perror("chdir");
        return 2;
    }
    remove(fileNameForX);
    if (!pil.empty()) {
        std::ofstream pf(fileNameForX, std::ofstream::binary);
        pf.write(pil.data(), (std::streamsize)pil.size());
    }

    std::vector<complexF> Y((size_t)rows * cols), Hconj((size_t)rows * K), X(K);
    std::ofstream o_h((out_prefix + ".hconj").c_str(), std::ofstream::binary);
    std::ofstream o_e((out_prefix + ".hsqrd").c_str(), std::ofstream::binary);
    std::ofstream o_c((out_prefix + ".comb").c_str(), std::ofstream::binary);

    shm_unlink(shmemID);
    buffPtr = new ShMemSymBuff(shmemID, 1); /* master: creates and initialises the ring */
    numTimes = 1;

    const std::chrono::steady_clock::time_point t_begin = std::chrono::steady_clock::now();
    for (int f = 0; f < F; f++) {
        std::complex<float> *frame = (std::complex<float> *)rx.data() + (size_t)f * S * slot;
        /* producer contract rx_and_corr.cpp:64-87: one slot per symbol, NoWait.
         * The ring cannot tell full from empty (writePtr==readPtr), so a single
         * thread must keep the writer one slot short until the reader caught up. */
        for (int s = 0; s < S - 1; s++) buffPtr->writeNextSymbolNoWait(frame + (size_t)s * slot);

        /* ---- firstVector with the pilot read restored ---- */
        matrix_readX(X.data(), K);                   /* cpuLS.hpp:249 */
        buffPtr->readNextSymbol(Y.data(), 0);        /* cpuLS.hpp:266-272 (restored) */
        for (int row = 0; row < rows; row++) fftOneRow(Y.data(), cols, row); /* :278-281 */
        for (int row = 0; row < rows; row++) {       /* :290-299 */
            memcpy(&Hconj[(size_t)row * K], &Y[(size_t)row * cols + 1], K * sizeof(complexF));
            divideOneRow(Hconj.data(), X.data(), K, row);
        }
        for (int i = 0; i < rows; i++)               /* :303-307 */
            for (int j = 0; j < K; j++) Hconj[(size_t)i * K + j].imag = -1 * Hconj[(size_t)i * K + j].imag;
        findDistSqrd(Hconj.data(), X.data(), rows, K); /* :311 -- X now holds sum|H|^2 in .real */

        o_h.write((const char *)Hconj.data(), (std::streamsize)(Hconj.size() * sizeof(complexF)));
        for (int j = 0; j < K; j++) o_e.write((const char *)&X[j].real, sizeof(float));

        /* ---- data symbols: the reference's doOneSymbol, verbatim ---- */
        for (int i = 1; i < S; i++) {                /* cpuLS_main.cpp:83-92 */
            if (i == S - 2) buffPtr->writeNextSymbolNoWait(frame + (size_t)(S - 1) * slot);
            doOneSymbol(Y.data(), Hconj.data(), X.data(), rows, cols, i);
        }
        std::vector<char> outc = slurp(file);        /* Output_cpu.dat, cpuLS.hpp:374-380 */
        if (outc.size() != (size_t)(S - 1) * K * sizeof(complexF)) {
            fprintf(stderr, "ref_driver: %s has %zu bytes\n", file.c_str(), outc.size());
            return 3;
        }
        o_c.write(outc.data(), (std::streamsize)outc.size());
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    /* wall time of the frame loop: ring writes/reads, CP strip, FFTs, LS, MRC, Output_cpu.dat I/O -- the
     * reference's whole per-frame CPU path (cpuLS_main.cpp:80-93), input file read excluded */
    printf("{\"frames\": %d, \"seconds\": %.6f}\n", F, secs);
    shm_unlink(shmemID);
    remove(file.c_str());
    remove(fileNameForX);
    return 0;
}
