// latency_main.cpp -- per-frame latency of the C ABI for one frame per call (BASELINE config c5:
// 64-point FFT, 16 antennas, 16 symbols, QPSK): steady_clock around
//   lsmrc_demod_frames_device(one device-resident frame) + lsmrc_sync          ("device")
//   lsmrc_demod_frames_host(one frame in pinned host memory, results to host)   ("host")
// under the three launch policies of lsmrc_set_oneshot (2 = default, 1 = fused kernel with staged copies,
// 0 = pilot + data kernel pair).  Prints one JSON line.
// Plain g++; links the C ABI only.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ofdm_lsmrc.h"

static void die(lsmrc_handle h, const char* what, int rc)
{
    fprintf(stderr, "latency_main: %s failed: %s (%s)\n", what, lsmrc_error_name(rc), lsmrc_last_error(h));
    exit(1);
}
#define OK(h, call)                         \
    do {                                    \
        const int rc_ = (call);             \
        if (rc_ < 0) die((h), #call, rc_);  \
    } while (0)

struct Stats {
    double p50, p99, mean;
};
static Stats stats(std::vector<double>& v)
{
    std::sort(v.begin(), v.end());
    double s = 0;
    for (double x : v) s += x;
    return {v[v.size() / 2], v[(size_t)(0.99 * (double)(v.size() - 1))], s / (double)v.size()};
}

int main(int argc, char** argv)
{
    int rows = 16, cols = 64, cp = 16, syms = 16, qam = 2, launches = 10000, warm = 500;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        const int v = atoi(argv[i + 1]);
        if (a == "--rows") rows = v;
        else if (a == "--cols") cols = v;
        else if (a == "--prefix") cp = v;
        else if (a == "--syms") syms = v;
        else if (a == "--qam") qam = v;
        else if (a == "--launches") launches = v;
        else if (a == "--warmup") warm = v;
        else {
            fprintf(stderr, "usage: %s [--rows A --cols N --prefix C --syms S --qam b --launches n --warmup w]\n", argv[0]);
            return 2;
        }
    }
    const int K = cols - 1;
    lsmrc_config c;
    c.n_ant = rows;
    c.fft_size = cols;
    c.cp_len = cp;
    c.n_sym = syms;
    c.qam_bits = qam;
    c.max_frames = 1;
    c.device = 0;
    c.n_lanes = 1;
    lsmrc_handle h = nullptr;
    OK(nullptr, lsmrc_create(&c, &h));

    const size_t n_rx = lsmrc_rx_frame_elems(&c);
    const size_t n_comb = (size_t)(syms - 1) * K;
    const size_t n_bits = (size_t)(syms - 1) * lsmrc_bits_row_bytes(cols, qam);
    void *h_rx = nullptr, *h_comb = nullptr, *h_bits = nullptr, *d_rx = nullptr, *d_comb = nullptr, *d_bits = nullptr;
    OK(h, lsmrc_host_alloc(h, n_rx * 8, &h_rx));
    OK(h, lsmrc_host_alloc(h, n_comb * 8, &h_comb));
    OK(h, lsmrc_host_alloc(h, n_bits, &h_bits));
    OK(h, lsmrc_dev_alloc(h, n_rx * 8, &d_rx));
    OK(h, lsmrc_dev_alloc(h, n_comb * 8, &d_comb));
    OK(h, lsmrc_dev_alloc(h, n_bits, &d_bits));
    unsigned long long s = 88172645463325252ULL;
    float* f = static_cast<float*>(h_rx);
    for (size_t i = 0; i < 2 * n_rx; ++i) {  // xorshift noise: any input exercises the same code path
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        f[i] = (float)((double)(s >> 11) / 9007199254740992.0 - 0.5);
    }
    std::vector<float> pilot((size_t)2 * K);
    for (int k = 0; k < K; ++k) {
        pilot[2 * k] = (k & 1) ? 0.70710678f : -0.70710678f;
        pilot[2 * k + 1] = (k & 2) ? 0.70710678f : -0.70710678f;
    }
    OK(h, lsmrc_set_pilot(h, pilot.data(), K));
    OK(h, lsmrc_copy_to_device(h, d_rx, h_rx, n_rx * 8));

    using clk = std::chrono::steady_clock;
    std::string out = "{";
    for (int oneshot = 2; oneshot >= 0; --oneshot) {
        OK(h, lsmrc_set_oneshot(h, oneshot));
        for (int path = 0; path < 2; ++path) {
            auto once = [&]() {
                if (path == 0) {
                    OK(h, lsmrc_demod_frames_device(h, d_rx, 1, nullptr, nullptr, d_comb, d_bits));
                    OK(h, lsmrc_sync(h));
                } else {
                    OK(h, lsmrc_demod_frames_host(h, h_rx, 1, nullptr, nullptr, h_comb, h_bits));
                }
            };
            if (oneshot == 1 && path == 0) continue;  // modes 1 and 2 differ only for host buffers
            for (int i = 0; i < warm; ++i) once();
            const long long l0 = lsmrc_launch_count(h);
            once();
            const long long per_frame = lsmrc_launch_count(h) - l0;
            std::vector<double> us((size_t)launches);
            for (int i = 0; i < launches; ++i) {
                const clk::time_point t0 = clk::now();
                once();
                us[(size_t)i] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
            }
            const Stats st = stats(us);
            char buf[256];
            snprintf(buf, sizeof buf, "%s\"%s_%s\": {\"p50_us\": %.3f, \"p99_us\": %.3f, \"mean_us\": %.3f, \"kernels_per_frame\": %lld}",
                     out.size() > 1 ? ", " : "", path == 0 ? "device" : "host",
                     oneshot == 2 ? (path == 0 ? "one_launch" : "one_launch_in_place") : oneshot == 1 ? "one_launch_staged" : "two_kernels", st.p50, st.p99,
                     st.mean, per_frame);
            out += buf;
        }
    }
    {
        // floor of the same call pattern: one trivial kernel (a 1-row roll) + sync through the same ABI
        void* d_tmp = nullptr;
        OK(h, lsmrc_dev_alloc(h, (size_t)K * 8, &d_tmp));
        std::vector<double> us((size_t)launches);
        for (int i = 0; i < warm + launches; ++i) {
            const clk::time_point t0 = clk::now();
            OK(h, lsmrc_stage_shift_rows(h, d_comb, d_tmp, 1));
            OK(h, lsmrc_sync(h));
            if (i >= warm) us[(size_t)(i - warm)] = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
        }
        const Stats st = stats(us);
        char buf[160];
        snprintf(buf, sizeof buf, ", \"floor_trivial_kernel\": {\"p50_us\": %.3f, \"p99_us\": %.3f, \"mean_us\": %.3f}", st.p50, st.p99, st.mean);
        out += buf;
        lsmrc_dev_free(h, d_tmp);
    }
    char tail[256];
    snprintf(tail, sizeof tail, ", \"launches\": %d, \"rows\": %d, \"cols\": %d, \"prefix\": %d, \"syms\": %d, \"qam\": %d}", launches, rows, cols,
             cp, syms, qam);
    out += tail;
    puts(out.c_str());
    lsmrc_dev_free(h, d_rx);
    lsmrc_dev_free(h, d_comb);
    lsmrc_dev_free(h, d_bits);
    lsmrc_host_free(h, h_rx);
    lsmrc_host_free(h, h_comb);
    lsmrc_host_free(h, h_bits);
    lsmrc_destroy(h);
    return 0;
}
