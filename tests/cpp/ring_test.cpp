// tests/cpp/ring_test.cpp -- CPU-only unit test of host/ShMemSymBuff.hpp + CSharedMemSimple.hpp.
// modes:
//   selftest <shm>            producer thread + consumer thread over one segment: layout, order,
//                             wrap-around, flow control, whole-frame views, CP strip
//   dump <shm> A N C L count file   consumer that writes every slot to a file
//   read <shm> A N C L count  consumer for a foreign producer (the reference's own ring class,
//                             see oracle/ref_ring_writer.cpp): checks `count` patterned symbols
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <unistd.h>

#include "ShMemBitsBuff.hpp"
#include "ShMemSymBuff.hpp"

static complexF pattern(int sym, int a, int n) { return complexF{(float)(sym * 1000 + a), (float)n}; }

static int fail(const char* msg)
{
    fprintf(stderr, "ring_test FAILED: %s\n", msg);
    return 1;
}

static int selftest(const char* shm)
{
    const int A = 3, N = 8, C = 2, L = 7, S = 5, FRAMES = 9;
    shm_unlink(shm);
    ShMemSymBuff master(shm, 1, A, N, C, L);
    if (ShMemSymBuff::segmentBytes(A, N, C, L) != 12 + (size_t)L * A * (N + C) * 8) return fail("segment size");
    if (master.slotBytes() != (size_t)A * (N + C) * 8) return fail("slot size");

    std::thread producer([&] {
        std::vector<complexF> sym((size_t)A * (N + C));
        for (int s = 0; s < FRAMES * S; ++s) {
            for (int a = 0; a < A; ++a)
                for (int n = 0; n < N + C; ++n) sym[(size_t)a * (N + C) + n] = pattern(s, a, n);
            master.writeNextSymbolWithWait(sym.data());
        }
    });
    int rc = 0;
    {
        ShMemSymBuff slave(shm, 0, A, N, C, L);
        std::vector<complexF> Y((size_t)A * N);
        int s = 0;
        // first two frames symbol by symbol (CP stripped), the rest as whole-frame views
        for (; s < 2 * S && !rc; ++s) {
            if (s % S == S - 1) slave.readLastSymbol(Y.data());
            else slave.readNextSymbol(Y.data(), s % S);
            for (int a = 0; a < A && !rc; ++a)
                for (int n = 0; n < N; ++n) {
                    const complexF w = pattern(s, a, n + C);
                    if (Y[(size_t)a * N + n].real != w.real || Y[(size_t)a * N + n].imag != w.imag) rc = fail("symbol content / CP strip");
                }
        }
        int wrapped = 0;
        for (int f = 2; f < FRAMES && !rc; ++f) {
            const complexF *first = nullptr, *second = nullptr;
            int n_first = 0;
            slave.waitFrame(S, &first, &n_first, &second);
            if (n_first < S) {
                wrapped++;
                if (!second) rc = fail("split frame without second piece");
            }
            for (int i = 0; i < S && !rc; ++i) {
                const complexF* p = (i < n_first) ? first + (size_t)i * slave.slotElems() : second + (size_t)(i - n_first) * slave.slotElems();
                const complexF w = pattern(f * S + i, A - 1, N + C - 1);
                const complexF g = p[(size_t)(A - 1) * (N + C) + N + C - 1];
                if (g.real != w.real || g.imag != w.imag) rc = fail("frame view content");
            }
            slave.releaseSlots(S);
        }
        if (!rc && wrapped == 0) rc = fail("test never exercised a wrapped frame");
        if (!rc && slave.available() != 0) rc = fail("ring not empty at the end");
    }
    producer.join();
    if (!rc && !master.readerGone()) rc = fail("slave destructor did not flag size=-1");
    if (!rc) printf("ring selftest ok\n");
    return rc;
}

static int read_foreign(const char* shm, int A, int N, int C, int L, int count)
{
    ShMemSymBuff slave(shm, 0, A, N, C, L);
    for (int s = 0; s < count; ++s) {
        const complexF* p = slave.peekSlot();
        for (int a = 0; a < A; ++a)
            for (int n = 0; n < N + C; n += (N + C - 1 > 0 ? N + C - 1 : 1)) {
                const complexF w = pattern(s, a, n), g = p[(size_t)a * (N + C) + n];
                if (g.real != w.real || g.imag != w.imag) { fprintf(stderr, "s=%d a=%d n=%d got (%g,%g) want (%g,%g)\n", s,a,n,g.real,g.imag,w.real,w.imag); return fail("foreign producer: content mismatch"); }
            }
        slave.releaseSlots(1);
    }
    printf("ring read ok (%d symbols)\n", count);
    return 0;
}

// consumer that appends every slot it reads to a file (checks multi-threaded producers: order and content)
static int dump_foreign(const char* shm, int A, int N, int C, int L, int count, const char* out_path)
{
    ShMemSymBuff slave(shm, 0, A, N, C, L);
    FILE* out = fopen(out_path, "wb");
    if (!out) return fail("cannot open the dump file");
    for (int s = 0; s < count; ++s) {
        const complexF* p = slave.peekSlot();
        fwrite(p, sizeof(complexF), slave.slotElems(), out);
        if (s % 7 == 3) usleep(200);  // a reader that stalls now and then: the producers must wait for free slots
        slave.releaseSlots(1);
    }
    fclose(out);
    printf("ring dump ok (%d symbols)\n", count);
    return 0;
}

// writer end of the return ring (ShMemBitsBuff, master): frame i carries bytes (i*131 + j*7) & 255
static int bits_write(const char* name, long frame_bytes, int slots, int frames)
{
    ShMemBitsBuff ring(name, 1, (size_t)frame_bytes, slots);
    std::vector<uint8_t> buf((size_t)frame_bytes);
    for (int i = 0; i < frames; ++i) {
        for (long j = 0; j < frame_bytes; ++j) buf[(size_t)j] = (uint8_t)((i * 131 + j * 7) & 255);
        if (!ring.writeFrame(buf.data())) {
            fprintf(stderr, "reader went away at frame %d\n", i);
            return 1;
        }
    }
    // give the reader time to drain before the segment name disappears with this process
    for (int spin = 0; spin < 2000 && ring.frameReady(); ++spin) usleep(1000);
    printf("bits ring wrote %d frames\n", frames);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 6 && std::strcmp(argv[1], "bitswrite") == 0) return bits_write(argv[2], atol(argv[3]), atoi(argv[4]), atoi(argv[5]));
    if (argc >= 3 && std::strcmp(argv[1], "selftest") == 0) return selftest(argv[2]);
    if (argc >= 8 && std::strcmp(argv[1], "read") == 0)
        return read_foreign(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
    if (argc >= 9 && std::strcmp(argv[1], "dump") == 0)
        return dump_foreign(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), argv[8]);
    fprintf(stderr, "usage: ring_test selftest <shm> | read <shm> A N C L count\n");
    return 2;
}
