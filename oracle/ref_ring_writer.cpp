/*
 * oracle/ref_ring_writer.cpp -- TEST INFRASTRUCTURE.  A producer built from the reference's OWN
 * ring classes (ShMemSymBuff.hpp + CSharedMemSimple.hpp, #included from /root/reference): it
 * creates the segment as master and writes `count` patterned symbols with
 * writeNextSymbolNoWait, exactly as rx_and_corr.cpp:83 does.  tests/test_ring_host.py lets the
 * new host/ShMemSymBuff.hpp attach as slave and read them, proving the segment layout and the
 * index protocol are wire-compatible with the reference producer.
 * usage: ring_writer_<case> count   (dims are the -D macros; ring name is shmemID "/blah")
 */
#include "CSharedMemSimple.hpp"
#include "ShMemSymBuff.hpp"

#include <unistd.h>
#include <vector>

int main(int argc, char** argv)
{
    const int count = argc > 1 ? atoi(argv[1]) : 1;
    const int A = numOfRows, W = dimension + prefix;
    /* never destroyed, like cpuLS_main.cpp:96 ("//delete buffPtr"): the reference master's
     * destructor loops on `while(size == -1) delete` once a reader has left (ShMemSymBuff.hpp:221-226) */
    ShMemSymBuff& ring = *new ShMemSymBuff(shmemID, 1);
    usleep(100000); /* a NoWait producer overruns late readers: give the reader time to attach */
    std::vector<std::complex<float> > sym((size_t)A * W);
    for (int s = 0; s < count; s++) {
        for (int a = 0; a < A; a++)
            for (int n = 0; n < W; n++) sym[(size_t)a * W + n] = std::complex<float>((float)(s * 1000 + a), (float)n);
        ring.writeNextSymbolNoWait(sym.data());
        usleep(2000); /* NoWait never blocks: pace the writes so the reader keeps up */
    }
    usleep(300000); /* keep the segment mapped while the reader drains it */
    shm_unlink(shmemID);
    return 0;
}
