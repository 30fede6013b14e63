// ShMemBitsBuff.hpp -- return ring: the receiver hands the decoded bits of whole frames back to
// another process over shared memory, symmetrical to the input ring (the
// reference only appends to Output_*.dat, cpuLS.hpp:374-380).
//
// The segment is an ordinary ShMemSymBuff segment -- 12-byte header {size, readPtr, writePtr}, then
// `len` slots, same single-producer/single-consumer protocol -- whose slot carries the packed bits of
// ONE frame ((S-1) rows of lsmrc_bits_row_bytes() bytes, padded to a multiple of 8).  Roles are those
// of the input ring: the writer (here the receiver) is the master and creates the segment, the
// reader (the decoder) attaches as slave and announces its departure through size = -1.
#ifndef LSMRC_HOST_SHMEMBITSBUFF_HPP_
#define LSMRC_HOST_SHMEMBITSBUFF_HPP_

#include <cstdint>
#include <cstring>
#include <string>

#include "ShMemSymBuff.hpp"

class ShMemBitsBuff {
   public:
    ShMemBitsBuff(const std::string& shm_uid, int isMaster, size_t frame_bytes, int len)
        : bytes_(frame_bytes), ring_(shm_uid, isMaster, 1, (int)((frame_bytes + sizeof(complexF) - 1) / sizeof(complexF)), 0, len)
    {
    }
    size_t frameBytes() const { return bytes_; }
    int slots() const { return ring_.slots(); }

    // writer: blocks while the ring is full; false when the reader has gone away
    bool writeFrame(const uint8_t* bits)
    {
        complexF* s = ring_.acquireWriteSlot();
        if (!s) return false;
        std::memcpy(s, bits, bytes_);
        ring_.commitWriteSlot();
        return true;
    }
    // reader: blocks until a frame is there
    void readFrame(uint8_t* bits)
    {
        const complexF* s = ring_.peekSlot();
        std::memcpy(bits, s, bytes_);
        ring_.releaseSlots(1);
    }
    bool frameReady() { return ring_.available() >= 1; }
    bool readerGone() { return ring_.readerGone(); }

   private:
    size_t bytes_;
    ShMemSymBuff ring_;
};

#endif
