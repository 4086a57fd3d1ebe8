// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), host code inside libd2p.so.
//
// TensorFlow's checkpoint files (tensor bundle, the format tf.train.Saver writes at reference
// trainer.py:114,182 and restores at trainer.py:145 / evaler.py:99) protect every SSTable
// block and every tensor with a masked CRC-32C; demo2program_b200/tf_checkpoint.py calls this
// for the 45 MB of parameters.  Slicing-by-8, tables built once.
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace {

struct Crc32cTables {
    uint32_t t[8][256];
    Crc32cTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xffu];
    }
};

const Crc32cTables& tables() {
    static const Crc32cTables tb;
    return tb;
}

}  // namespace

// crc = running value (0 to start); returns the CRC-32C of the bytes seen so far (unmasked).
extern "C" unsigned int d2p_crc32c(const void* data, size_t n, unsigned int crc) {
    const Crc32cTables& tb = tables();
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = ~crc;
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7u) != 0) {
        c = tb.t[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
        --n;
    }
    while (n >= 8) {
        uint64_t w;
        std::memcpy(&w, p, 8);
        w ^= c;   // little-endian host (x86-64 / aarch64 as configured for CUDA)
        c = tb.t[7][w & 0xff] ^ tb.t[6][(w >> 8) & 0xff] ^ tb.t[5][(w >> 16) & 0xff] ^ tb.t[4][(w >> 24) & 0xff] ^
            tb.t[3][(w >> 32) & 0xff] ^ tb.t[2][(w >> 40) & 0xff] ^ tb.t[1][(w >> 48) & 0xff] ^ tb.t[0][w >> 56];
        p += 8;
        n -= 8;
    }
    while (n > 0) {
        c = tb.t[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
        --n;
    }
    return ~c;
}
