// crc32c.cpp - CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), the checksum of the TF-V2 checkpoint ("tensor
// bundle") format: every table block of <prefix>.index and every tensor of <prefix>.data-* carries one.  Host-only
// helper behind wun_crc32c (include/wun.h) for TFCheckpoint.py; slicing-by-8, no ISA extensions needed.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace {

struct Tables {
    uint32_t t[8][256];
    Tables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xffu];
    }
};

const Tables& tables() {
    static const Tables T;
    return T;
}

}  // namespace

extern "C" uint32_t wun_crc32c(uint32_t crc, const void* data, uint64_t n) {
    const Tables& T = tables();
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t c = ~crc;
    while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu]; --n; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);                     // little-endian host (x86-64 / aarch64)
        const uint32_t lo = (uint32_t)w ^ c, hi = (uint32_t)(w >> 32);
        c = T.t[7][lo & 0xffu] ^ T.t[6][(lo >> 8) & 0xffu] ^ T.t[5][(lo >> 16) & 0xffu] ^ T.t[4][lo >> 24] ^
            T.t[3][hi & 0xffu] ^ T.t[2][(hi >> 8) & 0xffu] ^ T.t[1][(hi >> 16) & 0xffu] ^ T.t[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n) { c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu]; --n; }
    return ~c;
}
