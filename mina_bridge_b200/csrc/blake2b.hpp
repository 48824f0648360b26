// BLAKE2b-512 (RFC 7693), unkeyed.  Needed for the SRS hash-to-curve derivation
// (poly-commitment `SRS::create`, reached from AL/operator/mina/lib/src/lib.rs:34 and
// verifier_index.rs:204-208) and for the VRF tiebreak of the fork-choice rule
// (AL/operator/mina/lib/src/consensus_state.rs:140-146).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>

namespace pasta {
namespace host {

class Blake2b512 {
   public:
    Blake2b512() {
        for (int i = 0; i < 8; i++) h_[i] = iv(i);
        h_[0] ^= 0x01010000ull ^ 64ull;  // digest length 64, no key, fanout = depth = 1
    }
    void update(const uint8_t *data, size_t len) {
        while (len) {
            if (fill_ == 128) {  // only compress when more input follows (last block is special)
                counter_ += 128;
                compress(false);
                fill_ = 0;
            }
            size_t take = 128 - fill_ < len ? 128 - fill_ : len;
            std::memcpy(buf_ + fill_, data, take);
            fill_ += take;
            data += take;
            len -= take;
        }
    }
    std::array<uint8_t, 64> finish() {
        counter_ += fill_;
        std::memset(buf_ + fill_, 0, 128 - fill_);
        compress(true);
        std::array<uint8_t, 64> out;
        std::memcpy(out.data(), h_, 64);
        return out;
    }
    static std::array<uint8_t, 64> hash(const uint8_t *data, size_t len) {
        Blake2b512 b;
        b.update(data, len);
        return b.finish();
    }

   private:
    static uint64_t iv(int i) {
        static const uint64_t t[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                      0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                      0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return t[i];
    }
    static uint64_t ror(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static void mix(uint64_t *v, int a, int b, int c, int d, uint64_t x, uint64_t y) {
        v[a] += v[b] + x;
        v[d] = ror(v[d] ^ v[a], 32);
        v[c] += v[d];
        v[b] = ror(v[b] ^ v[c], 24);
        v[a] += v[b] + y;
        v[d] = ror(v[d] ^ v[a], 16);
        v[c] += v[d];
        v[b] = ror(v[b] ^ v[c], 63);
    }
    void compress(bool last) {
        static const uint8_t perm[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint64_t m[16], v[16];
        std::memcpy(m, buf_, 128);
        for (int i = 0; i < 8; i++) {
            v[i] = h_[i];
            v[8 + i] = iv(i);
        }
        v[12] ^= counter_;
        if (last) v[14] = ~v[14];
        for (int round = 0; round < 12; round++) {
            const uint8_t *s = perm[round % 10];
            mix(v, 0, 4, 8, 12, m[s[0]], m[s[1]]);
            mix(v, 1, 5, 9, 13, m[s[2]], m[s[3]]);
            mix(v, 2, 6, 10, 14, m[s[4]], m[s[5]]);
            mix(v, 3, 7, 11, 15, m[s[6]], m[s[7]]);
            mix(v, 0, 5, 10, 15, m[s[8]], m[s[9]]);
            mix(v, 1, 6, 11, 12, m[s[10]], m[s[11]]);
            mix(v, 2, 7, 8, 13, m[s[12]], m[s[13]]);
            mix(v, 3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[8 + i];
    }

    uint64_t h_[8];
    uint8_t buf_[128];
    size_t fill_ = 0;
    uint64_t counter_ = 0;
};

}  // namespace host
}  // namespace pasta
