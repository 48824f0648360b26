// SRS construction: g[i] = to_group(H(be32(i))), h = to_group(H("srs_misc" || be32(0))).
//
// Host-side replacement for poly-commitment `SRS::<G>::create(depth)` and groupmap `BWParameters`
// (lambdaclass/openmina-proof-systems @ 44e0d3b), which the reference evaluates lazily at first
// use: `SRS::<Vesta>::create(Fq::SRS_DEPTH)` at AL/operator/mina/lib/src/lib.rs:34 and
// `SRS::create(max_poly_size)` at AL/operator/mina/lib/src/verifier_index.rs:204-208.
// The committed srs/vesta.srs and srs/pallas.srs hold exactly these points (tests pin a SHA-256 of
// the derived arrays that was checked against those files).
#pragma once
#include <unistd.h>

#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "blake2b.hpp"
#include "host_field.hpp"

namespace pasta {
namespace host {

// Shallue-van de Woestijne map with u = 1 (groupmap `BWParameters::setup`).
template <class B>
struct GroupMap {
    Fe<B> fu, inv_three_u2, sqrt_neg_three_u2, sqrt_neg_three_u2_minus_u_over_2, five;
    GroupMap() {
        five = Fe<B>::from_u64(5);
        fu = Fe<B>::one() + five;  // u^3 + b with u = 1
        Fe<B> three = Fe<B>::from_u64(3);
        inv_three_u2 = three.inv();
        bool ok = (-three).sqrt(sqrt_neg_three_u2);
        (void)ok;
        sqrt_neg_three_u2_minus_u_over_2 = (sqrt_neg_three_u2 - Fe<B>::one()) * Fe<B>::from_u64(2).inv();
    }
    Affine<B> to_group(const Fe<B> &t) const {
        Fe<B> t2 = t.sqr();
        Fe<B> t2_fu = t2 + fu;
        Fe<B> alpha_inv = t2_fu * t2;
        Fe<B> alpha = alpha_inv.is_zero() ? alpha_inv : alpha_inv.inv();
        Fe<B> xs[3];
        xs[0] = sqrt_neg_three_u2_minus_u_over_2 - t2.sqr() * alpha * sqrt_neg_three_u2;
        xs[1] = -Fe<B>::one() - xs[0];
        xs[2] = Fe<B>::one() - t2_fu.sqr() * (alpha * t2_fu) * inv_three_u2;
        for (int k = 0; k < 3; k++) {
            Fe<B> y;
            if ((xs[k].sqr() * xs[k] + five).sqrt(y)) {
                Affine<B> p;
                p.x = xs[k];
                p.y = y;
                p.inf = false;
                return p;
            }
        }
        return Affine<B>::identity();  // unreachable for a valid map
    }
};

// poly-commitment `point_of_random_bytes`: first 31 digest bytes, bits LSB-first per byte, read as
// a big-endian bit string.
template <class B>
Fe<B> srs_hash_to_field(const uint8_t *msg, size_t len) {
    auto dg = Blake2b512::hash(msg, len);
    uint64_t limbs[4] = {0, 0, 0, 0};
    for (int i = 0; i < 31; i++)
        for (int j = 0; j < 8; j++)
            if ((dg[i] >> j) & 1) {
                int pos = 247 - (i * 8 + j);
                limbs[pos >> 6] |= 1ull << (pos & 63);
            }
    return Fe<B>::from_canonical(Fe<B>::raw(limbs[0], limbs[1], limbs[2], limbs[3]));
}

template <class B>
struct Srs {
    std::vector<Affine<B>> g;
    Affine<B> h;
};

template <class B>
Srs<B> srs_create(uint32_t depth, unsigned nthreads = 0) {
    Srs<B> srs;
    srs.g.resize(depth);
    GroupMap<B> gm;
    if (nthreads == 0) nthreads = std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min<unsigned>(nthreads, 64);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            for (uint32_t i = t; i < depth; i += nthreads) {
                uint8_t msg[4] = {(uint8_t)(i >> 24), (uint8_t)(i >> 16), (uint8_t)(i >> 8), (uint8_t)i};
                srs.g[i] = gm.to_group(srs_hash_to_field<B>(msg, 4));
            }
        });
    for (auto &th : pool) th.join();
    const uint8_t misc[12] = {'s', 'r', 's', '_', 'm', 'i', 's', 'c', 0, 0, 0, 0};
    srs.h = gm.to_group(srs_hash_to_field<B>(misc, 12));
    return srs;
}

// The committed SRS files (srs/vesta.srs, srs/pallas.srs in the reference tree):
// MessagePack `[ [bin(33) x n], bin(33) ]`, each 33-byte entry = ark-serialize 0.3 compressed point
// (32 B x little-endian, flag byte 0x80 <=> y > (p-1)/2, 0x40 = infinity).  `want` points are taken from
// the front (the Pallas side only ever uses the first 2^15 of the file's 2^16).
template <class B>
bool srs_decompress(const uint8_t *rec33, Affine<B> &out) {
    if (rec33[32] & 0x40) return false;  // infinity never occurs in an SRS
    if (rec33[32] & ~0x80u) return false;
    Fe<B> x;
    if (!Fe<B>::from_bytes_le(rec33, x)) return false;
    Fe<B> y;
    if (!(x.sqr() * x + Fe<B>::from_u64(5)).sqrt(y)) return false;
    if (y.is_lexicographically_large() != ((rec33[32] & 0x80) != 0)) y = -y;
    out.x = x;
    out.y = y;
    out.inf = false;
    return true;
}
template <class B>
bool srs_load_file(const std::string &path, uint32_t want, Srs<B> &out, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    for (size_t n; (n = std::fread(chunk, 1, sizeof chunk, f)) > 0;) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    size_t o = 0;
    auto need = [&](size_t k) { return o + k <= buf.size(); };
    if (!need(1) || buf[o++] != 0x92) {
        err = "not a 2-element MessagePack array";
        return false;
    }
    uint64_t count = 0;
    if (!need(1)) return false;
    uint8_t tag = buf[o++];
    if ((tag & 0xf0) == 0x90)
        count = tag & 0x0f;
    else if (tag == 0xdc && need(2)) {
        count = ((uint64_t)buf[o] << 8) | buf[o + 1];
        o += 2;
    } else if (tag == 0xdd && need(4)) {
        count = ((uint64_t)buf[o] << 24) | ((uint64_t)buf[o + 1] << 16) | ((uint64_t)buf[o + 2] << 8) | buf[o + 3];
        o += 4;
    } else {
        err = "bad array header";
        return false;
    }
    if (count < want) {
        err = "file holds fewer points than requested";
        return false;
    }
    if (!need((count + 1) * 35)) {
        err = "truncated file";
        return false;
    }
    const size_t first = o;
    for (uint64_t i = 0; i <= count; i++)
        if (buf[first + 35 * i] != 0xc4 || buf[first + 35 * i + 1] != 33) {
            err = "entry is not bin8(33)";
            return false;
        }
    out.g.resize(want);
    unsigned nthreads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 64);
    std::vector<std::thread> pool;
    std::vector<int> bad(nthreads, 0);
    for (unsigned t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            for (uint32_t i = t; i < want; i += nthreads)
                if (!srs_decompress<B>(&buf[first + 35 * (size_t)i + 2], out.g[i])) bad[t] = 1;
        });
    for (auto &th : pool) th.join();
    bool ok = srs_decompress<B>(&buf[first + 35 * count + 2], out.h);
    for (int b : bad) ok = ok && !b;
    if (!ok) err = "a point failed to decompress";
    return ok;
}

// Pinned digests of the derived SRS (Blake2b-512 over the (depth + 1) x 64-byte Montgomery payload:
// g[0..depth) then h).  The same points are pinned independently in tests/golden/srs_sha256.json
// against the reference's committed srs/{vesta,pallas}.srs.  The cache file is a trust root, so it
// is only accepted when its payload hashes to the pin compiled into this library.
template <class B>
inline const char *srs_pin(uint32_t depth) {
    if (B::ID == 1 && depth == (1u << 16))
        return "a1c1f9fab3f026a53700fcf42d67eaad8a01f1aad9c6938073764ed4f1950421d34838024998432b534d03b8a38f5a086e66ff1369eebb8797b6bab6430f9cb6";
    if (B::ID == 0 && depth == (1u << 15))
        return "7b928f3f7312de46acad01eb865ba564152c81eb8d2ba7d32ab7b06b055adf93c5fcd286b99b80e9219cfe4e7264614c101305a34f6d5cc63cf15a22d145e170";
    return nullptr;  // other depths (tests, config 2) are never cached
}
inline std::string hex_of(const uint8_t *b, size_t n) {
    static const char *d = "0123456789abcdef";
    std::string s;
    for (size_t i = 0; i < n; i++) {
        s += d[b[i] >> 4];
        s += d[b[i] & 15];
    }
    return s;
}
template <class B>
std::vector<uint64_t> srs_payload(const Srs<B> &srs) {
    std::vector<uint64_t> buf((srs.g.size() + 1) * 8);
    for (size_t i = 0; i <= srs.g.size(); i++) {
        const Affine<B> &p = i < srs.g.size() ? srs.g[i] : srs.h;
        std::memcpy(&buf[i * 8], p.x.l, 32);
        std::memcpy(&buf[i * 8 + 4], p.y.l, 32);
    }
    return buf;
}
template <class B>
bool srs_matches_pin(const Srs<B> &srs) {
    const char *pin = srs_pin<B>((uint32_t)srs.g.size());
    if (!pin) return false;
    auto buf = srs_payload(srs);
    auto dg = Blake2b512::hash(reinterpret_cast<const uint8_t *>(buf.data()), buf.size() * 8);
    return hex_of(dg.data(), 64) == pin;
}

// Flat cache: magic, field id, depth, 0, then the payload above.  Accepted only if the payload matches
// the compiled-in pin; written to a temporary name and renamed so a concurrent reader never sees a
// partial file.
template <class B>
bool srs_load_cache(const std::string &path, uint32_t depth, Srs<B> &out) {
    if (!srs_pin<B>(depth)) return false;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t hdr[4];
    bool ok = std::fread(hdr, sizeof hdr, 1, f) == 1 && hdr[0] == 0x53525342u && hdr[1] == (uint32_t)B::ID && hdr[2] == depth;
    std::vector<uint64_t> buf((size_t)(depth + 1) * 8);
    if (ok) ok = std::fread(buf.data(), 64, depth + 1, f) == depth + 1;
    std::fclose(f);
    if (!ok) return false;
    auto dg = Blake2b512::hash(reinterpret_cast<const uint8_t *>(buf.data()), buf.size() * 8);
    if (hex_of(dg.data(), 64) != srs_pin<B>(depth)) return false;
    out.g.resize(depth);
    for (uint32_t i = 0; i <= depth; i++) {
        Affine<B> &p = i < depth ? out.g[i] : out.h;
        std::memcpy(p.x.l, &buf[(size_t)i * 8], 32);
        std::memcpy(p.y.l, &buf[(size_t)i * 8 + 4], 32);
        p.inf = false;
    }
    return true;
}
template <class B>
bool srs_store_cache(const std::string &path, const Srs<B> &srs) {
    std::string tmp = path + ".tmp." + std::to_string((unsigned long)getpid());
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) return false;
    uint32_t hdr[4] = {0x53525342u, (uint32_t)B::ID, (uint32_t)srs.g.size(), 0};
    auto buf = srs_payload(srs);
    bool ok = std::fwrite(hdr, sizeof hdr, 1, f) == 1 && std::fwrite(buf.data(), 8, buf.size(), f) == buf.size();
    ok = ok && std::fflush(f) == 0 && fsync(fileno(f)) == 0;
    std::fclose(f);
    if (ok) ok = std::rename(tmp.c_str(), path.c_str()) == 0;
    if (!ok) std::remove(tmp.c_str());
    return ok;
}

}  // namespace host
}  // namespace pasta
