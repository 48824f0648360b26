// Host-only C ABI hooks: exercise the C++ host arithmetic (field, sqrt, group map, SRS derivation)
// without a GPU, so the CPU test tier can check it against the oracle.
#include <cstring>
#include <stdexcept>

#include "../../include/mina_b200.h"
#include "context.cuh"

using namespace pasta;

namespace {
template <class F>
int host_field_op_t(int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
    using E = host::Fe<F>;
    for (uint32_t i = 0; i < n; i++) {
        E a, b;
        if (!E::from_bytes_le(a32 + 32 * (size_t)i, a)) return -2;
        if (!E::from_bytes_le((b32 ? b32 : a32) + 32 * (size_t)i, b)) return -2;
        E r;
        switch (op) {
            case 0: r = a * b; break;
            case 1: r = a + b; break;
            case 2: r = a - b; break;
            case 3: r = a.inv(); break;
            case 4: r = a.sqr(); break;
            case 5:
                if (!a.sqrt(r)) r = E::zero();
                break;
            default: return -3;
        }
        r.to_bytes_le(out32 + 32 * (size_t)i);
    }
    return 0;
}
template <class B>
void host_srs_t(uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64) {
    host::GroupMap<B> gm;
    for (uint32_t k = 0; k < count; k++) {
        uint32_t i = first + k;
        uint8_t msg[4] = {(uint8_t)(i >> 24), (uint8_t)(i >> 16), (uint8_t)(i >> 8), (uint8_t)i};
        host::Affine<B> p = gm.to_group(host::srs_hash_to_field<B>(msg, 4));
        p.x.to_bytes_le(out64 + 64 * (size_t)k);
        p.y.to_bytes_le(out64 + 64 * (size_t)k + 32);
    }
    if (h64) {
        const uint8_t misc[12] = {'s', 'r', 's', '_', 'm', 'i', 's', 'c', 0, 0, 0, 0};
        host::Affine<B> p = gm.to_group(host::srs_hash_to_field<B>(misc, 12));
        p.x.to_bytes_le(h64);
        p.y.to_bytes_le(h64 + 32);
    }
}
}  // namespace

extern "C" {

int mina_b200_host_field_op(int field, int op, uint32_t n, const uint8_t *a32, const uint8_t *b32, uint8_t *out32) {
    if (field == 0) return host_field_op_t<FpParams>(op, n, a32, b32, out32);
    if (field == 1) return host_field_op_t<FqParams>(op, n, a32, b32, out32);
    return -1;
}

int mina_b200_host_srs_derive(int curve, uint32_t first, uint32_t count, uint8_t *out64, uint8_t *h64) {
    if (curve == 0)
        host_srs_t<FpParams>(first, count, out64, h64);
    else if (curve == 1)
        host_srs_t<FqParams>(first, count, out64, h64);
    else
        return -1;
    return 0;
}

int mina_b200_host_build_srs_cache(const char *cache_dir) {
    if (!cache_dir || !*cache_dir) return -1;
    try {
        std::string dir(cache_dir);
        {
            std::string path = dir + "/vesta_" + std::to_string(VESTA_SRS_DEPTH) + ".srsbin";
            host::Srs<FqParams> s;
            if (!host::srs_load_cache<FqParams>(path, VESTA_SRS_DEPTH, s)) {
                s = host::srs_create<FqParams>(VESTA_SRS_DEPTH);
                if (!host::srs_store_cache<FqParams>(path, s)) return -2;
            }
        }
        {
            std::string path = dir + "/pallas_" + std::to_string(PALLAS_SRS_DEPTH) + ".srsbin";
            host::Srs<FpParams> s;
            if (!host::srs_load_cache<FpParams>(path, PALLAS_SRS_DEPTH, s)) {
                s = host::srs_create<FpParams>(PALLAS_SRS_DEPTH);
                if (!host::srs_store_cache<FpParams>(path, s)) return -2;
            }
        }
    } catch (...) {
        return -3;
    }
    return 0;
}

int mina_b200_host_blake2b512(const uint8_t *data, size_t len, uint8_t out[64]) {
    auto dg = host::Blake2b512::hash(data, len);
    std::memcpy(out, dg.data(), 64);
    return 0;
}

}  // extern "C"
