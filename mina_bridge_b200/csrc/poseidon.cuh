// Poseidon permutation and Merkle-path fold for sm_100a (K3), table-driven -- see poseidon.hpp for
// the parity status of the constants (PARITY UNPINNED until a table passes merkle_verifier.rs:43-58).
//
// Replaces mina-poseidon `ArithmeticSponge` under `hash_with_kimchi` for batches of independent
// sponges: one thread per sponge; the permutation is 55 x (3 x x^7 + 3x3 MDS + round constant)
// = 55 x 21 field multiplications.  Parallelism comes from the batch (proofs x Merkle paths, proofs x
// protocol states), never from inside one sponge (the rounds are sequential).
// Reference call site: AL/operator/mina_account/lib/src/merkle_verifier.rs:9-35 (35 levels per proof).
#pragma once
#include <cuda_runtime.h>

#include "fe.cuh"

namespace pasta {

static constexpr int POSEIDON_WIDTH = 3, POSEIDON_ROUNDS = 55;
static constexpr int POSEIDON_TABLE_WORDS = 9 + 165;  // MDS row-major then rc[round][i], Montgomery

template <class F>
__device__ __forceinline__ void poseidon_permute(const fe *__restrict__ tab, fe st[3]) {
    for (int r = 0; r < POSEIDON_ROUNDS; r++) {
        fe sb[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            fe x2 = Fd<F>::sqr(st[i]);
            fe x4 = Fd<F>::sqr(x2);
            sb[i] = Fd<F>::mul(Fd<F>::mul(x4, x2), st[i]);
        }
#pragma unroll
        for (int i = 0; i < 3; i++)  // one MDS row = one dot product with a single reduction (Fd::dot3)
            st[i] = Fd<F>::add(Fd<F>::dot3(tab[3 * i], sb[0], tab[3 * i + 1], sb[1], tab[3 * i + 2], sb[2]), tab[9 + 3 * r + i]);
    }
}

// ---- a sponge on four lanes (lanes 0..2 hold one state element each, lane 3 idles) -----------------------------------
// One permutation is 55 dependent rounds; with the three S-boxes and the three MDS rows of a round on three
// lanes a round is 7 dependent multiplications instead of 21.
template <class F>
struct LaneSponge {
    fe st;            // this lane's state element, Montgomery
    bool absorbing;   // SpongeState::Absorbed(count) / Squeezed(count); uniform across the warp
    int count;
    const fe *tab;    // 9 MDS + 165 round constants, Montgomery
    uint32_t lane;    // 0..3 inside the group
    uint32_t base;    // first lane of the group inside the warp

    __device__ __forceinline__ fe from_lane(const fe &v, int l) const {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], base + l);
        return r;
    }
    __device__ void permute() {
        const uint32_t row = lane < 3 ? lane : 0;
#pragma unroll 1
        for (int r = 0; r < POSEIDON_ROUNDS; r++) {
            fe x2 = Fd<F>::sqr(st), x4 = Fd<F>::sqr(x2);
            fe sb = Fd<F>::mul(Fd<F>::mul(x4, x2), st);
            fe s0 = from_lane(sb, 0), s1 = from_lane(sb, 1), s2 = from_lane(sb, 2);
            st = Fd<F>::add(Fd<F>::dot3(tab[3 * row], s0, tab[3 * row + 1], s1, tab[3 * row + 2], s2), tab[9 + 3 * r + row]);
        }
    }
    // x: Montgomery, the same value on every lane of the group
    __device__ void absorb(const fe &x) {
        int slot;
        if (absorbing) {
            if (count == 2) {
                permute();
                slot = 0;
                count = 1;
            } else {
                slot = count;
                count++;
            }
        } else {
            slot = 0;
            absorbing = true;
            count = 1;
        }
        if ((int)lane == slot) st = Fd<F>::add(st, x);
    }
    // returns the squeezed element (Montgomery) on every lane of the group
    __device__ fe squeeze() {
        int slot;
        if (absorbing) {
            permute();
            absorbing = false;
            count = 1;
            slot = 0;
        } else if (count == 2) {
            permute();
            count = 1;
            slot = 0;
        } else {
            slot = count;
            count++;
        }
        return from_lane(st, slot);
    }
};

// states: n x 3 field elements, canonical in / canonical out (self-test + parity hook)
template <class F>
__global__ void __launch_bounds__(128) k_poseidon_permute(const fe *__restrict__ tab, fe *__restrict__ states, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe st[3];
#pragma unroll
    for (int k = 0; k < 3; k++) st[k] = Fd<F>::to_mont(states[(size_t)3 * i + k]);
    poseidon_permute<F>(tab, st);
#pragma unroll
    for (int k = 0; k < 3; k++) states[(size_t)3 * i + k] = Fd<F>::from_mont(st[k]);
}

// Merkle fold of nproofs independent paths.  nodes: [nproofs][max_depth] (tag in word 8 of a 48-byte
// record: 32 B canonical hash, u32 tag, padding); depths[p] = path length; leaves / roots canonical.
// prefix_states: [max_depth][3] Montgomery -- the sponge state after absorbing "MinaMklTree%03d" and
// squeezing once, precomputed per depth on the host.  ok[p] = (folded root == roots[p]).
struct MerkleNodeDev {
    uint32_t hash[8];
    uint32_t tag;
    uint32_t pad[3];
};
// Four lanes per path (LaneSponge above): 35 levels x 55 rounds x 7 dependent multiplications instead of x 21.  Every
// group of a warp walks max_depth levels (uniform control flow for the shuffles); a shorter path keeps its value.
template <class F>
__global__ void __launch_bounds__(128) k_merkle_fold(const fe *__restrict__ tab, const fe *__restrict__ prefix_states,
                                                     const MerkleNodeDev *__restrict__ nodes, const uint32_t *__restrict__ depths,
                                                     uint32_t max_depth, const fe *__restrict__ leaves, const fe *__restrict__ roots,
                                                     uint8_t *__restrict__ ok, fe *__restrict__ folded, uint32_t nproofs) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const uint32_t p = gid < nproofs ? gid : nproofs - 1;  // tail groups recompute the last path, no stores
    LaneSponge<F> sp;
    sp.lane = threadIdx.x & 3u;
    sp.base = threadIdx.x & 28u;
    sp.tab = tab;
    sp.absorbing = true;
    sp.count = 0;
    fe acc = Fd<F>::to_mont(leaves[p]);
    const uint32_t depth = depths[p];
    for (uint32_t d = 0; d < max_depth; d++) {
        const MerkleNodeDev &nd = nodes[(size_t)p * max_depth + (d < depth ? d : 0)];
        fe sib;
#pragma unroll
        for (int k = 0; k < 8; k++) sib.v[k] = nd.hash[k];
        sib = Fd<F>::to_mont(sib);
        // Left(sibling): [acc, sibling];  Right(sibling): [sibling, acc]   (merkle_verifier.rs:18-21)
        sp.st = prefix_states[3 * d + (sp.lane < 3 ? sp.lane : 0)];
        if (sp.lane == 0) sp.st = Fd<F>::add(sp.st, nd.tag ? sib : acc);
        if (sp.lane == 1) sp.st = Fd<F>::add(sp.st, nd.tag ? acc : sib);
        sp.permute();
        fe next = sp.from_lane(sp.st, 0);
        if (d < depth) acc = next;
    }
    if (gid >= nproofs || sp.lane != 0) return;
    fe out = Fd<F>::from_mont(acc);
    if (folded) folded[p] = out;
    ok[p] = fe_eq(out, roots[p]) ? 1 : 0;
}

void launch_poseidon_permute(int field, const fe *d_tab, fe *d_states, uint32_t n, cudaStream_t s);
void launch_merkle_fold(int field, const fe *d_tab, const fe *d_prefix_states, const MerkleNodeDev *d_nodes, const uint32_t *d_depths,
                        uint32_t max_depth, const fe *d_leaves, const fe *d_roots, uint8_t *d_ok, fe *d_folded, uint32_t nproofs,
                        cudaStream_t s);

}  // namespace pasta
