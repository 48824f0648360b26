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
        for (int i = 0; i < 3; i++) {
            fe acc = Fd<F>::mul(tab[3 * i], sb[0]);
            acc = Fd<F>::add(acc, Fd<F>::mul(tab[3 * i + 1], sb[1]));
            acc = Fd<F>::add(acc, Fd<F>::mul(tab[3 * i + 2], sb[2]));
            st[i] = Fd<F>::add(acc, tab[9 + 3 * r + i]);
        }
    }
}

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
template <class F>
__global__ void __launch_bounds__(64) k_merkle_fold(const fe *__restrict__ tab, const fe *__restrict__ prefix_states,
                                                    const MerkleNodeDev *__restrict__ nodes, const uint32_t *__restrict__ depths,
                                                    uint32_t max_depth, const fe *__restrict__ leaves, const fe *__restrict__ roots,
                                                    uint8_t *__restrict__ ok, fe *__restrict__ folded, uint32_t nproofs) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nproofs) return;
    fe acc = Fd<F>::to_mont(leaves[p]);
    const uint32_t depth = depths[p];
    for (uint32_t d = 0; d < depth; d++) {
        const MerkleNodeDev &nd = nodes[(size_t)p * max_depth + d];
        fe sib;
#pragma unroll
        for (int k = 0; k < 8; k++) sib.v[k] = nd.hash[k];
        sib = Fd<F>::to_mont(sib);
        fe st[3];
#pragma unroll
        for (int k = 0; k < 3; k++) st[k] = prefix_states[3 * d + k];
        // Left(sibling): [acc, sibling];  Right(sibling): [sibling, acc]   (merkle_verifier.rs:18-21)
        st[0] = Fd<F>::add(st[0], nd.tag ? sib : acc);
        st[1] = Fd<F>::add(st[1], nd.tag ? acc : sib);
        poseidon_permute<F>(tab, st);
        acc = st[0];
    }
    fe out = Fd<F>::from_mont(acc);
    if (folded) folded[p] = out;
    ok[p] = fe_eq(out, roots[p]) ? 1 : 0;
}

void launch_poseidon_permute(int field, const fe *d_tab, fe *d_states, uint32_t n, cudaStream_t s);
void launch_merkle_fold(int field, const fe *d_tab, const fe *d_prefix_states, const MerkleNodeDev *d_nodes, const uint32_t *d_depths,
                        uint32_t max_depth, const fe *d_leaves, const fe *d_roots, uint8_t *d_ok, fe *d_folded, uint32_t nproofs,
                        cudaStream_t s);

}  // namespace pasta
