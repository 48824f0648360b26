// Launchers for the Poseidon kernels (poseidon.cuh), dispatched on the field id.
#include <stdexcept>

#include "poseidon.cuh"

namespace pasta {

void launch_poseidon_permute(int field, const fe *d_tab, fe *d_states, uint32_t n, cudaStream_t s) {
    if (field != 0 && field != 1) throw std::runtime_error("bad field id");
    if (!n) return;
    dim3 g((n + 127) / 128);
    if (field == 0)
        k_poseidon_permute<FpParams><<<g, 128, 0, s>>>(d_tab, d_states, n);
    else
        k_poseidon_permute<FqParams><<<g, 128, 0, s>>>(d_tab, d_states, n);
}

void launch_merkle_fold(int field, const fe *d_tab, const fe *d_prefix_states, const MerkleNodeDev *d_nodes, const uint32_t *d_depths,
                        uint32_t max_depth, const fe *d_leaves, const fe *d_roots, uint8_t *d_ok, fe *d_folded, uint32_t nproofs,
                        cudaStream_t s) {
    if (field != 0 && field != 1) throw std::runtime_error("bad field id");
    if (!nproofs) return;
    dim3 g((4 * nproofs + 127) / 128);  // four lanes per path
    if (field == 0)
        k_merkle_fold<FpParams><<<g, 128, 0, s>>>(d_tab, d_prefix_states, d_nodes, d_depths, max_depth, d_leaves, d_roots, d_ok, d_folded, nproofs);
    else
        k_merkle_fold<FqParams><<<g, 128, 0, s>>>(d_tab, d_prefix_states, d_nodes, d_depths, max_depth, d_leaves, d_roots, d_ok, d_folded, nproofs);
}

}  // namespace pasta
