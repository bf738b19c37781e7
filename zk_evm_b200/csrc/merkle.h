// internal Merkle / Poseidon interface (see merkle.cu)
#pragma once
#include "internal.h"
namespace zk {
void poseidon_states(Ctx& c, uint64_t* dev_states, size_t count);
// digests[4*j..] = hash_or_noop(row j), row j = data[c*stride + j] over c < ncols
void leaf_hash(Ctx& c, const uint64_t* data, size_t stride, size_t ncols, size_t nrows, uint64_t* digests);
void merkle_layout(size_t nleaves, unsigned cap_height, std::vector<size_t>& off, std::vector<size_t>& cnt);
void merkle_inner_levels(Ctx& c, uint64_t* digests, const std::vector<size_t>& off, const std::vector<size_t>& cnt);
void merkle_build(Ctx& c, const uint64_t* rows_colmajor, size_t stride, size_t ncols, size_t nleaves,
                  unsigned cap_height, DevBuf& digests, std::vector<size_t>& off, std::vector<size_t>& cnt);
void merkle_export_plonky2(const std::vector<uint64_t>& levels_host, const std::vector<size_t>& off,
                           const std::vector<size_t>& cnt, uint64_t* out);
}  // namespace zk
