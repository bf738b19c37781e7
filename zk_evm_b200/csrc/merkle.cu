// Poseidon Merkle tree over the rows of a column-major, bit-reversed LDE.
//
// Replaces plonky2 1.0.0 hash/merkle_tree.rs `MerkleTree::new(leaves, cap_height)` + `PoseidonHash::hash_or_noop`
// / `two_to_one`, reached through PolynomialBatch::from_values at
// /root/reference/evm_arithmetization/src/prover.rs:100-107.
//
// Leaf kernel: one thread per leaf (LDE row).  Row j is lde[c*N + j] over columns c, so a warp's 32 lanes read 32
// consecutive 8-byte elements of every column: fully coalesced 256-byte requests, each column read exactly once.
// The 12-word sponge state lives in registers; the sponge overwrites state[0..8] with 8 columns per permutation.
// Inner levels: one thread per node, children are adjacent 32-byte digests.
#include "internal.h"
#include <stdlib.h>
#include "poseidon_fast.cuh"
#include "merkle.h"

namespace zk {

__global__ void __launch_bounds__(128) leaf_hash_kernel(const uint64_t* __restrict__ data, size_t stride, size_t ncols,
                                                        size_t nrows, uint64_t* __restrict__ digests) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nrows) return;
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    const uint64_t* p = data + j;
    if (ncols <= 4) {   // hash_or_noop: short rows are copied, zero padded
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((size_t)k < ncols) s[k] = p[k * stride];
    } else {
        // one permutation call site (the permutation body must stay resident in the instruction cache)
#pragma unroll 1
        for (size_t c = 0; c < ncols; c += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (c + k < ncols) s[k] = p[(c + k) * stride];
            pf_permute(s);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] = pf_canon(s[k]);   // the sponge state is non-canonical between permutations
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + 4 * j);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

__global__ void __launch_bounds__(128) merkle_level_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                           size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const ulonglong2* c = reinterpret_cast<const ulonglong2*>(in + 8 * i);
    ulonglong2 a = c[0], b = c[1], d = c[2], e = c[3];
    uint64_t s[12] = {a.x, a.y, b.x, b.y, d.x, d.y, e.x, e.y, 0, 0, 0, 0};
    pf_permute(s);
#pragma unroll
    for (int k = 0; k < 4; k++) s[k] = pf_canon(s[k]);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(out + 4 * i);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

// The top of the tree in ONE launch: block s finishes cap subtree s from the level that has 2 * blockDim.x of its nodes (or fewer)
// down to its cap entry.  Small levels are latency-bound — one permutation deep, a handful of warps wide — and there are 7-8 of them
// per tree and ~54 trees per segment proof: as separate launches they were 557 of a proof's 1136 launches (profiles/r1x).
struct TailArgs { uint64_t* digests; size_t off[12]; unsigned count0, nlevels; };   // off[l], l = 0 .. nlevels: input level, then outputs
__global__ void __launch_bounds__(128) merkle_tail_kernel(TailArgs a) {
    // level k of this subtree: count0 >> k nodes at digests + off[k] + 4 * blockIdx.x * (count0 >> k)
    for (unsigned k = 1; k <= a.nlevels; k++) {
        const unsigned cnt = a.count0 >> k;
        const uint64_t* in = a.digests + a.off[k - 1] + 4 * (size_t)blockIdx.x * (a.count0 >> (k - 1));
        uint64_t* out = a.digests + a.off[k] + 4 * (size_t)blockIdx.x * cnt;
        for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
            const ulonglong2* c = reinterpret_cast<const ulonglong2*>(in + 8 * i);
            ulonglong2 x = c[0], y = c[1], z = c[2], w = c[3];
            uint64_t s[12] = {x.x, x.y, y.x, y.y, z.x, z.y, w.x, w.y, 0, 0, 0, 0};
            pf_permute(s);
#pragma unroll
            for (int q = 0; q < 4; q++) s[q] = pf_canon(s[q]);
            ulonglong2* o = reinterpret_cast<ulonglong2*>(out + 4 * i);
            o[0] = make_ulonglong2(s[0], s[1]);
            o[1] = make_ulonglong2(s[2], s[3]);
        }
        __syncthreads();      // the level just written is read by other threads of this block only
    }
}

__global__ void poseidon_states_kernel(uint64_t* states, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
    pf_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) states[12 * i + k] = pf_canon(s[k]);
}

void poseidon_states(Ctx& c, uint64_t* dev_states, size_t count) {
    if (!count) return;
    poseidon_states_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c.stream>>>(dev_states, count);
    c.count_launch();
    c.check_launch("poseidon_states_kernel");
}

// Resident blocks per SM: 9 (56 registers x 128 threads).  Measured with a shared-memory occupancy limiter (tools/leafbench.py,
// profiles/r1o_leafbench.jsonl): throughput falls only 3 % from 9 to 5 resident blocks, i.e. five warps per sub-partition already
// keep the integer pipes busy, so uneven waves (Keccak: 2048 blocks on 148 x 9 slots) cost nothing and 9 is best for every shape.
void leaf_hash(Ctx& c, const uint64_t* data, size_t stride, size_t ncols, size_t nrows, uint64_t* digests) {
    KernelScope ks(c, KF_LEAF_HASH, (8.0 * ncols + 32.0) * nrows);
    static const char* env = getenv("ZKGPU_LEAF_BLOCKS");     // experiment knob of tools/leafbench.py: 5..8 resident blocks per SM
    const unsigned b = env && *env ? (unsigned)atoi(env) : 9;
    const size_t smem = b >= 9 || b < 1 ? 0 : (size_t)(227 * 1024) / (b + 1) + 1024;
    leaf_hash_kernel<<<(unsigned)((nrows + 127) / 128), 128, smem, c.stream>>>(data, stride, ncols, nrows, digests);
    c.count_launch();
    c.check_launch("leaf_hash_kernel");
}

void merkle_layout(size_t nleaves, unsigned cap_height, std::vector<size_t>& off, std::vector<size_t>& cnt) {
    off.clear(); cnt.clear();
    size_t o = 0, k = nleaves;
    for (;;) {
        off.push_back(o); cnt.push_back(k);
        o += 4 * k;
        if (k <= ((size_t)1 << cap_height)) break;
        k >>= 1;
    }
}

// digests: level 0 already filled with leaf digests
void merkle_inner_levels(Ctx& c, uint64_t* digests, const std::vector<size_t>& off, const std::vector<size_t>& cnt) {
    double nodes = 0;
    for (size_t l = 1; l < cnt.size(); l++) nodes += (double)cnt[l];
    KernelScope ks(c, KF_MERKLE_LEVELS, 96.0 * nodes);
    const size_t ncap = cnt.back();
    size_t l = 1;
    // wide levels: one launch each, one thread per node
    for (; l < off.size() && cnt[l - 1] / ncap > 256; l++) {
        merkle_level_kernel<<<(unsigned)((cnt[l] + 127) / 128), 128, 0, c.stream>>>(digests + off[l - 1], digests + off[l], cnt[l]);
        c.count_launch();
    }
    // the rest (each cap subtree has <= 256 nodes on the input level): one launch, one block per cap subtree
    if (l < off.size()) {
        TailArgs a;
        a.digests = digests;
        a.count0 = (unsigned)(cnt[l - 1] / ncap);
        a.nlevels = (unsigned)(off.size() - l);
        ZK_REQUIRE(a.nlevels < 12, "merkle tail: too many levels");
        for (size_t k = 0; k <= a.nlevels; k++) a.off[k] = off[l - 1 + k];
        merkle_tail_kernel<<<(unsigned)ncap, 128, 0, c.stream>>>(a);
        c.count_launch();
    }
    c.check_launch("merkle_level_kernel");
}

void merkle_build(Ctx& c, const uint64_t* rows_colmajor, size_t stride, size_t ncols, size_t nleaves,
                  unsigned cap_height, DevBuf& digests, std::vector<size_t>& off, std::vector<size_t>& cnt) {
    ZK_REQUIRE(((size_t)1 << cap_height) <= nleaves, "cap_height too large for the number of leaves");
    merkle_layout(nleaves, cap_height, off, cnt);
    digests = DevBuf(&c, (off.back() + 4 * cnt.back()) * 8);
    leaf_hash(c, rows_colmajor, stride, ncols, nleaves, digests.get());
    merkle_inner_levels(c, digests.get(), off, cnt);
}

// plonky2 `digests` layout from level arrays (host): per cap subtree, recursively
//   buf(left) | digest(left child) | digest(right child) | buf(right)
static void fill_rec(uint64_t* buf, const std::vector<uint64_t>& lv, const std::vector<size_t>& off, size_t lvl, size_t idx) {
    if (lvl == 0) return;
    size_t m = (size_t)1 << lvl, half_buf = m - 2;
    fill_rec(buf, lv, off, lvl - 1, 2 * idx);
    memcpy(buf + 4 * half_buf, &lv[off[lvl - 1] + 4 * (2 * idx)], 32);
    memcpy(buf + 4 * (half_buf + 1), &lv[off[lvl - 1] + 4 * (2 * idx + 1)], 32);
    fill_rec(buf + 4 * (half_buf + 2), lv, off, lvl - 1, 2 * idx + 1);
}
void merkle_export_plonky2(const std::vector<uint64_t>& levels_host, const std::vector<size_t>& off,
                           const std::vector<size_t>& cnt, uint64_t* out) {
    size_t ncap = cnt.back(), nleaves = cnt[0];
    size_t per = nleaves / ncap;
    if (per < 2) return;
    size_t sub = 2 * per - 2;
    for (size_t cidx = 0; cidx < ncap; cidx++) fill_rec(out + 4 * cidx * sub, levels_host, off, off.size() - 1, cidx);
}

}  // namespace zk
