// A commitment split over several devices: the pieces of PolynomialBatch::from_values
// (/root/reference/evm_arithmetization/src/prover.rs:100-107; plonky2 1.0.0 fri/oracle.rs, hash/merkle_tree.rs) that k devices
// compute side by side for ONE table of a segment (the table-sharded layout; DESIGN.md "Multi-GPU").
//
//   columns  ->  ifft + LDE of a column slice             zkgpu_lde_slice      (per-column transforms: split by columns)
//   rows     ->  leaf digests + Merkle levels of a block  zkgpu_merkle_block   (the sponge absorbs a row's columns in order: split by rows)
//   owner    ->  batch over the exchanged buffers         zkgpu_batch_assemble
//
// The exchange between the steps is the caller's (an NCCL all-gather over NVLink in zk_evm_b200/segment.py); nothing here knows
// about other devices.  Bit-identical to the one-device commitment: the same kernels run on sub-ranges.
#include "internal.h"
#include "ntt.h"
#include "merkle.h"
#include <memory>

namespace zk {
void init_batch(Ctx& c, Batch& b, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height);

// level l of a block: count[l] = (nleaves >> l) / nblocks digests at word offset off[l] of the packed block
static void block_layout(size_t nleaves, unsigned cap_height, unsigned nblocks, std::vector<size_t>& off, std::vector<size_t>& cnt) {
    ZK_REQUIRE(nblocks >= 1 && (nblocks & (nblocks - 1)) == 0 && nblocks <= ((size_t)1 << cap_height), "nblocks must be a power of two that divides the cap");
    ZK_REQUIRE(((size_t)1 << cap_height) <= nleaves, "cap_height too large for the number of leaves");
    std::vector<size_t> foff, fcnt;
    merkle_layout(nleaves, cap_height, foff, fcnt);
    off.clear(); cnt.clear();
    size_t o = 0;
    for (size_t k : fcnt) { off.push_back(o); cnt.push_back(k / nblocks); o += 4 * (k / nblocks); }
    off.push_back(o);      // total words
}
}  // namespace zk

using namespace zk;

extern "C" {

int zkgpu_lde_slice(zkgpu_ctx* h, const uint64_t* values, int mem_kind, size_t ncols, size_t n, uint32_t rate_bits,
                    uint64_t* values_out, uint64_t* coeffs_out, uint64_t* lde_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && (ncols == 0 || (values && coeffs_out && lde_out)), "null argument");
    ZK_REQUIRE(rate_bits >= 1, "the LDE buffer doubles as transform scratch: rate_bits >= 1");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    if (!ncols) return ZKGPU_OK;
    const unsigned L = log2_exact(n);
    const uint64_t* v = values;
    DevBuf tmp;
    if (mem_kind != ZKGPU_MEM_DEVICE) {
        uint64_t* dst = values_out;
        if (!dst) { tmp = DevBuf(&c, ncols * n * 8); dst = tmp.get(); }
        ZK_CUDA(cudaMemcpyAsync(dst, values, ncols * n * 8, mem_kind == ZKGPU_MEM_AUTO ? cudaMemcpyDefault : cudaMemcpyHostToDevice, c.stream));
        v = dst;
    } else if (values_out && values_out != values) {
        ZK_CUDA(cudaMemcpyAsync(values_out, values, ncols * n * 8, cudaMemcpyDeviceToDevice, c.stream));
    }
    intt_natural(c, v, lde_out, coeffs_out, ncols, L, 0);
    lde_bitrev(c, coeffs_out, lde_out, ncols, L, rate_bits, GL_GENERATOR);
    ZK_API_END
}

int zkgpu_merkle_block_words(size_t nleaves, uint32_t cap_height, uint32_t nblocks, size_t* words) {
    ZK_API_BEGIN
    ZK_REQUIRE(words, "null argument");
    std::vector<size_t> off, cnt;
    block_layout(nleaves, cap_height, nblocks, off, cnt);
    *words = off.back();
    ZK_API_END
}

int zkgpu_merkle_block(zkgpu_ctx* h, const uint64_t* lde, size_t stride, size_t ncols, size_t nleaves, uint32_t cap_height,
                       uint32_t nblocks, uint32_t block, uint64_t* packed_out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && lde && packed_out && ncols, "null argument");
    ZK_REQUIRE(block < nblocks, "block index out of range");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::vector<size_t> off, cnt;
    block_layout(nleaves, cap_height, nblocks, off, cnt);
    off.pop_back();
    leaf_hash(c, lde + (size_t)block * cnt[0], stride, ncols, cnt[0], packed_out);
    merkle_inner_levels(c, packed_out, off, cnt);
    ZK_API_END
}

int zkgpu_batch_assemble(zkgpu_ctx* h, const uint64_t* values, const uint64_t* coeffs, const uint64_t* lde, const uint64_t* packed,
                         uint32_t nblocks, size_t ncols, size_t n, uint32_t rate_bits, uint32_t cap_height, zkgpu_batch** out) {
    ZK_API_BEGIN
    ZK_REQUIRE(h && coeffs && lde && packed && out, "null argument");
    Ctx& c = h->c;
    ZK_CUDA(cudaSetDevice(c.device));
    std::unique_ptr<zkgpu_batch> hb(new zkgpu_batch());
    Batch& b = hb->b;
    init_batch(c, b, ncols, n, rate_bits, cap_height);
    if (values) b.values = DevBuf::borrowed(values, ncols * n * 8);
    b.coeffs = DevBuf::borrowed(coeffs, ncols * n * 8);
    b.lde = DevBuf::borrowed(lde, ncols * b.N * 8);
    std::vector<size_t> boff, bcnt;
    block_layout(b.N, cap_height, nblocks, boff, bcnt);
    const size_t words = boff.back();
    merkle_layout(b.N, cap_height, b.level_off, b.level_cnt);
    b.digests = DevBuf(&c, (b.level_off.back() + 4 * b.level_cnt.back()) * 8);
    // level l of the tree = the nblocks block slices of level l, block after block
    for (size_t l = 0; l < b.level_cnt.size(); l++) {
        const size_t w = 4 * bcnt[l] * 8;     // bytes of one block's slice of this level
        ZK_CUDA(cudaMemcpy2DAsync(b.digests.get() + b.level_off[l], w, packed + boff[l], words * 8, w, nblocks, cudaMemcpyDeviceToDevice, c.stream));
    }
    b.cap_host.resize(4 * b.level_cnt.back());
    c.d2h(b.cap_host.data(), b.cap_dev(), b.cap_host.size() * 8);
    *out = hb.release();
    ZK_API_END
}

}  // extern "C"
