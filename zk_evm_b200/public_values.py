"""PublicValues -> the field elements `observe_public_values` feeds the challenger, in order
(evm_arithmetization/src/get_challenges.rs:11-227; limb helpers util.rs:40-58, 101-126; structs proof.rs:68-91, 314-321, 357-364, 398-425, 471-488).

`prove_with_traces(ctx, traces, flatten_public_values(pv), ...)`: the C ABI takes the flattened list, so that the host keeps its own
PublicValues type.  2217 elements with the default `eth_mainnet` feature: TrieRootsTarget::SIZE * 2 + BlockMetadataTarget::SIZE +
BlockHashesTarget::SIZE + ExtraBlockDataTarget::SIZE = 24 * 2 + 97 + 2056 + 16 (proof.rs:652-655, 981, 1193, 1382, 1469); registers and the
memory caps are not observed (get_challenges.rs:202-227).  H256 = 32 bytes (big-endian integer), U256 = Python int, Address = 20 bytes."""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

P = 0xFFFFFFFF00000001


class IntegerTooLarge(ValueError):
    """ProgramError::IntegerTooLarge (util.rs:40-58)"""


def u256_limbs(x):
    """util.rs:101-113: the eight 32-bit limbs, little-endian"""
    if not 0 <= x < 1 << 256:
        raise IntegerTooLarge("not a U256")
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def h256_limbs(h):
    """util.rs:116-126 (and observe_root, get_challenges.rs:11-19, which walks the same limbs as u64 halves)"""
    h = bytes(h)
    if len(h) != 32:
        raise ValueError("H256 is 32 bytes")
    return u256_limbs(int.from_bytes(h, "big"))


def u256_to_u32(x):
    if not 0 <= x < 1 << 32:
        raise IntegerTooLarge("%d does not fit 32 bits" % x)
    return [x]


def u256_to_u64(x):
    if not 0 <= x < 1 << 64:
        raise IntegerTooLarge("%d does not fit 64 bits" % x)
    return [x & 0xFFFFFFFF, x >> 32]


ZERO_H256 = bytes(32)


@dataclass
class TrieRoots:
    state_root: bytes = ZERO_H256
    transactions_root: bytes = ZERO_H256
    receipts_root: bytes = ZERO_H256


@dataclass
class BlockMetadata:
    block_beneficiary: bytes = bytes(20)
    block_timestamp: int = 0
    block_number: int = 0
    block_difficulty: int = 0
    block_random: bytes = ZERO_H256
    block_gaslimit: int = 0
    block_chain_id: int = 0
    block_base_fee: int = 0
    block_gas_used: int = 0
    block_blob_gas_used: int = 0
    block_excess_blob_gas: int = 0
    parent_beacon_block_root: bytes = ZERO_H256
    block_bloom: List[int] = field(default_factory=lambda: [0] * 8)


@dataclass
class BlockHashes:
    prev_hashes: List[bytes] = field(default_factory=lambda: [ZERO_H256] * 256)
    cur_hash: bytes = ZERO_H256


@dataclass
class ExtraBlockData:
    checkpoint_state_trie_root: bytes = ZERO_H256
    checkpoint_consolidated_hash: List[int] = field(default_factory=lambda: [0] * 4)
    txn_number_before: int = 0
    txn_number_after: int = 0
    gas_used_before: int = 0
    gas_used_after: int = 0


@dataclass
class PublicValues:
    trie_roots_before: TrieRoots = field(default_factory=TrieRoots)
    trie_roots_after: TrieRoots = field(default_factory=TrieRoots)
    block_metadata: BlockMetadata = field(default_factory=BlockMetadata)
    block_hashes: BlockHashes = field(default_factory=BlockHashes)
    extra_block_data: ExtraBlockData = field(default_factory=ExtraBlockData)
    burn_addr: Optional[int] = None          # cdk_erigon only


def _trie_roots(r):
    return h256_limbs(r.state_root) + h256_limbs(r.transactions_root) + h256_limbs(r.receipts_root)


def _block_metadata(m, eth_mainnet):
    if len(m.block_beneficiary) != 20 or len(m.block_bloom) != 8:
        raise ValueError("beneficiary is 20 bytes, the bloom filter 8 words")
    out = u256_limbs(int.from_bytes(m.block_beneficiary, "big"))[:5]
    out += u256_to_u32(m.block_timestamp) + u256_to_u32(m.block_number) + u256_to_u32(m.block_difficulty)
    out += h256_limbs(m.block_random)
    out += u256_to_u32(m.block_gaslimit) + u256_to_u32(m.block_chain_id)
    out += u256_to_u64(m.block_base_fee)
    out += u256_to_u32(m.block_gas_used)
    if eth_mainnet:
        out += u256_to_u64(m.block_blob_gas_used) + u256_to_u64(m.block_excess_blob_gas) + h256_limbs(m.parent_beacon_block_root)
    for w in m.block_bloom:
        out += u256_limbs(w)
    return out


def _block_hashes(b):
    if len(b.prev_hashes) != 256:
        raise ValueError("256 previous block hashes")
    out = []
    for h in b.prev_hashes:
        out += h256_limbs(h)
    return out + h256_limbs(b.cur_hash)


def _extra_block_data(e):
    if len(e.checkpoint_consolidated_hash) != 4 or any(not 0 <= int(x) < P for x in e.checkpoint_consolidated_hash):
        raise ValueError("the consolidated hash is 4 canonical field elements")
    return (h256_limbs(e.checkpoint_state_trie_root) + [int(x) for x in e.checkpoint_consolidated_hash] + u256_to_u32(e.txn_number_before)
            + u256_to_u32(e.txn_number_after) + u256_to_u32(e.gas_used_before) + u256_to_u32(e.gas_used_after))


def flatten_public_values(pv, eth_mainnet=True, cdk_erigon=False):
    """observe_public_values (get_challenges.rs:202-227) as the list of observed elements -> uint64 array"""
    out = _trie_roots(pv.trie_roots_before) + _trie_roots(pv.trie_roots_after) + _block_metadata(pv.block_metadata, eth_mainnet) \
        + _block_hashes(pv.block_hashes) + _extra_block_data(pv.extra_block_data)
    if cdk_erigon:
        if pv.burn_addr is None:
            raise ValueError("There should be an address set in cdk_erigon.")
        out += u256_limbs(pv.burn_addr)
    return np.array(out, dtype=np.uint64)
