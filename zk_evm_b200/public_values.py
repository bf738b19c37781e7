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
class RegistersData:
    """proof.rs:537-550; not observed by the challenger (get_challenges.rs:202-227), but written to memory: see memory_extra_looking_values"""
    program_counter: int = 0
    is_kernel: int = 0
    stack_len: int = 0
    stack_top: int = 0
    context: int = 0
    gas_used: int = 0


@dataclass
class PublicValues:
    trie_roots_before: TrieRoots = field(default_factory=TrieRoots)
    trie_roots_after: TrieRoots = field(default_factory=TrieRoots)
    block_metadata: BlockMetadata = field(default_factory=BlockMetadata)
    block_hashes: BlockHashes = field(default_factory=BlockHashes)
    extra_block_data: ExtraBlockData = field(default_factory=ExtraBlockData)
    burn_addr: Optional[int] = None          # cdk_erigon only
    registers_before: RegistersData = field(default_factory=RegistersData)
    registers_after: RegistersData = field(default_factory=RegistersData)


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


# ---- the memory writes that come from the public values, not from Cpu rows (verifier.rs:315-512, 536-737) -----------------------
# GlobalMetadata in declaration order (cpu/kernel/constants/global_metadata.rs:11-115): a field's index inside Segment::GlobalMetadata
# is its position here (`unscale`)
GLOBAL_METADATA = (
    "LargestContext", "MemorySize", "TrieDataSize", "StateTrieRoot", "TransactionTrieRoot", "ReceiptTrieRoot",
    "StateTrieRootDigestBefore", "TransactionTrieRootDigestBefore", "ReceiptTrieRootDigestBefore",
    "StateTrieRootDigestAfter", "TransactionTrieRootDigestAfter", "ReceiptTrieRootDigestAfter",
    "BlockBeneficiary", "BlockTimestamp", "BlockNumber", "BlockDifficulty", "BlockRandom", "BlockGasLimit", "BlockChainId", "BlockBaseFee",
    "BlockBlobGasUsed", "BlockExcessBlobGas", "BlockGasUsed", "BlockGasUsedBefore", "BlockGasUsedAfter", "BlockCurrentHash",
    "ParentBeaconBlockRoot", "RefundCounter", "AccessedAddressesLen", "AccessedStorageKeysLen", "SelfDestructListLen", "JournalLen",
    "JournalDataLen", "CurrentCheckpoint", "TouchedAddressesLen", "AccessListDataCost", "ContractCreation", "IsPrecompileFromEoa",
    "CallStackDepth", "LogsLen", "LogsDataLen", "LogsPayloadLen", "TxnNumberBefore", "TxnNumberAfter", "CreatedContractsLen",
    "KernelHash", "KernelLen", "AccountsLinkedListNextAvailable", "StorageLinkedListNextAvailable", "InitialAccountsLinkedListLen",
    "InitialStorageLinkedListLen", "TransientStorageLen", "BlobVersionedHashesLen", "BurnAddr")
assert len(GLOBAL_METADATA) == 54            # GlobalMetadata::COUNT
# memory/segments.rs:10-91, unscaled
SEGMENT_GLOBAL_METADATA, SEGMENT_GLOBAL_BLOCK_BLOOM, SEGMENT_BLOCK_HASHES, SEGMENT_REGISTERS_STATES = 5, 24, 32, 33


def _h2u(h):
    h = bytes(h)
    if len(h) != 32:
        raise ValueError("H256 is 32 bytes")
    return int.from_bytes(h, "big")


def _registers(r):
    return [r.program_counter, r.is_kernel, r.stack_len, r.stack_top, r.context, r.gas_used]


def memory_extra_looking_values(pv, kernel_code_hash, kernel_code_len, eth_mainnet=True, cdk_erigon=False):
    """verifier.rs:547-737 `get_memory_extra_looking_values`: the Memory-table rows that no Cpu row sends — the kernel's writes of the
    block metadata, trie roots, kernel hash / length, block bloom, previous block hashes and the registers before / after — as the
    13 values of the memory lookup (is_read = 0, context 0, segment, index, eight 32-bit value limbs, timestamp 2), in the reference's
    order.  `kernel_code_hash` (32 bytes) and `kernel_code_len` stand for KERNEL.code_hash and KERNEL.code.len(): the kernel is
    assembled by the host, like the four kernel labels of the Cpu constraints."""
    m, e = pv.block_metadata, pv.extra_block_data
    fields = [("BlockBeneficiary", int.from_bytes(m.block_beneficiary, "big"))]
    if cdk_erigon:
        if pv.burn_addr is None:
            raise ValueError("There should be an address set in cdk_erigon.")
        fields.append(("BurnAddr", pv.burn_addr))
    fields += [("BlockTimestamp", m.block_timestamp), ("BlockNumber", m.block_number), ("BlockRandom", _h2u(m.block_random)),
               ("BlockDifficulty", m.block_difficulty), ("BlockGasLimit", m.block_gaslimit), ("BlockChainId", m.block_chain_id),
               ("BlockBaseFee", m.block_base_fee), ("BlockCurrentHash", _h2u(pv.block_hashes.cur_hash)), ("BlockGasUsed", m.block_gas_used)]
    if eth_mainnet:
        fields += [("BlockBlobGasUsed", m.block_blob_gas_used), ("BlockExcessBlobGas", m.block_excess_blob_gas),
                   ("ParentBeaconBlockRoot", _h2u(m.parent_beacon_block_root))]
    fields += [("TxnNumberBefore", e.txn_number_before), ("TxnNumberAfter", e.txn_number_after),
               ("BlockGasUsedBefore", e.gas_used_before), ("BlockGasUsedAfter", e.gas_used_after),
               ("StateTrieRootDigestBefore", _h2u(pv.trie_roots_before.state_root)),
               ("TransactionTrieRootDigestBefore", _h2u(pv.trie_roots_before.transactions_root)),
               ("ReceiptTrieRootDigestBefore", _h2u(pv.trie_roots_before.receipts_root)),
               ("StateTrieRootDigestAfter", _h2u(pv.trie_roots_after.state_root)),
               ("TransactionTrieRootDigestAfter", _h2u(pv.trie_roots_after.transactions_root)),
               ("ReceiptTrieRootDigestAfter", _h2u(pv.trie_roots_after.receipts_root)),
               ("KernelHash", _h2u(kernel_code_hash)), ("KernelLen", int(kernel_code_len))]

    def row(segment, index, val):        # add_extra_looking_row, verifier.rs:720-735 (add_data_write :492-512 builds the same row)
        return [0, 0, segment, index] + u256_limbs(val) + [2]
    rows = [row(SEGMENT_GLOBAL_METADATA, GLOBAL_METADATA.index(name), val) for name, val in fields]
    if len(m.block_bloom) != 8 or len(pv.block_hashes.prev_hashes) != 256:
        raise ValueError("the bloom filter is 8 words, there are 256 previous block hashes")
    rows += [row(SEGMENT_GLOBAL_BLOCK_BLOOM, i, m.block_bloom[i]) for i in range(8)]
    rows += [row(SEGMENT_BLOCK_HASHES, i, _h2u(pv.block_hashes.prev_hashes[i])) for i in range(256)]
    before, after = _registers(pv.registers_before), _registers(pv.registers_after)
    rows += [row(SEGMENT_REGISTERS_STATES, i, before[i]) for i in range(6)]
    rows += [row(SEGMENT_REGISTERS_STATES, 6 + i, after[i]) for i in range(6)]
    return rows


def memory_extra_looking_sum(pv, beta, gamma, kernel_code_hash, kernel_code_len, eth_mainnet=True, cdk_erigon=False):
    """verifier.rs:319-512 `get_memory_extra_looking_sum`: what the verifier adds to the looking side of the Memory lookup for one
    challenge (beta, gamma) — sum over those rows of 1 / (gamma + sum_i row_i beta^i)"""
    total = 0
    for r in memory_extra_looking_values(pv, kernel_code_hash, kernel_code_len, eth_mainnet, cdk_erigon):
        acc = 0
        for v in reversed(r):
            acc = (acc * beta + v) % P
        total = (total + pow((acc + gamma) % P, P - 2, P)) % P
    return total
