"""flatten_public_values against get_challenges.rs:11-227 FIELD BY FIELD.

Every 32-bit limb of every field of a hand-built PublicValues carries its own tag, and the expected list is written out here from the
reference's observe_* functions line by line (not produced by the code under test): position, limb order inside a field (H256 fields
through `into_uint().0` u64 limbs split low half first, get_challenges.rs:15-18 == h256_limbs util.rs:116-126; U256 fields little-endian
32-bit limbs util.rs:101-113; the 20-byte beneficiary as the low 5 limbs of its big-endian integer, get_challenges.rs:56-58; base fee /
blob gas as (low, high) 32-bit halves, util.rs:49-58), field order and section order (get_challenges.rs:210-214)."""
import numpy as np
import pytest
from zk_evm_b200.public_values import (PublicValues, TrieRoots, BlockMetadata, BlockHashes, ExtraBlockData, flatten_public_values,
                                       IntegerTooLarge)


def tagged_h256(tag):
    """32 bytes whose little-endian 32-bit limb i (of the big-endian integer) is (tag << 8) | i; returns (bytes, expected limbs)"""
    limbs = [(tag << 8) | i for i in range(8)]
    value = sum(l << (32 * i) for i, l in enumerate(limbs))
    return value.to_bytes(32, "big"), limbs


def tagged_u256(tag):
    limbs = [(tag << 8) | i for i in range(8)]
    return sum(l << (32 * i) for i, l in enumerate(limbs)), limbs


def test_flatten_public_values_field_by_field():
    exp = []
    # observe_trie_roots(before), observe_trie_roots(after): state, transactions, receipts (get_challenges.rs:21-29, 210-211)
    roots = {}
    for side, base in (("before", 0x10), ("after", 0x20)):
        parts = []
        for j in range(3):
            b, l = tagged_h256(base + j)
            parts.append(b)
            exp += l
        roots[side] = TrieRoots(*parts)
    # observe_block_metadata (get_challenges.rs:47-86)
    ben_limbs = [0xBE0000 | i for i in range(5)]
    beneficiary = sum(l << (32 * i) for i, l in enumerate(ben_limbs)).to_bytes(20, "big")
    exp += ben_limbs                                            # :56-58  low five limbs
    exp += [0x7157A0, 0xB10C40, 0xD1FF00]                       # :59-61  timestamp, number, difficulty (one u32 each)
    rnd, l = tagged_h256(0x31)
    exp += l                                                    # :62     block_random
    exp += [0x6A5117, 0xC4A1D0]                                 # :63-64  gaslimit, chain id
    exp += [0xBA5EFEE0, 0x000000B1]                             # :65-67  base fee (low, high)
    exp += [0x6A5ED0]                                           # :68     gas used
    exp += [0xB10B6A50, 0x000000B2]                             # :71-73  blob gas used (low, high)      [eth_mainnet]
    exp += [0xE8CE5500, 0x000000B3]                             # :74-76  excess blob gas (low, high)    [eth_mainnet]
    pbr, l = tagged_h256(0x32)
    exp += l                                                    # :77     parent beacon block root       [eth_mainnet]
    bloom = []
    for i in range(8):                                          # :79-81  eight U256 words
        v, l = tagged_u256(0x40 + i)
        bloom.append(v)
        exp += l
    meta = BlockMetadata(block_beneficiary=beneficiary, block_timestamp=0x7157A0, block_number=0xB10C40, block_difficulty=0xD1FF00,
                         block_random=rnd, block_gaslimit=0x6A5117, block_chain_id=0xC4A1D0, block_base_fee=(0xB1 << 32) | 0xBA5EFEE0,
                         block_gas_used=0x6A5ED0, block_blob_gas_used=(0xB2 << 32) | 0xB10B6A50,
                         block_excess_blob_gas=(0xB3 << 32) | 0xE8CE5500, parent_beacon_block_root=pbr, block_bloom=bloom)
    # observe_block_hashes (get_challenges.rs:172-185): 256 previous hashes in order, then the current one
    prev = []
    for i in range(256):
        b, l = tagged_h256(0x1000 + i)
        prev.append(b)
        exp += l
    cur, l = tagged_h256(0x2000)
    exp += l
    # observe_extra_block_data (get_challenges.rs:113-129)
    ck, l = tagged_h256(0x51)
    exp += l                                                    # :122 checkpoint_state_trie_root
    cons = [0xFFFFFFFF00000000, 0x1234567890ABCDEF, 7, 0]
    exp += cons                                                 # :123 four field elements as they are
    exp += [0xA1, 0xA2, 0xA3, 0xA4]                             # :124-127 txn_number_before/after, gas_used_before/after
    extra = ExtraBlockData(ck, cons, 0xA1, 0xA2, 0xA3, 0xA4)
    pv = PublicValues(roots["before"], roots["after"], meta, BlockHashes(prev, cur), extra)
    got = flatten_public_values(pv)
    assert got.dtype == np.uint64 and len(got) == 24 * 2 + 97 + 2056 + 16 == 2217      # Target sizes, proof.rs:652-655, 981, 1193, 1382, 1469
    assert [int(x) for x in got] == exp
    # sections start where the sizes say
    assert int(got[48]) == 0xBE0000 and int(got[48 + 97]) == (0x1000 << 8) and int(got[48 + 97 + 2056]) == (0x51 << 8)
    # without eth_mainnet the three blob fields are not observed (cfg at get_challenges.rs:69)
    assert len(flatten_public_values(pv, eth_mainnet=False)) == 2217 - 12
    # cdk_erigon: the burn address follows (get_challenges.rs:216-224), and must be set
    v, l = tagged_u256(0x61)
    pv.burn_addr = v
    assert [int(x) for x in flatten_public_values(pv, cdk_erigon=True)[-8:]] == l
    pv.burn_addr = None
    with pytest.raises(ValueError):
        flatten_public_values(pv, cdk_erigon=True)


def test_flatten_public_values_rejects_oversized_fields():
    # u256_to_u32 / u256_to_u64 fail with ProgramError::IntegerTooLarge (util.rs:40-58)
    for name, bad in (("block_timestamp", 1 << 32), ("block_number", 1 << 40), ("block_base_fee", 1 << 64), ("block_blob_gas_used", 1 << 64)):
        pv = PublicValues()
        setattr(pv.block_metadata, name, bad)
        with pytest.raises(IntegerTooLarge):
            flatten_public_values(pv)
    pv = PublicValues()
    pv.extra_block_data.gas_used_after = 1 << 32
    with pytest.raises(IntegerTooLarge):
        flatten_public_values(pv)


def test_h256_limb_order_is_observe_root_order():
    # observe_root (get_challenges.rs:11-19): into_uint().0 are the four u64 limbs, least significant first; each gives (low u32, high u32)
    h = bytes(range(1, 33))
    x = int.from_bytes(h, "big")
    u64 = [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    want = []
    for limb in u64:
        want += [limb & 0xFFFFFFFF, limb >> 32]
    pv = PublicValues()
    pv.trie_roots_before.state_root = h
    assert [int(v) for v in flatten_public_values(pv)[:8]] == want


def test_memory_extra_looking_values_field_by_field():
    """verifier.rs:547-737: (segment, index inside the segment) of every public-value write, the indices counted by hand from the
    GlobalMetadata declaration (cpu/kernel/constants/global_metadata.rs:11-115) and memory/segments.rs:10-91"""
    from zk_evm_b200.public_values import memory_extra_looking_values, memory_extra_looking_sum, RegistersData, GLOBAL_METADATA, P
    assert len(GLOBAL_METADATA) == len(set(GLOBAL_METADATA)) == 54
    pv = PublicValues()
    m, e = pv.block_metadata, pv.extra_block_data
    m.block_beneficiary = bytes(range(1, 21))
    m.block_timestamp, m.block_number, m.block_difficulty, m.block_gaslimit, m.block_chain_id = 101, 102, 103, 104, 105
    m.block_base_fee, m.block_gas_used, m.block_blob_gas_used, m.block_excess_blob_gas = 106, 107, 108, 109
    m.block_random, m.parent_beacon_block_root = bytes([7] * 32), bytes([8] * 32)
    m.block_bloom = [200 + i for i in range(8)]
    pv.block_hashes.prev_hashes = [i.to_bytes(32, "big") for i in range(1000, 1256)]
    pv.block_hashes.cur_hash = bytes([9] * 32)
    e.txn_number_before, e.txn_number_after, e.gas_used_before, e.gas_used_after = 110, 111, 112, 113
    tb, ta = pv.trie_roots_before, pv.trie_roots_after
    tb.state_root, tb.transactions_root, tb.receipts_root = bytes([1] * 32), bytes([2] * 32), bytes([3] * 32)
    ta.state_root, ta.transactions_root, ta.receipts_root = bytes([4] * 32), bytes([5] * 32), bytes([6] * 32)
    pv.registers_before = RegistersData(1, 2, 3, 4, 5, 6)
    pv.registers_after = RegistersData(11, 12, 13, 14, 15, 1 << 255)
    rows = memory_extra_looking_values(pv, bytes([10] * 32), 4242)
    val = lambda r: sum(v << (32 * i) for i, v in enumerate(r[4:12]))
    got = {(r[2], r[3]): val(r) for r in rows}
    assert len(got) == len(rows) == 301
    assert all(r[0] == 0 and r[1] == 0 and r[12] == 2 and all(0 <= v < 1 << 32 for v in r[4:12]) for r in rows)
    rep = lambda b: int.from_bytes(bytes([b] * 32), "big")
    want = {6: rep(1), 7: rep(2), 8: rep(3), 9: rep(4), 10: rep(5), 11: rep(6),                   # trie root digests before / after
            12: int.from_bytes(bytes(range(1, 21)), "big"), 13: 101, 14: 102, 15: 103, 16: rep(7), 17: 104, 18: 105, 19: 106,
            20: 108, 21: 109, 22: 107, 23: 112, 24: 113, 25: rep(9), 26: rep(8), 42: 110, 43: 111, 45: rep(10), 46: 4242}
    assert {k[1]: v for k, v in got.items() if k[0] == 5} == want                                   # Segment::GlobalMetadata = 5
    assert [got[(24, i)] for i in range(8)] == [200 + i for i in range(8)]                          # Segment::GlobalBlockBloom
    assert [got[(32, i)] for i in range(256)] == list(range(1000, 1256))                            # Segment::BlockHashes
    assert [got[(33, i)] for i in range(12)] == [1, 2, 3, 4, 5, 6, 11, 12, 13, 14, 15, 1 << 255]    # Segment::RegistersStates
    # without the eth_mainnet feature the three 4844 / 4788 fields are not written; cdk_erigon adds the burn address (index 53)
    assert len(memory_extra_looking_values(pv, bytes(32), 1, eth_mainnet=False)) == 298
    pv.burn_addr = 0xb0b
    r = memory_extra_looking_values(pv, bytes(32), 1, cdk_erigon=True)
    assert len(r) == 302 and r[1][:4] == [0, 0, 5, 53] and r[1][4] == 0xb0b
    # the sum is add_data_write's (verifier.rs:492-512): 1 / (gamma + sum_i row_i beta^i) per row
    beta, gamma = 0x1234567, 0x7654321
    s = 0
    for row in rows:
        s = (s + pow((sum(v * pow(beta, i, P) for i, v in enumerate(row)) + gamma) % P, P - 2, P)) % P
    pv.burn_addr = None
    assert memory_extra_looking_sum(pv, beta, gamma, bytes([10] * 32), 4242) == s
