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
