"""CPU: the segment trace container (zk_evm_b200/trace_file.py) round-trips a segment and refuses malformed files."""
import numpy as np
import pytest
from tests import traces
from zk_evm_b200 import trace_file as tf


def test_roundtrip_and_layout(tmp_path):
    tr = traces.valid_segment(seed=4)
    pv = np.arange(7, 7 + 2180, dtype=np.uint64)
    labels = (0x1234, 0x77, 0x4000, 0x5000)
    p = str(tmp_path / "seg.zktr")
    tf.save(p, tr, pv, labels)
    for mm in (True, False):
        seg = tf.load(p, mmap=mm, check_canonical=True)
        assert seg.labels == labels and np.array_equal(seg.public_values, pv)
        assert seg.table_in_use == [t is not None for t in tr]
        for a, b in zip(seg.traces, tr):
            assert (a is None) == (b is None)
            if b is not None:
                assert a.shape == b.shape and np.array_equal(a, b) and a.flags["C_CONTIGUOUS"]
    raw = open(p, "rb").read()
    assert raw[:8] == b"ZKSEGTR1" and len(raw) % 64 == 0


def test_rejects_malformed(tmp_path):
    tr = traces.valid_segment(seed=5)
    pv = np.arange(10, dtype=np.uint64)
    p = str(tmp_path / "seg.zktr")
    with pytest.raises(tf.TraceFileError):
        tf.save(p, [None] + tr[1:], pv, (1, 2, 3, 4))                      # Arithmetic is mandatory
    bad = [None if t is None else t.copy() for t in tr]
    bad[6][0, 0] = np.uint64(tf.P)
    with pytest.raises(tf.TraceFileError):
        tf.save(p, bad, pv, (1, 2, 3, 4))                                  # non-canonical element
    with pytest.raises(tf.TraceFileError):
        tf.save(p, [t if i != 2 else t[:, :48] for i, t in enumerate(tr)], pv, (1, 2, 3, 4))   # 48 rows: not a power of two
    tf.save(p, tr, pv, (1, 2, 3, 4))
    raw = bytearray(open(p, "rb").read())
    open(p, "wb").write(bytes(raw[: len(raw) // 2]))
    with pytest.raises(tf.TraceFileError):
        tf.load(p)                                                         # truncated
    raw[0] ^= 1
    open(p, "wb").write(bytes(raw))
    with pytest.raises(tf.TraceFileError):
        tf.load(p)                                                         # bad magic
