"""oracle/philox.py against the published Philox4x32-10 known-answer vectors (Random123 kat_vectors) and as N(0, 1)."""
import numpy as np

from oracle import philox

KAT = [  # (counter, key, expected) -- Random123 examples/kat_vectors, philox4x32 10 rounds
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want


def test_normal_field_is_standard_normal_and_addressable():
    z = philox.normal_field(1234, 20000, 19)
    assert z.dtype == np.float32 and z.shape == (20000, 19) and np.isfinite(z).all()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3.0) < 0.1
    # a pure function of (seed, row, j): a prefix of the rows is the same field, another seed is another field
    assert np.array_equal(philox.normal_field(1234, 100, 19), z[:100])
    assert not np.array_equal(philox.normal_field(1235, 100, 19), z[:100])
    assert np.array_equal(philox.normal_field(1234, 100, 19, sigma=0.5), np.float32(0.5) * z[:100])
    c = np.corrcoef(z[:, 0], z[:, 1])[0, 1]
    assert abs(c) < 0.03
