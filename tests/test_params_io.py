"""Parameters / VerifyingKey (de)serialisation (groth16/src/lib.rs:143-219,258-398): host-side
logic of the product, checked against the oracle's independent ZCash encoders and against the
byte count the reference pins (lib.rs:529: 2136 bytes for its 1-constraint circuit)."""
import random
import struct

import numpy as np
import pytest

from bellman_b200 import params_io
from oracle import o1
from oracle.oracle0 import fields as F

R = o1.FR_MODULUS


def _params(sizes, seed=1):
    rng = random.Random(seed)
    mk1 = lambda n: o1.g1_fixed_mul(o1.fr_from_ints([rng.randrange(1, R) for _ in range(n)]))
    mk2 = lambda n: o1.g2_fixed_mul(o1.fr_from_ints([rng.randrange(1, R) for _ in range(n)]))
    return dict(vk_g1=mk1(3), vk_g2=mk2(3), ic=mk1(sizes[0]), h=mk1(sizes[1]), l=mk1(sizes[2]), a=mk1(sizes[3]),
                b_g1=mk1(sizes[4]), b_g2=mk2(sizes[5]))


def test_reference_size_and_round_trip():
    # the reference's 1-constraint circuit: ic 2, h 3, l 2, a 3, b_g1 1, b_g2 1  -> 2136 bytes
    p = _params((2, 3, 2, 3, 1, 1))
    data = params_io.write_parameters(p)
    assert len(data) == 2136                                              # groth16/src/lib.rs:529
    q = params_io.read_parameters(data)
    for k in p:
        assert np.array_equal(np.asarray(p[k]).reshape(-1), q[k].reshape(-1)), k
    # encodings agree with the oracle's independent uncompressed encoder
    alpha = o1.g1_to_affine_ints(p["vk_g1"][0:1])[0]
    assert data[:96] == F.g1_uncompressed(alpha)
    beta2 = o1.g2_to_affine_ints(p["vk_g2"][0:1])[0]
    assert data[192:384] == F.g2_uncompressed(beta2)
    assert struct.unpack_from(">I", data, 864)[0] == 2                    # |ic| after the 864-byte key head


def test_rejections():
    p = _params((2, 3, 2, 3, 1, 1), seed=2)
    data = bytearray(params_io.write_parameters(p))
    with pytest.raises(EOFError):
        params_io.read_parameters(bytes(data[:-5]))
    bad = bytearray(data); bad[0] |= 0x80                                  # compressed flag in an uncompressed file
    with pytest.raises(ValueError):
        params_io.read_parameters(bytes(bad))
    bad = bytearray(data); bad[1:48] = b"\xff" * 47; bad[0] = 0x1f         # x >= p
    with pytest.raises(Exception):
        params_io.read_parameters(bytes(bad))
    # a point at infinity inside a vector: rejected when checked, kept as the identity otherwise
    p["a"][1] = 0
    data = params_io.write_parameters(p)
    with pytest.raises(ValueError):
        params_io.read_parameters(data, checked=True)
    q = params_io.read_parameters(data, checked=False)
    assert not q["a"][1].any() and np.array_equal(q["a"][0], p["a"][0])
