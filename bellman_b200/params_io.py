"""groth16::Parameters / VerifyingKey files (groth16/src/lib.rs:143-219, 258-398).

Format (all integers big-endian): VerifyingKey = alpha_g1, beta_g1 (96 B), beta_g2, gamma_g2 (192 B),
delta_g1 (96 B), delta_g2 (192 B), u32 |ic|, ic points (96 B each); Parameters = VerifyingKey, then
u32-length-prefixed h, l, a, b_g1 (G1) and b_g2 (G2).  Points use the uncompressed ZCash encoding
(`UncompressedEncoding::to_uncompressed`): x | y big-endian canonical coordinates, G2 as
x.c1 | x.c0 | y.c1 | y.c0; byte 0 carries the flags (bit 7 compressed = 0, bit 6 infinity).

`read_parameters` returns the arrays `bellman_b200.Parameters` uploads (Montgomery little-endian
limbs, identity = all-zero) and applies the reference's acceptance rules:

* encoding (both modes; what `from_uncompressed_unchecked` of the bls12_381 crate rejects): the
  compression and sort flags must be clear, coordinates must be canonical (< p), and an infinity
  flag requires every other bit of the encoding to be zero;
* the point at infinity is rejected in h, l, a, b_g1, b_g2 and ic in BOTH modes (lib.rs:199-207,
  303-315,337-345) and accepted for alpha, beta, gamma, delta (so that the prover can report the
  subversion CRS itself, prover.rs:320-324);
* curve equation and subgroup membership (`from_uncompressed`): always for the VerifyingKey
  (lib.rs:158-183 has no unchecked mode), for the five vectors when `checked=True`
  (lib.rs:294-298,328-332).  This is curve arithmetic, so it runs on the device
  (`bb_points_validate`): pass the `Worker` to use, or one is created on device 0 -- without a CUDA
  device the call fails, there is no CPU fallback.
"""
import ctypes as C
import struct

import numpy as np

from . import _check, load_library


def _coords_to_abi(raw, ncoord, g2):
    """raw: (n, ncoord*48) uint8 big-endian canonical coordinates -> (n, ncoord*6) uint64 Montgomery"""
    n = raw.shape[0]
    be = raw.reshape(n, ncoord, 48).copy()
    be[:, 0, 0] &= 0x1F                                   # strip the flag bits of byte 0
    le = np.ascontiguousarray(be[:, :, ::-1])             # little-endian bytes per coordinate
    if g2:                                                # file: x.c1 x.c0 y.c1 y.c0 -> ABI: x.c0 x.c1 y.c0 y.c1
        le = np.ascontiguousarray(le[:, [1, 0, 3, 2], :])
    limbs = le.reshape(n, ncoord * 48).view(np.uint64).reshape(n, ncoord * 6).copy()
    _check(load_library().bb_fp_convert(limbs.ctypes.data_as(C.c_void_p), C.c_size_t(n * ncoord), C.c_int(1)))
    return limbs


def _abi_to_coords(limbs, ncoord, g2):
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, ncoord * 6).copy()
    n = arr.shape[0]
    ident = ~arr.any(axis=1)
    _check(load_library().bb_fp_convert(arr.ctypes.data_as(C.c_void_p), C.c_size_t(n * ncoord), C.c_int(0)))
    le = arr.view(np.uint8).reshape(n, ncoord, 48)
    if g2:
        le = le[:, [1, 0, 3, 2], :]
    be = np.ascontiguousarray(le[:, :, ::-1]).reshape(n, ncoord * 48).copy()
    be[ident] = 0
    be[ident, 0] = 0x40
    return be


class InvalidData(ValueError):
    """io::ErrorKind::InvalidData of Parameters::read / VerifyingKey::read"""


def _read_points(buf, off, count, g2, what, allow_infinity, validate):
    """validate: None, or a callable (group_is_g2, points, what) applying the curve checks"""
    size = 192 if g2 else 96
    name = "G2" if g2 else "G1"
    end = off + count * size
    if end > len(buf):
        raise EOFError(f"{what}: file ends inside the vector")
    raw = np.frombuffer(buf, dtype=np.uint8, count=count * size, offset=off).reshape(count, size)
    inf = np.zeros(count, bool)
    if count:
        flags = raw[:, 0]
        if (flags & 0x80).any():
            raise InvalidData(f"invalid {name}: {what} holds a compressed encoding")
        if (flags & 0x20).any():
            raise InvalidData(f"invalid {name}: {what} has the sort flag set on an uncompressed point")
        inf = (flags & 0x40) != 0
        if inf.any():
            rest = raw[inf].copy()
            rest[:, 0] &= 0x3F & ~0x40
            if rest.any():
                raise InvalidData(f"invalid {name}: {what} has an infinity flag with non-zero coordinates")
            if not allow_infinity:
                raise InvalidData(f"point at infinity in {what}")                 # lib.rs:199-207,303-315
    try:
        out = _coords_to_abi(raw, 4 if g2 else 2, g2)                              # rejects coordinates >= p
    except Exception as e:
        raise InvalidData(f"invalid {name}: {what}: {e}") from None
    if count:
        out[inf] = 0
    if validate is not None and count:
        validate(g2, out, what)
    return out, end


def _read_u32(buf, off):
    if off + 4 > len(buf):
        raise EOFError("file ends inside a length prefix")
    return struct.unpack_from(">I", buf, off)[0], off + 4


class _Validator:
    """curve + subgroup checks on the device, with a Worker made on demand"""

    def __init__(self, worker):
        self.worker, self.own = worker, None

    def __call__(self, g2, points, what):
        if self.worker is None:
            from . import Worker
            self.own = self.worker = Worker(0)
        first, why = C.c_size_t(), C.c_int()
        pts = np.ascontiguousarray(points, dtype=np.uint64)
        _check(load_library().bb_points_validate(self.worker._h, C.c_int(2 if g2 else 1), pts.ctypes.data_as(C.c_void_p),
                                                 C.c_size_t(pts.shape[0]), C.c_int(1), C.byref(first), C.byref(why)))
        if why.value:
            reason = "not on the curve" if why.value == 1 else "not in the prime-order subgroup"
            raise InvalidData(f"invalid {'G2' if g2 else 'G1'}: {what}[{first.value}] is {reason}")

    def close(self):
        if self.own is not None:
            self.own.close()


def read_verifying_key(buf, off=0, worker=None, _validator=None):
    """VerifyingKey::read (lib.rs:158-218): always curve- and subgroup-checked."""
    val = _validator or _Validator(worker)
    try:
        vk = {}
        g1 = lambda o, w: _read_points(buf, o, 1, False, w, True, val)
        g2 = lambda o, w: _read_points(buf, o, 1, True, w, True, val)
        vk["alpha_g1"], off = g1(off, "alpha_g1")
        vk["beta_g1"], off = g1(off, "beta_g1")
        vk["beta_g2"], off = g2(off, "beta_g2")
        vk["gamma_g2"], off = g2(off, "gamma_g2")
        vk["delta_g1"], off = g1(off, "delta_g1")
        vk["delta_g2"], off = g2(off, "delta_g2")
        n, off = _read_u32(buf, off)
        vk["ic"], off = _read_points(buf, off, n, False, "ic", False, val)
    finally:
        if _validator is None:
            val.close()
    return vk, off


def read_parameters(data, checked=True, worker=None):
    """bytes of Parameters::write -> dict for bellman_b200.Parameters (+ 'gamma_g2', 'ic');
    Parameters::read(reader, checked), lib.rs:289-398."""
    buf = memoryview(data)
    val = _Validator(worker)
    try:
        vk, off = read_verifying_key(buf, 0, _validator=val)
        p = dict(vk_g1=np.concatenate([vk["alpha_g1"], vk["beta_g1"], vk["delta_g1"]]),
                 vk_g2=np.concatenate([vk["beta_g2"], vk["gamma_g2"], vk["delta_g2"]]), ic=vk["ic"])
        for name, is_g2 in (("h", False), ("l", False), ("a", False), ("b_g1", False), ("b_g2", True)):
            n, off = _read_u32(buf, off)
            p[name], off = _read_points(buf, off, n, is_g2, name, False, val if checked else None)
    finally:
        val.close()
    return p


def write_parameters(p):
    """inverse of read_parameters (Parameters::write, lib.rs:258-287)"""
    vk1 = np.asarray(p["vk_g1"], dtype=np.uint64).reshape(3, 12)
    vk2 = np.asarray(p["vk_g2"], dtype=np.uint64).reshape(3, 24)
    out = bytearray()
    out += _abi_to_coords(vk1[0:1], 2, False).tobytes()       # alpha_g1
    out += _abi_to_coords(vk1[1:2], 2, False).tobytes()       # beta_g1
    out += _abi_to_coords(vk2[0:1], 4, True).tobytes()        # beta_g2
    out += _abi_to_coords(vk2[1:2], 4, True).tobytes()        # gamma_g2
    out += _abi_to_coords(vk1[2:3], 2, False).tobytes()       # delta_g1
    out += _abi_to_coords(vk2[2:3], 4, True).tobytes()        # delta_g2
    for name, g2 in (("ic", False), ("h", False), ("l", False), ("a", False), ("b_g1", False), ("b_g2", True)):
        arr = np.asarray(p[name], dtype=np.uint64).reshape(-1, 24 if g2 else 12)
        out += struct.pack(">I", arr.shape[0])
        out += _abi_to_coords(arr, 4 if g2 else 2, g2).tobytes()
    return bytes(out)
