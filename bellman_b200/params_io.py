"""groth16::Parameters / VerifyingKey files (groth16/src/lib.rs:143-219, 258-398).

Format (all integers big-endian): VerifyingKey = alpha_g1, beta_g1 (96 B), beta_g2, gamma_g2 (192 B),
delta_g1 (96 B), delta_g2 (192 B), u32 |ic|, ic points (96 B each); Parameters = VerifyingKey, then
u32-length-prefixed h, l, a, b_g1 (G1) and b_g2 (G2).  Points use the uncompressed ZCash encoding
(`UncompressedEncoding::to_uncompressed`): x | y big-endian canonical coordinates, G2 as
x.c1 | x.c0 | y.c1 | y.c0; byte 0 carries the flags (bit 7 compressed = 0, bit 6 infinity).

`read_parameters` returns the arrays `bellman_b200.Parameters` uploads (Montgomery little-endian
limbs, identity = all-zero).  `checked=True` applies Parameters::read's rules that need no curve
arithmetic here: flags, canonical coordinates and "no point at infinity" for the vectors and the
vk (lib.rs:294-330); on-curve and subgroup membership are the loader's caller's business exactly
when the reference is called with checked=false.
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


def _read_points(buf, off, count, g2, checked, what):
    size = 192 if g2 else 96
    end = off + count * size
    if end > len(buf):
        raise EOFError(f"{what}: file ends inside the vector")
    raw = np.frombuffer(buf, dtype=np.uint8, count=count * size, offset=off).reshape(count, size)
    if count:
        flags = raw[:, 0]
        if (flags & 0x80).any():
            raise ValueError(f"{what}: compressed point in an uncompressed vector")
        inf = (flags & 0x40) != 0
        if checked and inf.any():
            raise ValueError(f"{what}: point at infinity")            # lib.rs:307-315
        if (flags & 0x20).any():
            raise ValueError(f"{what}: sort flag set on an uncompressed point")
    out = _coords_to_abi(raw, 4 if g2 else 2, g2)
    if count:
        out[inf] = 0
    return out, end


def _read_u32(buf, off):
    if off + 4 > len(buf):
        raise EOFError("file ends inside a length prefix")
    return struct.unpack_from(">I", buf, off)[0], off + 4


def read_verifying_key(buf, off=0, checked=True):
    vk = {}
    g1 = lambda o, w: _read_points(buf, o, 1, False, checked, w)
    g2 = lambda o, w: _read_points(buf, o, 1, True, checked, w)
    vk["alpha_g1"], off = g1(off, "alpha_g1")
    vk["beta_g1"], off = g1(off, "beta_g1")
    vk["beta_g2"], off = g2(off, "beta_g2")
    vk["gamma_g2"], off = g2(off, "gamma_g2")
    vk["delta_g1"], off = g1(off, "delta_g1")
    vk["delta_g2"], off = g2(off, "delta_g2")
    n, off = _read_u32(buf, off)
    vk["ic"], off = _read_points(buf, off, n, False, checked, "ic")
    return vk, off


def read_parameters(data, checked=True):
    """bytes of Parameters::write -> dict for bellman_b200.Parameters (+ 'gamma_g2', 'ic')."""
    buf = memoryview(data)
    vk, off = read_verifying_key(buf, 0, checked)
    p = dict(vk_g1=np.concatenate([vk["alpha_g1"], vk["beta_g1"], vk["delta_g1"]]),
             vk_g2=np.concatenate([vk["beta_g2"], vk["gamma_g2"], vk["delta_g2"]]), ic=vk["ic"])
    for name, is_g2 in (("h", False), ("l", False), ("a", False), ("b_g1", False), ("b_g2", True)):
        n, off = _read_u32(buf, off)
        p[name], off = _read_points(buf, off, n, is_g2, checked, name)
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
