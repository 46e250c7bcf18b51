"""ctypes binding for the C++ oracle (oracle/_build/liboracle1.so).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from bellman_b200/.

Arrays are numpy uint64: Fr = (n,4) Montgomery limbs, G1 affine = (n,12),
G2 affine = (n,24); all-zero rows are the identity.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle1.so")


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.o1_mimc_new.restype = C.c_void_p
        _lib.o1_h_poly.restype = C.c_long
        _lib.o1_window_size.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u64(shape):
    return np.zeros(shape, dtype=np.uint64)


FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FP_MODULUS = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
                 "1eabfffeb153ffffb9feffffffffaaab", 16)


def ints_to_limbs(vals, nlimbs):
    out = np.zeros((len(vals), nlimbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, arr.shape[-1])
    return [sum(int(arr[i, j]) << (64 * j) for j in range(arr.shape[1])) for i in range(arr.shape[0])]


def set_threads(n):
    lib().o1_set_threads(C.c_int(n))


def num_threads():
    return lib().o1_num_threads()


# ---- fields -------------------------------------------------------------------
def fr_from_ints(vals):
    c = ints_to_limbs([v % FR_MODULUS for v in vals], 4)
    out = _u64((len(vals), 4))
    lib().o1_fr_from_canonical(_p(c), _p(out), C.c_size_t(len(vals)))
    return out


def fr_to_ints(m):
    m = np.ascontiguousarray(m, dtype=np.uint64).reshape(-1, 4)
    out = _u64(m.shape)
    lib().o1_fr_to_canonical(_p(m), _p(out), C.c_size_t(m.shape[0]))
    return limbs_to_ints(out)


def fr_to_canonical(m):
    m = np.ascontiguousarray(m, dtype=np.uint64).reshape(-1, 4)
    out = _u64(m.shape)
    lib().o1_fr_to_canonical(_p(m), _p(out), C.c_size_t(m.shape[0]))
    return out


def fp_from_ints(vals):
    c = ints_to_limbs([v % FP_MODULUS for v in vals], 6)
    out = _u64((len(vals), 6))
    lib().o1_fp_from_canonical(_p(c), _p(out), C.c_size_t(len(vals)))
    return out


def fp_to_ints(m):
    m = np.ascontiguousarray(m, dtype=np.uint64).reshape(-1, 6)
    out = _u64(m.shape)
    lib().o1_fp_to_canonical(_p(m), _p(out), C.c_size_t(m.shape[0]))
    return limbs_to_ints(out)


def _binop(name, a, b, w):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, w)
    out = _u64(a.shape)
    getattr(lib(), name)(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def fr_mul(a, b): return _binop("o1_fr_mul", a, b, 4)
def fr_add(a, b): return _binop("o1_fr_add", a, b, 4)
def fr_sub(a, b): return _binop("o1_fr_sub", a, b, 4)
def fp_mul(a, b): return _binop("o1_fp_mul", a, b, 6)
def fp_add(a, b): return _binop("o1_fp_add", a, b, 6)
def fp_sub(a, b): return _binop("o1_fp_sub", a, b, 6)


def fr_random(seed, n):
    out = _u64((n, 4))
    lib().o1_fr_random(C.c_uint64(seed), _p(out), C.c_size_t(n))
    return out


# ---- curves -------------------------------------------------------------------
def g1_generator():
    out = _u64((1, 12)); lib().o1_g1_generator(_p(out)); return out


def g2_generator():
    out = _u64((1, 24)); lib().o1_g2_generator(_p(out)); return out


def g1_on_curve(p):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 12)
    return bool(lib().o1_g1_on_curve(_p(p), C.c_size_t(p.shape[0])))


def g2_on_curve(p):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 24)
    return bool(lib().o1_g2_on_curve(_p(p), C.c_size_t(p.shape[0])))


def g1_add(a, b): return _binop("o1_g1_add", a, b, 12)
def g2_add(a, b): return _binop("o1_g2_add", a, b, 24)


def g1_mul(bases, k):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 12)
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
    out = _u64(bases.shape)
    lib().o1_g1_mul(_p(bases), _p(k), _p(out), C.c_size_t(bases.shape[0]))
    return out


def g2_mul(bases, k):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 24)
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
    out = _u64(bases.shape)
    lib().o1_g2_mul(_p(bases), _p(k), _p(out), C.c_size_t(bases.shape[0]))
    return out


def g1_fixed_mul(k):
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
    out = _u64((k.shape[0], 12))
    lib().o1_g1_fixed_mul(_p(k), _p(out), C.c_size_t(k.shape[0]))
    return out


def g2_fixed_mul(k):
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
    out = _u64((k.shape[0], 24))
    lib().o1_g2_fixed_mul(_p(k), _p(out), C.c_size_t(k.shape[0]))
    return out


def g1_compress(p):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((p.shape[0], 48), dtype=np.uint8)
    lib().o1_g1_compress(_p(p), _p(out), C.c_size_t(p.shape[0]))
    return out


def g2_compress(p):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 24)
    out = np.zeros((p.shape[0], 96), dtype=np.uint8)
    lib().o1_g2_compress(_p(p), _p(out), C.c_size_t(p.shape[0]))
    return out


def g1_from_affine_ints(pts):
    """pts: list of (x, y) ints or None -> (n,12) Montgomery array"""
    out = _u64((len(pts), 12))
    for i, pt in enumerate(pts):
        if pt is not None:
            out[i] = fp_from_ints([pt[0], pt[1]]).reshape(-1)
    return out


def g1_to_affine_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 12)
    res = []
    for row in arr:
        if not row.any():
            res.append(None)
        else:
            x, y = fp_to_ints(row.reshape(2, 6))
            res.append((x, y))
    return res


def g2_from_affine_ints(pts):
    out = _u64((len(pts), 24))
    for i, pt in enumerate(pts):
        if pt is not None:
            (x0, x1), (y0, y1) = pt
            out[i] = fp_from_ints([x0, x1, y0, y1]).reshape(-1)
    return out


def g2_to_affine_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 24)
    res = []
    for row in arr:
        if not row.any():
            res.append(None)
        else:
            x0, x1, y0, y1 = fp_to_ints(row.reshape(4, 6))
            res.append(((x0, x1), (y0, y1)))
    return res


# ---- domain / multiexp -----------------------------------------------------------
FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3


def fft(data, mode):
    d = np.array(data, dtype=np.uint64).reshape(-1, 4)
    n = d.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().o1_fft(_p(d), C.c_uint32(log_n), C.c_int(mode))
    if rc:
        raise RuntimeError(f"o1_fft err {rc}")
    return d


def serial_fft(data, inverse=False):
    d = np.array(data, dtype=np.uint64).reshape(-1, 4)
    n = d.shape[0]
    log_n = n.bit_length() - 1
    lib().o1_serial_fft(_p(d), C.c_uint32(log_n), C.c_int(1 if inverse else 0))
    return d


def h_poly(a, b, c):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4)
    n = a.shape[0]
    m = 1
    while m < n:
        m *= 2
    out = _u64((m, 4))
    rc = lib().o1_h_poly(_p(a), _p(b), _p(c), C.c_size_t(n), _p(out))
    if rc < 0:
        raise RuntimeError(f"o1_h_poly err {-rc}")
    return out[: rc - 1]


ERR_NAMES = {0: "OK", 1: "PolynomialDegreeTooLarge", 2: "UnexpectedIdentity", 3: "IoError(UnexpectedEof)",
             4: "UnconstrainedVariable", 5: "DensityMismatch"}


def window_size(n):
    return lib().o1_window_size(C.c_size_t(n))


def multiexp(group, bases, offset, density, scalars):
    """Returns (err, affine_result_array)."""
    w = 12 if group == 1 else 24
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, w)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    dens = None if density is None else np.ascontiguousarray(density, dtype=np.uint8)
    if dens is not None:
        assert dens.shape[0] == scalars.shape[0]
    out = _u64((1, w))
    fn = lib().o1_multiexp_g1 if group == 1 else lib().o1_multiexp_g2
    rc = fn(_p(bases), C.c_size_t(bases.shape[0]), C.c_size_t(offset), _p(dens), _p(scalars),
            C.c_size_t(scalars.shape[0]), _p(out))
    return rc, out


def naive_multiexp(group, bases, scalars):
    w = 12 if group == 1 else 24
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, w)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = _u64((1, w))
    fn = lib().o1_naive_multiexp_g1 if group == 1 else lib().o1_naive_multiexp_g2
    fn(_p(bases), _p(scalars), C.c_size_t(bases.shape[0]), _p(out))
    return out


# ---- MiMC cases ---------------------------------------------------------------------
class Mimc:
    """MiMC circuit instance (groth16/tests/common/mod.rs) with `rounds` rounds."""

    def __init__(self, rounds, seed):
        self.h = C.c_void_p(lib().o1_mimc_new(C.c_size_t(rounds), C.c_uint64(seed)))
        shape = _u64(7)
        lib().o1_mimc_shape(self.h, _p(shape))
        (self.num_inputs, self.num_aux, self.num_constraints, self.m,
         self.a_aux_total, self.b_in_total, self.b_aux_total) = (int(x) for x in shape)

    def __del__(self):
        try:
            lib().o1_mimc_free(self.h)
        except Exception:
            pass

    def witness(self):
        n, ni, na = self.num_constraints, self.num_inputs, self.num_aux
        w = dict(a=_u64((n, 4)), b=_u64((n, 4)), c=_u64((n, 4)), inputs=_u64((ni, 4)), aux=_u64((na, 4)),
                 a_aux_density=np.zeros(na, np.uint8), b_input_density=np.zeros(ni, np.uint8),
                 b_aux_density=np.zeros(na, np.uint8))
        lib().o1_mimc_witness(self.h, _p(w["a"]), _p(w["b"]), _p(w["c"]), _p(w["inputs"]), _p(w["aux"]),
                              _p(w["a_aux_density"]), _p(w["b_input_density"]), _p(w["b_aux_density"]))
        return w

    def set_toxic(self, toxic_ints):
        t = fr_from_ints(list(toxic_ints))
        lib().o1_mimc_set_toxic(self.h, _p(t))

    def crs_scalars(self):
        nv = self.num_inputs + self.num_aux
        hk, ak, bk, ek = _u64((self.m - 1, 4)), _u64((nv, 4)), _u64((nv, 4)), _u64((nv, 4))
        lib().o1_mimc_crs_scalars(self.h, _p(hk), _p(ak), _p(bk), _p(ek))
        return dict(h=hk, a=ak, b=bk, ext=ek)

    def generate(self):
        rc = lib().o1_mimc_generate(self.h)
        if rc:
            raise RuntimeError(f"generate err {rc}")

    def export_params(self):
        sizes = _u64(6)
        lib().o1_mimc_param_sizes(self.h, _p(sizes))
        nic, nh, nl, na, nb1, nb2 = (int(x) for x in sizes)
        p = dict(vk_g1=_u64((3, 12)), vk_g2=_u64((3, 24)), ic=_u64((nic, 12)), h=_u64((nh, 12)), l=_u64((nl, 12)),
                 a=_u64((na, 12)), b_g1=_u64((nb1, 12)), b_g2=_u64((nb2, 24)))
        lib().o1_mimc_export_params(self.h, _p(p["vk_g1"]), _p(p["vk_g2"]), _p(p["ic"]), _p(p["h"]), _p(p["l"]),
                                    _p(p["a"]), _p(p["b_g1"]), _p(p["b_g2"]))
        return p

    def import_params(self, p):
        q = {k: np.ascontiguousarray(v, dtype=np.uint64) for k, v in p.items()}
        lib().o1_mimc_import_params(self.h, _p(q["vk_g1"]), _p(q["vk_g2"]),
                                    _p(q["h"]), C.c_size_t(q["h"].shape[0]), _p(q["l"]), C.c_size_t(q["l"].shape[0]),
                                    _p(q["a"]), C.c_size_t(q["a"].shape[0]), _p(q["b_g1"]), C.c_size_t(q["b_g1"].shape[0]),
                                    _p(q["b_g2"]), C.c_size_t(q["b_g2"].shape[0]))

    def prove(self, r_int, s_int):
        rs = fr_from_ints([r_int, s_int])
        out = np.zeros(192, np.uint8)
        rc = lib().o1_mimc_prove(self.h, _p(rs[0:1]), _p(rs[1:2]), _p(out))
        if rc:
            raise RuntimeError(f"prove err {rc}")
        return bytes(out)

    def expected_proof(self, r_int, s_int):
        rs = fr_from_ints([r_int, s_int])
        out = np.zeros(192, np.uint8)
        lib().o1_mimc_expected_proof(self.h, _p(rs[0:1]), _p(rs[1:2]), _p(out))
        return bytes(out)


def dummy_xordemo():
    out = np.zeros(48, np.uint32)
    rc = lib().o1_dummy_xordemo(_p(out))
    return rc, [int(x) for x in out]


def dummy_zero_coeff(one_var):
    return lib().o1_dummy_zero_coeff(C.c_int(1 if one_var else 0))
