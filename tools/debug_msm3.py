import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import bellman_b200 as bb
from oracle import o1
w = bb.Worker(0)
lib = bb.load_library()
R = o1.FR_MODULUS
def red(ks, K):
    pts = o1.g1_fixed_mul(o1.fr_from_ints(ks))
    pts[[i for i, k in enumerate(ks) if k == 0]] = 0
    out = np.zeros((1, 12), np.uint64)
    rc = lib.bb_selftest_bucket_reduce(w._h, pts.ctypes.data_as(C.c_void_p), C.c_uint32(len(ks)), C.c_uint32(K), out.ctypes.data_as(C.c_void_p))
    want = o1.g1_fixed_mul(o1.fr_from_ints([sum((i + 1) * k for i, k in enumerate(ks)) % R]))
    return rc, bool(np.array_equal(out, want))
print("single bucket0", red([5], 1))
print("bucket1 only", red([0, 5], 1))
print("bucket1 only K2", red([0, 5], 2))
print("bucket2 only", red([0, 0, 5, 0], 1))
print("two", red([3, 5], 1), red([3, 5], 2))
print("8 K1", red([3, 5, 7, 11, 13, 17, 19, 23], 1), "K2", red([3, 5, 7, 11, 13, 17, 19, 23], 2), "K8", red([3, 5, 7, 11, 13, 17, 19, 23], 8))
import random
rng = random.Random(1)
ks = [rng.randrange(R) for _ in range(4096)]
print("4096 K16", red(ks, 16), "K1", red(ks, 1))
# dbl through selftest op1 on the same point
P = o1.g1_fixed_mul(o1.fr_from_ints([5]))
out = np.zeros_like(P)
lib.bb_selftest_point(w._h, 1, 1, P.ctypes.data_as(C.c_void_p), np.zeros_like(P).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(1))
print("selftest 2P", bool(np.array_equal(out, o1.g1_fixed_mul(o1.fr_from_ints([10])))))
