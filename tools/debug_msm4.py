import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import bellman_b200 as bb
from oracle import o1
w = bb.Worker(0)
lib = bb.load_library()
pts = o1.g1_fixed_mul(o1.fr_from_ints([5]))
out = np.zeros((1, 12), np.uint64)
rc = lib.bb_selftest_bucket_reduce(w._h, pts.ctypes.data_as(C.c_void_p), C.c_uint32(1), C.c_uint32(1), out.ctypes.data_as(C.c_void_p))
print(rc, hex(int(pts[0,0])), hex(int(out[0,0])), bool(np.array_equal(pts, out)))
pts = o1.g1_fixed_mul(o1.fr_from_ints([5, 7]))
rc = lib.bb_selftest_bucket_reduce(w._h, pts.ctypes.data_as(C.c_void_p), C.c_uint32(2), C.c_uint32(2), out.ctypes.data_as(C.c_void_p))
print(rc, bool(np.array_equal(out, o1.g1_fixed_mul(o1.fr_from_ints([19])))))
