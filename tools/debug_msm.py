import sys, numpy as np
sys.path.insert(0, '.')
import bellman_b200 as bb
from oracle import o1
w = bb.Worker(0)
P = o1.g1_fixed_mul(o1.fr_from_ints([5]))
for c in (0, 2, 4, 8, 16):
    w.set_option("msm_window_bits", c)
    res = []
    for k in (1, 2, 3, 7, 8, 9, 15, 16, 17, 255, 256, 65535, 65536, 2**32 + 5, 2**200 + 12345, o1.FR_MODULUS - 1):
        got = bb.multiexp(w, (bb.Bases(w, bb.G1, P), 0), bb.FullDensity, o1.fr_from_ints([k])).wait()
        want = o1.g1_mul(P, o1.fr_from_ints([k]))
        res.append((k if k < 1 << 40 else hex(k)[:12], bool(np.array_equal(got, want))))
    print("c", c, res)
# two points
P2 = o1.g1_fixed_mul(o1.fr_from_ints([5, 11]))
for ks in ([2, 3], [2, 2], [16, 16], [17, 1], [3, 0]):
    w.set_option("msm_window_bits", 4)
    got = bb.multiexp(w, (bb.Bases(w, bb.G1, P2), 0), bb.FullDensity, o1.fr_from_ints(ks)).wait()
    want = o1.g1_fixed_mul(o1.fr_from_ints([5 * ks[0] + 11 * ks[1]]))
    print(ks, bool(np.array_equal(got, want)))
