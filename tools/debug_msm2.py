import sys, numpy as np
sys.path.insert(0, '.')
import bellman_b200 as bb
from oracle import o1
w = bb.Worker(0)
P = o1.g1_fixed_mul(o1.fr_from_ints([5]))
print("P x0", hex(int(P[0,0])))
for k in (1, 2, 3, 5):
    print("expected", k, hex(int(o1.g1_mul(P, o1.fr_from_ints([k]))[0, 0])))
w.set_option("msm_window_bits", 4)
for k in (2, 3, 0x35):
    print("== k", k)
    got = bb.multiexp(w, (bb.Bases(w, bb.G1, P), 0), bb.FullDensity, o1.fr_from_ints([k])).wait()
    print("got", hex(int(got[0, 0])))
