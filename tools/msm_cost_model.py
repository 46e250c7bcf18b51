#!/usr/bin/env python
"""Field-multiplication count of one MSM job under the two MSM forms, from the same rules the
library applies (msm.cu: choose_window, choose_affine_rounds).  A prediction of WORK, not of time: it says nothing
about gather locality, occupancy or launch latency -- those are what bb_groth16_autotune measures on the device.

    python tools/msm_cost_model.py            # the 2^20 prove's jobs
"""
import math


def choose_window(n):
    return 4 if n < 32 else 8 if n < 1 << 13 else 15 if n < 1 << 19 else 16 if n < 1 << 23 else 20


def rounds(pairs, entries, nb, unified, rows_log=3):
    if unified:
        if pairs < 1 << 13:
            return 0
        fill = entries // nb
        lg = int(math.floor(math.log2(max(1, fill + fill // 2))))          # nearest power of two
        return max(0, min(8, lg - rows_log))
    if pairs < 1 << 15 or entries < 3 << 20:
        return 0
    avg = entries // nb
    return 3 if avg >= 24 else 2 if avg >= 10 else 1 if avg >= 5 else 0


def job(n, table_len, unified, fp2=False):
    """multiplications of the base field F (Fp or Fp2) for n scalars over a vector of table_len bases"""
    c = choose_window(table_len if unified else min(n, table_len + 1))
    W, D = 255 // c + 1, 1 << (c - 1)
    nb = D if unified else W * D
    entries = n * W
    R = rounds(n, entries, nb, unified)
    fill = entries / nb
    affine = fill * (1 - 2.0 ** -R)                     # additions of the halving rounds, per bucket
    xyzz = fill * 2.0 ** -R                             # what the XYZZ stage adds
    per_bucket = 6.2 * affine + 10 * xyzz + 2 * 14      # + summation by parts: two full XYZZ additions per bucket
    return dict(c=c, W=W, buckets=nb, fill=fill, R=R, muls=per_bucket * nb * (3 if fp2 else 1), reduce_share=28 / per_bucket)


if __name__ == "__main__":
    m = 1 << 20
    jobs = [("h", m - 1, m - 1, False), ("l", m - 1, m - 1, False), ("a_aux", m - 1, m, False), ("b_g1_aux", m // 2 - 1, m // 2, False),
            ("b_g2_aux", m // 2 - 1, m // 2, True)]
    tot = {False: 0.0, True: 0.0}
    print("job        form        c   W  buckets   fill  R   M Fp-mul   reduction share")
    for name, n, tl, fp2 in jobs:
        for unified in (False, True):
            r = job(n, tl, unified, fp2)
            tot[unified] += r["muls"]
            print(f"{name:10s} {'one set' if unified else 'per-window':10s} {r['c']:3d} {r['W']:3d} {r['buckets']:8d} {r['fill']:6.0f} {r['R']:2d} {r['muls'] / 1e6:10.1f}   {100 * r['reduce_share']:5.1f} %")
    print(f"sum of the five large jobs: per-window {tot[False] / 1e6:.0f} M, one bucket set {tot[True] / 1e6:.0f} M Fp-mul "
          f"({100 * (tot[True] / tot[False] - 1):+.1f} %); at the measured 31.2 G Fp-mul/s ceiling: {tot[False] / 31.2e6:.1f} -> {tot[True] / 31.2e6:.1f} ms")
