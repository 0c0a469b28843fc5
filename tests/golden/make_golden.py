"""Generate the committed golden vectors from the oracle (LAPACK replay of the reference's call
sequence).  The reference itself is Julia and cannot be imported here (no julia in the image),
so these vectors pin the ORACLE's outputs on the reference's own test sizes; the oracle in turn
is pinned to the reference's fixtures in tests/test_oracle.py.
Run:  python tests/golden/make_golden.py   (writes tests/golden/*.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mak_oracle as O  # noqa: E402


def main():
    for dtype in ("f64", "c128"):
        for (m, n) in ((54, 37), (54, 54), (54, 63)):  # test/decompositions/qr.jl:21-22
            A = O.randn_matrix(m, n, dtype, seed=123)
            Q, R = O.qr_compact(A)
            Qf, Rf = O.qr_full(A)
            U, S, Vh = O.svd_compact(A)
            out = dict(A=A, Q=Q, R=R, Qf=Qf, Rf=Rf, U=U, S=S, Vh=Vh)
            if m >= n:
                W, P = O.left_polar(A)
                out.update(W=W, P=P)
            if m == n:
                H = O.rand_hermitian(n, dtype, seed=123)
                w, V = O.eigh_full(H)
                out.update(H=H, w=w, V=V)
            np.savez_compressed(os.path.join(HERE, f"golden_{dtype}_{m}x{n}.npz"), **out)
    # doctest KAT and fixed-spectrum fixtures of the reference
    K = np.array([[2.0, 1, 0], [1, 3, 1], [0, 1, 4]])
    np.savez_compressed(os.path.join(HERE, "kat_eigh3.npz"), A=K, w=np.array([3 - np.sqrt(3), 3, 3 + np.sqrt(3)]))


if __name__ == "__main__":
    main()
