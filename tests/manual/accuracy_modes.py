"""Measured parity per arithmetic mode (not a test): relative Frobenius error of H and Y and relative error of the
objective against the reference's float64 goldens.  Usage: python tests/manual/accuracy_modes.py [modes...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_inputs, load_golden, rel_fro  # noqa: E402
from exemplars_vc_b200 import ExemplarDictionary  # noqa: E402

CASES = ["kl_13x32x8_tol1e-4", "kl_201x777x37_tol1e-4", "kl_513x2000x64_tol1e-4", "speech_sf1_tf1_100162"]


def main():
    modes = sys.argv[1:] or ["fp32", "3xtf32", "tf32", "bf16"]
    for name in CASES:
        g = load_golden(name)
        X, A, B = (g["X"], g["A"], g["B"]) if name.startswith("speech") else golden_inputs(g)
        for mode in modes:
            with ExemplarDictionary(A, B, mode=mode) as d:
                act = d.solve(X, tol=float(g["tol"]), max_iter=int(g["max_iter"]))
                H = d.to_host(act.H)
                Y = d.to_host(d.convert(act.H))
            obj = float(g["objective"])
            print(f"{name:28s} {mode:7s} n_iter {act.n_iter:3d} (ref {int(g['n_iter']):3d})  relF(H) {rel_fro(H, g['W']):.2e}  "
                  f"relF(Y) {rel_fro(Y, g['Y']):.2e}  rel(obj) {abs(act.objective - obj) / obj:.2e}", flush=True)


if __name__ == "__main__":
    main()
