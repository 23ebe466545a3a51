"""Writes the reference-held fixtures of the hot path into tests/golden/ (run in the build container, which has
/root/reference; the outputs are committed because /root/reference does not exist on the GPU box):

  qr_rank_deficient_c64.npz   the matrix literal of `test_rank_deficient`
                              (faer/src/linalg/qr/no_pivoting/factor.rs:540-4751; block size 20, Q R ~ A at 1e-10, 4752-4787)
  svd_zink.json               the bidiagonal of `test_zink` (faer/src/linalg/svd/mod.rs:985-1028; recursion_threshold 8,
                              the last singular value must not come out as exactly zero, 1051)
  svd_bidiag_<name>.npz       faer/test_data/svd/<name>.txt (diag / subdiag lists parsed as bidiag_svd.rs:1526-1561 does;
                              used by test_qr_algorithm 1562-1606 and test_divide_and_conquer 1607-...: U S V^H ~ B)
Only INPUTS exist in the reference for these tests; the checks are reconstruction identities with the tolerances cited.
"""
import json
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/faer"


def rank_deficient_matrix():
    src = open(f"{REF}/src/linalg/qr/no_pivoting/factor.rs").read()
    body = src[src.index("fn test_rank_deficient()"):]
    body = body[body.index("let A = mat!["):body.index("let (m, n) = A.shape();")]
    rows = re.findall(r"\[\s*((?:[^\[\]]+?))\s*\]", body[body.index("mat![") + 5:])
    out = []
    for r in rows:
        ents = [e.strip() for e in r.split(",") if e.strip()]
        row = []
        for e in ents:
            mm = re.fullmatch(r"(-?[0-9.e+-]+?)\s*([+-])\s*([0-9.e+-]+?)\s*\*\s*i", e)
            assert mm, e
            re_, sg, im_ = float(mm.group(1)), mm.group(2), float(mm.group(3))
            row.append(complex(re_, im_ if sg == "+" else -im_))
        out.append(row)
    A = np.array(out, dtype=np.complex128)
    return A


def zink():
    src = open(f"{REF}/src/linalg/svd/mod.rs").read()
    body = src[src.index("fn test_zink()"):]
    def grab(name):
        seg = body[body.index(f"let {name} = ["):]
        seg = seg[seg.index("[") + 1:seg.index("];")]
        return [float(x) for x in seg.replace("\n", " ").split(",") if x.strip()]
    return grab("diag"), grab("subdiag")


def parse_bidiag(path):
    diag, sub, cur = [], [], None
    for line in open(path).read().splitlines():
        if line.startswith("diag"):
            cur = diag; continue
        if line.startswith("subdiag"):
            cur = sub; continue
        line = line.strip().rstrip(",")
        if line:
            cur.append(float(line))
    assert len(diag) == len(sub)
    return np.array(diag), np.array(sub)


if __name__ == "__main__":
    A = rank_deficient_matrix()
    np.savez_compressed(os.path.join(HERE, "qr_rank_deficient_c64.npz"), A=A)
    print("qr_rank_deficient_c64", A.shape, "numerical rank", np.linalg.matrix_rank(A))
    d, s = zink()
    json.dump({"source": "faer/src/linalg/svd/mod.rs:985-1028", "diag": d, "subdiag": s, "recursion_threshold": 8},
              open(os.path.join(HERE, "svd_zink.json"), "w"), indent=1)
    print("svd_zink", len(d), len(s))
    for f in sorted(os.listdir(f"{REF}/test_data/svd")):
        d, s = parse_bidiag(f"{REF}/test_data/svd/{f}")
        name = f[:-4]
        np.savez_compressed(os.path.join(HERE, f"svd_bidiag_{name}.npz"), diag=d, subdiag=s)
        print(name, d.size)
