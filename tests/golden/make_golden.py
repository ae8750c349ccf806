"""Regenerate the golden fixtures from the reference's own test data.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

issue143_matrix.npy : the literal 71x71 symmetric matrix of KrylovKit issue #143
                      (/root/reference/test/issues.jl:39-116), a known-answer fixture of the
                      reference test-suite (its spectrum must be reproduced, issues.jl:118-121).
Only DATA is extracted (numeric literals); no reference code is copied.
"""
import os
import re

import numpy as np

REF = "/root/reference/test/issues.jl"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    txt = open(REF).read()
    start = txt.index('@testset "Issue #143"')
    a0 = txt.index("A = [", start) + len("A = [")
    a1 = txt.index("]", a0)
    rows = [ln.strip() for ln in txt[a0:a1].strip().splitlines() if ln.strip()]
    A = np.array([[float(t) for t in re.split(r"\s+", ln.rstrip(";"))] for ln in rows])
    assert A.shape == (71, 71), A.shape
    assert np.allclose(A, A.T, rtol=1e-6, atol=1e-3)
    np.save(os.path.join(HERE, "issue143_matrix.npy"), A)
    print("issue143_matrix.npy", A.shape, "asym:", np.abs(A - A.T).max())


if __name__ == "__main__":
    main()
