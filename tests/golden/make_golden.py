"""Regenerates tests/golden/reference_tests.json.

The reference's own tests (Tests/multiplication.py, assignment.py, filter.py,
max.py) are print-and-compare scripts whose oracle is NumPy on
`numpy.random.seed(42)` + `randint(1, 100, size=(100,))` int64 vectors.  The
reference extension itself cannot be imported here (CPython 2 + LLVM 3.9), so
the fixtures are the NumPy side of each script — the value the script prints
next to the thunk result and compares with `str(...) == str(...)`.  NumPy's
legacy RandomState stream is frozen, so these are the numbers a py2.7 run
produced.  Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy


def draw(n_vectors):
    numpy.random.seed(42)
    return [numpy.random.randint(1, 100, size=(100,)).astype(numpy.int64) for _ in range(n_vectors)]


def main():
    out = {}

    # Tests/multiplication.py:6-38
    a, b, c, d, e, f = draw(6)
    res = d * e * f
    res2 = numpy.sort(res) * (a * b * c)
    out["multiplication"] = {
        "inputs": {k: v.tolist() for k, v in zip("abcdef", (a, b, c, d, e, f))},
        "res": res.tolist(), "res2": res2.tolist(),
        "unsorted_product": (res * (a * b * c)).tolist(),
    }

    # Tests/assignment.py:6-20 (c uses the ORIGINAL a; a_nump itself is untouched)
    a, b = draw(2)
    out["assignment"] = {"inputs": {"a": a.tolist(), "b": b.tolist()},
                         "assign": {"index": 0, "value": 1000},
                         "c": (a * b).tolist()}

    # Tests/filter.py:7-19
    (a,) = draw(1)
    res = a[a * 2 >= 50]
    chunks = [(12 * j, 100 if j == 7 else 12 * (j + 1)) for j in range(8)]   # scheduler.cpp:133-147
    out["filter"] = {"inputs": {"a": a.tolist()}, "res": res.tolist(),
                     "chunk_counts": [int((a[s:e] * 2 >= 50).sum()) for s, e in chunks]}

    # Tests/max.py:7-22 (despite its name it exercises sum)
    (a,) = draw(1)
    s = a[a >= 50].sum()
    out["max"] = {"inputs": {"a": a.tolist()}, "res": int(s), "res2": (s * a).tolist(),
                  "filtered": a[a >= 50].tolist(),
                  "chunk_partials": [int(a[st:en][a[st:en] >= 50].sum()) for st, en in chunks]}

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tests.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("wrote", path)


if __name__ == "__main__":
    main()
