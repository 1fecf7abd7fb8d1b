"""Extracts the hand-written corridor instance of the reference's formulation demo
(/root/reference/faster/other/gurobi_continuous.cpp: x0/xf :220-222, limits :207-212, A1/b1 :318-344,
A2/b2 :359-379, A3/b3 :381-401) into tests/golden/corridor_continuous.json.

The demo records INPUTS only (no expected outputs), so this fixture pins inputs; the expected
outputs stored next to it come from the restated model (oracle + HiGHS) and are labelled as such.
Run in the build container (needs /root/reference); the JSON travels, the reference does not.
"""
import json
import os
import re
import sys

SRC = "/root/reference/faster/other/gurobi_continuous.cpp"


def numbers(block):
    block = re.sub(r"/+", " ", block)
    return [float(t) for t in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", block)]


def grab(text, name, start=0):
    m = re.compile(r"\b%s\s*<<" % name).search(text, start)
    end = text.index(";", m.end())
    return numbers(text[m.end():end]), end


def main():
    text = open(SRC).read()
    polys = []
    for k, F in ((1, 12), (2, 10), (3, 10)):
        A, e = grab(text, "A%d" % k)
        b, _ = grab(text, "b%d" % k)
        assert len(A) == 3 * F and len(b) == F, (k, len(A), len(b))
        polys.append({"A": [A[3 * i:3 * i + 3] for i in range(F)], "b": b})
    x0 = numbers(re.search(r"x0\s*=\s*\{([^}]*)\}", text).group(1))
    xf = numbers(re.search(r"xf\s*=\s*\{([^}]*)\}", text).group(1))
    out = {
        "source": "faster/other/gurobi_continuous.cpp (inputs only)",
        "N": 10, "dt": 0.5, "lim": [5.0, 3.0, 5.0],   # vmax, amax, umax (:207-212)
        "x0": x0, "xf": xf, "polys": polys,
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "corridor_continuous.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    sys.exit(main())
