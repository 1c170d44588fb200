#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the hot path into tests/golden/reference_goldens.json.

The reference (pure Julia) cannot be executed in the build container, so the only bit-level fixtures are the
numbers printed in its doctests and the `≈` fixtures in its unit tests.  This script parses them
out of /root/reference (build container only) -- nothing at test time reads the reference tree.

  G1  src/lattice.jl:33-39      LatticeRule32([1,0x22], 2, 0x37), get_point(., 2) == [0.75, 0.5]
  G2  src/lattice.jl:61-84      LatticeRule32(K_3600_32, 16), get_point(., 123) == 16 printed values
  G3  src/lattice.jl:101-128    get_point(ShiftedLatticeRule(rule16), 0): 16 printed values (== the shift itself)
  G4  src/distribution.jl:57-81 x = rand(10) and transform(Normal(), x): 10 printed pairs (shortest round-trip repr)
  G5  test/distribution.jl:36-40   Normal quantiles at 0.1:0.1:0.9          (reference tolerance: isapprox, rtol sqrt(eps))
  G6  test/distribution.jl:61-65   TruncatedNormal(0,1,-2,2) at 0.1:0.1:0.9  (isapprox)
  G7  test/distribution.jl:82-86   Weibull(2, sqrt(2)) at 0.1:0.1:0.9         (isapprox)

usage: python tests/golden/make_golden.py [reference_root]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))


def lines(rel, lo, hi):
    with open(os.path.join(ref, rel)) as f:
        all_lines = f.read().split("\n")
    return all_lines[lo - 1:hi]


def floats_in(block):
    out = []
    for ln in block:
        s = ln.strip()
        if re.fullmatch(r"-?\d+\.\d+(e-?\d+)?", s):
            out.append(s)
    return out


g = {}

# G1
blk = lines("src/lattice.jl", 33, 39)
assert "0x00000001,0x00000022" in blk[0] and "0x00000037" in blk[0], blk[0]
g["G1"] = {"source": "src/lattice.jl:33-39", "z": [1, 0x22], "s": 2, "nmax": 0x37, "k": 2,
           "point": floats_in(blk)}
assert g["G1"]["point"] == ["0.75", "0.5"]

# G2
blk = lines("src/lattice.jl", 61, 84)
assert "K_3600_32.txt" in blk[0]
vals = floats_in(blk)
assert len(vals) == 16
g["G2"] = {"source": "src/lattice.jl:61-84", "genvec": "K_3600_32", "s": 16, "k": 123, "point": vals}

# G3
blk = lines("src/lattice.jl", 101, 128)
vals = floats_in(blk)
assert len(vals) == 16
g["G3"] = {"source": "src/lattice.jl:101-128", "genvec": "K_3600_32", "s": 16, "k": 0,
           "point_equals_shift": vals}

# G4
blk = lines("src/distribution.jl", 57, 81)
vals = floats_in(blk)
assert len(vals) == 20
g["G4"] = {"source": "src/distribution.jl:57-81", "dist": "Normal(0,1)", "x": vals[:10], "y": vals[10:]}


def y_vector(rel, lo, hi):
    """the `y = [ ... ]` literal between lines lo..hi of a reference test file, as the printed strings"""
    blk = "\n".join(lines(rel, lo, hi))
    m = re.search(r"y = \[(.*?)\]", blk, re.S)
    vals = [tok.strip() for tok in m.group(1).replace("\n", " ").split(",")]
    assert len(vals) == 9, vals
    return vals


# G5-G7: x = 0.1:0.1:0.9 (Julia range: i/10 exactly as the literal i*0.1 rounded, i.e. float("0.i"))
for key, (lo, hi, dist, kind, pars) in {
        "G5": (35, 39, "Normal()", 1, [0.0, 1.0, 0.0, 0.0]),
        "G6": (60, 64, "TruncatedNormal()", 2, [0.0, 1.0, -2.0, 2.0]),
        "G7": (81, 85, "Weibull()", 3, [2.0, 2.0 ** 0.5, 0.0, 0.0])}.items():
    g[key] = {"source": "test/distribution.jl:%d-%d" % (lo, hi), "dist": dist, "kind": kind, "pars": pars,
              "x": ["0.%d" % i for i in range(1, 10)], "y": y_vector("test/distribution.jl", lo, hi),
              "tolerance": "isapprox: rtol = sqrt(eps(Float64)) ~ 1.49e-8"}

with open(os.path.join(here, "reference_goldens.json"), "w") as f:
    json.dump(g, f, indent=1)
print("wrote reference_goldens.json with", sorted(g))
