"""CPU tests of the drop-in boundary: libmle_cuda.so loads without a GPU, exports every symbol that
include/mle_cuda.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mle_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mle_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    import mle_b200
    lib = mle_b200._lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libmle_cuda.so does not export %s" % n
        assert n in mle_b200._lib.SIGNATURES, "no ctypes prototype for %s" % n
    assert lib.mle_abi_version() == 1


def _split_top(text):
    """split on commas that are not nested in (), [] or {}"""
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(text, start):
    """text[start] == '(' -> index one past its matching ')'"""
    depth = 0
    for i in range(start, len(text)):
        depth += text[i] == "("
        depth -= text[i] == ")"
        if depth == 0:
            return i + 1
    raise ValueError("unbalanced")


def test_julia_binding_matches_the_header():
    """julia/MLECuda.jl cannot run here (no Julia), so check it statically: every ccall names a symbol the header declares, its
    type tuple has one entry per C parameter with the right width, as many values as types follow, and the entry points of
    the hot path are all bound"""
    import re
    hdr = open(os.path.join(ROOT, "include", "mle_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(mle_[a-z0-9_]+)\(([^;]*?)\);", hdr, flags=re.S):
        params = [] if m.group(2).strip() == "void" else _split_top(" ".join(m.group(2).split()))
        protos[m.group(1)] = params
    src = open(os.path.join(ROOT, "julia", "MLECuda.jl")).read()
    src = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in src.splitlines())
    width = {"Cint": ("int", "int32_t"), "UInt64": ("uint64_t",), "Csize_t": ("size_t",), "Cstring": ("const char *",)}
    seen = set()
    for m in re.finditer(r"ccall\(\(:(mle_[a-z0-9_]+), libmle\), (\w+),\s*", src):
        name = m.group(1)
        assert name in protos, "%s is not declared in include/mle_cuda.h" % name
        seen.add(name)
        call_end = _balanced(src, m.start() + len("ccall"))
        rest = src[m.end():call_end - 1]
        tup_end = _balanced(rest, rest.index("("))
        types = _split_top(rest[rest.index("(") + 1:tup_end - 1])
        values = _split_top(rest[tup_end:].lstrip().lstrip(","))
        want = protos[name]
        assert len(types) == len(want), (name, types, want)
        assert len(values) == len(want), (name, values, want)
        assert m.group(2) == ("Cstring" if name == "mle_last_error" else "Cint")
        for t, c in zip(types, want):
            if t in width:   # scalars must have the declared width; pointers are checked by count only
                assert "*" not in c.replace("const char *", "") and any(c.startswith(w) for w in width[t]), (name, t, c)
            else:
                assert t.startswith(("Ptr{", "Ref{")) and ("*" in c or c.split()[0].startswith("mle_")), (name, t, c)
    for name in ("mle_init", "mle_init_multi", "mle_destroy", "mle_last_error", "mle_lattice_create", "mle_dists_create", "mle_model_create",
                 "mle_model_set_basis", "mle_points", "mle_sample_batch", "mle_update_samples", "mle_sample_unbiased", "mle_moments_merge",
                 "mle_lattice_destroy", "mle_dists_destroy", "mle_model_destroy", "mle_abi_version", "mle_genvec"):
        assert name in seen, "julia/MLECuda.jl does not bind %s" % name
    assert "const ABI_VERSION = 1" in src and "#define MLE_ABI_VERSION 1" in open(os.path.join(ROOT, "include", "mle_cuda.h")).read()
    # struct mle_dist {int32 kind; int32 reserved; double p[4]} <-> MLEDist
    assert re.search(r"struct MLEDist\s+kind::Int32\s+reserved::Int32\s+p::NTuple\{4, Float64\}\s+end", src)


def test_no_torch_types_in_the_boundary():
    hdr = open(os.path.join(ROOT, "include", "mle_cuda.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and "std::" not in hdr


def test_embedded_generating_vectors_match_the_oracle(oracle):
    import mle_b200
    lib = mle_b200._lib.load()
    for name, n in (("K_3600_32", 3600), ("CKN_250_20", 250)):
        p, ln = C.POINTER(C.c_uint32)(), C.c_int()
        assert lib.mle_genvec(name.encode(), C.byref(p), C.byref(ln)) == 0 and ln.value == n
        z = np.ctypeslib.as_array(p, shape=(n,))
        assert np.array_equal(z, oracle.genvec(name))
        assert z[0] == 1
    z = oracle.genvec("K_3600_32")
    assert z.max() == 521301 and oracle.genvec("CKN_250_20").max() < 2**20


def test_moments_merge_is_chan(oracle):
    import mle_b200
    rng = np.random.default_rng(0)
    x = rng.normal(3.0, 2.0, size=1000)
    parts = np.split(x, [10, 11, 500])
    acc = np.zeros((1, 3))
    for p in parts:
        acc = mle_b200.moments_merge(acc, np.array([[len(p), p.mean(), ((p - p.mean())**2).sum()]]))
    assert acc[0, 0] == 1000
    assert abs(acc[0, 1] - x.mean()) < 1e-14 * abs(x.mean())
    assert abs(acc[0, 2] - ((x - x.mean())**2).sum()) < 1e-12 * acc[0, 2]


def test_fails_loudly_without_a_gpu():
    import torch
    import mle_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mle_b200.MLEError) as e:
        mle_b200.Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_uses_the_oracle():
    """the product path must not import, include, link or dlopen anything under oracle/ (comments may cite it)"""
    pkg = os.path.join(ROOT, "multilevelestimators.jl_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b|#include\s*[\"<][^\">]*oracle|libmle_oracle|\borc_[a-z])", re.M)
    seen = 0
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                seen += 1
                src = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(src), fn
    assert seen >= 5


_NULL_CALLS = r'''
import json
ROOT = %r
import sys, ctypes as C
sys.path.insert(0, ROOT)
import mle_b200
lib = mle_b200._lib.load()
N = None
z = (C.c_uint32 * 4)(1, 2, 3, 4)
h = C.c_void_p()
dbl = (C.c_double * 16)()
i32 = (C.c_int32 * 2)(0, 0)
res = {}
res["lattice_create"] = lib.mle_lattice_create(N, z, 4, 100, C.byref(h))
res["lattice_destroy"] = lib.mle_lattice_destroy(N, N)
res["dists_create"] = lib.mle_dists_create(N, N, 1, C.byref(h))
res["dists_destroy"] = lib.mle_dists_destroy(N, N)
res["model_create"] = lib.mle_model_create(N, b"expsum", 1, dbl, 2, C.byref(h))
res["model_set_basis"] = lib.mle_model_set_basis(N, N, 0, 2, 2, dbl)
q = C.c_int()
res["model_nqoi"] = lib.mle_model_nqoi(N, C.byref(q))
res["model_destroy"] = lib.mle_model_destroy(N, N)
res["points"] = lib.mle_points(N, N, N, N, 1, 0, 1, 1, 0, dbl)
res["points_dev"] = lib.mle_points_dev(N, N, N, N, 1, 0, 1, 1, 0, dbl)
res["sample_batch"] = lib.mle_sample_batch(N, N, N, N, i32, 1, N, 1, 0, 1, 1, 0, dbl, N, N)
res["sample_batch_dev"] = lib.mle_sample_batch_dev(N, N, N, N, i32, 1, N, 1, 0, 1, 1, 0, dbl, N)
res["update_samples"] = lib.mle_update_samples(N, N, N, N, 1, i32, 1, N, N, N, N, N, 0, N)
res["sample_unbiased"] = lib.mle_sample_unbiased(N, N, N, N, i32, 1, N, 1, 0, 1, 1, 0, dbl, dbl, dbl, N)
res["sync"] = lib.mle_sync(N)
res["stream"] = lib.mle_stream(N, C.byref(h))
cnt = C.c_uint64()
res["launch_count"] = lib.mle_launch_count(N, C.byref(cnt))
sm, cc, hb = C.c_int(), C.c_int(), C.c_size_t()
res["device_info"] = lib.mle_device_info(N, C.byref(sm), C.byref(hb), C.byref(cc))
res["destroy"] = lib.mle_destroy(N)
res["moments_merge_null"] = lib.mle_moments_merge(N, N, 1)
res["genvec_unknown"] = lib.mle_genvec(b"nope", C.byref(C.POINTER(C.c_uint32)()), C.byref(q))
print(json.dumps(res))
'''


def test_every_entry_point_refuses_null_handles_without_crashing():
    """no exception or crash crosses the boundary: with a NULL context (what a caller holds after a failed mle_init) every
    entry point returns MLE_ERR_ARG; the destroy functions accept NULL like free()"""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, "-c", _NULL_CALLS % ROOT], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert len(res) == 21
    for name, rc in res.items():
        assert rc == (0 if name.endswith("destroy") else -1), (name, rc)
