"""GPU parity: Factor `*` / sum / normalize! / getindex / fused eliminate through the C ABI vs the CPU oracle.

Tolerance (north_star): <= 1e-12 relative in fp64.  Products and binary sums are expected -- and asserted --
bit-exact; longer sums are exact too because the kernel accumulates in the reference's column-major order.
"""
import numpy as np
import pytest

from oracle import capi
from oracle.factors_np import NpFactor

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def F(bn, dims, arr):
    return bn.Factor(list(dims), np.asarray(arr, np.float64))


def test_product_golden_324(bn, golden):
    """test/test_factors.jl:162-199, exact equality, both operand orders."""
    g = golden["factor_product"]
    a = np.arange(1.0, 109.0).reshape(g["phi1"]["card"], order="F")
    b = np.arange(1.0, 13.0).reshape(g["phi2"]["card"], order="F")
    p12 = F(bn, g["phi1"]["dims"], a) * F(bn, g["phi2"]["dims"], b)
    assert p12.dimensions == ["X", "C", "Y", "A", "Z", "B"]
    assert p12.potential.ravel(order="F").tolist() == g["true_pot"]
    p21 = F(bn, g["phi2"]["dims"], b) * F(bn, g["phi1"]["dims"], a)
    assert p21.dimensions == p12.dimensions
    assert p21.potential.ravel(order="F").tolist() == g["true_pot"]


def test_product_dimension_mismatch(bn, golden):
    g = golden["factor_product_mismatch"]
    with pytest.raises(bn.DimensionMismatch):
        F(bn, g["phi1"]["dims"], np.zeros(g["phi1"]["card"])) * F(bn, g["phi2"]["dims"], np.zeros(g["phi2"]["card"]))


def test_constructor_and_basic_queries(bn):
    """test/test_factors.jl:6-40."""
    with pytest.raises(ValueError):
        bn.Factor(["X", "X"], [2, 3])
    with pytest.raises(bn.DimensionMismatch):
        bn.Factor(["A", "X"], [3])
    with pytest.raises(bn.DimensionMismatch):
        bn.Factor(["Y", "A", "X"], [2, 3])
    phi = bn.Factor(["X", "Y"], [3, 3], 16)
    assert (phi.potential == 16).all()
    phi = bn.Factor(["X", "Y", "Z"], [2, 3, 4])
    assert phi.ndims == 3 and phi.size("X") == 2 and phi.size() == (2, 3, 4) and phi.size("Y", "X") == (3, 2)
    assert "X" in phi and "W" not in phi and len(phi) == 24
    phi = bn.Factor(["l1", "l2"], [2, 3])
    assert phi.pattern("l1").ravel().tolist() == [1, 2, 1, 2, 1, 2]        # test_factors.jl:59
    assert phi["l2"].ravel().tolist() == [1, 1, 2, 2, 3, 3]                # :60
    assert phi.pattern().tolist() == [[1, 1], [2, 1], [1, 2], [2, 2], [1, 3], [2, 3]]  # :62


def test_reducedim_golden(bn, golden):
    """test/test_factors.jl:111-130."""
    g = golden["reducedim"]
    pot = np.array(g["potential"], np.float64).reshape(g["card"], order="F")
    phi = F(bn, g["dims"], pot)
    with pytest.raises(ValueError):
        phi.sum("waldo")
    assert np.array_equal(phi.potential, pot)
    z = phi.broadcast_mul("Z", 0.0).sum(phi.names())
    assert z.potential.shape == () and float(z.potential) == 0.0
    s = phi.broadcast_mul("Z", g["broadcast_mul_Z"]).sum(g["sum_over"])
    assert s.dimensions == ["X"] and s.potential.tolist() == g["expected"]
    with pytest.raises(ValueError):
        phi.sum(["Y", "K", "Z", "waldo"])


def test_broadcast_golden(bn):
    """test/test_factors.jl:87-97 (the `*` cases)."""
    phi = F(bn, ["X", "Y"], np.array([[1, 2], [3, 4], [5, 6]], float))
    out = phi.broadcast_mul(["Y", "X"], [[10, 0.1], 100.0])
    np.testing.assert_allclose(out.potential, [[1000, 20], [3000, 40], [5000, 60]], rtol=1e-15)
    with pytest.raises(ValueError):
        phi.broadcast_mul(["X", "Z"], [[10, 1, 0.1], [1, 2, 3]])
    with pytest.raises(ValueError):
        phi.broadcast_mul(["Z", "X", "A"], [2, [10, 1, 0.1], [1, 2, 3]])
    with pytest.raises(bn.DimensionMismatch):
        phi.broadcast_mul("X", [2016, 58.0])


def test_normalize_golden(bn, golden):
    """test/test_factors.jl:71-81."""
    g = golden["normalize"]
    pot = np.array(g["potential_colmajor"], np.float64).reshape(g["card"], order="F")
    phi = F(bn, ["a", "b"], pot)
    phi2 = phi.normalize(p=1)
    np.testing.assert_allclose(phi2.potential.ravel(order="F"), g["p1"], rtol=1e-15)
    assert np.array_equal(phi.potential, pot)
    with pytest.raises(ValueError):
        phi.normalize("waldo")
    phi.normalize_(p=2)
    np.testing.assert_allclose(phi.potential.ravel(order="F"), g["p2"], rtol=1e-15)
    with pytest.raises(ValueError):
        phi.normalize_(p=3)


def test_getindex_golden(bn, golden):
    """test/test_factors.jl:138-147."""
    g = golden["getindex"]
    phi = F(bn, g["dims"], np.array(g["potential"], np.float64).reshape(g["card"], order="F"))
    with pytest.raises(TypeError):
        phi[{"Y": "waldo"}]
    with pytest.raises(IndexError):
        phi[{"Y": 16}]
    s = phi[g["assignment_1based"]]
    assert s.dimensions == ["X"] and s.potential.tolist() == g["expected"]
    full = phi[{"X": 2, "Y": 1, "Z": 2}]
    assert full.potential.shape == () and float(full.potential) == 6.0


@pytest.mark.parametrize("seed", range(24))
def test_random_mixed_radix_ops_bit_exact(bn, seed):
    """Random dims / cards / orders: product, sum (1..all dims), slice vs the oracle, exact equality."""
    rng = np.random.default_rng(seed)
    nvars = 9
    cards = rng.integers(1, 6, nvars)
    names = [f"v{seed}_{i}" for i in range(nvars)]
    i1 = list(rng.permutation(nvars)[:rng.integers(0, 7)])
    i2 = list(rng.permutation(nvars)[:rng.integers(0, 7)])
    a = rng.random([cards[i] for i in i1])
    b = rng.random([cards[i] for i in i2])
    f = F(bn, [names[i] for i in i1], a) * F(bn, [names[i] for i in i2], b)
    d, o = capi.factor_product(i1, a, i2, b)
    assert f.dimensions == [names[i] for i in d]
    assert np.array_equal(f.potential, o)
    if d:
        k = int(rng.integers(1, len(d) + 1))
        sd = [int(x) for x in rng.permutation(d)[:k]]
        s = f.sum([names[i] for i in sd])
        ds, os_ = capi.factor_sum(d, o, sd)
        assert s.dimensions == [names[i] for i in ds]
        assert np.array_equal(s.potential, os_)
        # slice some dims
        k = int(rng.integers(1, len(d) + 1))
        sl = {int(x): int(rng.integers(0, cards[x])) for x in rng.permutation(d)[:k]}
        g = f[{names[i]: v + 1 for i, v in sl.items()}]
        ref = NpFactor(d, o).slice(sl)
        assert g.dimensions == [names[i] for i in ref.dimensions]
        assert np.array_equal(g.potential, ref.potential)


@pytest.mark.parametrize("seed", range(8))
def test_fused_eliminate_equals_fold_then_sum(bn, seed):
    """sum(reduce(*, fs), h) in one kernel == the reference's materialised left fold (src/Inference/exact.jl:27)."""
    import ctypes as C
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.factors import Factor, var_id
    rng = np.random.default_rng(100 + seed)
    nvars = 8
    cards = rng.integers(2, 5, nvars)
    names = [f"e{seed}_{i}" for i in range(nvars)]
    h = int(rng.integers(0, nvars))
    k = int(rng.integers(2, 5))
    fs, nps = [], []
    for _ in range(k):
        others = [i for i in rng.permutation(nvars)[:rng.integers(1, 5)] if i != h]
        dims = list(rng.permutation([h] + others))
        arr = rng.random([cards[i] for i in dims])
        fs.append(F(bn, [names[i] for i in dims], arr))
        nps.append(NpFactor([names[i] for i in dims], arr))
    fold = nps[0]
    for x in nps[1:]:
        fold = fold * x
    ref = fold.sum(names[h])
    hs = (_lib.h64 * k)(*[f.handle for f in fs])
    sd = np.array([var_id(names[h])], np.int32)
    out = _lib.h64(0)
    _lib.check(_lib.lib().bn_factor_eliminate(hs, C.c_int32(k), _lib.ptr(sd, _lib.i32p), C.c_int32(1), C.byref(out)))
    got = Factor._from_handle(out)
    assert got.dimensions == ref.dimensions
    assert np.array_equal(got.potential, ref.potential)
    # unfused on the device: same thing
    acc = fs[0]
    for x in fs[1:]:
        acc = acc * x
    assert np.array_equal(acc.sum(names[h]).potential, ref.potential)


@pytest.mark.parametrize("case", ["P1", "P2", "P3", "S1", "S2", "S3", "S4", "E1"])
def test_binary_sweep_shapes_vs_oracle(bn, case):
    """BASELINE config C5 shapes at 2^20 (oracle finishes in seconds), exact equality."""
    k = 20
    rng = np.random.default_rng(7)
    names = [f"b{i}" for i in range(k + 2)]
    ids = list(range(k + 2))
    A = rng.random((2,) * k)
    fa = F(bn, names[:k], A)
    if case == "P1":
        sub = list(range(0, k, 2))
        B = rng.random((2,) * len(sub))
        got = fa * F(bn, [names[i] for i in sub], B)
        d, o = capi.factor_product(ids[:k], A, sub, B)
    elif case == "P2":
        A2 = rng.random((2,) * (k - 1))
        B = rng.random((2,) * (k - 1))
        got = F(bn, names[:k - 1], A2) * F(bn, names[1:k], B)
        d, o = capi.factor_product(ids[:k - 1], A2, ids[1:k], B)
    elif case == "P3":
        perm = list(rng.permutation(k))
        B = rng.random((2,) * k)
        got = fa * F(bn, [names[i] for i in perm], B)
        d, o = capi.factor_product(ids[:k], A, perm, B)
    elif case in ("S1", "S2", "S3", "S4"):
        sd = {"S1": [0], "S2": [k - 1], "S3": [k // 2], "S4": [1, 5, 11, k - 2]}[case]
        got = fa.sum([names[i] for i in sd])
        d, o = capi.factor_sum(ids[:k], A, sd)
    else:  # E1: three factors sharing var 0
        B = rng.random((2,) * 12)
        Cc = rng.random((2,) * 10)
        dB = [0] + list(range(8, 19))
        dC = [0] + list(range(1, 10))
        ref = (NpFactor(ids[:k], A) * NpFactor(dB, B) * NpFactor(dC, Cc)).sum(0)
        import ctypes as C
        from bayesnets_jl_b200 import _lib
        from bayesnets_jl_b200.factors import Factor, var_id
        fs = [fa, F(bn, [names[i] for i in dB], B), F(bn, [names[i] for i in dC], Cc)]
        hs = (_lib.h64 * 3)(*[f.handle for f in fs])
        sd = np.array([var_id(names[0])], np.int32)
        out = _lib.h64(0)
        _lib.check(_lib.lib().bn_factor_eliminate(hs, C.c_int32(3), _lib.ptr(sd, _lib.i32p), C.c_int32(1), C.byref(out)))
        got = Factor._from_handle(out)
        d, o = ref.dimensions, ref.potential
    assert got.dimensions == [names[i] for i in d]
    assert np.array_equal(got.potential, o)


def test_normalize_large_vs_oracle(bn):
    rng = np.random.default_rng(3)
    for n_dims, p in [(20, 1), (21, 2), (0, 1)]:
        A = rng.random((2,) * n_dims) if n_dims else np.array(3.0)
        f = F(bn, [f"n{i}" for i in range(n_dims)], A)
        f.normalize_(p=p)
        ref = capi.factor_normalize(A, p)
        np.testing.assert_allclose(f.potential, ref, rtol=RTOL)
    odd = rng.random((3, 5, 7))
    f = F(bn, ["o1", "o2", "o3"], odd).normalize_()
    np.testing.assert_allclose(f.potential, odd / odd.sum(), rtol=RTOL)
    assert abs(f.potential.sum() - 1) < 1e-13


def test_full_size_properties_2_24(bn):
    """BASELINE size (2^24 entries), checked through size-independent properties instead of the oracle:
    linearity of sum, sum of a product with a constant-1 factor, marginal consistency, idempotent normalise."""
    k = 24
    rng = np.random.default_rng(11)
    names = [f"q{i}" for i in range(k)]
    A = rng.random(1 << k)
    fa = bn.Factor(names, A.reshape((2,) * k, order="F"))
    # S1: innermost pairs; S2: outermost halves; both must sum to the same grand total
    s1 = fa.sum(names[0]).potential.reshape(-1, order="F")
    assert np.array_equal(s1, A[0::2] + A[1::2])
    s2 = fa.sum(names[-1]).potential.reshape(-1, order="F")
    assert np.array_equal(s2, A[:1 << (k - 1)] + A[1 << (k - 1):])
    mid = fa.sum(names[12]).potential.reshape(-1, order="F")
    A3 = A.reshape(1 << 12, 2, 1 << 11, order="F")
    assert np.array_equal(mid, (A3[:, 0, :] + A3[:, 1, :]).reshape(-1, order="F"))
    # product with an interleaved 12-dim factor, then check a few thousand random entries exactly
    sub = list(range(0, k, 2))
    B = rng.random(1 << 12)
    fb = bn.Factor([names[i] for i in sub], B.reshape((2,) * 12, order="F"))
    prod = (fa * fb).potential.reshape(-1, order="F")
    idx = rng.integers(0, 1 << k, 4096)
    bidx = np.zeros_like(idx)
    for j, d in enumerate(sub):
        bidx |= ((idx >> d) & 1) << j
    assert np.array_equal(prod[idx], A[idx] * B[bidx])
    # all-ones factor is the identity of `*`
    ones = bn.Factor([names[3], names[17]], np.ones((2, 2)))
    assert np.array_equal((fa * ones).potential.reshape(-1, order="F"), A)
    # normalise: sums to one, second normalise is a no-op up to rounding
    n1 = fa.normalize().potential.reshape(-1, order="F")
    assert abs(n1.sum() - 1.0) < 1e-12
    np.testing.assert_allclose(n1, A / A.sum(), rtol=RTOL)


@pytest.mark.parametrize("seed", range(6))
def test_product_with_reordered_large_operand_row_vector_path(bn, seed):
    """Products whose second operand is large and ordered differently from the output take the row-vector gather
    path (the operand's two fastest dims become the tile rows); mixed radix incl. a 4-state leading dim, a third
    small operand, and operands that lack the output's first dim.  Bit-exact vs the oracle's fold."""
    rng = np.random.default_rng(50 + seed)
    nd = 17
    cards = [2] * nd
    if seed % 2:
        cards[int(rng.integers(9, nd))] = 4
        cards[int(rng.integers(0, 6))] = 3
    names = [f"rv{seed}_{i}" for i in range(nd)]
    A = rng.random(cards)
    perm = [int(x) for x in rng.permutation(nd)]
    if seed % 2:  # put the 4-state dim first in the second operand
        four = cards.index(4)
        perm.remove(four)
        perm.insert(0, four)
    if seed >= 4:  # second operand lacks the output's first dim
        perm.remove(0)
    B = rng.random([cards[i] for i in perm])
    fa, fb = F(bn, names, A), F(bn, [names[i] for i in perm], B)
    ref = NpFactor(list(range(nd)), A) * NpFactor(perm, B)
    got = fa * fb
    assert got.dimensions == [names[i] for i in ref.dimensions]
    assert np.array_equal(got.potential, ref.potential)
    # a third, small operand through the fused n-ary product
    small = sorted(int(x) for x in rng.choice(nd, size=3, replace=False))
    Cc = rng.random([cards[i] for i in small])
    fc = F(bn, [names[i] for i in small], Cc)
    ref3 = ref * NpFactor(small, Cc)
    got3 = bn.eliminate([fa, fb, fc], [])
    assert got3.dimensions == [names[i] for i in ref3.dimensions]
    assert np.array_equal(got3.potential, ref3.potential)


@pytest.mark.parametrize("seed", range(8))
def test_large_streaming_shapes_bit_exact(bn, seed):
    """2^22..2^24-entry leading operands (thousands of tiles): 512-entry leading block, mixed radix, summed states
    inside and outside the contiguous span, row dims the leading operand lacks, 1-3 operands.  Bit-exact vs the oracle."""
    rng = np.random.default_rng(900 + seed)
    lead_cards = [[2] * 9, [4, 2, 2, 2, 2, 2, 2, 2], [2, 4, 4, 2, 2, 2, 2], [4, 4, 4, 2, 2, 2]][seed % 4]
    tail = [2] * (22 - int(np.log2(np.prod(lead_cards))))
    if seed % 2:
        tail[int(rng.integers(0, len(tail)))] = 3
        tail[int(rng.integers(0, len(tail)))] = 4
    cards = lead_cards + tail
    nd = len(cards)
    names = [f"tm{seed}_{i}" for i in range(nd)]
    ids = list(range(nd))
    A = rng.random(cards)
    fa = F(bn, names, A)
    na = NpFactor(ids, A)

    def both(fn):
        return fn()

    # sums: inside the span (dim 0), outside (middle / last), both, and a 3-dim sum
    for sd in ([0], [nd - 1], [len(lead_cards) + 2], [0, nd - 2], [1, len(lead_cards), nd - 1]):
        got = both(lambda: fa.sum([names[i] for i in sd]))
        ref = na.sum(sd)
        assert got.dimensions == [names[i] for i in ref.dimensions]
        assert np.array_equal(got.potential, ref.potential), sd
    # products: small interleaved operand; operand with a dim the leading operand lacks
    sub = [int(x) for x in sorted(rng.choice(nd, size=6, replace=False))]
    B = rng.random([cards[i] for i in sub])
    got = both(lambda: fa * F(bn, [names[i] for i in sub], B))
    assert np.array_equal(got.potential, (na * NpFactor(sub, B)).potential)
    extra = f"tm{seed}_x"
    Bx = rng.random([cards[i] for i in sub] + [2])
    got = both(lambda: fa * F(bn, [names[i] for i in sub] + [extra], Bx))
    assert np.array_equal(got.potential, (na * NpFactor(sub + [nd], Bx)).potential)
    # fused eliminate: 3 operands sharing dim 0 (inside the span) / a tail dim (outside)
    for h in (0, nd - 3):
        d1 = [h] + [int(x) for x in rng.choice([i for i in ids if i != h], size=5, replace=False)]
        d2 = [int(x) for x in rng.choice([i for i in ids if i != h], size=4, replace=False)] + [h]
        B1, B2 = rng.random([cards[i] for i in d1]), rng.random([cards[i] for i in d2])
        fs = [fa, F(bn, [names[i] for i in d1], B1), F(bn, [names[i] for i in d2], B2)]
        got = both(lambda: bn.eliminate(fs, [names[h]]))
        ref = (na * NpFactor(d1, B1) * NpFactor(d2, B2)).sum(h)
        assert got.dimensions == [names[i] for i in ref.dimensions]
        assert np.array_equal(got.potential, ref.potential), h


@pytest.mark.parametrize("seed", range(16))
def test_sum_streaming_kernel_shapes_bit_exact(bn, seed, monkeypatch):
    """csrc/factor_stream.cu: sum(phi, dims) of one factor (factors_dims.jl:60-85,100).  Mixed-radix shapes that hit
    both modes (innermost dim kept -> 128-bit columns; innermost dim summed -> output pairs), 1..4 summed dims up to 16
    joint states, odd segments that must fall back, sizes that leave a ragged last CTA.  Equal to the oracle bit for
    bit, and to the contraction kernel (BN_FACTOR_NO_STREAM)."""
    rng = np.random.default_rng(4200 + seed)
    nd = int(rng.integers(3, 12))
    cards = [int(rng.choice([2, 2, 2, 3, 4, 5, 6])) for _ in range(nd)]
    while np.prod(cards) > (1 << 21):
        cards[int(rng.integers(0, nd))] = 2 if np.prod(cards) < (1 << 24) else 1
    if seed % 2 == 0:
        cards[0] = int(rng.choice([2, 4, 6]))  # even innermost dim: vector columns or pairs
    names = [f"ss{seed}_{i}" for i in range(nd)]
    A = rng.random(cards)
    fa, na = F(bn, names, A), NpFactor(list(range(nd)), A)
    picks = [[0], [nd - 1], [nd // 2], [0, 1], [0, nd - 1], [1, nd - 2]]
    for _ in range(4):
        picks.append(sorted(int(x) for x in rng.choice(nd, size=int(rng.integers(1, min(5, nd))), replace=False)))
    for sd in picks:
        sd = sorted(set(sd))
        if len(sd) == nd:
            continue
        ref = na.sum(sd)
        n0 = bn.launch_count()
        got = fa.sum([names[i] for i in sd])
        assert bn.launch_count() > n0
        assert got.dimensions == [names[i] for i in ref.dimensions]
        assert np.array_equal(got.potential, ref.potential), (cards, sd)
        monkeypatch.setenv("BN_FACTOR_NO_STREAM", "1")
        gen = fa.sum([names[i] for i in sd])
        monkeypatch.delenv("BN_FACTOR_NO_STREAM")
        assert np.array_equal(gen.potential, got.potential), (cards, sd)


@pytest.mark.parametrize("seed", range(12))
def test_eliminate_big_times_small_bit_exact(bn, seed):
    """The variable-elimination step sum(A * small..., dims) (src/Inference/exact.jl:27) where A spans every dim and the
    small operands are contiguous stretches of A's dims plus the summed dims, in any operand order: innermost dim summed
    / kept, 1-2 summed dims, mixed radix.  Equal to the oracle's left fold + sum bit for bit."""
    rng = np.random.default_rng(7300 + seed)
    nd = int(rng.integers(9, 15))
    cards = [int(rng.choice([2, 4])) for _ in range(2)] + [int(rng.choice([2, 2, 2, 2, 3, 4])) for _ in range(nd - 2)]
    while np.prod(cards) < (1 << 17):
        i = int(rng.integers(2, nd))
        cards[i] = 4
        if all(x == 4 for x in cards[2:]):
            break
    while np.prod(cards) > (1 << 21):
        cards[2 + int(np.argmax(cards[2:]))] = 2
    assert (1 << 16) <= np.prod(cards) <= (1 << 21)
    names = [f"se{seed}_{i}" for i in range(nd)]
    ids = list(range(nd))
    A = rng.random(cards)
    fa, na = F(bn, names, A), NpFactor(ids, A)
    for trial in range(4):
        h = [0] if trial % 2 == 0 else [int(rng.integers(2, nd))]
        if trial == 3 and nd > 6:
            h = sorted({h[0], int(rng.integers(2, nd))})
        smalls = []
        for _ in range(int(rng.integers(1, 3))):
            lo = int(rng.integers(1, nd - 2))
            hi = int(rng.integers(lo + 1, min(nd, lo + 6) + 1))
            run = [i for i in range(lo, hi) if i not in h]
            if np.prod([cards[i] for i in run] + [1]) * np.prod([cards[i] for i in h]) > (1 << 16):
                run = run[:3]
            sd = h if rng.random() < 0.8 else h[:1]
            dims = (sd + run) if rng.random() < 0.6 else (run + sd)
            if not dims:
                dims = list(h)
            smalls.append((dims, rng.random([cards[i] for i in dims])))
        order = [("A", None)] + [("S", i) for i in range(len(smalls))]
        if rng.random() < 0.5:
            order = order[1:2] + order[0:1] + order[2:]
        fs, ref = [], None
        for kind, i in order:
            f, n = (fa, na) if kind == "A" else (F(bn, [names[j] for j in smalls[i][0]], smalls[i][1]),
                                                 NpFactor(smalls[i][0], smalls[i][1]))
            fs.append(f)
            ref = n if ref is None else ref * n
        ref = ref.sum(h)
        got = bn.eliminate(fs, [names[i] for i in h])
        assert got.dimensions == [names[i] for i in ref.dimensions]
        assert np.array_equal(got.potential, ref.potential), (cards, h, [s[0] for s in smalls], order)


def test_block_cache_reuse_and_trim(bn):
    """Freed factor tables are parked inside the library and reused by the next allocation of the same size on the same
    stream (csrc/core.cu pool_alloc / pool_free); bn_pool_trim hands them back to the driver.  Results are unaffected by
    reuse: a chain of sums over recycled blocks equals the oracle's."""
    import gc
    rng = np.random.default_rng(5)
    k = 18
    names = [f"bc{i}" for i in range(k)]
    A = rng.random((2,) * k)
    fa, na = F(bn, names, A), NpFactor(list(range(k)), A)
    bn.pool_trim()
    for rep in range(3):  # same sizes every round: rounds 2 and 3 run entirely on recycled blocks
        f, n = fa, na
        for d in (0, 5, 9, 3):
            f, n = f.sum([names[d]]), n.sum(d)
            assert np.array_equal(f.potential, n.potential)
        del f
        gc.collect()
    parked = bn.pool_trim()
    assert parked >= 8 * (1 << (k - 1))  # at least the largest intermediate was parked
    assert bn.pool_trim() == 0
    assert np.array_equal(fa.sum([names[0]]).potential, na.sum(0).potential)


def test_sum_over_more_states_than_one_launch_takes(bn):
    """sum(phi, dims) over 4^6 = 4096 and 2^13 = 8192 joint states (one launch takes 2048): summed in stages, dims in
    the factor's order.  Same terms as the reference's single pass, re-associated: within 1e-12 (north_star) of the
    oracle; sums within the limit stay bit-identical (the other tests)."""
    rng = np.random.default_rng(31)
    for cards, sd in (([4] * 6 + [2, 3], [0, 1, 2, 3, 4, 5]), ([2] * 16, list(range(1, 14))), ([3, 5, 7, 4, 6, 8, 2], [6, 4, 2, 1, 5, 0])):
        names = [f"big{len(cards)}_{i}" for i in range(len(cards))]
        A = rng.random(cards)
        got = F(bn, names, A).sum([names[i] for i in sd])
        ref = NpFactor(list(range(len(cards))), A).sum(sd)
        assert got.dimensions == [names[i] for i in ref.dimensions]
        np.testing.assert_allclose(got.potential, ref.potential, rtol=RTOL, atol=0)
    with pytest.raises(ValueError):
        F(bn, ["u1", "u2"], rng.random((2, 2))).sum(["u1", "u1"])


def test_push_streaming_kernel_vs_oracle(bn, monkeypatch):
    """push!(phi, dim, length) (factors_main.jl:148-158) through the duplicate kernel: even and odd table lengths (odd
    falls back to the contraction), lengths 1..5, a ragged last CTA."""
    rng = np.random.default_rng(77)
    for i, cards in enumerate([[2] * 15, [3, 5, 7, 11], [4, 3, 1001], [2, 3, 5, 7, 9, 4], [6]]):
        for length in (1, 2, 5):
            names = [f"pu{i}_{length}_{j}" for j in range(len(cards))]
            A = rng.random(cards)
            f, o = F(bn, names, A), NpFactor(names, np.asfortranarray(A.copy()))
            f.push_("new", length)
            o2 = o.push("new", length)
            assert f.dimensions == o2.dimensions and np.array_equal(f.potential, o2.potential)


# ------------------------------------------------------------------- the rest of the Factor algebra (csrc/factor_ops.cu)
def test_broadcast_reference_suite(bn):
    """test/test_factors.jl:87-105, same assertions."""
    phi = F(bn, ["X", "Y"], [[1.0, 2], [3, 4], [5, 6]])
    got = phi.broadcast("*", ["Y", "X"], [[10, 0.1], 100.0]).potential
    np.testing.assert_allclose(got, [[1000, 20], [3000, 40], [5000, 60]], rtol=1e-15)
    assert np.array_equal(phi.potential, [[1.0, 2], [3, 4], [5, 6]])  # broadcast copies, broadcast! mutates
    with pytest.raises(ValueError):
        phi.broadcast("*", ["X", "Z"], [[10, 1, 0.1], [1, 2, 3]])
    with pytest.raises(ValueError):
        phi.broadcast("*", ["Z", "X", "A"], [2, [10, 1, 0.1], [1, 2, 3]])
    with pytest.raises(bn.DimensionMismatch):
        phi.broadcast("*", "X", [2016, 58.0])
    p3 = F(bn, ["X", "Y", "Z"], np.array([1.0, 2, 3, 2, 3, 4, 4, 6, 7, 8, 10, 16]).reshape((3, 2, 2), order="F"))
    two = p3.broadcast("+", "Z", [10, 0.1]).broadcast("+", "X", 10.0)
    one = p3.broadcast("+", ["X", "Z"], [10.0, [10, 0.1]])
    assert np.array_equal(two.potential, one.potential)
    red = p3.broadcast("*", "Z", [1, 10.0]).sum(["Y", "Z"])           # :121-129
    assert red.dimensions == ["X"] and np.array_equal(red.potential, [123.0, 165.0, 237.0])
    p3.broadcast_("*", "Z", 0.0)
    assert not p3.potential.any()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_join_reduce_normalize_ops_equal_oracle(bn, seed):
    """/ + - joins (with the reference's larger-first swap), prod / maximum / minimum, normalize!(phi, dims; p) and
    every broadcast op on random factors with ragged cardinalities: identical to the numpy restatement."""
    rng = np.random.default_rng(seed)
    names = list("abcdefg")
    da = [str(x) for x in rng.choice(names, size=int(rng.integers(1, 6)), replace=False)]
    db = [str(x) for x in rng.choice(names, size=int(rng.integers(1, 5)), replace=False)]
    card = {n: int(rng.integers(1, 5)) for n in names}
    A = rng.random([card[d] for d in da]) + 0.25
    B = rng.random([card[d] for d in db]) + 0.25
    if seed == 3:
        A.flat[0] = np.nan  # max / min propagate NaN like Base.max
    fa, fb = F(bn, da, A), F(bn, db, B)
    oa, ob = NpFactor(da, A), NpFactor(db, B)
    pairs = [("/", fa / fb, oa.join("/", ob)), ("/", fb / fa, ob.join("/", oa)), ("+", fa + fb, oa.join("+", ob)),
             ("-", fa - fb, oa.join("-", ob)), ("-", fb - fa, ob.join("-", oa)), ("*", fa.join("*", fb), oa.join("*", ob))]
    for op, got, want in pairs:
        assert got.dimensions == want.dimensions
        assert np.array_equal(got.potential, want.potential, equal_nan=True), op
    for k in range(1, len(da) + 1):
        dims = [str(x) for x in rng.choice(da, size=k, replace=False)]
        for op, fn in (("*", fa.prod), ("max", fa.maximum), ("min", fa.minimum)):
            got, want = fn(dims), oa.reducedim(op, dims)
            assert got.dimensions == want.dimensions
            assert np.array_equal(got.potential, want.potential, equal_nan=True), (op, dims)
        assert np.array_equal(fa.sum(dims).potential, oa.reducedim("+", dims).potential, equal_nan=True)
        for p in (1, 2):
            got, want = fa.normalize(dims, p=p), oa.normalize_dims(dims, p=p)
            assert np.array_equal(got.potential, want.potential, equal_nan=True), (dims, p)
    for op in ("*", "/", "+", "-", "max", "min"):
        d = da[int(rng.integers(0, len(da)))]
        v = rng.random(card[d]) + 0.5
        assert np.array_equal(fa.broadcast(op, d, v).potential, oa.broadcast(op, d, v).potential, equal_nan=True), op
        assert np.array_equal(fa.broadcast(op, d, 0.75).potential, oa.broadcast(op, d, 0.75).potential, equal_nan=True), op
    with pytest.raises(ValueError):
        fa.prod("waldo")
    with pytest.raises(ValueError):
        fa.normalize("waldo")
    with pytest.raises(RuntimeError):
        fa.join("*", fb, kind="inner")


def test_factor_ops_large_binary(bn):
    """2^20-entry binary factors: merged-dim addressing of the gather kernels, in-place normalize over inner and outer dims."""
    rng = np.random.default_rng(5)
    dims = [f"v{i}" for i in range(20)]
    A = rng.random([2] * 20) + 0.1
    sub = [dims[i] for i in (3, 0, 19, 7)]
    B = rng.random([2] * 4) + 0.1
    fa, fb = F(bn, dims, A), F(bn, sub, B)
    oa, ob = NpFactor(dims, A), NpFactor(sub, B)
    assert np.array_equal((fa / fb).potential, oa.join("/", ob).potential)
    assert np.array_equal((fb - fa).potential, ob.join("-", oa).potential)
    for rd in (["v0", "v1", "v2"], ["v19"], ["v5", "v18", "v9"]):
        assert np.array_equal(fa.maximum(rd).potential, oa.reducedim("max", rd).potential)
        assert np.array_equal(fa.prod(rd).potential, oa.reducedim("*", rd).potential)
        assert np.array_equal(fa.normalize(rd).potential, oa.normalize_dims(rd).potential)


# ------------------------------------------------------------------ setindex!, permutedims, push! (row a9)
def test_setindex_golden(bn, golden):
    """test/test_factors.jl:147-152 through the mirror: one cell, then a slab over the unassigned dim."""
    g = golden["setindex"]
    pot = np.array(g["potential"], np.float64).reshape(g["card"], order="F")
    f = F(bn, g["dims"], pot)
    want = NpFactor(g["dims"], pot.copy())
    for st in g["steps"]:
        f[st["assignment_1based"]] = st["value"]
        want.setindex(st["value"], {k: v - 1 for k, v in st["assignment_1based"].items()})
        assert np.array_equal(f.potential, want.potential)
    assert f.potential[1, 0, 1] == 1600.0
    assert f.potential[0, 1, :].tolist() == [2016.0, 2016.0]
    with pytest.raises(IndexError):
        f[{"Y": 16}] = 1.0
    with pytest.raises(TypeError):
        f[{"Y": "waldo"}] = 1.0
    before = f.potential
    f[{"X": 3, "K": 16}] = np.array([[1.0, 2.0], [3.0, 4.0]])  # `.= v` with an array of the slab's shape; K ignored
    before[2, :, :] = [[1.0, 2.0], [3.0, 4.0]]
    assert np.array_equal(f.potential, before)
    with pytest.raises(bn.DimensionMismatch):
        f[{"X": 3}] = np.array([1.0, 2.0, 3.0])
    f[{}] = 0.5  # every dim a Colon: fill
    assert (f.potential == 0.5).all()


@pytest.mark.parametrize("seed", range(6))
def test_setindex_permutedims_push_random_vs_oracle(bn, seed):
    """Random mixed-radix factors: slab stores (scalar and array), permutations and push! are pure data movement, so
    the device tables must equal the oracle's bit for bit."""
    rng = np.random.default_rng(seed)
    nd = int(rng.integers(1, 7))
    card = [int(rng.integers(1, 6)) for _ in range(nd)]
    dims = [f"v{seed}_{i}" for i in range(nd)]
    pot = rng.random(card)
    f, o = F(bn, dims, pot), NpFactor(dims, np.asfortranarray(pot.copy()))
    for _ in range(4):
        k = int(rng.integers(0, nd + 1))
        pick = list(rng.choice(nd, k, replace=False))
        a = {dims[i]: int(rng.integers(1, card[i] + 1)) for i in pick}
        slab = [card[i] for i in range(nd) if i not in pick]
        v = float(rng.random()) if rng.random() < 0.5 else rng.random(slab)
        f[a] = v
        o.setindex(v, {d: s - 1 for d, s in a.items()})
        assert np.array_equal(f.potential, o.potential)
    perm = list(rng.permutation(nd))
    fp, op = f.permutedims([p + 1 for p in perm]), o.permutedims(perm)
    assert fp.dimensions == op.dimensions and np.array_equal(fp.potential, op.potential)
    assert np.array_equal(f.potential, o.potential)  # permutedims leaves phi alone
    f.permutedims_([p + 1 for p in perm])
    assert f.dimensions == op.dimensions and np.array_equal(f.potential, op.potential)
    length = int(rng.integers(1, 5))
    f.push_("new", length)
    o2 = op.push("new", length)
    assert f.dimensions == o2.dimensions and np.array_equal(f.potential, o2.potential)
    with pytest.raises(RuntimeError):
        f.push_("new", 2)
    with pytest.raises(ValueError):
        f.permutedims([1] * len(f.dimensions) if len(f.dimensions) > 1 else [2])
    assert f.indexin("new") == len(f.dimensions) and f.indexin(["new", "waldo"]) == [len(f.dimensions), None]
    s = f.similar()
    assert s.dimensions == f.dimensions and s.size() == f.size()
    f.sum_("new")  # sum! rebinds the table
    assert f.dimensions == op.dimensions
    np.testing.assert_allclose(f.potential, op.potential * length, rtol=1e-14)


def test_permutedims_and_push_large_binary(bn):
    """2^22-entry binary factor: reverse all dims / rotate (the gather side of the contraction kernel), push a 4-state
    dim (2^24 outputs); a permutation applied twice with its inverse returns the table."""
    rng = np.random.default_rng(11)
    k = 22
    dims = [f"b{i}" for i in range(k)]
    a = rng.random((2,) * k)
    f = F(bn, dims, a)
    rev = list(range(k, 0, -1))
    r = f.permutedims(rev)
    assert r.dimensions == dims[::-1]
    assert np.array_equal(r.potential, np.transpose(a, [p - 1 for p in rev]))
    assert np.array_equal(r.permutedims(rev).potential, a)
    rot = list(range(8, k + 1)) + list(range(1, 8))
    assert np.array_equal(f.permutedims(rot).potential, np.transpose(a, [p - 1 for p in rot]))
    f.push_("q", 4)
    got = f.potential
    assert got.shape == (2,) * k + (4,)
    for j in range(4):
        assert np.array_equal(got[..., j], a)


@pytest.mark.parametrize("seed", range(10))
def test_permutedims_tiled_transpose_random_shapes(bn, seed):
    """The tiled transpose (csrc/permute.cu) over shapes that exercise its planner: many small mixed-radix dims (tile
    and remaining dims unmerged), size-1 dims, a random permutation of 20 binary dims (13+ remaining dims, one lane
    each), large odd cards, identity-like permutations; plus the contraction-kernel fallback (BN_NO_TILED_PERMUTE is
    read once per process, so the fallback is reached through shapes the tiling refuses: a 255 x 255 leading pair)."""
    rng = np.random.default_rng(100 + seed)
    if seed == 0:
        card = [2] * 20
    elif seed == 1:
        card = [255, 255, 3, 2]          # A x B tile would be 65025 entries -> fallback path
    elif seed == 2:
        card = [7, 1, 5, 3, 1, 11, 2, 13, 3, 2]
    elif seed == 3:
        card = [64, 3, 64, 5, 2]
    else:
        nd = int(rng.integers(2, 13))
        card = [int(rng.integers(1, 6)) for _ in range(nd)]
    nd = len(card)
    dims = [f"t{seed}_{i}" for i in range(nd)]
    a = rng.random(card)
    f = F(bn, dims, a)
    perms = [list(rng.permutation(nd)) for _ in range(3)]
    perms.append(list(range(nd)))                      # identity
    perms.append(list(range(nd))[::-1])                # full reversal
    perms.append([0] + list(range(nd - 1, 0, -1)))     # leading dim stays
    for perm in perms:
        g = f.permutedims([int(x) + 1 for x in perm])
        assert g.dimensions == [dims[i] for i in perm]
        assert np.array_equal(g.potential, np.transpose(a, perm))


def test_permutedims_fuzz_against_numpy(bn):
    """400 random shapes x 3 random permutations (binary chains, small mixed radix, a few wide dims, many tiny dims)
    through the tiled transpose, its pair-vectorised variant and the contraction fallback: identical to numpy.transpose
    (tools/probe/permute_fuzz.py ran 9000 such permutations on the B200 without a mismatch)."""
    rng = np.random.default_rng(12)
    for case in range(400):
        kind = case % 5
        if kind == 0:
            card = [2] * int(rng.integers(1, 19))
        elif kind == 1:
            card = [int(rng.integers(1, 8)) for _ in range(int(rng.integers(1, 9)))]
        elif kind == 2:
            card = [int(rng.choice([1, 2, 3, 4, 16, 31, 32, 33, 64, 100, 255])) for _ in range(int(rng.integers(1, 4)))]
        elif kind == 3:
            card = [int(rng.choice([2, 4, 8])) for _ in range(int(rng.integers(2, 9)))]
        else:
            card = [int(rng.integers(1, 4)) for _ in range(int(rng.integers(8, 15)))]
        while int(np.prod(card)) > 1 << 20:
            card.pop()
        nd = len(card)
        a = rng.random(card)
        f = F(bn, [f"fz{case}_{i}" for i in range(nd)], a)
        for _ in range(3):
            perm = [int(x) for x in rng.permutation(nd)]
            g = f.permutedims([p + 1 for p in perm])
            assert np.array_equal(g.potential, np.transpose(a, perm)), (card, perm)
