"""Pin the CPU oracle against the reference's own golden vectors -- runs without a GPU.

Every expected value here comes from tests/golden/reference_vectors.json, which make_golden.py extracted from
the reference's test files (file:line in the JSON), or is a closed form derived from CPTs cited there.
"""
import itertools

import numpy as np
import pytest

from oracle import capi
from oracle.factors_np import (DimensionMismatch, NpFactor, brute_force_posterior, cpd_factor, exact_infer,
                               make_net)


def net_from(nodes):
    return make_net([(n, p, r) for n, p, r in nodes])


# ------------------------------------------------------------------------------------- Philox
def test_philox_known_answers(golden):
    for v in golden["philox4x32_10_kat"]["vectors"]:
        out = capi.philox4x32_10(v["ctr"], v["key"])
        assert [int(x) for x in out] == v["out"]


# ------------------------------------------------------------------------------------- Factors
def test_factor_product_golden_324(golden):
    """test/test_factors.jl:162-199 -- exact equality, both oracle implementations."""
    g = golden["factor_product"]
    a = np.arange(1.0, 109.0).reshape(g["phi1"]["card"], order="F")
    b = np.arange(1.0, 13.0).reshape(g["phi2"]["card"], order="F")
    f = NpFactor(g["phi1"]["dims"], a) * NpFactor(g["phi2"]["dims"], b)
    assert f.dimensions == ["X", "C", "Y", "A", "Z", "B"]
    assert f.potential.ravel(order="F").tolist() == g["true_pot"]
    ids = {n: i for i, n in enumerate("XCYAZB")}
    d, o = capi.factor_product([ids[x] for x in g["phi1"]["dims"]], a, [ids[x] for x in g["phi2"]["dims"]], b)
    assert d == [ids[x] for x in "XCYAZB"]
    assert o.ravel(order="F").tolist() == g["true_pot"]
    # operand order does not matter: the larger factor always leads (factors_dims.jl:189-191)
    d2, o2 = capi.factor_product([ids[x] for x in g["phi2"]["dims"]], b, [ids[x] for x in g["phi1"]["dims"]], a)
    assert d2 == d and np.array_equal(o2, o)


def test_factor_product_dimension_mismatch(golden):
    g = golden["factor_product_mismatch"]
    f1 = NpFactor(g["phi1"]["dims"], np.zeros(g["phi1"]["card"]))
    f2 = NpFactor(g["phi2"]["dims"], np.zeros(g["phi2"]["card"]))
    with pytest.raises(DimensionMismatch):
        f1 * f2
    with pytest.raises(ValueError):
        capi.factor_product([0, 1, 2], f1.potential, [3, 0, 1], f2.potential)


def test_reducedim_golden(golden):
    """test/test_factors.jl:111-130: broadcast(*, phi, :Z, [1,10]) then sum!(phi, [:Y,:Z]) == [123,165,237]."""
    g = golden["reducedim"]
    pot = np.array(g["potential"], np.float64).reshape(g["card"], order="F")
    phi = NpFactor(g["dims"], pot) * NpFactor(["Z"], np.array(g["broadcast_mul_Z"]))
    assert phi.dimensions == g["dims"]
    s = phi.sum(g["sum_over"])
    assert s.dimensions == ["X"] and s.potential.tolist() == g["expected"]
    d, o = capi.factor_sum([0, 1, 2], phi.potential, [1, 2])
    assert d == [0] and o.tolist() == g["expected"]
    z = (NpFactor(g["dims"], pot) * NpFactor(["Z"], np.zeros(2))).sum(g["dims"])
    assert z.potential.shape == () and float(z.potential) == 0.0  # zero-dim result, :121
    d, o = capi.factor_sum([0, 1, 2], np.zeros(g["card"]), [0, 1, 2])
    assert d == [] and o.shape == () and float(o) == 0.0
    with pytest.raises(ValueError):  # :116,:132
        NpFactor(g["dims"], pot).sum(["Y", "K", "Z", "waldo"])
    with pytest.raises(ValueError):
        capi.factor_sum([0, 1, 2], pot, [1, 7])


def test_normalize_golden(golden):
    """test/test_factors.jl:71-81 (p=2 divides by sum(abs2), no sqrt)."""
    g = golden["normalize"]
    pot = np.array(g["potential_colmajor"], np.float64).reshape(g["card"], order="F")
    f = NpFactor(["a", "b"], pot)
    np.testing.assert_allclose(f.normalize(1).potential.ravel(order="F"), g["p1"], rtol=1e-15)
    np.testing.assert_allclose(f.normalize(2).potential.ravel(order="F"), g["p2"], rtol=1e-15)
    np.testing.assert_allclose(capi.factor_normalize(pot, 1).ravel(order="F"), g["p1"], rtol=1e-15)
    np.testing.assert_allclose(capi.factor_normalize(pot, 2).ravel(order="F"), g["p2"], rtol=1e-15)
    with pytest.raises(ValueError):
        capi.factor_normalize(pot, 3)
    assert np.array_equal(f.potential, pot)  # normalize() leaves phi alone, :75


def test_getindex_golden(golden):
    g = golden["getindex"]
    pot = np.array(g["potential"], np.float64).reshape(g["card"], order="F")
    a = {k: v - 1 for k, v in g["assignment_1based"].items()}
    s = NpFactor(g["dims"], pot).slice(a)
    assert s.dimensions == ["X"] and s.potential.tolist() == g["expected"]
    with pytest.raises(IndexError):
        NpFactor(g["dims"], pot).slice({"Y": 15})
    with pytest.raises(TypeError):
        NpFactor(g["dims"], pot).slice({"Y": "waldo"})


def test_setindex_golden(golden):
    """test/test_factors.jl:147-152: phi[X=>2, Y=>1, Z=>2] = 1600.0 sets one cell; phi[X=>1, Y=>2] = 2016 fills the
    slab over the unassigned dim."""
    g = golden["setindex"]
    f = NpFactor(g["dims"], np.array(g["potential"], np.float64).reshape(g["card"], order="F"))
    untouched = f.potential.copy()
    st = g["steps"][0]
    f.setindex(st["value"], {k: v - 1 for k, v in st["assignment_1based"].items()})
    i = tuple(x - 1 for x in st["check_index_1based"])
    assert f.potential[i] == st["expected"]
    untouched[i] = st["expected"]
    assert np.array_equal(f.potential, untouched)
    st = g["steps"][1]
    f.setindex(st["value"], {k: v - 1 for k, v in st["assignment_1based"].items()})
    assert f.potential[0, 1, :].tolist() == st["expected_slab"]
    untouched[0, 1, :] = st["expected_slab"]
    assert np.array_equal(f.potential, untouched)
    with pytest.raises(IndexError):
        f.setindex(1.0, {"Y": 15})
    with pytest.raises(TypeError):
        f.setindex(1.0, {"Y": "waldo"})
    f.setindex(7.0, {"K": 3})  # a dim phi lacks is ignored: every dim is a Colon -> fill
    assert (f.potential == 7.0).all()


def test_permutedims_and_push_restatements():
    """permutedims (factors_main.jl:160-166) and push! / duplicate (:148-158, auxiliary.jl:50-65, whose docstring
    example repeats [1 3; 2 4] three times along a new last dim)."""
    f = NpFactor(["A", "B"], np.array([[1.0, 3.0], [2.0, 4.0]]))
    d = f.push("C", 3)
    assert d.dimensions == ["A", "B", "C"] and d.potential.shape == (2, 2, 3)
    for k in range(3):
        assert np.array_equal(d.potential[:, :, k], f.potential)
    with pytest.raises(RuntimeError):
        f.push("A", 2)
    g = NpFactor(["X", "Y", "Z"], np.arange(24, dtype=np.float64).reshape((2, 3, 4), order="F"))
    h = g.permutedims([2, 0, 1])
    assert h.dimensions == ["Z", "X", "Y"] and h.potential.shape == (4, 2, 3)
    assert h.potential[3, 1, 2] == g.potential[1, 2, 3]
    assert np.array_equal(h.permutedims([1, 2, 0]).potential, g.potential)
    with pytest.raises(ValueError):
        g.permutedims([0, 0, 1])


def test_factor_constructor_errors():
    """test/test_factors.jl:6-8."""
    with pytest.raises(ValueError):
        NpFactor(["X", "X"], np.zeros((2, 3)))
    with pytest.raises(ValueError):
        NpFactor(["A", "X"], np.zeros(3))


@pytest.mark.parametrize("seed", range(6))
def test_c_and_numpy_factor_ops_agree(seed):
    """Two independent restatements (C odometer vs numpy broadcasting) must agree bit for bit."""
    rng = np.random.default_rng(seed)
    nvars = 7
    cards = rng.integers(1, 5, nvars)
    d1 = list(rng.permutation(nvars)[:rng.integers(0, 6)])
    d2 = list(rng.permutation(nvars)[:rng.integers(0, 6)])
    a = rng.random([cards[i] for i in d1])
    b = rng.random([cards[i] for i in d2])
    f = NpFactor(d1, a) * NpFactor(d2, b)
    d, o = capi.factor_product(d1, a, d2, b)
    assert d == [int(x) for x in f.dimensions]
    assert np.array_equal(o, f.potential)
    if d:
        k = int(rng.integers(1, len(d) + 1))
        sd = [int(x) for x in rng.permutation(d)[:k]]
        s = f.sum(sd)
        ds, os_ = capi.factor_sum(d, o, sd)
        assert ds == [int(x) for x in s.dimensions]
        assert np.array_equal(os_, s.potential)


# ------------------------------------------------------------------------------------- ExactInference
def test_cpd_to_factor_layout(golden):
    """convert(Factor, cpd): dims [node, parents...], first parent fastest (factors_main.jl:62-68; the reference
    checks it against table() to 1e-13, test_factors.jl:44-51 -- here against pdf by enumeration)."""
    net = net_from(golden["inference_student"]["nodes"])
    g = cpd_factor(net, net.index["G"])
    assert g.dimensions == ["G", "D", "I"] and g.size == (3, 2, 2)
    rows = golden["inference_student"]["nodes"][2][2]
    for d, i in itertools.product(range(2), range(2)):
        assert g.potential[:, d, i].tolist() == rows[d + 2 * i]


def test_exact_inference_reference_posteriors(golden):
    """test/test_inference.jl:36-110 for ExactInference (exact values; the atol there is for the samplers)."""
    g = golden["inference_acb"]
    net = net_from(g["nodes"])
    np.testing.assert_allclose(exact_infer(net, "a").potential, g["P_a"], atol=1e-15)
    np.testing.assert_allclose(exact_infer(net, "c").potential, g["P_c"], atol=1e-15)
    bc = exact_infer(net, ["b", "c"])
    np.testing.assert_allclose(bc.permuted_to(["b", "c"]).ravel(order="F"), g["P_bc_colmajor"], atol=1e-15)
    s = golden["inference_student"]
    net = net_from(s["nodes"])
    np.testing.assert_allclose(exact_infer(net, "D").potential, s["P_D"], rtol=1e-14)
    np.testing.assert_allclose(exact_infer(net, "G", {"D": 0, "I": 0}).potential, s["P_G_given_d1_i1"], rtol=1e-14)
    with pytest.raises(ValueError):
        exact_infer(net, ["D", "waldo"])
    with pytest.raises(ValueError):
        exact_infer(net, "D", {"D": 1})


def xdsl_net(golden):
    cpts = golden["xdsl"]["cpts"]
    nodes = []
    for c in cpts:
        k = len(c["states"])
        p = c["probabilities"]
        nodes.append([c["id"], c["parents"][::-1], [p[i:i + k] for i in range(0, len(p), k)]])
    return net_from(nodes)


def test_xdsl_joint_and_posteriors(golden):
    """test/test_io.jl:13-18 joint pdfs; derived P(Success | Forecast) closed forms."""
    net = xdsl_net(golden)
    for s, f, p in golden["xdsl"]["joint_1based"]:
        joint = brute_force_posterior(net, ["Success", "Forecast"])
        assert joint[s - 1, f - 1] == pytest.approx(p, rel=1e-14)
    np.testing.assert_allclose(exact_infer(net, "Forecast").potential, [0.16, 0.32, 0.52], rtol=1e-14)
    np.testing.assert_allclose(exact_infer(net, "Success", {"Forecast": 0}).potential, [0.5, 0.5], rtol=1e-14)
    np.testing.assert_allclose(exact_infer(net, "Success", {"Forecast": 1}).potential, [0.25, 0.75], rtol=1e-14)
    np.testing.assert_allclose(exact_infer(net, "Success", {"Forecast": 2}).potential, [1 / 13, 12 / 13], rtol=1e-14)


def test_logpdf_table_on_xdsl_joint(golden):
    """logpdf(bn, df) restatement (ProbabilisticGraphicalModels.jl:83-99) against the six pdf(bn, ...) products of
    test/test_io.jl:13-18: a table holding those six assignments has log-likelihood sum(log(p)); single rows give
    log(p) each; a table that repeats rows adds them; zero rows give 0.0."""
    import math
    net = xdsl_net(golden)
    idx = {n: i for i, n in enumerate(net.names)}
    rows = golden["xdsl"]["joint_1based"]
    tab = np.zeros((2, len(rows)), np.uint8)
    for r, (s, f, p) in enumerate(rows):
        tab[idx["Success"], r], tab[idx["Forecast"], r] = s - 1, f - 1
        one = np.ascontiguousarray(tab[:, r:r + 1])
        assert capi.logpdf_table(net, one) == pytest.approx(math.log(p), rel=1e-14)
    want = math.fsum(math.log(p) for _, _, p in rows)
    assert capi.logpdf_table(net, tab) == pytest.approx(want, rel=1e-14)
    assert capi.logpdf_table(net, np.tile(tab, (1, 7))) == pytest.approx(7 * want, rel=1e-13)
    assert capi.logpdf_table(net, tab[:, :0], n_rows=0, pitch=16) == 0.0
    # rows marked 0xFF in their first byte are skipped, like orc_statistics
    marked = tab.copy()
    marked[0, 2] = 0xFF
    assert capi.logpdf_table(net, marked) == pytest.approx(want - math.log(rows[2][2]), rel=1e-14)


def test_logpdf_table_equals_counts_dot_log_cpt():
    """The identity the device path uses: sum over rows of log P(row) = sum_b N[b] * log(cpt[b]) over bins seen."""
    import math
    rng = np.random.default_rng(5)
    nodes = [["a", [], [[0.2, 0.5, 0.3]]], ["b", ["a"], [[0.1, 0.9], [0.6, 0.4], [0.5, 0.5]]],
             ["c", ["a", "b"], [list(rng.dirichlet(np.ones(4))) for _ in range(6)]]]
    net = net_from(nodes)
    tab, _, _ = capi.sample(net, 5000, seed=3)
    counts = capi.statistics(net, tab)
    cpt = np.asarray(net.cpt)
    want = math.fsum(int(c) * math.log(p) for c, p in zip(counts, cpt) if c)
    assert capi.logpdf_table(net, tab) == pytest.approx(want, rel=1e-12)


def test_exact_matches_enumeration_on_asia_and_student(golden):
    asia = net_from(golden["asia"]["nodes"])
    np.testing.assert_allclose(exact_infer(asia, "T", {"X": 1}).potential, [0.907589116841376, 0.092410883158624], rtol=1e-12)
    np.testing.assert_allclose(exact_infer(asia, "L", {"D": 1, "S": 1}).potential, [0.856116995114274, 0.143883004885726], rtol=1e-12)
    np.testing.assert_allclose(exact_infer(asia, "B", {"D": 1}).potential, [0.285265293007839, 0.714734706992161], rtol=1e-12)
    for q, ev in [(["T", "L"], {"D": 1}), (["A", "B", "X"], {}), (["S"], {"X": 0, "D": 1})]:
        ve = exact_infer(asia, q, ev)
        np.testing.assert_allclose(ve.permuted_to(q), brute_force_posterior(asia, q, ev), rtol=1e-12)
    st = net_from(golden["inference_student"]["nodes"])
    np.testing.assert_allclose(exact_infer(st, "G").potential, [0.447, 0.2714, 0.2816], rtol=1e-12)
    np.testing.assert_allclose(exact_infer(st, "I", {"L": 0, "S": 1}).potential,
                               [0.075174112754398, 0.924825887245602], rtol=1e-12)


# ------------------------------------------------------------------------------------- samplers
@pytest.mark.parametrize("mode", [capi.REFERENCE_SCAN, capi.DEVICE_SPEC])
def test_lw_oracle_reference_posteriors(golden, mode):
    """test/test_inference.jl:45-110 for LikelihoodWeightingInference, at the reference's tolerances but with
    far more particles than its default 500, plus a 4-sigma bound against ExactInference."""
    g = golden["inference_acb"]
    net = net_from(g["nodes"])
    c, sw, sw2 = capi.lw_infer(net, None, [0], 20000, seed=1, mode=mode)
    np.testing.assert_allclose(c / c.sum(), g["P_a"], atol=g["atol"])
    c, _, _ = capi.lw_infer(net, None, [1, 2], 20000, seed=2, mode=mode)
    np.testing.assert_allclose((c / c.sum()).ravel(order="F"), g["P_bc_colmajor"], atol=g["atol"])
    s = golden["inference_student"]
    net = net_from(s["nodes"])
    ev = np.array([0, 0, -1, -1, -1], np.int8)
    c, sw, sw2 = capi.lw_infer(net, ev, [2], 200000, seed=3, mode=mode)
    p = c / c.sum()
    n_eff = sw * sw / sw2
    exact = np.array(s["P_G_given_d1_i1"])
    assert np.all(np.abs(p - exact) < 4 * np.sqrt(exact * (1 - exact) / n_eff))
    # evidence below the query: weights matter
    ev = np.array([-1, -1, -1, 0, 1], np.int8)
    c, sw, sw2 = capi.lw_infer(net, ev, [1], 400000, seed=4, mode=mode)
    p = c / c.sum()
    exact = np.array([0.075174112754398, 0.924825887245602])
    n_eff = sw * sw / sw2
    assert np.all(np.abs(p - exact) < 4 * np.sqrt(exact * (1 - exact) / n_eff))


def test_device_spec_tracks_reference_scan(golden):
    """Quantising the CDF to 31 bits changes a draw only when u falls within 2^-32 of a CDF step."""
    net = net_from(golden["asia"]["nodes"])
    a, _, _ = capi.sample(net, 200000, kind=capi.ANCESTRAL, seed=9, mode=capi.REFERENCE_SCAN)
    b, _, _ = capi.sample(net, 200000, kind=capi.ANCESTRAL, seed=9, mode=capi.DEVICE_SPEC)
    assert (a != b).any(axis=0).mean() < 1e-5


def test_samplers_on_deterministic_net(golden):
    """test/test_sampling.jl:9-20 and test/test_gibbs.jl:100-121 (A -> C <- B with deterministic roots)."""
    net = net_from(golden["inference_acb"]["nodes"])
    st, w, ok = capi.sample(net, 64, kind=capi.ANCESTRAL, seed=5)
    assert (st[0] == 0).all() and (st[1] == 1).all() and (st[2] == 0).all()  # a=1, b=2, c=1 in 1-based terms
    ev = np.array([-1, -1, 0], np.int8)
    st, w, ok = capi.sample(net, 64, evidence=ev, kind=capi.REJECTION, seed=6)
    assert ok.all() and (st[2] == 0).all()
    ev_bad = np.array([1, -1, -1], np.int8)  # a=2 has probability 0 -> never accepted
    st, w, ok = capi.sample(net, 8, evidence=ev_bad, kind=capi.REJECTION, seed=7, max_attempts=20)
    assert not ok.any()
    st, w, ok = capi.sample(net, 64, evidence=ev, kind=capi.WEIGHTED, seed=8)
    assert (st[2] == 0).all() and np.allclose(w, 1.0)  # P(c=1 | a=1, b=2) = 1
    out = capi.gibbs_sample(net, np.array([-1, -1, -1], np.int8), 3, 5, 100, thinning=5, seed=1)
    assert out.shape == (3, 15)
    assert (out[0] == 0).all() and (out[1] == 1).all() and (out[2] == 0).all()
    out = capi.gibbs_sample(net, ev, 2, 5, 100, thinning=5, seed=2, log_domain=True)
    assert (out[0] == 0).all() and (out[1] == 1).all() and (out[2] == 0).all()


@pytest.mark.parametrize("full", [False, True])
def test_gibbs_oracle_reference_posteriors(golden, full):
    """test/test_inference.jl:45-110 for GibbsSamplingNodewise / GibbsSamplingFull."""
    g = golden["inference_acb"]
    net = net_from(g["nodes"])
    c, _ = capi.gibbs_infer(net, None, [0], 64, 2000, 500, 3, full=full, seed=1)
    np.testing.assert_allclose(c / c.sum(), g["P_a"], atol=g["atol"])
    c, _ = capi.gibbs_infer(net, None, [1, 2], 64, 2000, 500, 3, full=full, seed=2)
    np.testing.assert_allclose((c / c.sum()).ravel(order="F"), g["P_bc_colmajor"], atol=g["atol"])
    s = golden["inference_student"]
    net = net_from(s["nodes"])
    c, _ = capi.gibbs_infer(net, None, [0], 256, 2000, 500, 3, full=full, seed=3)
    np.testing.assert_allclose(c / c.sum(), s["P_D"], atol=0.02)
    ev = np.array([0, 0, -1, -1, -1], np.int8)
    c, _, pcc = capi.gibbs_infer(net, ev, [2], 256, 2000, 500, 3, full=full, seed=4, per_chain=True)
    np.testing.assert_allclose(c / c.sum(), s["P_G_given_d1_i1"], atol=0.02)
    assert np.array_equal(pcc.sum(axis=0), c.ravel(order="F"))
    kept = -(-(2000 - 500) // 3)
    assert c.sum() == 256 * kept  # ceil((nsamples-burn_in)/thin) samples per chain, gibbs.jl:74
    with pytest.raises(ValueError):
        capi.gibbs_infer(net, None, [0], 1, 100, 100, 3)  # burn_in < nsamples, gibbs.jl:70-71


def test_gibbs_state_resume(golden):
    """im.state carries the chain across calls (gibbs.jl:64,83-88): 2 x 1000 iterations == the states reached
    by seeding the second call with the first call's final states."""
    net = net_from(golden["asia"]["nodes"])
    c1, st1 = capi.gibbs_infer(net, None, [2], 8, 1000, 0, 1, seed=11)
    assert st1.shape == (8, 7)
    c2, st2 = capi.gibbs_infer(net, None, [2], 8, 1000, 0, 1, seed=12, state=st1)
    assert c2.sum() == 8 * 1000 and st2.shape == st1.shape


def test_sweep_permutation_is_permutation():
    for m in [0, 1, 2, 7, 180]:
        p = capi.sweep_permutation(3, 5, m)
        assert sorted(p.tolist()) == list(range(m))
    assert not np.array_equal(capi.sweep_permutation(3, 5, 50), capi.sweep_permutation(3, 6, 50))


# --------------------------------------------------------------------------------------------- statistics / scores
def _flat_structure(names, edges, cards):
    """structure-only flat net for the oracle: parents ascending by node index (inneighbors), nodes in `names` order."""
    from oracle.factors_np import FlatNet
    idx = {n: i for i, n in enumerate(names)}
    plist = [sorted(idx[p] for p, c in edges if c == n) for n in names]
    par_ptr, par_idx, cpt_off = [0], [], [0]
    for i, ps in enumerate(plist):
        par_idx += ps
        par_ptr.append(len(par_idx))
        cpt_off.append(cpt_off[-1] + cards[i] * int(np.prod([cards[p] for p in ps])) if ps else cpt_off[-1] + cards[i])
    return FlatNet(list(names), np.array(cards, np.int32), np.array(par_ptr, np.int32), np.array(par_idx, np.int32),
                   np.array(cpt_off, np.int64), np.zeros(cpt_off[-1]), idx)


def test_statistics_and_bayesian_scores_match_reference_tests(golden):
    """test/test_discrete_bayes_nets.jl:40-44 (five score values), test/test_learning.jl:87-96 (count matrices)."""
    from oracle import capi
    from oracle.scoring_py import bayesian_score_component, bdeu_alpha, split_counts, uniform_alpha
    g = golden["statistics_ab"]
    net = _flat_structure(g["names"], g["edges"], [2, 2])
    data = np.array([g["data"][n] for n in g["names"]], np.uint8) - 1
    Ns = split_counts(net.card, net.par_ptr, net.par_idx, net.cpt_off, capi.statistics(net, data))
    assert Ns[0].ravel().tolist() == g["count_a"]
    for key, cnt in g["count_b_given_a_1based"].items():
        a, b = (int(x) for x in key.split(","))
        assert Ns[1][b - 1, a - 1] == cnt
    assert np.isclose(bayesian_score_component(Ns[0], uniform_alpha(2, 1)), g["score_1_uniform"], rtol=1e-13)
    assert np.isclose(bayesian_score_component(Ns[0], uniform_alpha(2, 1, 2.0)), g["score_1_uniform2"], rtol=1e-13)
    assert np.isclose(bayesian_score_component(Ns[0], bdeu_alpha(2, 1)), g["score_1_bdeu"], rtol=1e-13)
    assert np.isclose(bayesian_score_component(Ns[0], bdeu_alpha(2, 1, 2.0)), g["score_1_bdeu2"], rtol=1e-13)
    assert np.isclose(bayesian_score_component(Ns[1], uniform_alpha(2, 2)), g["score_2_uniform"], rtol=1e-13)
    g = golden["statistics_abc"]
    net = _flat_structure(g["names"], g["edges"], [3, 2, 2])
    data = np.array([g["data"][n] for n in g["names"]], np.uint8) - 1
    Ns = split_counts(net.card, net.par_ptr, net.par_idx, net.cpt_off, capi.statistics(net, data))
    assert Ns[0].tolist() == g["stats_A"]
    assert Ns[2].tolist() == g["stats_C"]
    # skipped rows (0xFF marker) and a padded pitch
    padded = np.full((3, 16), 0, np.uint8)
    padded[:, :12] = data
    padded[:, 5] = 0xFF
    c = capi.statistics(net, padded, n_rows=12, pitch=16)
    keep = np.delete(data, 5, axis=1)
    assert np.array_equal(c, capi.statistics(net, keep))


# ------------------------------------------------------------------------------------- LoopyBelief
def test_loopy_oracle_reference_posteriors(golden):
    """test/test_inference.jl:45-63,94-110 with im = LoopyBelief(): the four posteriors the reference pins for it
    (atol 0.02 / 0.05 there; the restatement lands on them to 1e-12 because both nets are polytrees whose
    unobserved roots have at most uniform lambdas coming back)."""
    from oracle.loopy_np import loopy_infer
    g = golden["inference_acb"]
    net = net_from(g["nodes"])
    np.testing.assert_allclose(loopy_infer(net, net.index["a"]), g["P_a"], atol=1e-12)
    np.testing.assert_allclose(loopy_infer(net, net.index["c"]), g["P_c"], atol=1e-12)
    s = golden["inference_student"]
    net = net_from(s["nodes"])
    np.testing.assert_allclose(loopy_infer(net, net.index["D"]), s["P_D"], rtol=1e-12)
    ev = np.array([0, 0, -1, -1, -1], np.int8)
    np.testing.assert_allclose(loopy_infer(net, net.index["G"], ev), s["P_G_given_d1_i1"], rtol=1e-12)
    # without evidence belief propagation is exact on a polytree: every marginal of the Student net
    for name in net.names:
        np.testing.assert_allclose(loopy_infer(net, net.index[name]), exact_infer(net, name).potential, rtol=1e-12)


def test_loopy_oracle_stopping_rule_and_aliasing(golden):
    """Early stop needs iters_for_convergence quiet iterations (loopy_belief.jl:94-96,192-198); tol < 0 never stops;
    the in-place update of a root's prior (:125,:170-175) is visible once evidence sits below a root with two
    children: the reference's answer is NOT the exact posterior there, and the oracle reproduces that."""
    from oracle.loopy_np import loopy_infer
    net = net_from(golden["inference_student"]["nodes"])
    _, it = loopy_infer(net, 0, return_iters=True)
    assert 6 <= it < 20
    _, it = loopy_infer(net, 0, iters_for_convergence=1, return_iters=True)
    assert it < 8
    _, it = loopy_infer(net, 0, nsamples=37, tol=-1.0, return_iters=True)
    assert it == 37
    out0 = loopy_infer(net, 1, nsamples=0)                       # no iteration: prior x uniform lambdas
    np.testing.assert_allclose(out0, [0.7, 0.3], rtol=1e-15)
    ev = np.array([-1, -1, -1, 0, 1], np.int8)                   # L = 1, S = 2 (1-based)
    lbp = loopy_infer(net, net.index["I"], ev)
    exact = exact_infer(net, "I", {"L": 0, "S": 1}).potential
    assert abs(lbp[0] - exact[0]) > 0.05 and lbp[0] < 1e-12      # the root's prior was multiplied in every iteration


# ------------------------------------------------------------------------------------- the rest of the Factor algebra
def test_broadcast_reduce_normalize_dims_golden():
    """test/test_factors.jl:87-105 (broadcast), :111-130 (reducedim after broadcast), :71-81 (normalize)."""
    phi = NpFactor(["X", "Y"], np.array([[1.0, 2], [3, 4], [5, 6]]))
    got = phi.broadcast("*", ["Y", "X"], [[10, 0.1], 100.0]).potential
    np.testing.assert_allclose(got, [[1000, 20], [3000, 40], [5000, 60]], rtol=1e-15)           # :90-92
    with pytest.raises(ValueError):
        phi.broadcast("*", ["X", "Z"], [[10, 1, 0.1], [1, 2, 3]])                               # :94
    with pytest.raises(ValueError):
        phi.broadcast("*", ["Z", "X", "A"], [2, [10, 1, 0.1], [1, 2, 3]])                       # :96
    with pytest.raises(DimensionMismatch):
        phi.broadcast("*", "X", [2016, 58.0])                                                   # :98
    p3 = NpFactor(["X", "Y", "Z"], np.array([1.0, 2, 3, 2, 3, 4, 4, 6, 7, 8, 10, 16]).reshape((3, 2, 2), order="F"))
    two = p3.broadcast("+", "Z", [10, 0.1]).broadcast("+", "X", 10.0)
    one = p3.broadcast("+", ["X", "Z"], [10.0, [10, 0.1]])
    assert np.array_equal(two.potential, one.potential)                                         # :104-105
    red = p3.broadcast("*", "Z", [1, 10.0]).reducedim("+", ["Y", "Z"])
    assert red.dimensions == ["X"] and np.array_equal(red.potential, [123.0, 165.0, 237.0])     # :121-129
    assert np.array_equal(p3.broadcast("*", "Z", 0.0).reducedim("+", ["X", "Y", "Z"]).potential, np.array(0.0))  # :119
    with pytest.raises(ValueError):
        p3.reducedim("*", "waldo")                                                              # :113
    n = NpFactor(["a", "b"], np.array([[1.0, 2], [3, 4]]))
    np.testing.assert_allclose(n.normalize_dims(["a", "b"], p=1).potential, [[0.1, 0.2], [0.3, 0.4]], rtol=1e-15)
    np.testing.assert_allclose(n.normalize_dims(["a", "b"], p=2).potential, [[1 / 30, 2 / 30], [1 / 10, 4 / 30]], rtol=1e-15)
    np.testing.assert_allclose(n.normalize_dims("a").potential, [[0.25, 2 / 6], [0.75, 4 / 6]], rtol=1e-15)
    with pytest.raises(ValueError):
        n.normalize_dims("waldo")                                                               # :77


def test_join_ops_follow_the_reference_swap():
    """join(op, phi1, phi2) swaps to larger-first before applying op (factors_dims.jl:189-191): a / b with a smaller
    than b is b ./ a; `*` agrees with the pinned product."""
    rng = np.random.default_rng(0)
    a = NpFactor(["x"], rng.random(3) + 0.5)
    b = NpFactor(["y", "x"], rng.random((2, 3)) + 0.5)
    assert np.array_equal(a.join("*", b).potential, (a * b).potential)
    d = a.join("/", b)
    assert d.dimensions == ["y", "x"]
    np.testing.assert_array_equal(d.potential, b.potential / a.potential[None, :])
    np.testing.assert_array_equal(b.join("-", a).potential, b.potential - a.potential[None, :])
    np.testing.assert_array_equal(a.join("-", b).potential, b.potential - a.potential[None, :])  # the swap
    np.testing.assert_array_equal(a.join("+", b).potential, b.potential + a.potential[None, :])


@pytest.mark.parametrize("seed,n_nodes,kw", [(0, 12, dict()), (1, 30, dict(nsamples=40)), (2, 60, dict(nsamples=15, tol=-1.0)),
                                             (3, 8, dict(iters_for_convergence=1, tol=1e-3)), (4, 100, dict(nsamples=6))])
def test_loopy_c_and_numpy_restatements_agree(seed, n_nodes, kw):
    """The numpy oracle keeps the reference's aliasing literally, the C one states its consequences: bit-identical
    beliefs and iteration counts on random nets and evidence (incl. impossible evidence -> NaN)."""
    import bayesnets_jl_b200 as bn
    from conftest import oracle_net_from_layout
    from oracle.loopy_np import loopy_infer
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=seed)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    rng = np.random.default_rng(seed)
    B = 6
    ev = np.full((B, lay.n), -1, np.int8)
    q = np.zeros(B, np.int32)
    for b in range(B):
        picks = rng.choice(lay.n, size=min(lay.n, 1 + (b * n_nodes) // 12), replace=False)
        q[b] = picks[0]
        for i in picks[1:]:
            ev[b, i] = rng.integers(0, lay.card[i])
    out, its = capi.loopy_infer(on, q, ev, **kw)
    for b in range(B):
        want, it = loopy_infer(on, int(q[b]), ev[b], return_iters=True, **kw)
        assert it == its[b]
        assert np.array_equal(out[b, :len(want)], want, equal_nan=True), (b, out[b], want)
        assert not out[b, len(want):].any()
