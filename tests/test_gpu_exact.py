"""GPU parity: infer(ExactInference(), ...) -- variable elimination on device factors -- vs the oracle.

Bar (north_star): exact posteriors within 1e-12 relative of the reference's ExactInference (here: the numpy
restatement of src/Inference/exact.jl and brute-force enumeration).
"""
import numpy as np
import pytest

from conftest import make_mirror_net, oracle_net_from_layout
from oracle.factors_np import brute_force_posterior, exact_infer
from test_gpu_sampling import random_case

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def by_name(phi, query):
    return np.transpose(phi.potential, [phi.dimensions.index(q) for q in query])


@pytest.mark.parametrize("fused", [True, False])
def test_exact_reference_suite(bn, golden, fused):
    """test/test_inference.jl:45-110 for ExactInference, plus the derived closed forms of SURVEY.md 8c."""
    im = bn.ExactInference(fused=fused)
    g = golden["inference_acb"]
    acb = make_mirror_net(bn, g["nodes"])
    phi = bn.infer(im, acb, "a")
    assert len(phi) == 2
    assert float(phi[{"a": 1}].potential) == pytest.approx(1.0, abs=1e-15)
    assert float(phi[{"a": 2}].potential) == pytest.approx(0.0, abs=1e-15)
    np.testing.assert_allclose(bn.infer(im, acb, "c").potential, g["P_c"], atol=1e-15)
    bc = bn.infer(im, acb, ["b", "c"])
    assert bc.size() == (2, 2)
    np.testing.assert_allclose(by_name(bc, ["b", "c"]).ravel(order="F"), g["P_bc_colmajor"], atol=1e-15)
    s = golden["inference_student"]
    stu = make_mirror_net(bn, s["nodes"])
    np.testing.assert_allclose(bn.infer(im, stu, "D").potential, s["P_D"], rtol=RTOL)
    np.testing.assert_allclose(bn.infer(im, stu, "G", evidence={"D": 1, "I": 1}).potential, s["P_G_given_d1_i1"], rtol=RTOL)
    np.testing.assert_allclose(bn.infer(stu, "G").potential, [0.447, 0.2714, 0.2816], rtol=RTOL)  # default method
    np.testing.assert_allclose(bn.infer(im, stu, "I", evidence={"L": 1, "S": 2}).potential,
                               [0.075174112754398, 0.924825887245602], rtol=RTOL)
    asia = bn.get_asia_bn()
    np.testing.assert_allclose(bn.infer(im, asia, "T", evidence={"X": 2}).potential, [0.907589116841376, 0.092410883158624], rtol=RTOL)
    np.testing.assert_allclose(bn.infer(im, asia, "L", evidence={"D": 2, "S": 2}).potential, [0.856116995114274, 0.143883004885726], rtol=RTOL)
    np.testing.assert_allclose(bn.infer(im, asia, "B", evidence={"D": 2}).potential, [0.285265293007839, 0.714734706992161], rtol=RTOL)
    with pytest.raises(ValueError):
        bn.infer(im, asia, ["T", "waldo"])
    with pytest.raises(ValueError):
        bn.infer(im, asia, "T", evidence={"T": 1})


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("seed,n_nodes,n_ev,n_q", [(0, 12, 2, 2), (1, 11, 0, 3), (2, 10, 4, 1), (3, 16, 3, 2)])
def test_exact_vs_enumeration_and_oracle_ve(bn, fused, seed, n_nodes, n_ev, n_q):
    net, lay, on, query, evidence = random_case(bn, n_nodes, seed, n_ev, n_q)
    phi = bn.infer(bn.ExactInference(fused=fused), net, query, evidence=evidence)
    ev0 = {k: v - 1 for k, v in evidence.items()}
    np.testing.assert_allclose(by_name(phi, query), brute_force_posterior(on, query, ev0), rtol=RTOL)
    np.testing.assert_allclose(by_name(phi, query), exact_infer(on, query, ev0).permuted_to(query), rtol=RTOL)
    assert abs(phi.potential.sum() - 1.0) < 1e-13


def test_exact_100_node_net_vs_oracle_ve(bn):
    """The C2 network itself (100 nodes, 10 evidence): intermediates reach ~8e7 entries."""
    net, lay, on, query, evidence = random_case(bn, 100, 0, 10)
    ev0 = {k: v - 1 for k, v in evidence.items()}
    ref = exact_infer(on, query, ev0).permuted_to(query)
    for fused in (True, False):
        phi = bn.infer(bn.ExactInference(fused=fused), net, query, evidence=evidence)
        np.testing.assert_allclose(by_name(phi, query), ref, rtol=RTOL)


def binary_chain_with_hub(bn, n_parents, seed=0):
    """C5-style net: binary nodes, one node with `n_parents` parents (CPT = 2^(n_parents+1) entries)."""
    rng = np.random.default_rng(seed)
    net = bn.BayesNet()
    roots = []
    for i in range(n_parents):
        ps = [roots[-1]] if roots and i % 3 == 0 else []
        net.push(bn.rand_cpd(net, 2, f"R{i}", ps, rng=rng))
        roots.append(f"R{i}")
    q = 1 << n_parents
    p0 = rng.random(q)
    net.push(bn.CategoricalCPD("HUB", roots, [2] * n_parents, list(np.stack([p0, 1 - p0], axis=1))))
    net.push(bn.rand_cpd(net, 2, "LEAF", ["HUB", roots[0]], rng=rng))
    return net


def test_exact_hub_network_vs_oracle(bn):
    net = binary_chain_with_hub(bn, 16)
    on = oracle_net_from_layout(bn.flatten(net))
    for query, ev in [(["HUB"], {"LEAF": 2}), (["R3", "R9"], {"HUB": 1}), (["LEAF"], {})]:
        ref = exact_infer(on, query, {k: v - 1 for k, v in ev.items()}).permuted_to(query)
        for fused in (True, False):
            phi = bn.infer(bn.ExactInference(fused=fused), net, query, evidence=ev)
            np.testing.assert_allclose(by_name(phi, query), ref, rtol=RTOL)


def test_factor_from_bn_matches_cpd_layout(bn, golden):
    """Factor(bn, name[, evidence]): test/test_factors.jl:44-51 (vs pdf) and factors_main.jl:79-83."""
    net = bn.rand_discrete_bn(10, 4, seed=5)
    name = "N5"
    phi = bn.Factor.from_bn(net, name)
    cpd = net.get(name)
    assert phi.dimensions == [name] + cpd.parents
    pat = phi.pattern()
    pot = phi.potential.ravel(order="F")
    for row, p in zip(pat, pot):
        a = {d: int(v) for d, v in zip(phi.dimensions, row)}
        assert abs(cpd.pdf(a) - p) < 1e-13
    if cpd.parents:
        ev = {cpd.parents[0]: 1, "unrelated": 3}
        sl = bn.Factor.from_bn(net, name, ev)
        assert sl.dimensions == [name] + cpd.parents[1:]
        assert np.array_equal(sl.potential, np.take(phi.potential, 0, axis=1))


def test_exact_c5_at_its_stated_size(bn):
    """BASELINE configs[4] at full size: ExactInference on a 30-node binary net whose hub has 23 parents (a 2^24-entry
    CPT, 128 MiB) -- src/Inference/exact.jl:6-33.  The numpy oracle materialises every product (minutes at this size),
    so here: fused == the reference's pairwise product-then-sum to 1e-12, the posterior is a distribution, evidence
    on a root changes it the way Bayes' rule says (P(q) = sum_e P(q | e) P(e)); the oracle itself is compared on the
    same construction with 19 parents (2^20 entries)."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import path_bench
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import DeviceNet

    def run(lay, dn, ev, q, fused):
        o, od, oc = np.zeros(2), np.zeros(1, np.int32), np.zeros(1, np.int32)
        _lib.check(_lib.lib().bn_exact_infer(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p), C.c_int32(1),
                                             C.c_int32(fused), _lib.ptr(od, _lib.i32p), _lib.ptr(oc, _lib.i32p),
                                             _lib.ptr(o, _lib.f64p)))
        assert int(od[0]) == int(q[0]) and int(oc[0]) == 2
        return o

    lay = path_bench.c5_layout(bn, k=23)
    assert lay.n == 30 and int(lay.cpt_off[-1] - lay.cpt_off[-2]) == 1 << 24
    dn = DeviceNet(lay, 0)
    q = np.array([lay.n - 1], np.int32)
    ev = np.full(lay.n, -1, np.int8)
    ev[3] = 1
    pf, pu = run(lay, dn, ev, q, 1), run(lay, dn, ev, q, 0)
    np.testing.assert_allclose(pf, pu, rtol=RTOL)
    assert abs(pf.sum() - 1.0) < 1e-12 and (pf > 0).all()
    # total probability over the evidence node: P(hub) = sum_s P(hub | B4 = s) P(B4 = s)
    none = np.full(lay.n, -1, np.int8)
    prior_q = run(lay, dn, none, q, 1)
    prior_e = run(lay, dn, none, np.array([3], np.int32), 1)
    ev0 = none.copy()
    ev0[3] = 0
    mix = prior_e[0] * run(lay, dn, ev0, q, 1) + prior_e[1] * pf
    np.testing.assert_allclose(mix, prior_q, rtol=1e-11)
    dn.close()
    # the oracle on the 2^20 sub-case
    lay_s = path_bench.c5_layout(bn, k=19)
    dn_s = DeviceNet(lay_s, 0)
    got = run(lay_s, dn_s, ev, np.array([lay_s.n - 1], np.int32), 1)
    ref = exact_infer(oracle_net_from_layout(lay_s), [lay_s.names[-1]], {lay_s.names[3]: 1}).permuted_to([lay_s.names[-1]])
    np.testing.assert_allclose(got, ref, rtol=RTOL)
    dn_s.close()
