"""GPU parity: likelihood weighting and the table samplers through the C ABI.

Two bars (north_star):
  * vs the oracle's DEVICE_SPEC mode on the same seeds: per-particle states and weights are bit-identical, so
    unweighted counts match exactly and weighted sums to 1e-12 (fp64 addition order differs);
  * vs ExactInference: every posterior cell within 4 sigma, sigma^2 = p(1-p)/N_eff, N_eff = (sum w)^2/sum w^2.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import make_mirror_net, oracle_net_from_layout
from oracle import capi
from oracle.factors_np import exact_infer

pytestmark = pytest.mark.gpu


def lw_raw(bn, net, query, evidence, n, seed=0, first=0):
    im = bn.LikelihoodWeightingInference(nsamples=n, seed=seed, first_particle=first)
    inf = bn.InferenceState(net, query, evidence)
    counts = im.raw_counts(inf)
    return counts, im.last_stats


@pytest.fixture
def lw_mode(bn):
    yield bn.set_lw_mode
    bn.set_lw_mode("auto", True)


def random_case(bn, n_nodes, seed, n_ev, n_q=2):
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=seed)
    lay = bn.flatten(net)
    rng = np.random.default_rng(1000 + seed)
    while True:
        picks = [str(x) for x in rng.choice(lay.names, size=n_q + n_ev, replace=False)]
        query, evn = picks[:n_q], picks[n_q:]
        evidence = {e: int(rng.integers(1, net.ncategories(e) + 1)) for e in evn}
        on = oracle_net_from_layout(lay)
        c, sw, _ = capi.lw_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), 2000, seed=1)
        if sw > 0:  # reject evidence with P(e) = 0
            return net, lay, on, query, evidence


@pytest.mark.parametrize("seed,n_nodes,n_ev", [(0, 100, 10), (1, 100, 10), (2, 37, 5), (3, 200, 20), (4, 8, 0), (5, 12, 11)])
def test_lw_matches_oracle_device_spec(bn, seed, n_nodes, n_ev):
    n_q = 1 if n_nodes - n_ev < 2 else 2
    net, lay, on, query, evidence = random_case(bn, n_nodes, seed, n_ev, n_q)
    n = 200_000
    counts, (sw, sw2) = lw_raw(bn, net, query, evidence, n, seed=seed)
    oc, osw, osw2 = capi.lw_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), n, seed=seed,
                                  mode=capi.DEVICE_SPEC)
    np.testing.assert_allclose(counts, oc.ravel(order="F"), rtol=1e-12, atol=0)
    assert sw == pytest.approx(osw, rel=1e-12) and sw2 == pytest.approx(osw2, rel=1e-12)
    if not evidence:  # all weights are 1: counts are integers and must match exactly
        assert np.array_equal(counts, oc.ravel(order="F")) and sw == n


def test_lw_shards_reproduce_single_stream(bn):
    """Particle ids are global Philox counters: 3 ragged shards == one call."""
    net, lay, on, query, evidence = random_case(bn, 100, 0, 10)
    n = 300_001
    whole, (sw, _) = lw_raw(bn, net, query, evidence, n, seed=5)
    parts = np.zeros_like(whole)
    psw = 0.0
    for first, cnt in [(0, 100_000), (100_000, 7), (100_007, n - 100_007)]:
        c, (s, _) = lw_raw(bn, net, query, evidence, cnt, seed=5, first=first)
        parts += c
        psw += s
    np.testing.assert_allclose(parts, whole, rtol=1e-12)
    assert psw == pytest.approx(sw, rel=1e-12)
    empty, (s0, _) = lw_raw(bn, net, query, evidence, 0, seed=5)
    assert not empty.any() and s0 == 0.0


def test_lw_posterior_within_4_sigma_of_exact(bn):
    """C2-shaped case: 100 nodes, 10 evidence, 2 query; 2e7 particles; ExactInference (oracle VE) as truth."""
    net, lay, on, query, evidence = random_case(bn, 100, 0, 10)
    n = 20_000_000
    counts, (sw, sw2) = lw_raw(bn, net, query, evidence, n, seed=0)
    p = counts / counts.sum()
    n_eff = sw * sw / sw2
    ev0 = {k: v - 1 for k, v in evidence.items()}
    exact = exact_infer(on, query, ev0).permuted_to(query).ravel(order="F")
    sigma = np.sqrt(np.maximum(exact * (1 - exact), 1e-12) / n_eff)
    assert n_eff > 1000
    assert np.all(np.abs(p - exact) < 4 * sigma), (p, exact, sigma, n_eff)
    # and through the public API
    phi = bn.infer(bn.LikelihoodWeightingInference(nsamples=n, seed=0), net, query, evidence=evidence)
    assert phi.dimensions == query
    np.testing.assert_allclose(phi.potential.ravel(order="F"), p, rtol=1e-12)


def test_lw_reference_suite(bn, golden):
    """test/test_inference.jl:45-110 for LikelihoodWeightingInference (default 500 samples, reference atol)."""
    g = golden["inference_acb"]
    net = make_mirror_net(bn, g["nodes"])
    phi = bn.infer(bn.LikelihoodWeightingInference(), net, "a")
    assert len(phi) == 2
    assert float(phi[{"a": 1}].potential) == pytest.approx(1.0, abs=g["atol"])
    assert float(phi[{"a": 2}].potential) == pytest.approx(0.0, abs=g["atol"])
    phi = bn.infer(bn.LikelihoodWeightingInference(), net, "c")
    assert float(phi[{"c": 1}].potential) == pytest.approx(1.0, abs=g["atol"])
    phi = bn.infer(bn.LikelihoodWeightingInference(), net, ["b", "c"])
    assert phi.size() == (2, 2)
    np.testing.assert_allclose(phi.potential.ravel(order="F"), g["P_bc_colmajor"], atol=g["atol"])
    s = golden["inference_student"]
    net = make_mirror_net(bn, s["nodes"])
    phi = bn.infer(bn.LikelihoodWeightingInference(nsamples=20000, seed=3), net, "D")
    np.testing.assert_allclose(phi.potential, s["P_D"], atol=s["atol"])
    phi = bn.infer(bn.LikelihoodWeightingInference(nsamples=20000, seed=4), net, "G", evidence={"D": 1, "I": 1})
    assert phi.size() == (3,)
    np.testing.assert_allclose(phi.potential, s["P_G_given_d1_i1"], atol=s["atol"])
    # impossible evidence: every weight is 0 -> 0/0 = NaN, as likelihood.jl:56 does
    acb = make_mirror_net(bn, g["nodes"])
    phi = bn.infer(bn.LikelihoodWeightingInference(100), acb, "c", evidence={"a": 2})
    assert np.isnan(phi.potential).all()


def test_xdsl_config_c1(bn, golden, tmp_path):
    """BASELINE config C1: sample_bn.xdsl, LW 100k samples, 1 evidence var, vs ExactInference."""
    from test_host import write_xdsl
    path = str(tmp_path / "sample_bn.xdsl")
    write_xdsl(path, golden)
    net = bn.readxdsl(path)
    for f, exact in [(1, [0.5, 0.5]), (2, [0.25, 0.75]), (3, [1 / 13, 12 / 13])]:
        im = bn.LikelihoodWeightingInference(nsamples=100_000, seed=f)
        phi = bn.infer(im, net, "Success", evidence={"Forecast": f})
        sw, sw2 = im.last_stats
        ex = np.array(exact)
        assert np.all(np.abs(phi.potential - ex) < 4 * np.sqrt(ex * (1 - ex) / (sw * sw / sw2)))
        np.testing.assert_allclose(bn.infer(bn.ExactInference(), net, "Success", evidence={"Forecast": f}).potential,
                                   ex, rtol=1e-12)


@pytest.mark.parametrize("mode", ["generic", "specialized"])
@pytest.mark.parametrize("kind", ["ancestral", "weighted", "rejection"])
def test_sample_tables_match_oracle(bn, lw_mode, kind, mode):
    """rand(bn, n) / weighted rows / rejection rows from the interpreter kernel and from the network-specialised
    NVRTC kernel are the same bytes as the oracle's."""
    net, lay, on, query, evidence = random_case(bn, 40, 7, 2)
    from bayesnets_jl_b200.sampling import _sample
    lw_mode(mode, True)
    n = 50_001
    k = {"ancestral": 0, "rejection": 1, "weighted": 2}[kind]
    _, st, w, ok = _sample(net, n, evidence, k, 100, 21, None)
    ost, ow, ook = capi.sample(on, n, evidence=lay.evidence_vector(evidence), kind=k, seed=21, max_attempts=100)
    assert np.array_equal(ok, ook)
    assert np.array_equal(st, ost)
    if kind == "weighted":
        assert np.array_equal(w, ow)  # per-particle fp64 products are bit-identical
    if kind == "rejection":
        evv = lay.evidence_vector(evidence)
        good = ok.astype(bool)
        assert good.any()
        for i in np.nonzero(evv >= 0)[0]:
            assert (st[i, good] == evv[i]).all()
        assert (st[:, ~good] == 0xFF).all()


def test_rand_reference_suite(bn, golden):
    """test/test_sampling.jl:1-27 on A -> C <- B."""
    net = make_mirror_net(bn, golden["inference_acb"]["nodes"])
    assert bn.rand(net) == {"a": 1, "b": 2, "c": 1}
    t1 = bn.rand(net, 5)
    assert t1.shape == (5, 3) and t1["a"].tolist() == [1] * 5 and t1["b"].tolist() == [2] * 5
    t2 = bn.rand(net, 5, {"c": 1})
    assert t2.shape == (5, 3) and t2["c"].tolist() == [1] * 5
    t3 = bn.rand(net, 5, c=1, b=2)
    assert t3.shape == (5, 3)
    t4 = bn.rand(net, bn.LikelihoodWeightedSampler({"c": 1}), 5)
    assert t4.shape == (5, 4) and t4["potential"].sum() == pytest.approx(1.0)
    with pytest.raises(KeyError):
        bn.rand(net, 3, {"a": 2})  # P(a=2) = 0: no consistent sample in 100 tries
    a, w = bn.get_weighted_sample(net, {"c": 1})
    assert a == {"a": 1, "b": 2, "c": 1} and w == 1.0


def test_ancestral_marginals_and_ragged_sizes(bn):
    net = bn.get_asia_bn()
    for n in [1, 31, 1025]:
        assert bn.rand(net, n, seed=n).shape == (n, 7)
    df = bn.rand(net, 400_000, seed=2)
    p_s = (df["S"] == 2).mean()
    assert abs(p_s - 0.5) < 4 * np.sqrt(0.25 / 400_000)
    p_l = (df["L"] == 2).mean()  # 0.5*0.01 + 0.5*0.1
    assert abs(p_l - 0.055) < 4 * np.sqrt(0.055 * 0.945 / 400_000)


def test_error_paths(bn):
    net = bn.get_asia_bn()
    with pytest.raises(ValueError):
        bn.infer(bn.LikelihoodWeightingInference(10), net, "nope")
    with pytest.raises(ValueError):
        bn.infer(bn.LikelihoodWeightingInference(10), net, "T", evidence={"T": 1})
    with pytest.raises(IndexError):
        bn.infer(bn.LikelihoodWeightingInference(10), net, "T", evidence={"A": 3})


@pytest.mark.parametrize("mode,prune", [("generic", True), ("specialized", True), ("specialized", False)])
@pytest.mark.parametrize("seed,n_nodes,n_ev,n_q", [(0, 100, 10, 2), (3, 200, 20, 2), (6, 60, 6, 4), (7, 30, 0, 1)])
def test_lw_kernel_variants_equal_oracle(bn, lw_mode, mode, prune, seed, n_nodes, n_ev, n_q):
    """The generic interpreter, the NVRTC network-specialised kernel with barren-node pruning and the same kernel
    with every node kept alive all implement ONE stream spec: counts equal the oracle's (n_q = 4 -> up to 256
    query cells exercises the shared-memory-atomic accumulation path, <= 16 cells the register path)."""
    net, lay, on, query, evidence = random_case(bn, n_nodes, seed, n_ev, n_q)
    lw_mode(mode, prune)
    n = 150_001
    launches = bn.launch_count()
    counts, (sw, sw2) = lw_raw(bn, net, query, evidence, n, seed=seed + 11, first=12345)
    assert bn.launch_count() == launches + 1
    oc, osw, osw2 = capi.lw_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), n, seed=seed + 11,
                                  mode=capi.DEVICE_SPEC, first=12345)
    np.testing.assert_allclose(counts, oc.ravel(order="F"), rtol=1e-12, atol=0)
    assert sw == pytest.approx(osw, rel=1e-12) and sw2 == pytest.approx(osw2, rel=1e-12)
    if not evidence:
        assert np.array_equal(counts, oc.ravel(order="F"))


def test_lw_specialized_kernel_is_cached_per_evidence_mask(bn, lw_mode):
    """Evidence VALUES are run-time data: changing them must reuse the compiled kernel (fast second call) and still
    be exact; changing the evidence MASK compiles a new one."""
    import time
    net, lay, on, query, evidence = random_case(bn, 100, 0, 10)
    lw_mode("specialized", True)
    lw_raw(bn, net, query, evidence, 1000, seed=1)
    ev2 = dict(evidence)
    k = next(iter(ev2))
    ev2[k] = ev2[k] % net.ncategories(k) + 1
    t0 = time.perf_counter()
    counts, (sw, _) = lw_raw(bn, net, query, ev2, 50_000, seed=2)
    assert time.perf_counter() - t0 < 0.2  # no NVRTC compile on this call
    oc, osw, _ = capi.lw_infer(on, lay.evidence_vector(ev2), lay.query_indices(query), 50_000, seed=2,
                               mode=capi.DEVICE_SPEC)
    np.testing.assert_allclose(counts, oc.ravel(order="F"), rtol=1e-12, atol=1e-300)
    assert sw == pytest.approx(osw, rel=1e-12)


@pytest.mark.parametrize("mode", ["generic", "specialized"])
def test_edge_cardinalities_and_deterministic_tables(bn, lw_mode, mode):
    """Single-state nodes, a 7-state node (two threshold words per row), deterministic CPT rows (thresholds 0 and
    2^31, i.e. 'always' / 'never') and impossible evidence (every weight 0 -> NaN posterior like the reference's 0/0)."""
    rng = np.random.default_rng(11)
    net = bn.BayesNet()
    net.push(bn.CategoricalCPD("one", [], [], [[1.0]]))
    net.push(bn.CategoricalCPD("det", ["one"], [1], [[0.0, 1.0, 0.0]]))
    net.push(bn.CategoricalCPD("wide", ["det"], [3], [rng.dirichlet(np.ones(7)) for _ in range(3)]))
    net.push(bn.CategoricalCPD("leaf", ["wide", "det"], [7, 3], [rng.dirichlet(np.ones(2)) for _ in range(21)]))
    net.push(bn.CategoricalCPD("never", ["leaf"], [2], [[1.0, 0.0], [1.0, 0.0]]))
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    lw_mode(mode, True)
    n = 40_000
    for query, evidence in [(["wide"], {"leaf": 2}), (["wide", "leaf"], {}), (["det", "one"], {"wide": 4})]:
        counts, (sw, sw2) = lw_raw(bn, net, query, evidence, n, seed=4)
        oc, osw, osw2 = capi.lw_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), n, seed=4)
        np.testing.assert_allclose(counts, oc.ravel(order="F"), rtol=1e-12)
        assert sw == pytest.approx(osw, rel=1e-12)
    # det is always state 2: P(det) = [0, 1, 0] exactly
    c, _ = lw_raw(bn, net, ["det"], {}, n, seed=5)
    assert c.tolist() == [0.0, float(n), 0.0]
    # impossible evidence: all weights are 0 -> normalize! gives NaN, as likelihood.jl:56 would
    phi = bn.infer(bn.LikelihoodWeightingInference(nsamples=1000, seed=1), net, "wide", evidence={"never": 2})
    assert np.isnan(phi.potential).all()
    # tables from both kernels are identical to the oracle's
    from bayesnets_jl_b200.sampling import _sample
    _, st, w, ok = _sample(net, 5000, {"leaf": 1}, 2, 1, 9, None)
    ost, ow, _ = capi.sample(on, 5000, evidence=lay.evidence_vector({"leaf": 1}), kind=2, seed=9)
    assert np.array_equal(st, ost) and np.array_equal(w, ow)


def test_lw_mc_bound_at_benchmark_scale(bn, lw_mode):
    """north_star: "posteriors inside the stated MC bound" at the benchmark's size, not at a toy size: the C2 case of
    bench.py (100 nodes, 10 evidence, 2 query, every node sampled), 4 calls x 1e9 particles with distinct seeds pooled
    into one 4e9-particle estimate; every cell within 4 sigma (sigma^2 = p(1-p)/N_eff) of ExactInference on the device.
    31-bit thresholds bias a probability by <= 2^-32, far below 4 sigma ~ 3e-5 here."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    def check(net, lay, query, evidence):
        im = bn.LikelihoodWeightingInference(nsamples=4096, seed=1)
        im.raw_counts(bn.InferenceState(net, query, evidence))
        return im.last_stats[0] > 0

    net, lay, query, evidence = bench.build_case(bn, check=check)
    lw_mode("specialized", False)
    c, sw, sw2 = 0.0, 0.0, 0.0
    for seed in range(4):
        counts, (a, b) = lw_raw(bn, net, query, evidence, 1_000_000_000, seed=seed)
        c, sw, sw2 = c + counts, sw + a, sw2 + b
    p = c / c.sum()
    exact = bench.exact_posterior(bn, net, query, evidence, 0)
    n_eff = sw * sw / sw2
    assert n_eff > 1e8
    z = bench.z_scores(p, exact, n_eff)
    assert np.abs(z).max() < 4.0, z
