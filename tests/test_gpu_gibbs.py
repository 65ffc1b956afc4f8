"""GPU parity: batched Gibbs (infer Nodewise/Full, gibbs_sample) through the C ABI.

Chains are driven by the same Philox streams as the oracle's device spec and use the same fp64 product /
cumulative-scan arithmetic, so trajectories -- and therefore integer counts and sample tables -- must match the
oracle EXACTLY.  Statistical bar: posterior cells within 4 sigma of ExactInference, sigma from the between-chain
variance (north_star).
"""
import numpy as np
import pytest

from conftest import make_mirror_net, oracle_net_from_layout
from oracle import capi
from oracle.factors_np import exact_infer
from test_gpu_sampling import random_case

pytestmark = pytest.mark.gpu


@pytest.fixture
def gibbs_mode(bn):
    yield bn.set_gibbs_mode
    bn.set_gibbs_mode("auto")


@pytest.mark.parametrize("mode", ["generic", "specialized"])
@pytest.mark.parametrize("full", [False, True])
@pytest.mark.parametrize("seed,n_nodes,n_ev", [(0, 30, 3), (1, 200, 20), (2, 9, 0)])
def test_gibbs_infer_counts_equal_oracle(bn, gibbs_mode, mode, full, seed, n_nodes, n_ev):
    """Both Gibbs kernels -- the interpreter (csrc/gibbs.cu) and the NVRTC network-specialised one
    (csrc/specialize.cu) -- reproduce the oracle's chains exactly: counts, final states, per-chain counts."""
    gibbs_mode(mode)
    net, lay, on, query, evidence = random_case(bn, n_nodes, seed, n_ev)
    cls = bn.GibbsSamplingFull if full else bn.GibbsSamplingNodewise
    nsamples, burn, thin = (60, 20, 3) if full else (1500, 500, 5)
    im = cls(nsamples=nsamples, burn_in=burn, thin=thin, n_chains=300, seed=seed, keep_chain_counts=True)
    counts = im.raw_counts(bn.InferenceState(net, query, evidence))
    oc, ost, opc = capi.gibbs_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), 300, nsamples, burn,
                                    thin, full=full, seed=seed, per_chain=True)
    assert np.array_equal(counts, oc.ravel(order="F"))
    assert np.array_equal(im.chain_states, ost)
    assert np.array_equal(im.last_chain_counts, opc)
    assert counts.sum() == 300 * -(-(nsamples - burn) // thin)


@pytest.mark.parametrize("mode", ["generic", "specialized"])
def test_gibbs_state_resume_equals_oracle(bn, gibbs_mode, mode):
    """im.state semantics (src/Inference/gibbs.jl:64,83-88): a second call continues the chains."""
    gibbs_mode(mode)
    net, lay, on, query, evidence = random_case(bn, 25, 4, 2)
    im = bn.GibbsSamplingNodewise(nsamples=400, burn_in=100, thin=2, n_chains=64, seed=9)
    inf = bn.InferenceState(net, query, evidence)
    c1 = im.raw_counts(inf)
    st1 = im.chain_states.copy()
    im.seed = 10
    c2 = im.raw_counts(inf)
    evv, qi = lay.evidence_vector(evidence), lay.query_indices(query)
    o1, ost1 = capi.gibbs_infer(on, evv, qi, 64, 400, 100, 2, seed=9)
    o2, ost2 = capi.gibbs_infer(on, evv, qi, 64, 400, 100, 2, seed=10, state=ost1)
    assert np.array_equal(st1, ost1) and np.array_equal(c1, o1.ravel(order="F"))
    assert np.array_equal(im.chain_states, ost2) and np.array_equal(c2, o2.ravel(order="F"))
    assert im.state == {n: int(ost2[0, i]) + 1 for i, n in enumerate(lay.names)}


@pytest.mark.parametrize("rows", ["0", "16", "4096"])
@pytest.mark.parametrize("full", [False, True])
def test_gibbs_tabulated_conditionals_equal_oracle(bn, gibbs_mode, monkeypatch, rows, full):
    """csrc/gibbs_tables.cu: nodes with small Markov blankets draw from integer threshold tables (R_v = the smallest
    32-bit draw for which the fp64 cumulative scan of src/Inference/gibbs.jl:36-50,110-112 passes state v).  Every row
    limit -- no tables, a few nodes, every node of the 30-node net -- must walk the oracle's chains exactly."""
    monkeypatch.setenv("BN_GIBBS_TABLE_ROWS", rows)
    gibbs_mode("specialized")
    net, lay, on, query, evidence = random_case(bn, 30, 11, 3)
    cls = bn.GibbsSamplingFull if full else bn.GibbsSamplingNodewise
    nsamples, burn, thin = (40, 10, 3) if full else (1200, 300, 7)
    im = cls(nsamples=nsamples, burn_in=burn, thin=thin, n_chains=200, seed=21, keep_chain_counts=True)
    counts = im.raw_counts(bn.InferenceState(net, query, evidence))
    oc, ost, opc = capi.gibbs_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), 200, nsamples, burn,
                                    thin, full=full, seed=21, per_chain=True)
    assert np.array_equal(counts, oc.ravel(order="F"))
    assert np.array_equal(im.chain_states, ost)
    assert np.array_equal(im.last_chain_counts, opc)


def test_gibbs_tables_with_deterministic_cpds_fall_back(bn, gibbs_mode, monkeypatch):
    """CPDs with exact zeros put the threshold 2^32 ("never") into some table rows; it does not fit 32 bits, so the
    library keeps the fp64 kernel for such a net -- and still equals the oracle.  Cards 2..5 cover every row width."""
    monkeypatch.setenv("BN_GIBBS_TABLE_ROWS", "4096")
    gibbs_mode("specialized")
    rng = np.random.default_rng(3)
    nodes = [["a", [], [[0.2, 0.8]]],
             ["b", ["a"], [[1.0, 0.0, 0.0], [0.0, 0.5, 0.5]]],
             ["c", ["a", "b"], [list(x) for x in rng.dirichlet(np.ones(4), 6)]],
             ["d", ["c"], [[0.0, 1.0, 0.0, 0.0, 0.0], [0.1, 0.2, 0.3, 0.2, 0.2], [0.0, 0.0, 0.0, 0.0, 1.0], [0.2] * 5]],
             ["e", ["b", "d"], [list(x) for x in rng.dirichlet(np.ones(2), 15)]]]
    net = make_mirror_net(bn, nodes)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    query, evidence = ["c", "d"], {"e": 1}
    im = bn.GibbsSamplingNodewise(nsamples=900, burn_in=100, thin=3, n_chains=128, seed=5)
    counts = im.raw_counts(bn.InferenceState(net, query, evidence))
    oc, ost = capi.gibbs_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), 128, 900, 100, 3, seed=5)
    assert np.array_equal(counts, oc.ravel(order="F")) and np.array_equal(im.chain_states, ost)


def test_gibbs_large_cardinality_path(bn):
    """Nodes with more than 8 states take the chunked conditional path."""
    rng = np.random.default_rng(0)
    net = bn.BayesNet()
    net.push(bn.CategoricalCPD("r", [], [], [rng.dirichlet(np.ones(11))]))
    net.push(bn.CategoricalCPD("m", ["r"], [11], [rng.dirichlet(np.ones(19)) for _ in range(11)]))
    net.push(bn.CategoricalCPD("l", ["m", "r"], [19, 11], [rng.dirichlet(np.ones(3)) for _ in range(19 * 11)]))
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    im = bn.GibbsSamplingNodewise(nsamples=600, burn_in=100, thin=1, n_chains=128, seed=3)
    c = im.raw_counts(bn.InferenceState(net, ["m"], {"l": 2}))
    oc, _ = capi.gibbs_infer(on, lay.evidence_vector({"l": 2}), lay.query_indices(["m"]), 128, 600, 100, 1, seed=3)
    assert np.array_equal(c, oc.ravel(order="F"))


def test_gibbs_posterior_within_4_sigma_between_chain(bn):
    net, lay, on, query, evidence = random_case(bn, 40, 11, 4, n_q=1)
    im = bn.GibbsSamplingFull(nsamples=300, burn_in=100, thin=2, n_chains=4096, seed=1, keep_chain_counts=True)
    phi = bn.infer(im, net, query, evidence=evidence)
    pc = im.last_chain_counts
    pc = pc / pc.sum(axis=1, keepdims=True)
    mean = pc.mean(axis=0)
    sigma = pc.std(axis=0, ddof=1) / np.sqrt(pc.shape[0])
    exact = exact_infer(on, query, {k: v - 1 for k, v in evidence.items()}).potential.ravel(order="F")
    np.testing.assert_allclose(phi.potential.ravel(order="F"), mean, rtol=1e-12)
    assert np.all(np.abs(mean - exact) < 4 * sigma + 1e-4), (mean, exact, sigma)  # 1e-4: residual burn-in bias


def test_gibbs_reference_suite(bn, golden):
    """test/test_inference.jl:45-110 for GibbsSamplingNodewise / GibbsSamplingFull with the default
    nsamples=2000, burn_in=500, thin=3 and a single chain, at the reference's tolerances."""
    g = golden["inference_acb"]
    acb = make_mirror_net(bn, g["nodes"])
    s = golden["inference_student"]
    stu = make_mirror_net(bn, s["nodes"])
    for cls in (bn.GibbsSamplingNodewise, bn.GibbsSamplingFull):
        phi = bn.infer(cls(), acb, "a")
        np.testing.assert_allclose(phi.potential, g["P_a"], atol=g["atol"])
        phi = bn.infer(cls(), acb, "c")
        np.testing.assert_allclose(phi.potential, g["P_c"], atol=g["atol"])
        phi = bn.infer(cls(), acb, ["b", "c"])
        assert phi.size() == (2, 2)
        np.testing.assert_allclose(phi.potential.ravel(order="F"), g["P_bc_colmajor"], atol=g["atol"])
        phi = bn.infer(cls(n_chains=64), stu, "D")
        np.testing.assert_allclose(phi.potential, s["P_D"], atol=s["atol"])
        phi = bn.infer(cls(n_chains=64), stu, "G", evidence={"D": 1, "I": 1})
        np.testing.assert_allclose(phi.potential, s["P_G_given_d1_i1"], atol=s["atol"])
        with pytest.raises(ValueError):
            bn.infer(cls(nsamples=10, burn_in=10), stu, "D")


@pytest.mark.parametrize("mode", ["generic", "specialized"])
@pytest.mark.parametrize("order", [None, "fixed"])
def test_gibbs_sample_tables_equal_oracle(bn, gibbs_mode, order, mode):
    """gibbs_sample tables from the interpreter kernel and from the NVRTC network-specialised one (per-node blocks behind
    a switch on the sweep order) are identical to the oracle's: random and fixed sweep orders, thinning, burn-in."""
    gibbs_mode(mode)
    net, lay, on, query, evidence = random_case(bn, 30, 5, 3)
    evv = lay.evidence_vector(evidence)
    free = [n for n in lay.names if evv[lay.index[n]] < 0]
    vo_names = None
    vo = None
    if order == "fixed":
        rng = np.random.default_rng(0)
        perm = rng.permutation(len(free))
        vo = perm.astype(np.int32)
        vo_names = [free[i] for i in perm] + list(evidence)  # must mention every variable (src/gibbs.jl:302-311)
    df = bn.gibbs_sample(net, 7, 20, thinning=2, consistent_with=evidence, variable_order=vo_names, n_chains=96, seed=8)
    ost = capi.gibbs_sample(on, evv, 96, 7, 20, thinning=2, variable_order=vo, seed=8)
    assert df.shape == (96 * 7, lay.n)
    for n in free:
        assert np.array_equal(df[n].to_numpy() - 1, ost[lay.index[n]])
    for n, v in evidence.items():
        assert (df[n] == float(v)).all()


def test_gibbs_sample_reference_suite(bn, golden):
    """test/test_gibbs.jl:100-121 (discrete nets) and test/test_sampling.jl:25-27."""
    d_bn = make_mirror_net(bn, golden["inference_acb"]["nodes"])
    t12 = bn.gibbs_sample(d_bn, 5, 100, thinning=5)
    assert t12.shape == (5, 3)
    assert t12["a"].tolist() == [1] * 5 and t12["b"].tolist() == [2] * 5 and t12["c"].tolist() == [1] * 5
    t13 = bn.gibbs_sample(d_bn, 5, 100, thinning=5, consistent_with={"c": 1})
    assert t13.shape == (5, 3)
    assert t13["a"].tolist() == [1] * 5 and t13["b"].tolist() == [2] * 5 and t13["c"].tolist() == [1.0] * 5
    t5 = bn.rand(d_bn, bn.GibbsSampler({"c": 1}, burn_in=5), 5)
    assert t5["a"].tolist() == [1] * 5 and t5["b"].tolist() == [2] * 5
    # test/test_gibbs.jl:22-49: c = 2(b-1) + a deterministically
    bn2 = bn.BayesNet()
    bn2.push(bn.DiscreteCPD("a", [0.5, 0.5]))
    bn2.push(bn.DiscreteCPD("b", [0.5, 0.5]))
    bn2.push(bn.CategoricalCPD("c", ["a", "b"], [2, 2], [[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]]))
    t6 = bn.gibbs_sample(bn2, 5, 100, thinning=5, consistent_with={"c": 1})
    assert t6["a"].tolist() == [1] * 5 and t6["b"].tolist() == [1] * 5
    t7 = bn.gibbs_sample(bn2, 5, 100, thinning=5, consistent_with={"c": 2})
    assert t7["a"].tolist() == [2] * 5 and t7["b"].tolist() == [1] * 5 and t7["c"].tolist() == [2.0] * 5
    # initial_sample given: no likelihood-weighted start
    t8 = bn.gibbs_sample(bn2, 4, 0, initial_sample={"a": 2, "b": 2, "c": 4}, consistent_with={"c": 4})
    assert t8["a"].tolist() == [2] * 4 and t8["b"].tolist() == [2] * 4
    # impossible evidence: no initial sample with non-zero probability (src/gibbs.jl:332-334)
    with pytest.raises(RuntimeError):
        bn.gibbs_sample(d_bn, 5, 10, consistent_with={"a": 2})


def test_gibbs_specialized_falls_back_for_large_cardinalities(bn, gibbs_mode):
    """Nodes with more than 8 states are outside the specialised generator: "auto" silently uses the interpreter
    kernel, "specialized" refuses loudly (NotImplementedError <- BN_UNSUPPORTED)."""
    rng = np.random.default_rng(0)
    net = bn.BayesNet()
    net.push(bn.CategoricalCPD("r", [], [], [rng.dirichlet(np.ones(11))]))
    net.push(bn.CategoricalCPD("m", ["r"], [11], [rng.dirichlet(np.ones(3)) for _ in range(11)]))
    gibbs_mode("specialized")
    with pytest.raises(NotImplementedError):
        bn.GibbsSamplingNodewise(nsamples=50, burn_in=10, thin=1, n_chains=32).raw_counts(bn.InferenceState(net, ["m"], {}))
    gibbs_mode("auto")
    c = bn.GibbsSamplingNodewise(nsamples=50, burn_in=10, thin=1, n_chains=32, seed=1).raw_counts(bn.InferenceState(net, ["m"], {}))
    assert c.sum() == 32 * 40


@pytest.mark.parametrize("mode", ["generic", "specialized"])
def test_gibbs_wide_query_and_cards_up_to_8(bn, gibbs_mode, mode):
    """> 16 query cells (MATCH.ANY-grouped atomics in the specialised kernel) and 5..8-state nodes."""
    rng = np.random.default_rng(3)
    net = bn.BayesNet()
    net.push(bn.CategoricalCPD("a", [], [], [rng.dirichlet(np.ones(5))]))
    net.push(bn.CategoricalCPD("b", ["a"], [5], [rng.dirichlet(np.ones(8)) for _ in range(5)]))
    net.push(bn.CategoricalCPD("c", ["a", "b"], [5, 8], [rng.dirichlet(np.ones(6)) for _ in range(40)]))
    net.push(bn.CategoricalCPD("d", ["c"], [6], [rng.dirichlet(np.ones(2)) for _ in range(6)]))
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    gibbs_mode(mode)
    for full in (False, True):
        cls = bn.GibbsSamplingFull if full else bn.GibbsSamplingNodewise
        im = cls(nsamples=300, burn_in=60, thin=2, n_chains=200, seed=7)
        c = im.raw_counts(bn.InferenceState(net, ["a", "b"], {"d": 1}))
        oc, _ = capi.gibbs_infer(on, lay.evidence_vector({"d": 1}), lay.query_indices(["a", "b"]), 200, 300, 60, 2,
                                 full=full, seed=7)
        assert np.array_equal(c, oc.ravel(order="F"))


def test_gibbs_c4_at_its_stated_size(bn, gibbs_mode):
    """BASELINE configs[3] at full size on one GPU: GibbsSamplingNodewise(2000, 1000, 5), 200-node net, 20 evidence,
    65 536 chains in ONE launch (src/Inference/gibbs.jl:66-145).  The oracle runs the first 256 chains: their
    integer counts, per-chain counts and final states must be identical; every chain must have kept exactly
    ceil((nsamples - burn_in) / thin) samples (:74)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gibbs_bench
    gibbs_mode("specialized")
    net, lay, query, evidence = gibbs_bench.build_case(bn)
    assert lay.n == 200 and len(evidence) == 20
    n_chains, nsamples, burn_in, thin = 65536, 2000, 1000, 5
    im = bn.GibbsSamplingNodewise(nsamples=nsamples, burn_in=burn_in, thin=thin, n_chains=n_chains, seed=0,
                                  keep_chain_counts=True)
    counts = im.raw_counts(bn.InferenceState(net, query, evidence))
    kept = -(-(nsamples - burn_in) // thin)
    assert np.array_equal(im.last_chain_counts.sum(axis=1), np.full(n_chains, float(kept)))
    assert counts.sum() == n_chains * kept and np.array_equal(counts, im.last_chain_counts.sum(axis=0))
    on = oracle_net_from_layout(lay)
    evv, qi = lay.evidence_vector(evidence), lay.query_indices(query)
    oc, ost, opc = capi.gibbs_infer(on, evv, qi, 256, nsamples, burn_in, thin, seed=0, per_chain=True)
    assert np.array_equal(im.last_chain_counts[:256], opc)
    assert np.array_equal(im.chain_states[:256], ost)
    for i, e in enumerate(evv):  # evidence stays clamped in every chain
        if e >= 0:
            assert (im.chain_states[:, i] == e).all()
