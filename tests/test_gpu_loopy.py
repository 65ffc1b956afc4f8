"""GPU parity: infer(LoopyBelief(), ...) -- the batched message-passing kernel (csrc/loopy.cu) -- vs the oracle
(oracle/loopy_np.py, a numpy restatement of src/Inference/loopy_belief.jl:24-218 with the same array aliasing).

Bar: beliefs and iteration counts IDENTICAL to the oracle (the kernel follows its arithmetic order term by term and is
compiled without fused multiply-adds); the reference's own four LoopyBelief posteriors within its tolerance.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import make_mirror_net, oracle_net_from_layout
from oracle.loopy_np import loopy_infer

pytestmark = pytest.mark.gpu


def random_instances(bn, n_nodes, seed, n_inst, max_ev):
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=seed)
    lay = bn.flatten(net)
    rng = np.random.default_rng(77 + seed)
    queries, evidences = [], []
    for b in range(n_inst):
        k = int(rng.integers(0, max_ev + 1)) if b else 0  # instance 0: no evidence
        picks = [str(x) for x in rng.choice(lay.names, size=min(k + 1, lay.n), replace=False)]
        queries.append(picks[0])
        evidences.append({e: int(rng.integers(1, net.ncategories(e) + 1)) for e in picks[1:]})
    return net, lay, oracle_net_from_layout(lay), queries, evidences


def oracle_batch(on, lay, queries, evidences, **kw):
    outs, its = [], []
    for q, e in zip(queries, evidences):
        o, it = loopy_infer(on, lay.index[q], lay.evidence_vector(e), return_iters=True, **kw)
        outs.append(o)
        its.append(it)
    return outs, its


def test_loopy_reference_suite(bn, golden):
    """test/test_inference.jl:45-63,94-110 with im = LoopyBelief(), same assertions and tolerances."""
    im = bn.LoopyBelief()
    g = golden["inference_acb"]
    acb = make_mirror_net(bn, g["nodes"])
    phi = bn.infer(im, acb, "a")
    assert len(phi) == 2
    assert float(phi[{"a": 1}].potential) == pytest.approx(1.0, abs=g["atol"])
    assert float(phi[{"a": 2}].potential) == pytest.approx(0.0, abs=g["atol"])
    phi = bn.infer(im, acb, "c")
    assert len(phi) == 2
    assert float(phi[{"c": 1}].potential) == pytest.approx(1.0, abs=g["atol"])
    assert float(phi[{"c": 2}].potential) == pytest.approx(0.0, abs=g["atol"])
    s = golden["inference_student"]
    stu = make_mirror_net(bn, s["nodes"])
    phi = bn.infer(im, stu, "D")
    assert phi.size() == (2,)
    np.testing.assert_allclose(phi.potential, s["P_D"], atol=s["atol"])
    phi = bn.infer(im, stu, "G", evidence={"D": 1, "I": 1})
    assert phi.size() == (3,)
    np.testing.assert_allclose(phi.potential, s["P_G_given_d1_i1"], atol=s["atol"])
    with pytest.raises(ValueError):  # loopy_belief.jl:26
        bn.infer(im, stu, ["D", "G"])
    with pytest.raises(ValueError):  # InferenceState
        bn.infer(im, stu, "waldo")
    with pytest.raises(ValueError):
        bn.infer(im, stu, "D", evidence={"D": 1})


@pytest.mark.parametrize("seed,n_nodes,n_inst,max_ev,kw", [
    (0, 12, 24, 4, dict()),
    (1, 9, 16, 3, dict(nsamples=40, tol=-1.0)),
    (2, 37, 10, 6, dict(nsamples=60)),
    (3, 100, 4, 12, dict(nsamples=25)),
    (4, 200, 2, 20, dict(nsamples=6, iters_for_convergence=2)),
    (5, 5, 40, 2, dict(iters_for_convergence=1, tol=1e-3)),
    (6, 1, 3, 0, dict()),
])
def test_loopy_batch_identical_to_oracle(bn, seed, n_nodes, n_inst, max_ev, kw):
    net, lay, on, queries, evidences = random_instances(bn, n_nodes, seed, n_inst, max_ev)
    im = bn.LoopyBelief(**kw)
    post = im.infer_batch(net, queries, evidences)
    outs, its = oracle_batch(on, lay, queries, evidences, **kw)
    assert list(im.last_iters) == its
    for b, o in enumerate(outs):
        c = len(o)
        assert np.array_equal(post[b, :c], o, equal_nan=True), (b, post[b, :c], o)
        assert not post[b, c:].any()


def test_loopy_single_calls_equal_batch(bn):
    net, lay, on, queries, evidences = random_instances(bn, 20, 11, 6, 4)
    im = bn.LoopyBelief(nsamples=80)
    post = im.infer_batch(net, queries, evidences)
    for b, (q, e) in enumerate(zip(queries, evidences)):
        phi = bn.infer(im, net, q, evidence=e)
        assert phi.dimensions == [q]
        assert np.array_equal(phi.potential, post[b, :net.ncategories(q)], equal_nan=True)


def test_loopy_exact_on_polytree_without_evidence(bn, golden):
    """No evidence -> all lambdas stay uniform, roots keep their priors, and BP is exact on a polytree."""
    stu = make_mirror_net(bn, golden["inference_student"]["nodes"])
    im = bn.LoopyBelief()
    for name, want in [("D", [0.6, 0.4]), ("I", [0.7, 0.3]), ("G", [0.447, 0.2714, 0.2816])]:
        np.testing.assert_allclose(bn.infer(im, stu, name).potential, want, rtol=1e-12)
        np.testing.assert_allclose(bn.infer(im, stu, name).potential,
                                   bn.infer(bn.ExactInference(), stu, name).potential, rtol=1e-12)


def test_loopy_impossible_evidence_and_edge_parameters(bn, golden):
    """Evidence of probability zero -> 0/0 = NaN, as in the reference; nsamples = 0 returns prior x uniform lambdas."""
    acb = make_mirror_net(bn, golden["inference_acb"]["nodes"])
    lay = bn.flatten(acb)
    on = oracle_net_from_layout(lay)
    im = bn.LoopyBelief()
    got = bn.infer(im, acb, "c", evidence={"a": 2}).potential  # P(a = 2) = 0
    want = loopy_infer(on, lay.index["c"], lay.evidence_vector({"a": 2}))
    assert np.array_equal(got, want, equal_nan=True)
    stu = make_mirror_net(bn, golden["inference_student"]["nodes"])
    np.testing.assert_array_equal(bn.infer(bn.LoopyBelief(nsamples=0), stu, "I").potential, [0.7, 0.3])
    with pytest.raises(ValueError):
        bn.infer(bn.LoopyBelief(iters_for_convergence=0), stu, "I")
    with pytest.raises(IndexError):
        bn.infer(im, stu, "I", evidence={"D": 3})


def test_loopy_wide_cards_and_many_parents(bn):
    """7-8 state nodes and a 5-parent node: the digit table and the generic parent loop."""
    rng = np.random.default_rng(3)
    net = bn.BayesNet()
    cards = [7, 8, 2, 3, 2, 5]
    for i, c in enumerate(cards):
        net.push(bn.rand_cpd(net, c, f"R{i}", [f"R{i - 1}"] if i in (2, 4) else [], rng=rng))
    net.push(bn.rand_cpd(net, 4, "H", ["R0", "R1", "R3", "R4", "R5"], rng=rng))
    net.push(bn.rand_cpd(net, 3, "L1", ["H", "R2"], rng=rng))
    net.push(bn.rand_cpd(net, 2, "L2", ["H"], rng=rng))
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    queries = ["H", "R0", "R3", "L1", "R5"]
    evidences = [{"L1": 2}, {"L2": 1, "R1": 8}, {"H": 4, "R0": 7}, {}, {"L1": 3, "L2": 2, "R2": 1}]
    kw = dict(nsamples=30)
    im = bn.LoopyBelief(**kw)
    post = im.infer_batch(net, queries, evidences)
    outs, its = oracle_batch(on, lay, queries, evidences, **kw)
    assert list(im.last_iters) == its
    for b, o in enumerate(outs):
        assert np.array_equal(post[b, :len(o)], o, equal_nan=True), (b, post[b], o)


def test_loopy_device_entry_point(bn):
    """bn_loopy_infer_dev on device pointers (torch only as the allocator) == the host-buffer call."""
    torch = pytest.importorskip("torch")
    net, lay, on, queries, evidences = random_instances(bn, 30, 21, 64, 5)
    im = bn.LoopyBelief(nsamples=50)
    want = im.infer_batch(net, queries, evidences)
    dn = bn.device_net(net)
    ev = np.stack([lay.evidence_vector(e) for e in evidences])
    q = np.array([lay.index[x] for x in queries], np.int32)
    stride = int(lay.card.max())
    dev = torch.device(f"cuda:{dn.device}")
    ev_d = torch.from_numpy(ev).to(dev)
    q_d = torch.from_numpy(q).to(dev)
    out_d = torch.empty((len(queries), stride), dtype=torch.float64, device=dev)
    it_d = torch.empty(len(queries), dtype=torch.int32, device=dev)
    from bayesnets_jl_b200 import _lib
    _lib.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        _lib.check(_lib.lib().bn_loopy_infer_dev(_lib.h64(dn.handle), C.c_void_p(ev_d.data_ptr()), C.c_void_p(q_d.data_ptr()),
                                                 C.c_uint64(len(queries)), C.c_int32(50), C.c_double(1e-8), C.c_int32(6),
                                                 C.c_int32(stride), C.c_void_p(out_d.data_ptr()), C.c_void_p(it_d.data_ptr())))
        torch.cuda.synchronize(dev)
    finally:
        _lib.set_stream(None)
    assert np.array_equal(out_d.cpu().numpy(), want, equal_nan=True)
    assert np.array_equal(it_d.cpu().numpy(), im.last_iters)


def test_loopy_placements_and_large_net(bn, monkeypatch):
    """Plan / CPTs / message planes in shared or global memory (BN_LOOPY_MODE bits 0 / 1 / 2) are the same kernel:
    every placement is identical to the oracle; a 700-node net, too large for shared memory, runs from global memory."""
    net, lay, on, queries, evidences = random_instances(bn, 37, 2, 10, 6)
    outs, its = oracle_batch(on, lay, queries, evidences, nsamples=40)
    for mode in (0, 1, 2, 3, 4, 5, 7):
        monkeypatch.setenv("BN_LOOPY_MODE", str(mode))
        im = bn.LoopyBelief(nsamples=40)
        got = im.infer_batch(net, queries, evidences)
        assert list(im.last_iters) == its, mode
        for b, o in enumerate(outs):
            assert np.array_equal(got[b, :len(o)], o, equal_nan=True), mode
    monkeypatch.delenv("BN_LOOPY_MODE")
    net, lay, on, queries, evidences = random_instances(bn, 700, 8, 2, 30)
    kw = dict(nsamples=3)
    im = bn.LoopyBelief(**kw)
    got = im.infer_batch(net, queries, evidences)
    outs, its = oracle_batch(on, lay, queries, evidences, **kw)
    assert list(im.last_iters) == its
    for b, o in enumerate(outs):
        assert np.array_equal(got[b, :len(o)], o, equal_nan=True)


def test_loopy_c2_batch_identical_to_c_oracle(bn):
    """The bench case itself: 100-node net, 10 evidence nodes per pair, 768 pairs (24 CTAs) -- every belief and iteration
    count identical to the C restatement (which the CPU tests hold bit-identical to the numpy oracle)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import loopy_bench
    from oracle import capi
    net, lay, ev, q = loopy_bench.build_case(bn, 768)
    on = oracle_net_from_layout(lay)
    for kw in (dict(nsamples=30, tol=-1.0), dict(nsamples=12, tol=0.05, iters_for_convergence=2)):
        im = bn.LoopyBelief(**kw)
        queries = [lay.names[int(i)] for i in q]
        evidences = [{lay.names[i]: int(e[i]) + 1 for i in np.nonzero(e >= 0)[0]} for e in ev]
        got = im.infer_batch(net, queries, evidences)
        want, its = capi.loopy_infer(on, q, ev, **kw)
        assert np.array_equal(got, want, equal_nan=True)
        assert np.array_equal(im.last_iters, its)
