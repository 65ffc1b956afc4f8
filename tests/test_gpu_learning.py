"""GPU parity: sufficient statistics (bn_statistics[_dev], csrc/stats.cu), Bayesian score and MLE fit vs the CPU
oracle and the reference's own golden values.  Counts are integers: the bar is exact equality."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

from conftest import oracle_net_from_layout
from oracle import capi
from oracle.scoring_py import bayesian_score_component as orc_score, split_counts, uniform_alpha, bdeu_alpha

pytestmark = pytest.mark.gpu


def test_reference_golden_statistics_scores_and_fit(bn, golden):
    """test/test_discrete_bayes_nets.jl:13-67 and test/test_learning.jl:40-96 through the public API."""
    g = golden["statistics_ab"]
    data = pd.DataFrame(g["data"])
    net = bn.fit(data, tuple(tuple(e) for e in g["edges"]))
    assert net.names() == ["a", "b"]
    mat = np.array([g["data"]["a"], g["data"]["b"]])
    plist, ncat = [[], [1]], [2, 2]
    assert np.isclose(bn.bayesian_score_component(1, plist[0], ncat, mat, bn.UniformPrior()), g["score_1_uniform"], rtol=1e-13)
    assert np.isclose(bn.bayesian_score_component(1, plist[0], ncat, mat, bn.UniformPrior(2.0)), g["score_1_uniform2"], rtol=1e-13)
    assert np.isclose(bn.bayesian_score_component(1, plist[0], ncat, mat, bn.BDeuPrior()), g["score_1_bdeu"], rtol=1e-13)
    assert np.isclose(bn.bayesian_score_component(1, plist[0], ncat, mat, bn.BDeuPrior(2.0)), g["score_1_bdeu2"], rtol=1e-13)
    assert np.isclose(bn.bayesian_score_component(2, plist[1], ncat, mat, bn.UniformPrior()), g["score_2_uniform"], rtol=1e-13)
    tot = g["score_1_uniform"] + g["score_2_uniform"]
    assert np.isclose(bn.bayesian_score(plist, ncat, mat, bn.UniformPrior()), tot, rtol=1e-13)
    assert np.isclose(bn.bayesian_score(net, data), tot, rtol=1e-13)
    comps = bn.bayesian_score_components(net, data, bn.UniformPrior())
    assert np.allclose(comps, [g["score_1_uniform"], g["score_2_uniform"]], rtol=1e-13)
    # fitted CPTs = relative frequencies (test_learning.jl test_disc_bn counts)
    assert np.allclose(net.get("a").distributions[0], [0.5, 0.5])
    assert np.allclose(net.get("b").distributions[0], [0.5, 0.5]) and np.allclose(net.get("b").distributions[1], [0.75, 0.25])

    g = golden["statistics_abc"]
    data = pd.DataFrame(g["data"])
    net = bn.fit(data, tuple(tuple(e) for e in g["edges"]))
    assert bn.statistics(net, "A", data).tolist() == g["stats_A"]
    assert bn.statistics(net, "C", data).tolist() == g["stats_C"]
    allN = bn.statistics(net, data)
    assert [x.tolist() for x in allN] == [bn.statistics(net, n, data).tolist() for n in ("A", "B", "C")]
    assert bn.bayesian_score(net, data) < 0.0
    with pytest.raises(bn.DimensionMismatch):
        bn.statistics(net, data[["A", "B"]])
    with pytest.raises(IndexError):  # state outside 1:ncategories
        bn.statistics([[], [1]], [2, 2], np.array([[1, 2, 3], [1, 1, 1]]))
    # parents listed child-first (non-topological labelling) and a cyclic parent list still count per node
    mat = np.array([g["data"][n] for n in g["names"]])
    fwd = bn.statistics([[], [1], [1, 2]], [3, 2, 2], mat)
    rev = bn.statistics([[2, 3], [3], []], [2, 2, 3], mat[::-1])
    assert rev[0].tolist() == bn.statistics(1, [2, 3], [2, 2, 3], mat[::-1]).tolist()
    assert fwd[2].sum() == rev[0].sum() == 12
    cyc = bn.statistics([[2], [1]], [3, 2], mat[:2])
    assert cyc[0].tolist() == bn.statistics(1, [2], [3, 2], mat[:2]).tolist()
    assert cyc[1].tolist() == bn.statistics(2, [1], [3, 2], mat[:2]).tolist()


@pytest.mark.parametrize("seed,n_nodes,n_rows", [(0, 12, 1), (1, 30, 1000), (2, 100, 4096), (3, 100, 50_001),
                                                 (4, 200, 300_017), (5, 7, 2_000_003)])
def test_statistics_host_table_equals_oracle(bn, seed, n_nodes, n_rows):
    """Random nets and random (not model-generated) tables: ragged last tile, padded pitch, every node at once."""
    rng = np.random.default_rng(seed)
    net = bn.rand_discrete_bn(n_nodes, 3, 5, seed=seed)
    lay = bn.flatten(net)
    table = np.stack([rng.integers(0, int(c), n_rows, dtype=np.uint8) for c in lay.card])
    out = np.zeros(int(lay.cpt_off[-1]), np.int64)
    from bayesnets_jl_b200 import _lib
    dn = bn.device_net(net)
    _lib.check(_lib.lib().bn_statistics(_lib.h64(dn.handle), _lib.ptr(table, _lib.u8p), C.c_uint64(n_rows), _lib.ptr(out, _lib.i64p)))
    ref = capi.statistics(oracle_net_from_layout(lay), table)
    assert np.array_equal(out, ref)
    for i in range(lay.n):  # every node's table sums to the number of rows
        assert out[lay.cpt_off[i]:lay.cpt_off[i + 1]].sum() == n_rows
    bad = table.copy()
    bad[min(3, n_nodes - 1), n_rows // 2] = 250
    with pytest.raises(IndexError):
        _lib.check(_lib.lib().bn_statistics(_lib.h64(dn.handle), _lib.ptr(bad, _lib.u8p), C.c_uint64(n_rows), _lib.ptr(out, _lib.i64p)))


def test_wide_cards_large_count_table_uses_global_atomics(bn):
    """A node with 255 states and two 40-state parents: 408 000 bins > shared memory -> global-atomic path."""
    rng = np.random.default_rng(9)
    a = bn.CategoricalCPD("a", [], [], [np.full(40, 1 / 40)])
    b = bn.CategoricalCPD("b", [], [], [np.full(40, 1 / 40)])
    c = bn.CategoricalCPD("c", ["a", "b"], [40, 40], [np.full(255, 1 / 255)] * 1600)
    net = bn.BayesNet([a, b, c])
    lay = bn.flatten(net)
    n_rows = 200_000
    table = np.stack([rng.integers(0, int(k), n_rows, dtype=np.uint8) for k in lay.card])
    out = np.zeros(int(lay.cpt_off[-1]), np.int64)
    from bayesnets_jl_b200 import _lib
    dn = bn.device_net(net)
    _lib.check(_lib.lib().bn_statistics(_lib.h64(dn.handle), _lib.ptr(table, _lib.u8p), C.c_uint64(n_rows), _lib.ptr(out, _lib.i64p)))
    assert np.array_equal(out, capi.statistics(oracle_net_from_layout(lay), table))


def test_sample_then_count_on_device_recovers_the_cpts(bn):
    """rand(bn, n) stays in HBM (bn_sample_dev) -> bn_statistics_dev -> MLE fit: counts equal the oracle's on the
    same table, unaligned row counts take the plain-load path, rejection-failed rows (0xFF) are skipped, and the
    fitted CPTs lie within 5 sigma of the generating ones."""
    import torch
    from bayesnets_jl_b200 import _lib
    net = bn.rand_discrete_bn(40, 3, 4, seed=3)
    lay = bn.flatten(net)
    dn = bn.device_net(net)
    L = _lib.lib()
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    for n_rows in (1 << 20, (1 << 18) + 5):
        tab = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
        _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), None, C.c_int32(0), C.c_int32(1), C.c_uint64(0), C.c_uint64(n_rows),
                                   C.c_uint64(11), C.c_void_p(tab.data_ptr()), None, None))
        lay2, counts = bn.statistics_device(net, tab.data_ptr(), n_rows)
        ref = capi.statistics(oracle_net_from_layout(lay), tab.cpu().numpy())
        assert np.array_equal(counts, ref)
    # MLE from the last table vs the true CPT
    fitted = bn.fit_from_device_table(net, tab.data_ptr(), n_rows)
    assert sorted(fitted.names()) == sorted(net.names())  # (the constructor may pick another topological order)
    Ns = split_counts(lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, counts)
    for i, N in enumerate(Ns):
        src, est_cpd = net.get(lay.names[i]), fitted.get(lay.names[i])
        assert est_cpd.parents == src.parents and est_cpd.parental_ncategories == src.parental_ncategories
        tot = np.maximum(N.sum(axis=0), 1)
        for j in range(N.shape[1]):
            p, est = src.distributions[j], est_cpd.distributions[j]
            if N[:, j].sum() > 0:
                assert np.array_equal(est, N[:, j] / N[:, j].sum())
                assert np.all(np.abs(est - p) <= 5 * np.sqrt(p * (1 - p) / tot[j]) + 1e-12)
            else:
                assert np.allclose(est, 1.0 / len(est))
    # rejection sampling with evidence that sometimes fails: failed rows are 0xFF and must not be counted
    ev = lay.evidence_vector({lay.names[5]: 1, lay.names[17]: 2})
    n_rows = 1 << 16
    tab = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
    ok = torch.empty(n_rows, dtype=torch.uint8, device="cuda:0")
    _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), C.c_int32(1), C.c_int32(2), C.c_uint64(0),
                               C.c_uint64(n_rows), C.c_uint64(5), C.c_void_p(tab.data_ptr()), None, C.c_void_p(ok.data_ptr())))
    n_ok = int(ok.sum().item())
    assert 0 < n_ok < n_rows
    _, counts = bn.statistics_device(net, tab.data_ptr(), n_rows)
    assert counts[lay.cpt_off[0]:lay.cpt_off[1]].sum() == n_ok
    assert np.array_equal(counts, capi.statistics(oracle_net_from_layout(lay), tab.cpu().numpy()))


@pytest.mark.parametrize("n_rows", [0, 1, 257, 100_003])
def test_statistics_many_parents_and_empty_tables(bn, n_rows):
    """Nodes with more parents than the kernel keeps in registers (generic tail loop), a parentless single-state node,
    zero rows, and the same structure with a count table too large for shared memory (global-atomic path)."""
    from bayesnets_jl_b200 import _lib
    rng = np.random.default_rng(21)
    for wide in (False, True):
        cards = [3, 2, 4, 2, 3, 1, 2] + ([200, 150] if wide else [])
        cpds = []
        for i, k in enumerate(cards):
            if i < 5 or i == 5:
                parents = []
            elif i == 6:
                parents = [0, 1, 2, 3, 4]          # 5 parents: 3 in registers + 2 from the program
            else:
                parents = [0, 2, 4, 6][: (4 if i == 7 else 2)] + ([7] if i == 8 else [])  # 4 parents; 255-ish wide tables
            pn = [cards[p] for p in parents]
            q = int(np.prod(pn)) if parents else 1
            cpds.append(bn.CategoricalCPD(f"m{i}", [f"m{p}" for p in parents], pn, [np.full(k, 1.0 / k)] * q))
        net = bn.BayesNet(cpds)
        lay = bn.flatten(net)
        table = np.stack([rng.integers(0, int(c), n_rows, dtype=np.uint8) for c in lay.card]) if n_rows else \
            np.zeros((lay.n, 0), np.uint8)
        out = np.full(int(lay.cpt_off[-1]), -1, np.int64)
        dn = bn.device_net(net)
        _lib.check(_lib.lib().bn_statistics(_lib.h64(dn.handle), _lib.ptr(table, _lib.u8p) if n_rows else None,
                                            C.c_uint64(n_rows), _lib.ptr(out, _lib.i64p)))
        ref = capi.statistics(oracle_net_from_layout(lay), table, n_rows=n_rows, pitch=max(n_rows, 1)) if n_rows else \
            np.zeros_like(out)
        assert np.array_equal(out, ref), (wide, n_rows)
