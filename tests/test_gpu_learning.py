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


# ------------------------------------------------------------------------------------------ logpdf(bn, df)
# Tolerance: the oracle adds n_rows x n_nodes logs one after the other like the reference (rounding error grows with
# the number of terms, ~sqrt(terms) * 1e-16 relative); the device adds a few thousand count-weighted terms.  Both are
# held to 1e-12 relative of each other at the sizes below (<= 3e7 terms) and the device value to 1e-13 of an exactly
# rounded sum (math.fsum) of the same count-weighted terms.
def _exact_logl(lay, table):
    import math
    counts = capi.statistics(oracle_net_from_layout(lay), table)
    return math.fsum(int(c) * math.log(p) for c, p in zip(counts, np.asarray(lay.cpt)) if c)


def test_logpdf_reference_values(bn, golden, tmp_path):
    """pdf / logpdf through the public API on the reference's XDSL net (test/test_io.jl:13-18 joint values)."""
    import math
    from test_host import write_xdsl
    path = str(tmp_path / "sample_bn.xdsl")
    write_xdsl(path, golden)
    net = bn.readxdsl(path)
    rows = golden["xdsl"]["joint_1based"]
    for s, f, p in rows:
        assert bn.pdf(net, {"Success": s, "Forecast": f}) == pytest.approx(p, rel=1e-14)
        assert bn.logpdf(net, {"Success": s, "Forecast": f}) == pytest.approx(math.log(p), rel=1e-14)
        one = pd.DataFrame({"Success": [s], "Forecast": [f]})
        assert bn.logpdf(net, one) == pytest.approx(math.log(p), rel=1e-14)
        assert bn.pdf(net, one) == pytest.approx(p, rel=1e-13)
    df = pd.DataFrame({"Forecast": [f for _, f, _ in rows] * 3, "extra": 0, "Success": [s for s, _, _ in rows] * 3})
    want = 3 * math.fsum(math.log(p) for _, _, p in rows)
    assert bn.logpdf(net, df) == pytest.approx(want, rel=1e-14)  # columns by name, extra column ignored
    assert bn.pdf(net, df) == pytest.approx(math.exp(want), rel=1e-12)
    assert bn.logpdf(net, df.iloc[:0]) == 0.0
    with pytest.raises(KeyError):
        bn.logpdf(net, df[["Success"]])
    # a leaf value outside its support has pdf 0 (-> -Inf); as a parent it is a BoundsError
    leaf = [n for n in net.names() if not net.children(n)][0]
    root = [n for n in net.names() if net.children(n)][0]
    bad = df.copy()
    bad.loc[2, leaf] = 9
    assert bn.logpdf(net, bad) == -math.inf and bn.pdf(net, bad) == 0.0
    bad = df.copy()
    bad.loc[2, root] = 9
    with pytest.raises(IndexError):
        bn.logpdf(net, bad)


@pytest.mark.parametrize("seed,n_nodes,n_rows", [(0, 12, 1), (1, 30, 1000), (2, 100, 4096), (3, 100, 50_001),
                                                 (4, 200, 150_017), (5, 7, 2_000_003)])
def test_logpdf_host_table_equals_oracle(bn, seed, n_nodes, n_rows):
    """Random nets, tables sampled from the net by the oracle (so no zero-probability rows), through bn_logpdf."""
    from bayesnets_jl_b200 import _lib
    net = bn.rand_discrete_bn(n_nodes, 3, 5, seed=seed)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    table, _, _ = capi.sample(on, n_rows, seed=seed + 20)
    dn = bn.device_net(net)
    out = np.zeros(1)
    _lib.check(_lib.lib().bn_logpdf(_lib.h64(dn.handle), _lib.ptr(table, _lib.u8p), C.c_uint64(n_rows), _lib.ptr(out, _lib.f64p)))
    assert np.isfinite(out[0]) and out[0] < 0
    assert out[0] == pytest.approx(capi.logpdf_table(on, table), rel=1e-12)
    assert out[0] == pytest.approx(_exact_logl(lay, table), rel=1e-13)
    # the DataFrame front end (1-based states, by name) gives the same value
    if n_rows <= 50_001:
        df = pd.DataFrame({nm: table[i].astype(np.int64) + 1 for i, nm in enumerate(lay.names)})
        assert bn.logpdf(net, df) == out[0]
    bad = table.copy()
    bad[min(3, n_nodes - 1), n_rows // 2] = 250
    with pytest.raises(IndexError):
        _lib.check(_lib.lib().bn_logpdf(_lib.h64(dn.handle), _lib.ptr(bad, _lib.u8p), C.c_uint64(n_rows), _lib.ptr(out, _lib.f64p)))


def test_logpdf_zero_probability_rows_and_empty_table(bn):
    """A row the net gives probability 0 makes the sum -Inf (log(0) in the reference's loop); bins with cpt = 0 that
    no row visits contribute nothing; zero rows give 0.0."""
    import math
    from bayesnets_jl_b200 import _lib
    a = bn.DiscreteCPD("a", [0.5, 0.5, 0.0])
    b = bn.DiscreteCPD("b", ["a"], [3], [[1.0, 0.0], [0.25, 0.75], [0.5, 0.5]])
    net = bn.BayesNet([a, b])
    ok = pd.DataFrame({"a": [1, 2, 2, 1], "b": [1, 2, 1, 1]})
    assert bn.logpdf(net, ok) == pytest.approx(4 * math.log(0.5) + math.log(0.75) + math.log(0.25), rel=1e-14)
    assert bn.logpdf(net, pd.DataFrame({"a": [1, 2, 1], "b": [1, 2, 2]})) == -math.inf  # P(b=2 | a=1) = 0
    assert bn.logpdf(net, pd.DataFrame({"a": [3], "b": [1]})) == -math.inf            # P(a=3) = 0
    assert bn.pdf(net, pd.DataFrame({"a": [3], "b": [1]})) == 0.0
    dn = bn.device_net(net)
    out = np.full(1, 7.0)
    _lib.check(_lib.lib().bn_logpdf(_lib.h64(dn.handle), None, C.c_uint64(0), _lib.ptr(out, _lib.f64p)))
    assert out[0] == 0.0


def test_sample_then_logpdf_on_device(bn):
    """rand(bn, n) left in HBM -> bn_logpdf_dev: equals the oracle on the same table (aligned and unaligned row
    counts, pitched table), skips rejection-failed rows, and -logl / n_rows estimates the net's entropy."""
    import torch
    from bayesnets_jl_b200 import _lib
    net = bn.rand_discrete_bn(40, 3, 4, seed=3)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    dn = bn.device_net(net)
    L = _lib.lib()
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    per_row = []
    for n_rows in (1 << 19, (1 << 17) + 5):
        tab = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
        _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), None, C.c_int32(0), C.c_int32(1), C.c_uint64(0), C.c_uint64(n_rows),
                                   C.c_uint64(11), C.c_void_p(tab.data_ptr()), None, None))
        got = bn.logpdf_device(net, tab.data_ptr(), n_rows)
        host = tab.cpu().numpy()
        assert got == pytest.approx(capi.logpdf_table(on, host), rel=1e-12)
        assert got == pytest.approx(_exact_logl(lay, host), rel=1e-13)
        per_row.append(got / n_rows)
        # enqueue-only form writes the same double into a caller's device buffer
        out = torch.zeros(1, dtype=torch.float64, device="cuda:0")
        assert bn.logpdf_device(net, tab.data_ptr(), n_rows, out_dev_ptr=out.data_ptr()) is None
        assert float(out.cpu()[0]) == got
    assert per_row[0] == pytest.approx(per_row[1], rel=5e-3)  # both estimate -H(bn)
    # pitched table: rows 0..n_rows-1 of a wider allocation
    n_rows, pitch = 10_000, 10_240
    wide = torch.zeros((lay.n, pitch), dtype=torch.uint8, device="cuda:0")
    wide[:, :n_rows] = tab[:, :n_rows]
    got = bn.logpdf_device(net, wide.data_ptr(), n_rows, pitch=pitch)
    assert got == pytest.approx(capi.logpdf_table(on, np.ascontiguousarray(tab[:, :n_rows].cpu().numpy())), rel=1e-12)
    # rejection sampling with evidence that sometimes fails: failed rows (0xFF) are not part of the likelihood
    ev = lay.evidence_vector({lay.names[5]: 1, lay.names[17]: 2})
    n_rows = 1 << 16
    tab = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
    ok = torch.empty(n_rows, dtype=torch.uint8, device="cuda:0")
    _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), C.c_int32(1), C.c_int32(2), C.c_uint64(0),
                               C.c_uint64(n_rows), C.c_uint64(5), C.c_void_p(tab.data_ptr()), None, C.c_void_p(ok.data_ptr())))
    assert 0 < int(ok.sum().item()) < n_rows
    got = bn.logpdf_device(net, tab.data_ptr(), n_rows)
    assert got == pytest.approx(capi.logpdf_table(on, tab.cpu().numpy()), rel=1e-12)


def test_sharded_table_statistics_single_rank(bn):
    """distributed.ShardedTableStatistics at world size 1 (no communicator: the all-reduce is the identity): sample ->
    count -> logpdf on the device equals the oracle on the same table; a shard [first, first + count) of the global
    row space holds exactly those rows of the whole table (what the 2-rank tests and tools/multi_gpu_check.py rely on).
    No torch anywhere: the device buffers are bn_dev_alloc allocations."""
    from bayesnets_jl_b200 import DeviceBuffer
    from bayesnets_jl_b200.distributed import ShardedTableStatistics
    net = bn.rand_discrete_bn(40, 3, 4, seed=3)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    n_rows = 100_003
    sh = ShardedTableStatistics(net, n_rows, seed=21, device=0)
    sh.step()
    counts = sh.counts_host()
    whole, _, _ = capi.sample(on, n_rows, seed=21)
    assert np.array_equal(sh.table.download(np.uint8, lay.n * n_rows).reshape(lay.n, n_rows), whole)
    assert np.array_equal(counts, capi.statistics(on, whole))
    assert sh.logpdf() == pytest.approx(capi.logpdf_table(on, whole), rel=1e-12)
    # a middle shard of the same global row space
    sh.first, sh.count = 33_334, 33_335
    sh.pitch = sh.count
    sh.table = DeviceBuffer(lay.n * sh.count, 0)
    sh.sample()
    part = sh.table.download(np.uint8, lay.n * sh.count).reshape(lay.n, sh.count)
    assert np.array_equal(part, whole[:, 33_334:33_334 + 33_335])
    sh.launch()
    assert np.array_equal(sh.counts_host(), capi.statistics(on, np.ascontiguousarray(whole[:, 33_334:33_334 + 33_335])))


def test_count_reference_values(bn, golden):
    """Base.count: test/test_discrete_bayes_nets.jl:18-29 (a -> b, 8 rows) and test/test_learning.jl:5-38,85 (A -> B,
    A -> C, B -> C, 12 rows: 7 observed (A, B, C) rows, in order of first appearance)."""
    net = bn.BayesNet()
    net.push(bn.DiscreteCPD("a", [0.4, 0.6]))
    net.push(bn.DiscreteCPD("b", ["a"], [2], [[0.5, 0.5], [0.2, 0.8]]))
    data = pd.DataFrame({"a": [1, 1, 1, 1, 2, 2, 2, 2], "b": [1, 2, 1, 2, 1, 1, 1, 2]})
    T = bn.count(net, "a", data)
    assert list(T.columns) == ["a", "count"] and T["a"].tolist() == [1, 2] and T["count"].tolist() == [4, 4]
    T = bn.count(net, "b", data)
    assert len(T) == 4 and list(T.columns) == ["a", "b", "count"]
    got = {(int(r.a), int(r.b)): int(r.count) for r in T.itertuples()}
    assert got == {(1, 1): 2, (2, 1): 3, (1, 2): 2, (2, 2): 1}
    assert T[["a", "b"]].values.tolist() == [[1, 1], [1, 2], [2, 1], [2, 2]]  # unique(): order of first appearance
    g = golden["statistics_abc"]
    data = pd.DataFrame(g["data"])
    net3 = bn.fit(data, tuple(tuple(e) for e in g["edges"]))
    TA = bn.count(net3, "A", data)
    assert TA.shape == (3, 2) and TA["count"].tolist() == [4, 4, 4]
    TC = bn.count(net3, "C", data)
    assert TC.shape == (7, 4)
    rows = {(int(r.A), int(r.B), int(r.C)): int(r.count) for r in TC.itertuples()}
    assert rows == {(1, 1, 1): 3, (1, 2, 2): 1, (2, 2, 1): 2, (2, 2, 2): 1, (2, 1, 1): 1, (3, 1, 1): 3, (3, 2, 2): 1}
    assert len(bn.count(net3, data)) == 3
    assert int(TC["count"].sum()) == len(data)
