"""Host-side logic of the mirror (no GPU needed): containers, device layout, validation, formats, ABI surface."""
import ctypes
import io
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT, make_mirror_net, oracle_net_from_layout


def test_library_exports_every_declared_symbol(bn):
    """include/bn_cuda.h is the contract: the built .so must export each declared entry point, and the ctypes
    table must bind the same set."""
    hdr = open(os.path.join(ROOT, "include", "bn_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert os.path.exists(bn.LIB_PATH), "build libbn_cuda.so first (python bayesnets.jl_b200/build.py)"
    lib = ctypes.CDLL(bn.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in bn_cuda.h but not exported"
    assert declared == set(bn.EXPORTS)
    assert lib.bn_version() == 100


def test_no_cpu_fallback(bn):
    """Without a GPU every compute entry point must fail loudly, never route to the oracle."""
    if bn.device_count.__module__ and _has_gpu(bn):
        pytest.skip("GPU present")
    net = bn.get_asia_bn()
    with pytest.raises(bn.BNCudaError):
        bn.infer(bn.LikelihoodWeightingInference(10), net, "T")
    with pytest.raises(bn.BNCudaError):
        bn.Factor(["X"], np.ones(3))
    src = ""
    for f in os.listdir(os.path.join(ROOT, "bayesnets.jl_b200")):
        if f.endswith(".py"):
            src += open(os.path.join(ROOT, "bayesnets.jl_b200", f)).read()
    assert "import oracle" not in src and "from oracle" not in src


def _has_gpu(bn):
    try:
        return bn.device_count() > 0
    except Exception:
        return False


def test_topological_order_matches_graphs_dfs(bn):
    """src/bayes_nets.jl:26-48: reverse DFS post-order, re-run on every push! (:258-276).  Sorting a, b, c
    (c <- a, c <- b) in one go gives b, a, c; pushing them one at a time re-sorts [b, a] + c into a, b, c."""
    cpds = [bn.DiscreteCPD("a", [1.0, 0.0]), bn.DiscreteCPD("b", [0.0, 1.0]),
            bn.CategoricalCPD("c", ["a", "b"], [2, 2], [[0.1, 0.9], [0.2, 0.8], [1.0, 0.0], [0.4, 0.6]])]
    assert bn.BayesNet(cpds).names() == ["b", "a", "c"]
    net = make_mirror_net(bn, [["a", [], [[1.0, 0.0]]], ["b", [], [[0.0, 1.0]]],
                               ["c", ["a", "b"], [[0.1, 0.9], [0.2, 0.8], [1.0, 0.0], [0.4, 0.6]]]])
    assert net.names() == ["a", "b", "c"]
    assert net.parents("c") == ["a", "b"] and net.children("a") == ["c"]
    assert net.markov_blanket("a") == {"b", "c"}
    with pytest.raises(ValueError):
        net.push(bn.DiscreteCPD("a", [0.5, 0.5]))
    assert net.pdf({"a": 1, "b": 2, "c": 1}) == 1.0 and net.pdf(a=2, b=2, c=1) == 0.0


def test_device_layout_student(bn, golden):
    net = make_mirror_net(bn, golden["inference_student"]["nodes"])
    lay = bn.flatten(net)
    assert sorted(lay.names) == ["D", "G", "I", "L", "S"]
    g = lay.index["G"]
    assert lay.card[g] == 3 and [lay.names[p] for p in lay.parents(g)] == ["D", "I"]
    # first parent fastest: row q = d + 2*i (src/CPDs/categorical_cpd.jl:36-46)
    rows = golden["inference_student"]["nodes"][2][2]
    tab = lay.cpt[lay.cpt_off[g]:lay.cpt_off[g + 1]].reshape(3, 4, order="F")
    for q in range(4):
        assert tab[:, q].tolist() == rows[q]
    for i in range(lay.n):  # topological
        assert all(p < i for p in lay.parents(i))
    ev = lay.evidence_vector({"D": 1, "I": 2, "nope": 3})
    assert ev[lay.index["D"]] == 0 and ev[lay.index["I"]] == 1 and (ev >= 0).sum() == 2
    with pytest.raises(IndexError):
        lay.evidence_vector({"D": 3})
    with pytest.raises(TypeError):
        lay.evidence_vector({"D": "x"})


def test_inference_state_validation(bn):
    """test/test_inference.jl:5-31."""
    net = bn.rand_discrete_bn(seed=1)
    with pytest.raises(ValueError):
        bn.InferenceState(net, ["waldo", "N3"])
    with pytest.raises(ValueError):
        bn.InferenceState(net, ["N2", "N3"], {"N1": 1, "N3": 3, "N7": 2016})
    assert bn.InferenceState(net, ["N3", "N5", "N2", "N3"]).query == ["N3", "N5", "N2"]
    assert bn.InferenceState(net, "N6").query == ["N6"]
    with pytest.raises(ValueError):
        bn.InferenceState(net, "N1", {"N1": 2})


def test_rand_discrete_bn_invariants(bn):
    """test/test_genbn.jl:10-28."""
    net = bn.rand_discrete_bn(10, 4, 5, seed=0)
    assert set(net.names()) == {f"N{i}" for i in range(1, 11)}
    for n in net.names():
        assert 2 <= net.ncategories(n) <= 5 and len(net.parents(n)) <= 4
        for d in net.get(n).distributions:
            assert abs(d.sum() - 1) < 1e-12
    with pytest.raises(ValueError):
        bn.rand_bn_inference(net, 20, 16)
    q, ev = bn.rand_bn_inference(net, 3, 5, seed=0)
    assert len(q) == 3 and len(ev) == 5 and not set(q) & set(ev)
    big = bn.rand_discrete_bn(100, 3, 4, seed=0)
    lay = bn.flatten(big)
    assert lay.n == 100 and lay.card.max() <= 4 and np.diff(lay.par_ptr).max() <= 3
    assert (np.diff(lay.par_ptr)[1:] >= 1).all()  # connected=true: every node but the first has a parent


def write_xdsl(path, golden):
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<smile version="1.0" id="Unnamed">\n <nodes>\n')
        for c in golden["xdsl"]["cpts"]:
            f.write(f'  <cpt id="{c["id"]}">\n')
            for s in c["states"]:
                f.write(f'   <state id="{s}" />\n')
            if c["parents"]:
                f.write(f'   <parents>{" ".join(c["parents"])}</parents>\n')
            f.write(f'   <probabilities>{" ".join(repr(p) for p in c["probabilities"])}</probabilities>\n  </cpt>\n')
        f.write(" </nodes>\n</smile>\n")


def test_readxdsl_and_text_io(bn, golden, tmp_path):
    """test/test_io.jl:7-55."""
    path = str(tmp_path / "sample_bn.xdsl")
    write_xdsl(path, golden)
    net = bn.readxdsl(path)
    assert sorted(net.names()) == ["Forecast", "Success"]
    assert net.parents("Success") == [] and net.parents("Forecast") == ["Success"]
    for s, f, p in golden["xdsl"]["joint_1based"]:
        assert net.pdf(Success=s, Forecast=f) == pytest.approx(p, rel=1e-14)
    with pytest.raises(ValueError):
        bn.readxdsl(str(tmp_path / "x.txt"))
    net.push(bn.DiscreteCPD("Test", ["Forecast"], [3], [[1.0, 0.0], [0.0, 1.0], [0.1234569, 1 - 0.1234569]]))
    buf = io.StringIO()
    bn.write_text(buf, net)
    lines = buf.getvalue().splitlines()
    assert [l.strip() for l in lines] == golden["text_io"]["lines"]
    net2 = bn.read_text(io.StringIO(buf.getvalue()))
    for s in (1, 2):
        for f in (1, 2, 3):
            for t in (1, 2):
                assert net2.pdf(Success=s, Forecast=f, Test=t) == pytest.approx(net.pdf(Success=s, Forecast=f, Test=t))


def test_xdsl_reverses_parent_order(bn, tmp_path):
    """src/DiscreteBayesNet/io.jl:67: SMILE varies the FIRST listed parent slowest."""
    p = tmp_path / "two.xdsl"
    p.write_text("""<smile><nodes>
      <cpt id="A"><state id="1"/><state id="2"/><probabilities>0.3 0.7</probabilities></cpt>
      <cpt id="B"><state id="1"/><state id="2"/><state id="3"/><probabilities>0.2 0.3 0.5</probabilities></cpt>
      <cpt id="C"><state id="1"/><state id="2"/><parents>A B</parents>
        <probabilities>0.1 0.9 0.2 0.8 0.3 0.7 0.4 0.6 0.5 0.5 0.6 0.4</probabilities></cpt>
    </nodes></smile>""")
    net = bn.readxdsl(str(p))
    assert net.parents("C") == ["B", "A"]
    # SMILE order: (A=1,B=1),(A=1,B=2),(A=1,B=3),(A=2,B=1)...
    assert net.get("C")({"A": 1, "B": 3})[0] == 0.3
    assert net.get("C")({"A": 2, "B": 1})[0] == 0.4


def test_gibbs_sample_argument_validation(bn):
    """test/test_gibbs.jl:123-231 -- every check happens before the device is touched."""
    net = make_mirror_net(bn, [["a", [], [[1.0, 0.0]]], ["b", [], [[0.0, 1.0]]],
                               ["c", ["a", "b"], [[0.1, 0.9], [0.2, 0.8], [1.0, 0.0], [0.4, 0.6]]]])
    kw = dict(thinning=5)
    for bad in [dict(nsamples=0, burn_in=100), dict(nsamples=-1, burn_in=100), dict(nsamples=5, burn_in=-1)]:
        with pytest.raises(ValueError):
            bn.gibbs_sample(net, bad["nsamples"], bad["burn_in"], **kw)
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, thinning=-1)
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, variable_order=["a", "b"])
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, variable_order=["a", "b", "c", "d"])
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, time_limit=0)
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, initial_sample={"a": 1, "b": 2})
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, consistent_with={"c": 1}, initial_sample={"a": 1, "b": 2, "c": 2})
    with pytest.raises(ValueError):
        bn.gibbs_sample(net, 5, 100, initial_sample={"a": 2, "b": 2, "c": 1})  # pdf == 0
    with pytest.raises(ValueError):
        bn.infer(bn.GibbsSamplingNodewise(nsamples=100, burn_in=100), net, "a")


def test_shard_ranges_tile_exactly():
    from bayesnets_jl_b200.distributed import shard_range
    for n in [0, 1, 7, 10**11, 65536]:
        for world in [1, 2, 3, 4, 8]:
            nxt = 0
            for r in range(world):
                lo, cnt = shard_range(n, r, world)
                assert lo == nxt and cnt >= 0
                nxt = lo + cnt
            assert nxt == n
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_layout_feeds_oracle_identically(bn, golden, oracle):
    """The oracle consumes the very arrays that are uploaded to the device."""
    from oracle.factors_np import exact_infer
    net = make_mirror_net(bn, golden["asia"]["nodes"])
    on = oracle_net_from_layout(bn.flatten(net))
    np.testing.assert_allclose(exact_infer(on, "B", {"D": 1}).potential, [0.285265293007839, 0.714734706992161],
                               rtol=1e-12)


def test_specialised_lw_source_compiles_without_a_gpu(bn):
    """The network-specialised LW kernel is generated and NVRTC-compiled for sm_100a on the host; no device is
    needed for that, so the code generator is covered by the CPU suite (the numerics are covered by -m gpu)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from dump_lw_spec import spec_source
    net = bn.rand_discrete_bn(40, 3, 4, seed=3)
    lay = bn.flatten(net)
    evidence = {"N5": 1, "N17": 2}
    pruned, cub_p, live_p = spec_source(lay, evidence, ["N9", "N30"], prune=True)
    full, cub_f, live_f = spec_source(lay, evidence, ["N9", "N30"], prune=False)
    assert cub_p > 0 and cub_f > 0                      # NVRTC produced a cubin for sm_100a
    assert live_f == lay.n and 4 <= live_p <= live_f    # barren nodes dropped, query + evidence kept
    assert "lw_spec" in pruned and "cp.async.bulk" in pruned and "chk" in full and "chk" not in pruned
    # every state the query cell reads is defined
    for q in ("N9", "N30"):
        assert f"u32 s{lay.index[q]};" in pruned
    # > 16 query cells switches to shared-memory atomics
    big, _, _ = spec_source(lay, {}, ["N1", "N2", "N3", "N4", "N6"], prune=True, compile=False)
    ncell = int(np.prod([lay.card[lay.index[f"N{i}"]] for i in (1, 2, 3, 4, 6)]))
    assert ("atomicAdd(&bins[cell], w)" in big) == (ncell > 16)


def _gibbs_source(bn, lay, evidence, query, full=False, threads=256, compile=True):
    from bayesnets_jl_b200 import _lib
    import ctypes as C
    ev, q = lay.evidence_vector(evidence), lay.query_indices(query)
    n_len, cub = C.c_int64(0), C.c_int64(0)
    args = [C.c_int32(lay.n), _lib.ptr(lay.card, _lib.i32p), _lib.ptr(lay.par_ptr, _lib.i32p),
            _lib.ptr(lay.par_idx, _lib.i32p), _lib.ptr(lay.cpt_off, _lib.i64p), _lib.ptr(lay.cpt, _lib.f64p),
            _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p), C.c_int32(q.size), C.c_int32(1 if full else 0), C.c_int32(threads)]
    _lib.check(_lib.lib().bn_gibbs_spec_source(*args, None, C.c_int64(0), C.byref(n_len), None))
    buf = C.create_string_buffer(n_len.value + 1)
    _lib.check(_lib.lib().bn_gibbs_spec_source(*args, buf, C.c_int64(n_len.value + 1), C.byref(n_len),
                                               C.byref(cub) if compile else None))
    return buf.value.decode(), cub.value


def test_specialised_gibbs_source_tables_and_compile_without_a_gpu(bn, monkeypatch):
    """The Gibbs generator on the host (csrc/specialize.cu, csrc/gibbs_tables.cu layout): small-blanket nodes get the
    integer-threshold table block, the others the fp64 block; the record step is one out-of-line function; the Philox
    refill is inlined exactly when the number of free nodes is a multiple of 4; BN_GIBBS_TABLE_ROWS=0 removes every
    table block; and NVRTC compiles the result for sm_100a."""
    net = bn.rand_discrete_bn(30, 3, 4, seed=5)
    lay = bn.flatten(net)
    src, cub = _gibbs_source(bn, lay, {"N3": 1, "N11": 2}, ["N7"])      # 28 free nodes: a multiple of 4
    assert cub > 0 and "gibbs_spec" in src
    n_tab, n_fp = src.count("tabulated conditional"), src.count("own CPD")
    assert n_tab > 0 and n_tab + n_fp == 28                             # every free node has exactly one block
    assert "__ldg(tp)" in src or "const uint2 t = __ldg" in src or "const uint4 t = __ldg" in src
    assert src.count("__noinline__ void record(") == 1 and "#define RECORD record(" in src
    assert "#define REFILL_ATTR __forceinline__" in src
    src3, _ = _gibbs_source(bn, lay, {"N3": 1, "N11": 2, "N20": 1}, ["N7"], compile=False)  # 27 free nodes
    assert "#define REFILL_ATTR __forceinline__" not in src3
    monkeypatch.setenv("BN_GIBBS_TABLE_ROWS", "0")
    src0, _ = _gibbs_source(bn, lay, {"N3": 1, "N11": 2}, ["N7"], compile=False)
    assert "tabulated conditional" not in src0 and src0.count("own CPD") == 28
    monkeypatch.setenv("BN_GIBBS_TABLE_ROWS", "4")
    src4, _ = _gibbs_source(bn, lay, {"N3": 1, "N11": 2}, ["N7"], compile=False)
    assert 0 <= src4.count("tabulated conditional") <= n_tab           # a tighter limit tables fewer nodes


def test_sample_weighted_dataframe_scan(bn):
    """src/sampling.jl:182-197: cumulative scan `while c < u && i < n`; the last row absorbs rounding."""
    import pandas as pd
    df = pd.DataFrame({"a": [1, 2, 3], "b": [2, 2, 1], "potential": [0.2, 0.5, 0.3]})
    assert bn.sample_weighted_dataframe(df, u=0.0) == {"a": 1, "b": 2}
    assert bn.sample_weighted_dataframe(df, u=0.2) == {"a": 1, "b": 2}   # c < u is false at c == u
    assert bn.sample_weighted_dataframe(df, u=0.21) == {"a": 2, "b": 2}
    assert bn.sample_weighted_dataframe(df, u=0.71) == {"a": 3, "b": 1}
    assert bn.sample_weighted_dataframe(df, u=1.5) == {"a": 3, "b": 1}
    a = {"z": 9}
    assert bn.sample_weighted_dataframe(df, a) is a and a["z"] == 9 and set(a) == {"z", "a", "b"}
    with pytest.raises(IndexError):
        bn.sample_weighted_dataframe(df.iloc[:0])


def test_table_matches_reference_values(bn):
    """test/test_discrete_bayes_nets.jl:5-16: table(bn, :a), table(bn, :b) -- parents fastest, own state slowest."""
    net = bn.BayesNet()
    net.push(bn.DiscreteCPD("a", [0.4, 0.6]))
    net.push(bn.DiscreteCPD("b", ["a"], [2], [[0.5, 0.5], [0.2, 0.8]]))
    ta = bn.table(net, "a")
    assert list(ta.columns) == ["a", "potential"] and ta["a"].tolist() == [1, 2] and ta["potential"].tolist() == [0.4, 0.6]
    tb = bn.table(net, "b")
    assert list(tb.columns) == ["a", "b", "potential"]
    assert tb["a"].tolist() == [1, 2, 1, 2] and tb["b"].tolist() == [1, 1, 2, 2]
    assert tb["potential"].tolist() == [0.5, 0.2, 0.5, 0.8]
    assert bn.table(net, "b", a=2)["potential"].tolist() == [0.2, 0.8]
    assert bn.table(net, "b", {"a": 1, "b": 2})["potential"].tolist() == [0.5]


def test_mle_cpds_from_counts_match_reference_cpd_tests(bn, oracle):
    """test/test_cpds.jl:21-33,41-45,51-60: fit(DiscreteCPD, ...) = relative frequencies per parental instantiation.
    Host logic only (learning._cpds_from_counts); the counts come from the oracle here, from the statistics kernel
    on the device (tests/test_gpu_learning.py holds the two equal)."""
    from bayesnets_jl_b200.learning import _cpds_from_counts, _split, _structure_layout
    from conftest import oracle_net_from_layout

    def fit_counts(plist, ncat, mat, names):
        lay = _structure_layout(plist, ncat, names)
        counts = oracle.capi.statistics(oracle_net_from_layout(lay), np.ascontiguousarray(np.asarray(mat) - 1, np.uint8))
        return _cpds_from_counts(names, plist, ncat, _split(lay, counts))

    (cpd,) = fit_counts([[]], [3], [[1, 2, 1, 2, 3]], ["a"])
    assert cpd.parents == [] and cpd.distributions[0].tolist() == [0.4, 0.4, 0.2]
    (cpd,) = fit_counts([[]], [3], [[1, 2, 1, 2, 2]], ["a"])               # ncategories=3, state 3 never seen
    assert cpd.pdf({"a": 1}) == 0.4 and cpd.pdf({"a": 2}) == pytest.approx(0.6) and cpd.pdf({"a": 3}) == 0.0
    a, b = fit_counts([[], [0]], [3, 2], [[1, 2, 1, 2, 3], [1, 1, 2, 1, 2]], ["a", "b"])
    assert b.parents == ["a"] and b.parental_ncategories == [3]
    assert b({"a": 1}).tolist() == [0.5, 0.5] and b({"a": 2}).tolist() == [1.0, 0.0] and b({"a": 3}).tolist() == [0.0, 1.0]
    # parental_ncategories=[3], target_ncategories=5 -> 15 parameters; an instantiation never seen is uniform
    a, b = fit_counts([[], [0]], [3, 5], [[1, 2, 1, 2, 2], [1, 1, 2, 1, 2]], ["a", "b"])
    assert sum(len(d) for d in b.distributions) == 15
    assert b({"a": 3}).tolist() == [0.2] * 5


def test_cubin_disk_cache_serves_a_second_process(bn, tmp_path):
    """The specialised kernels' cubins persist in BN_CUBIN_CACHE keyed by a hash of the generated source (+ toolchain),
    with the source stored beside the cubin and compared on load: a second PROCESS compiles nothing (the cold-start
    path of bench.py's e2e_cached).  Host-only: NVRTC needs no device."""
    import json
    import subprocess
    code = (
        "import os, sys, json; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tools'))\n"
        "import bayesnets_jl_b200 as bn\n"
        "from dump_lw_spec import spec_source\n"
        "net = bn.rand_discrete_bn(25, 3, 4, seed=8); lay = bn.flatten(net)\n"
        "src, cub, live = spec_source(lay, {'N4': 1}, ['N9'], prune=True)\n"
        "src2, cub2, _ = spec_source(lay, {'N4': 1}, ['N9'], prune=True)\n"
        "print(json.dumps({'cubin': cub, 'same': cub == cub2, 'stats': bn.spec_cache_stats()}))\n" % (ROOT, ROOT))
    env = dict(os.environ, BN_CUBIN_CACHE=str(tmp_path / "cubins"))
    runs = []
    for _ in range(2):
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    first, second = runs
    assert first["cubin"] > 0 and first["same"] and first["cubin"] == second["cubin"]
    assert first["stats"]["nvrtc_compiles"] == 1 and first["stats"]["disk_hits"] == 0   # 2nd request: in-process table
    assert second["stats"]["nvrtc_compiles"] == 0 and second["stats"]["disk_hits"] == 1  # fresh process: from disk
    files = [f for f in os.listdir(tmp_path / "cubins") if f.endswith(".cubin")]
    assert len(files) == 1
    # a corrupted cache entry is ignored (source mismatch / short file), not trusted
    p = tmp_path / "cubins" / files[0]
    p.write_bytes(p.read_bytes()[:100])
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    third = json.loads(r.stdout.strip().splitlines()[-1])
    assert third["stats"]["nvrtc_compiles"] == 1 and third["stats"]["disk_hits"] == 0 and third["cubin"] == first["cubin"]


def test_evidence_state_beyond_int8_is_an_error_not_free(bn):
    """Evidence travels as int8 (-1 = free): a state above 128 of a wide node must raise, never wrap to 'free'."""
    rng = np.random.default_rng(0)
    net = bn.BayesNet()
    net.push(bn.CategoricalCPD("wide", [], [], [rng.dirichlet(np.ones(200))]))
    net.push(bn.CategoricalCPD("leaf", ["wide"], [200], [rng.dirichlet(np.ones(2)) for _ in range(200)]))
    lay = bn.flatten(net)
    assert lay.evidence_vector({"wide": 128})[0] == 127
    with pytest.raises(IndexError):
        lay.evidence_vector({"wide": 129})
    with pytest.raises(IndexError):
        lay.evidence_vector({"wide": 201})


def test_julia_ccalls_match_the_header():
    """julia/src/**: every `ccall((:bn_*, libbn_cuda), Ret, (Types...), args...)` must bind a prototype of
    include/bn_cuda.h with the same arity and C types, and so must the ctypes table (tools/abi_table.py).  Julia cannot
    run in this image; this is what keeps the glue a maintainer would add from drifting away from the ABI."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import abi_table
    protos, calls, errs = abi_table.check()
    assert not errs, "\n".join(errs)
    assert len(calls) >= 40 and len({c["name"] for c in calls}) >= 37
    # the entry points the hot path needs are all bound from Julia
    bound = {c["name"] for c in calls}
    for need in ("bn_net_create", "bn_net_create_comm", "bn_comm_init_all", "bn_comm_init_rank", "bn_lw_infer", "bn_gibbs_infer",
                 "bn_gibbs_sample", "bn_sample", "bn_exact_infer", "bn_factor_product", "bn_factor_sum", "bn_factor_normalize",
                 "bn_factor_eliminate", "bn_loopy_infer", "bn_statistics", "bn_logpdf"):
        assert need in bound, need
    # INTEGRATION.md carries the table generated from this parse
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert abi_table.markdown(protos, calls) in doc, "regenerate: python tools/abi_table.py (INTEGRATION.md section 2)"


def test_julia_and_python_hosts_expose_the_same_surface(bn):
    """The two host mirrors must not drift: the reference's keyword surface of gibbs_sample / GibbsSampler
    (src/gibbs.jl:289-297,410-433), its validation messages (:299-323) and the samplers are in both."""
    import inspect
    jl = open(os.path.join(ROOT, "julia", "src", "Inference", "CUDABackend", "CUDABackend.jl")).read()
    py = inspect.signature(bn.gibbs_sample).parameters
    for kw in ("thinning", "consistent_with", "variable_order", "time_limit", "error_if_time_out", "initial_sample",
               "max_cache_size"):
        assert kw in py and re.search(rf"\b{kw}::", jl), kw
    for msg in ("nsamples parameter less than 1", "Negative burn_in parameter", "Negative thinning parameter",
                "Gibbs sample variable_order must contain all variables in the Bayes Net",
                "Gibbs sample variable_order contains a variable not in the Bayes Net",
                "Gibbs sample initial_sample must be an assignment with all variables in the Bayes Net",
                "Gibbs sample initial_sample was inconsistent with consistent_with",
                "Gibbs sample initial_sample has a pdf value of zero"):
        assert msg in jl and msg in inspect.getsource(bn.gibbs_sample), msg
    for name in ("cuda_get_weighted_sample", "CUDAGibbsSampler", "CUDALikelihoodWeightedSampler", "CUDARejectionSampler",
                 "CUDADirectSampler", "CUDALikelihoodWeighting", "CUDAGibbsNodewise", "CUDAGibbsFull", "CUDAExact",
                 "CUDALoopyBelief", "Communicator"):
        assert name in jl, name
    for name in ("get_weighted_sample", "GibbsSampler", "LikelihoodWeightedSampler", "RejectionSampler", "DirectSampler",
                 "LikelihoodWeightingInference", "GibbsSamplingNodewise", "GibbsSamplingFull", "ExactInference", "LoopyBelief",
                 "Communicator"):
        assert hasattr(bn, name), name
    assert "Float64" in jl[jl.index("function cuda_gibbs_sample"):jl.index("mutable struct CUDAGibbsSampler")]  # evidence columns
