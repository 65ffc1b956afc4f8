"""World-size-2 `gloo` test of the multi-GPU plumbing on CPU.

The sharding contract is host logic: rank g runs particle ids [g*N//G, (g+1)*N//G) of ONE global Philox counter space
and the ranks all-reduce the (cells + 2) doubles.  On the GPU box both halves live inside libbn_cuda.so (csrc/comm.cu:
bn_comm_shard + ncclAllReduce).  Here the shard boundaries come from the library's own bn_comm_shard (host-only, no GPU
needed), each rank's shard is computed by the oracle (the GPU kernels are bit-identical to it per
tests/test_gpu_sampling.py), the reduction is gloo's, and the result is compared with the single-process one: integer
(unweighted) counts must match exactly, weighted sums to 1e-12.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loopy_pairs(lay):
    rng = np.random.default_rng(3)
    pairs = []
    for _ in range(7):  # odd: ragged shards
        picks = rng.choice(lay.n, size=3, replace=False)
        ev = np.full(lay.n, -1, np.int8)
        for i in picks[1:]:
            ev[i] = rng.integers(0, lay.card[i])
        pairs.append((int(picks[0]), ev))
    return pairs


def _worker(rank, world, port, n_total, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        import bayesnets_jl_b200 as bn
        from bayesnets_jl_b200.distributed import allreduce_sum_, comm_shard
        from bayesnets_jl_b200.distributed import shard_range as py_shard_range

        def shard_range(n, r, w):  # the C library's rule, checked against its Python restatement
            lo, cnt = comm_shard(0, n, r, w)
            assert (lo, cnt) == py_shard_range(n, r, w)
            return lo, cnt
        from oracle import capi
        from oracle.factors_np import FlatNet
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        net = bn.rand_discrete_bn(40, 3, 4, seed=5)
        lay = bn.flatten(net)
        on = FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))
        evidence, query = {"N3": 1, "N20": 2}, ["N7", "N33"]
        ev, qi = lay.evidence_vector(evidence), lay.query_indices(query)
        first, count = shard_range(n_total, rank, world)
        c, sw, sw2 = capi.lw_infer(on, ev, qi, count, seed=9, first=first, mode=capi.DEVICE_SPEC)
        t = torch.from_numpy(np.concatenate([c.ravel(order="F"), [sw, sw2]]))
        allreduce_sum_(t)
        # Gibbs: chains sharded the same way, integer counts
        gfirst, gcount = shard_range(64, rank, world)
        gc, _ = capi.gibbs_infer(on, ev, qi, gcount, 300, 100, 3, seed=4, first_chain=gfirst)
        g = torch.from_numpy(gc.astype(np.float64).ravel(order="F").copy())
        allreduce_sum_(g)
        # LoopyBelief: (query, evidence) pairs sharded the same way, beliefs all-gathered in pair order
        from bayesnets_jl_b200.distributed import allgather_rows_
        from oracle.loopy_np import loopy_infer
        pairs = _loopy_pairs(lay)
        lfirst, lcount = shard_range(len(pairs), rank, world)
        rows = np.zeros((lcount, int(lay.card.max())))
        for r, (qn, e) in enumerate(pairs[lfirst:lfirst + lcount]):
            o = loopy_infer(on, qn, e, nsamples=8)
            rows[r, :len(o)] = o
        lb = allgather_rows_(torch.from_numpy(rows), [shard_range(len(pairs), g, world)[1] for g in range(world)])
        # table rows sharded the same way: rank g draws rows [first, first + count) of one global row index space and
        # counts them; the int64 counts are all-reduced (ShardedTableStatistics' contract)
        tfirst, tcount = shard_range(10_001, rank, world)
        rows_u8, _, _ = capi.sample(on, tcount, seed=6, first=tfirst)
        tc = torch.from_numpy(capi.statistics(on, rows_u8))
        allreduce_sum_(tc)
        if rank == 0:
            q.put((t.numpy().copy(), g.numpy().copy(), lb.numpy().copy(), tc.numpy().copy()))
        dist.destroy_process_group()
    except Exception as e:  # surface worker failures in the parent
        if rank == 0:
            q.put(e)
        raise


def test_two_rank_gloo_allreduce_equals_single_stream():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import bayesnets_jl_b200 as bn
    from oracle import capi
    from oracle.factors_np import FlatNet
    capi.build()
    n_total = 100_001  # odd: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if isinstance(got, Exception):
        raise got
    lw2, g2, lb2, tc2 = got
    net = bn.rand_discrete_bn(40, 3, 4, seed=5)
    lay = bn.flatten(net)
    on = FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))
    ev, qi = lay.evidence_vector({"N3": 1, "N20": 2}), lay.query_indices(["N7", "N33"])
    c, sw, sw2 = capi.lw_infer(on, ev, qi, n_total, seed=9, mode=capi.DEVICE_SPEC)
    np.testing.assert_allclose(lw2[:-2], c.ravel(order="F"), rtol=1e-12)
    assert lw2[-2] == pytest.approx(sw, rel=1e-12) and lw2[-1] == pytest.approx(sw2, rel=1e-12)
    gc, _ = capi.gibbs_infer(on, ev, qi, 64, 300, 100, 3, seed=4)
    assert np.array_equal(g2, gc.astype(np.float64).ravel(order="F"))
    whole, _, _ = capi.sample(on, 10_001, seed=6)
    assert tc2.dtype == np.int64 and np.array_equal(tc2, capi.statistics(on, whole))
    from oracle.loopy_np import loopy_infer
    pairs = _loopy_pairs(lay)
    assert lb2.shape == (len(pairs), int(lay.card.max()))
    for b, (qn, e) in enumerate(pairs):
        o = loopy_infer(on, qn, e, nsamples=8)
        assert np.array_equal(lb2[b, :len(o)], o) and not lb2[b, len(o):].any()


def test_comm_api_on_cpu():
    """Without a GPU the communicator entry points fail loudly (no fallback); the sharding rule is host-only."""
    sys.path.insert(0, ROOT)
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200.distributed import Communicator, comm_shard
    assert comm_shard(5, 10, 1, 4) == (5 + 2, 3)  # [2, 5) of 10 -> 3 indices
    assert sum(comm_shard(0, 10**11 + 7, g, 8)[1] for g in range(8)) == 10**11 + 7
    with pytest.raises(ValueError):
        comm_shard(0, 10, 4, 4)
    if bn.device_count_or_zero() == 0:
        with pytest.raises(bn.BNCudaError):
            Communicator.init_all([0])


# ------------------------------------------------------------------------------------------ split factors (8e)
def _np_backend():
    """The numpy oracle's factors behind ShardedFactor's backend interface (CPU, gloo)."""
    import torch
    import torch.distributed as dist
    from bayesnets_jl_b200.sharded_factors import FactorBackend

    class NpBackend(FactorBackend):
        def slice(self, f, assignment):
            return f.slice(assignment)

        def allreduce_sum(self, f):
            t = torch.from_numpy(f.potential) if f.potential.ndim else torch.from_numpy(f.potential.reshape(1))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return f

        def norm(self, f, p):
            return float(np.abs(f.potential).sum() if p == 1 else (f.potential ** 2).sum())

        def div_scalar(self, f, total):
            f.potential /= total
            return f
    return NpBackend()


def _split_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from bayesnets_jl_b200.sharded_factors import ShardedFactor
        from oracle.factors_np import NpFactor
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        be = _np_backend()
        rng = np.random.default_rng(3)  # same operands on every rank
        cards = {"a": 3, "b": 2, "c": 4, "d": 2, "e": 3, "z": 2, "y": 2}
        A = NpFactor(["a", "b", "c", "d", "y", "z"], rng.random([cards[k] for k in "abcdyz"]))
        B = NpFactor(["c", "z", "e"], rng.random([4, 2, 3]))        # mentions the split variable z
        Cc = NpFactor(["b", "a"], rng.random([2, 3]))                # does not
        D = NpFactor(["d", "y", "z", "a"], rng.random([2, 2, 2, 3]))  # split alike below
        split = ["y", "z"] if world == 4 else ["z"]
        sa = ShardedFactor.from_replicated(A, split, rank, world, be)
        sd = ShardedFactor.from_replicated(D, split, rank, world, be)
        prod = (sa * B) * Cc * sd                                     # local everywhere
        part = prod.sum(["c", "a"])                                   # local
        fused = sa.eliminate([B, Cc, sd], ["c"])                      # local, one pass
        total = part.sum(split)                                       # THE exchange step: all-reduce
        norm = (sa * B).normalize_()
        out = {"prod": prod.gather(), "prod_dims": prod.dimensions, "part": part.gather(), "part_dims": part.dimensions,
               "fused": fused.gather(), "fused_dims": fused.dimensions, "total": total.potential.copy(),
               "total_dims": list(total.dimensions), "norm": norm.gather(), "norm_dims": norm.dimensions}
        if rank == 0:
            q.put(out)
        dist.destroy_process_group()
    except Exception as e:
        if rank == 0:
            q.put(e)
        raise


@pytest.mark.parametrize("world", [2, 4])
def test_split_factor_algebra_equals_single_process(world):
    """A factor split over 2 / 4 ranks along its outermost variables: local products (replicated operands sliced at
    the rank's state, operands split alike), local sums, the fused eliminate, the all-reduce sum over the split
    variables and the two-step normalise all equal the unsplit computation (products exactly, sums to 1e-12)."""
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle.factors_np import NpFactor
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if isinstance(got, Exception):
        raise got
    rng = np.random.default_rng(3)
    cards = {"a": 3, "b": 2, "c": 4, "d": 2, "e": 3, "z": 2, "y": 2}
    A = NpFactor(["a", "b", "c", "d", "y", "z"], rng.random([cards[k] for k in "abcdyz"]))
    B = NpFactor(["c", "z", "e"], rng.random([4, 2, 3]))
    Cc = NpFactor(["b", "a"], rng.random([2, 3]))
    D = NpFactor(["d", "y", "z", "a"], rng.random([2, 2, 2, 3]))
    split = ["y", "z"] if world == 4 else ["z"]

    def same(arr, dims, ref, exact):
        perm = [ref.dimensions.index(d) for d in dims]
        want = np.transpose(ref.potential, perm)
        if exact:
            assert np.array_equal(arr, want)
        else:
            np.testing.assert_allclose(arr, want, rtol=1e-12)
    prod = ((A * B) * Cc) * D
    same(got["prod"], got["prod_dims"], prod, True)
    same(got["part"], got["part_dims"], prod.sum(["c", "a"]), False)
    same(got["fused"], got["fused_dims"], prod.sum(["c"]), False)
    same(got["total"], got["total_dims"], prod.sum(["c", "a"] + split), False)
    nrm = A * B
    same(got["norm"], got["norm_dims"], NpFactor(nrm.dimensions, nrm.potential / np.abs(nrm.potential).sum()), False)
