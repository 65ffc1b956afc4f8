"""GPU parity of the multi-GPU path that lives inside libbn_cuda.so (csrc/comm.cu; include/bn_cuda.h "multi-GPU").

infer(im, bn, query; evidence) is one call for the whole machine (src/ProbabilisticGraphicalModels/inference.jl:48-49):
on a net created with bn_net_create_comm the library shards the global particle / chain range over the GPUs of its NCCL
communicator and all-reduces the counts itself.  Checked here:
  * a 1-GPU communicator (runs on any GPU box): same counts as a plain net and as the oracle;
  * ONE process driving 2 GPUs (ncclCommInitAll) and one process PER GPU (ncclCommInitRank, id shipped through a file):
    LW counts equal to the 1-GPU counts to 1e-12 (the same particles, fp64 addition order differs), Gibbs counts
    identical (integers) -- skipped when fewer than 2 devices are visible.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import oracle_net_from_layout
from oracle import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def case(bn, n_nodes=60, seed=11):
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=seed)
    lay = bn.flatten(net)
    on = oracle_net_from_layout(lay)
    rng = np.random.default_rng(500 + seed)
    while True:
        picks = [str(x) for x in rng.choice(lay.names, size=6, replace=False)]
        query, evn = picks[:2], picks[2:]
        evidence = {e: int(rng.integers(1, net.ncategories(e) + 1)) for e in evn}
        _, sw, _ = capi.lw_infer(on, lay.evidence_vector(evidence), lay.query_indices(query), 2000, seed=1)
        if sw > 0:
            return net, lay, on, query, evidence


def run_jobs(bn, comm, net, query, evidence, n, n_chains):
    """LW counts (+ sum w, sum w^2) and Gibbs counts of one job, through the host-buffer entry points."""
    from bayesnets_jl_b200.device_layout import device_net
    im = bn.LikelihoodWeightingInference(nsamples=n, seed=5)
    g = bn.GibbsSamplingNodewise(nsamples=500, burn_in=100, thin=3, n_chains=n_chains, seed=6, keep_chain_counts=True)
    inf = bn.InferenceState(net, query, evidence)
    bn.set_default_comm(comm)
    try:
        assert (device_net(net).comm is comm)
        c = im.raw_counts(inf)
        gc = g.raw_counts(inf)
    finally:
        bn.set_default_comm(None)
    return c, im.last_stats, gc, g


@pytest.mark.parametrize("mode", ["generic", "specialized"])
def test_one_gpu_communicator_equals_plain_net_and_oracle(bn, mode):
    from bayesnets_jl_b200.distributed import Communicator
    net, lay, on, query, evidence = case(bn)
    bn.set_lw_mode(mode, True)
    bn.set_gibbs_mode(mode)
    try:
        comm = Communicator.init_all([0])
        assert (comm.world, comm.first_rank, comm.n_local) == (1, 0, 1)
        n, nch = 300_001, 300
        c, (sw, sw2), gc, g = run_jobs(bn, comm, net, query, evidence, n, nch)
        c1, st1, gc1, g1 = run_jobs(bn, None, net, query, evidence, n, nch)
    finally:
        bn.set_lw_mode("auto", True)
        bn.set_gibbs_mode("auto")
    evv, qi = lay.evidence_vector(evidence), lay.query_indices(query)
    oc, osw, osw2 = capi.lw_infer(on, evv, qi, n, seed=5, mode=capi.DEVICE_SPEC)
    np.testing.assert_allclose(c, oc.ravel(order="F"), rtol=1e-12)
    np.testing.assert_allclose(c, c1, rtol=1e-12)
    assert sw == pytest.approx(osw, rel=1e-12) and sw2 == pytest.approx(osw2, rel=1e-12)
    ogc, ost = capi.gibbs_infer(on, evv, qi, nch, 500, 100, 3, seed=6)
    assert np.array_equal(gc, ogc.ravel(order="F")) and np.array_equal(gc, gc1)
    assert np.array_equal(g.chain_states, ost) and np.array_equal(g.last_chain_counts, g1.last_chain_counts)
    comm.barrier()
    x = np.array([3.0, 4.0])
    assert np.array_equal(comm.allreduce_(x, "max"), [3.0, 4.0])
    comm.close()


def test_shard_rule_tiles_the_range(bn):
    """bn_comm_shard (host-only) == the Python restatement; shards tile [first, first + n) for ragged n."""
    from bayesnets_jl_b200.distributed import comm_shard, shard_range
    for n, world in [(100_001, 2), (10**11, 8), (7, 8), (0, 4), (2**63 + 12345, 3)]:
        at = 17
        for g in range(world):
            lo, cnt = comm_shard(17, n, g, world)
            assert (lo - 17, cnt) == shard_range(n, g, world)
            assert lo == at
            at += cnt
        assert at == 17 + n


def _two_gpus(bn):
    if bn.device_count() < 2:
        pytest.skip("needs 2 visible GPUs")


def test_one_process_two_gpus_equals_one_gpu(bn):
    """ncclCommInitAll: one host thread drives both GPUs; state / per-chain rows come back from both."""
    _two_gpus(bn)
    from bayesnets_jl_b200.distributed import Communicator
    net, lay, on, query, evidence = case(bn)
    comm = Communicator.init_all([0, 1])
    assert (comm.world, comm.n_local) == (2, 2)
    for mode in ("generic", "specialized"):
        bn.set_lw_mode(mode, True)
        bn.set_gibbs_mode(mode)
        try:
            n, nch = 400_001, 301  # odd: ragged shards
            c, (sw, sw2), gc, g = run_jobs(bn, comm, net, query, evidence, n, nch)
            c1, (sw_1, sw2_1), gc1, g1 = run_jobs(bn, None, net, query, evidence, n, nch)
        finally:
            bn.set_lw_mode("auto", True)
            bn.set_gibbs_mode("auto")
        np.testing.assert_allclose(c, c1, rtol=1e-12)
        assert sw == pytest.approx(sw_1, rel=1e-12) and sw2 == pytest.approx(sw2_1, rel=1e-12)
        assert np.array_equal(gc, gc1)
        assert np.array_equal(g.chain_states, g1.chain_states) and np.array_equal(g.last_chain_counts, g1.last_chain_counts)
    comm.close()


def test_process_per_gpu_equals_one_gpu(bn, tmp_path):
    """ncclCommInitRank: two processes (RANK / WORLD_SIZE / LOCAL_RANK like torchrun), the NCCL id shipped through a
    file; every rank gets the whole job's counts."""
    _two_gpus(bn)
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29711", BN_COMM_DIR=str(tmp_path),
               BN_COMM_JOB=f"pytest{os.getpid()}")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--json", "--quick"],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["world"] == 2 and res["lw_equal_single_gpu_1e12"] and res["gibbs_counts_identical"] and res["table_counts_identical"]
