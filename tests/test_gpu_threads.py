"""GPU: the threading contract of include/bn_cuda.h -- any number of host threads / streams may use one net (and the
library's caches) concurrently.  Every call is deterministic given its seed, so each thread's results must equal the
ones computed serially before the threads start.  Covers the per-call staging of the samplers, the locked kernel cache
(NVRTC-specialised and interpreter kernels), the Gibbs conditional tables (built once per net under the net's mutex) and
the block cache in front of the stream-ordered pool (factor tables allocated and freed from many threads at once).
"""
import threading

import numpy as np
import pytest

from oracle.factors_np import NpFactor
from test_gpu_sampling import random_case

pytestmark = pytest.mark.gpu


def _work(bn, net, query, evidence, seed, fa_names, A):
    """One thread's job: LW, Gibbs and a chain of factor ops; returns everything it computed."""
    lw = bn.LikelihoodWeightingInference(nsamples=150_000, seed=1000 + seed)
    c_lw = lw.raw_counts(bn.InferenceState(net, query, evidence))
    gb = bn.GibbsSamplingNodewise(nsamples=600, burn_in=100, thin=3, n_chains=96, seed=2000 + seed)
    c_gb = gb.raw_counts(bn.InferenceState(net, query, evidence))
    f = bn.Factor(fa_names, A)
    d = (seed % (len(fa_names) - 2)) + 1
    g = (f * f.sum([fa_names[d]])).sum([fa_names[0], fa_names[-1]])
    g.normalize_()
    return c_lw, lw.last_stats, c_gb, gb.chain_states.copy(), g.potential


@pytest.mark.parametrize("mode", ["specialized", "generic"])
def test_concurrent_host_threads_equal_serial_results(bn, mode):
    import torch
    bn.set_lw_mode(mode, True)
    bn.set_gibbs_mode(mode)
    try:
        net, lay, on, query, evidence = random_case(bn, 40, 17, 4)
        rng = np.random.default_rng(99)
        k = 16
        names = [f"th{mode}_{i}" for i in range(k)]
        A = rng.random((2,) * k)
        n_threads, rounds = 8, 4
        expected = [_work(bn, net, query, evidence, s, names, A) for s in range(n_threads)]
        # the factor chain against the oracle once (the threads then compare with `expected`)
        na = NpFactor(list(range(k)), A)
        d0 = (0 % (k - 2)) + 1
        ref = (na * na.sum(d0)).sum([0, k - 1])
        np.testing.assert_allclose(expected[0][4], ref.potential / ref.potential.sum(), rtol=1e-12)
        errors = []

        def run(seed):
            try:
                stream = torch.cuda.Stream()
                bn.set_stream(stream.cuda_stream)  # thread-local
                bn.set_lw_mode(mode, True)         # kernel selection is thread-local too
                bn.set_gibbs_mode(mode)
                for _ in range(rounds):
                    got = _work(bn, net, query, evidence, seed, names, A)
                    stream.synchronize()
                    for name, a, b in zip(("lw counts", "lw stats", "gibbs counts", "gibbs states", "factor chain"),
                                          got, expected[seed]):
                        a, b = np.asarray(a), np.asarray(b)
                        # LW sums fp64 weights with atomics (order varies run to run): 1e-12; everything else is exact
                        same = np.allclose(a, b, rtol=1e-12, atol=0) if name.startswith("lw") else np.array_equal(a, b)
                        if not same:
                            raise AssertionError(f"thread {seed}: {name} differ from the serial result")
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))
            finally:
                bn.set_stream(None)

        ts = [threading.Thread(target=run, args=(s,)) for s in range(n_threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors
    finally:
        bn.set_lw_mode("auto", True)
        bn.set_gibbs_mode("auto")
