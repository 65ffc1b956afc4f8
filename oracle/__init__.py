"""CPU oracle for the BayesNets.jl discrete-inference hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``bayesnets.jl_b200``) never does; it fails loudly when
``libbn_cuda.so`` is missing instead of falling back to anything here.

Two halves:
  * ``oracle.capi``       ctypes view of ``libbn_oracle.so`` (``bn_oracle.c``): samplers, Gibbs, and the
                          C factor ops that are timed as the CPU baseline
  * ``oracle.factors_np`` numpy restatement of ``src/Factors`` + ``ExactInference`` + brute-force
                          joint enumeration (independent second opinion on the C code)
  * ``oracle.loopy_np``   numpy restatement of ``infer(::LoopyBelief)`` that keeps the reference's message aliasing
                          literally (``capi.loopy_infer`` is the C twin, bit-identical, timed as the CPU baseline)
  * ``oracle.scoring_py`` Dirichlet priors + ``bayesian_score_component`` on the count tables of
                          ``capi.statistics`` (``structure_scoring.jl:17-42``), pinned to the five score
                          values of ``test/test_discrete_bayes_nets.jl:40-44``

Parity status: the reference cannot run in this image (no Julia; deps un-vendored, ``Project.toml:5-25``),
so the oracle is pinned against the reference's golden vectors (``tests/test_oracle_golden.py``).
Random-stream parity is unpinned by the reference's own tests -> statistical checks.
"""
from . import capi, factors_np  # noqa: F401
