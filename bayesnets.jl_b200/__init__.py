"""bayesnets.jl_b200 -- B200-native (sm_100a) discrete Bayesian-network inference behind the BayesNets.jl API.

Only the data-parallel inference hot path is here: likelihood weighting, ancestral /
rejection / weighted sampling, batched Gibbs, and the Factor product / sum / normalize / slice kernels that
drive exact variable elimination -- hand-written CUDA in csrc/, exported through the C ABI of
include/bn_cuda.h (libbn_cuda.so).  This Python package mirrors the reference's Julia surface one for one
(same names, argument meaning, 1-based states, error behaviour) because Julia is not available in the build
image; julia/ holds the equivalent `ccall` glue a BayesNets.jl maintainer would add.

There is NO CPU fallback: importing works anywhere, but any compute call raises unless libbn_cuda.so is
built and a B200 is visible.
"""
from ._lib import (BNCudaError, DeviceBuffer, DimensionMismatch, EXPORTS, LIB_PATH, device_count, device_count_or_zero, launch_count, pool_trim,  # noqa: F401
                   set_gibbs_mode, set_lw_mode, set_stream, spec_cache_stats, synchronize)
from .bayes_net import (Assignment, BayesNet, CategoricalCPD, DiscreteBayesNet, DiscreteCPD, consistent,  # noqa: F401
                        get_asia_bn, get_sat_fail_bn, get_sprinkler_bn, rand_bn_inference, rand_cpd,
                        rand_discrete_bn)
from .device_layout import (DeviceLayout, DeviceNet, default_comm, device_net, flatten, set_default_comm,  # noqa: F401
                            set_default_device)
from .distributed import Communicator  # noqa: F401
from .factors import Factor, eliminate  # noqa: F401
from .inference import (ExactInference, GibbsSamplingFull, GibbsSamplingNodewise, InferenceMethod,  # noqa: F401
                        InferenceState, LikelihoodWeightingInference, LoopyBelief, infer)
from .io import read_text, readxdsl, write_text  # noqa: F401
from .learning import (BDeuPrior, DirichletPrior, UniformPrior, bayesian_score, bayesian_score_component,  # noqa: F401
                       bayesian_score_components, count, fit, fit_from_device_table, logpdf, logpdf_device, pdf, statistics,
                       table,
                       statistics_device)
from .sampling import (BayesNetSampler, DirectSampler, GibbsSampler, LikelihoodWeightedSampler,  # noqa: F401
                       RejectionSampler, get_weighted_dataframe, get_weighted_sample, gibbs_sample, rand,
                       sample_weighted_dataframe)

__all__ = [n for n in dir() if not n.startswith("_")]
