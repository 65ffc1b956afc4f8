"""Where does the host-side time of one public infer() call go?  (flatten / bn_net_create / launch+sync / destroy)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesnets_jl_b200 as bn
from bayesnets_jl_b200.device_layout import flatten, DeviceNet
import bench

net, lay, query, evidence = bench.build_case(bn)
bn.set_lw_mode("specialized", False)
def T(f, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3, r
for n in (1000, int(1e9)):
    im = bn.LikelihoodWeightingInference(nsamples=n, seed=1)
    bn.infer(im, net, query, evidence=evidence)
    print("n", n)
    print(" flatten ms", T(lambda: flatten(net))[:2])
    l = flatten(net)
    print(" DeviceNet create+destroy ms", T(lambda: DeviceNet(l).close())[:2])
    print(" infer cached net ms", T(lambda: bn.infer(im, net, query, evidence=evidence))[:2])
    def fresh():
        net._device_cache = None
        return bn.infer(im, net, query, evidence=evidence)
    print(" infer fresh net ms", T(fresh)[:2])
    st = bn.InferenceState(net, query, evidence)
    print(" raw_counts ms", T(lambda: im.raw_counts(st))[:2])
