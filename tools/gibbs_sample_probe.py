import ctypes as C, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bayesnets_jl_b200 as bn
from bayesnets_jl_b200 import _lib
from bayesnets_jl_b200.device_layout import device_net
import bench
net, lay, query, evidence = bench.build_case(bn)
dn = device_net(net, 0); L = _lib.lib()
bn.set_gibbs_mode(os.environ.get("BN_GIBBS_KERNEL", "specialized"))
ev = lay.evidence_vector(evidence)
for nch, ns, bi, th in [(4096,16,50,1),(4096,16,0,0),(4096,1,0,0),(65536,16,50,1),(65536,1,0,0),(65536,16,0,0)]:
    buf = np.zeros((lay.n, nch*ns), np.uint8)
    def run():
        _lib.check(L.bn_gibbs_sample(_lib.h64(dn.handle), _lib.ptr(ev,_lib.i8p), C.c_uint64(0), C.c_uint64(nch), C.c_int64(ns), C.c_int64(bi), C.c_int64(th), None, C.c_uint64(5), None, _lib.ptr(buf,_lib.u8p)))
    run(); run()
    ts=[]
    for _ in range(5):
        t0=time.perf_counter(); run(); ts.append(time.perf_counter()-t0)
    sweeps = bi + ns*(th+1)
    print(nch, ns, bi, th, "ms", round(min(ts)*1e3,3), "updates/s %.3g" % (nch*sweeps*90/min(ts)))
