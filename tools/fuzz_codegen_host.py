"""Offline check of the device code generator without a device (CPU only): the tick functors generated for N random graphs,
compiled for the host (see tests/test_codegen_host.py), against the oracle / the netlist evaluator, bit for bit.

    python tools/fuzz_codegen_host.py <seed> <N>"""
import sys, random, ctypes, subprocess, tempfile, os
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in ('', 'oracle', 'tests', os.path.join('tests', 'golden')): sys.path.insert(0, os.path.join(ROOT, _p))
import numpy as np
import zignal_b200 as zg, flowz_oracle as fo, netlist_flowz as nl
import test_codegen_host as th
from test_fuzz_frontend import _gen
rng=random.Random(int(sys.argv[1])); N=int(sys.argv[2])
consts=["0.5f","0.25f","-0.75f","0x1p-1f","1.5f","-1.0f","$0","$1"]
exprs=[]
while len(exprs)<N:
    e=_gen(rng, rng.randint(2,6), rng.randint(1,4), consts=consts, ops="+-*+-*/")
    try: g=zg.compile(e)
    except zg.ZgError: continue
    if g.all_f32 and g.n_in>=1 and g.n_out>=1 and g.n_state<=64: exprs.append(e)
d=tempfile.mkdtemp()
parts=[th.HOST_PRELUDE]+[th._tick_struct(zg,e,f"Tick{i}") for i,e in enumerate(exprs)]
parts.append('extern "C" void zg_host_run(int which, const float* const* in, float* const* out, long n, float* state, const float* params) {\n switch (which) {\n'+"".join(f" case {i}: run_tick<Tick{i}>(in,out,n,state,params); break;\n" for i in range(N))+" }\n}\n")
open(d+"/t.cpp","w").write("\n".join(parts))
subprocess.check_call(["g++","-std=c++17","-O1","-ffp-contract=off","-fPIC","-shared","-o",d+"/t.so",d+"/t.cpp"])
lib=ctypes.CDLL(d+"/t.so"); P=ctypes.POINTER(ctypes.c_float)
bad=ext=0; T=64
for i,e in enumerate(exprs):
    g=zg.compile(e)
    x=[np.ascontiguousarray(fo.noise(1,T,seed=9+7*i+k)[0]) for k in range(g.n_in)]
    y=[np.zeros(T,np.float32) for _ in range(g.n_out)]
    state=np.zeros(max(g.n_state,1),np.float32); params=np.array([0.5,-0.25],np.float32)
    for t0,n in ((0,23),(23,T-23)):
        ins=(P*max(g.n_in,1))(*[a[t0:].ctypes.data_as(P) for a in x]); outs=(P*g.n_out)(*[a[t0:].ctypes.data_as(P) for a in y])
        lib.zg_host_run(i,ins,outs,ctypes.c_long(n),state.ctypes.data_as(P),params.ctypes.data_as(P))
    try:
        o=fo.Oracle(e, params=list(params)); o.tick(*([0.0]*g.n_in))
        o=fo.Oracle(e, params=list(params))
        want=[np.zeros(T,np.float32) for _ in range(g.n_out)]
        for t in range(T):
            r=o.tick(*[float(a[t]) for a in x])
            for k in range(g.n_out): want[k][t]=r[k][1][0]
    except Exception:
        net=nl.Netlist(e, params=list(params)); ext+=1
        want=[np.zeros(T,np.float32) for _ in range(g.n_out)]
        for t in range(T):
            r=net.tick(*[float(a[t]) for a in x])
            for k in range(g.n_out): want[k][t]=r[k][1]
    for k in range(g.n_out):
        same=(y[k].view(np.uint32)==want[k].view(np.uint32))|(np.isnan(y[k])&np.isnan(want[k]))
        if not same.all():
            bad+=1
            if bad<6: print("BAD", e, k, np.argwhere(~same)[:3].ravel().tolist(), y[k][~same][:3], want[k][~same][:3])
            break
print(N,"graphs",ext,"beyond the reference",bad,"bad")
