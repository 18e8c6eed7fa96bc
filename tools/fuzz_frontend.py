"""Offline fuzzing of the host front end (CPU only): random flowz graphs with parameters, divisions and mixed int / float /
double arguments; product vs oracle on validity, arity, delays, canonical form and ticks, product vs the netlist evaluator
(tests/netlist_flowz.py) for the graphs the reference cannot compile.

    python tools/fuzz_frontend.py <first seed> <seconds>

The committed test (tests/test_fuzz_frontend.py) is the small deterministic version of this."""
import sys, random, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in ('', 'oracle', 'tests', os.path.join('tests', 'golden')): sys.path.insert(0, os.path.join(ROOT, _p))
import zignal_b200 as zg, flowz_oracle as fo, netlist_flowz as nl
import test_fuzz_frontend as tf
consts = tf.CONSTS + ["$0", "$1", "$0", "1e-3", "7"]
PV = [0.5, -0.25]
bad=tot=ext=0
t0=time.time()
seed0=int(sys.argv[1]); budget=float(sys.argv[2])
seed=seed0
while time.time()-t0 < budget:
    seed+=1
    rng=random.Random(seed)
    for _ in range(20):
        expr=tf._gen(rng, rng.randint(1,6), rng.randint(1,5), consts=consts)
        n_par = 2 if "$" in expr else 0
        try:
            g=zg.compile(expr); pe=None
        except zg.ZgError as e:
            g=None; pe="limit" if e.status==zg.ZG_ERR_UNSUPPORTED else str(e)
        if pe=="limit": continue
        try:
            o=fo.Oracle(expr, params=PV); n_in=fo.input_arity(fo.parse(expr))
            res=fo.Oracle(expr, params=PV).tick(*([0.0]*n_in))
            if any(r is fo.BOTTOM for r in res): raise ValueError("fed-back wire nothing assigns")
            oe=None
        except Exception as e:
            o=None; oe=str(e)
        tot+=1
        try:
            if g is not None:
                v=g.voice()
                for k in range(g.n_params): v.set_param(k, PV[k])
                dt=[rng.choice([0,1,2]) for _ in range(g.n_in)]
            if g is not None and o is None:
                ext+=1
                net=nl.Netlist(expr, params=PV)
                for t in range(5):
                    xs=[float(rng.randint(-3,3)) for _ in range(g.n_in)]
                    r=net.tick(*xs, dtype=dt)
                    assert tf._same(tuple(float(y) for y in v.tick(*xs, dtypes=dt)), tuple(float(x) for _,x in r)), "ext values"
                    assert v.out_dtypes==tuple(d for d,_ in r), "ext dtypes"
                continue
            assert (g is None)==(o is None), ("validity", pe, oe)
            if g is None: continue
            e=fo.parse(expr)
            assert zg.arity(expr)==(fo.input_arity(e), fo.output_arity(e))
            assert zg.delays(expr)==fo.max_input_delays(e) and zg.delays(expr,minimum=True)==fo.min_input_delays(e)
            c=zg.canonical(expr); assert c==zg.canonical(str(fo.make_canonical(e))) and zg.canonical(c)==c
            for t in range(5):
                xs=[float(rng.randint(-3,3)) for _ in range(g.n_in)]
                r=o.tick(*xs, dtype=dt)
                want=tuple(float(val[0]) for _,val in r)
                assert tf._same(tuple(float(y) for y in v.tick(*xs, dtypes=dt)), want), "values"
                assert v.out_dtypes==tuple(d for d,_ in r), "dtypes"
        except Exception as ex:
            bad+=1
            if bad<10: print("BAD", expr, repr(ex)[:300], flush=True)
print("graphs", tot, "beyond-reference", ext, "bad", bad, "seeds", seed-seed0, flush=True)
