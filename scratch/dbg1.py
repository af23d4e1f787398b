import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import orc
from sylt2d_b200 import scenes, World
sc = scenes.pyramid(20, 10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
o = orc.OracleWorld(sc.gravity, sc.iterations, sincos_mode=1); o.load_scene(sc)
F=("x","y","rotation","vx","vy","angular_velocity")
for s in range(300):
    w.step(sc.dt)
    keys, jo = w.solve_order()
    try:
        o.step_ordered(sc.dt, keys, jo)
    except Exception as e:
        ga,_=w.get_arbiters(); o2a,_=o.get_arbiters()
        print("step",s,"order rejected:",e, "gpu A",len(ga),"oracle A",len(o2a), "stats", w.get_stats())
        gk=set(zip(ga["id1"].tolist(),ga["id2"].tolist())); ok=set(zip(o2a["id1"].tolist(),o2a["id2"].tolist()))
        print("gpu-only", sorted(gk-ok)[:10], "oracle-only", sorted(ok-gk)[:10])
        break
    g,r=w.get_bodies(),o.get_bodies()
    bad=[f for f in F if not np.array_equal(g[f],r[f])]
    ga,gc=w.get_arbiters()
    props=w.get_body_props()
    okc=True
    for c in np.unique(ga["color"]):
        sel=ga[ga["color"]==c]; b=np.concatenate([sel["body1"],sel["body2"]]); b=b[props[b,1]!=0]
        if np.unique(b).size!=b.size: okc=False
    if bad or not okc or s%50==0:
        print("step",s,"bad",bad,"colour ok",okc, w.get_stats(), "KE", float((g["vx"]**2+g["vy"]**2).sum()))
    if bad or not okc: break
