import sys; sys.path.insert(0,'/root/repo')
import numpy as np, os
from sylt2d_b200 import scenes, World
sc = scenes.pyramid(20, 10)
w = World(sc.gravity, sc.iterations); w.load_scene(sc)
orders=[]; colors=[]
for s in range(300):
    w.step(sc.dt)
    arb,_=w.get_arbiters()
    o=np.lexsort((arb["id2"],arb["id1"],arb["color"]))
    orders.append(np.stack([arb["id1"][o],arb["id2"][o],arb["color"][o].astype(np.uint64)],1))
os.makedirs("gpurun_out",exist_ok=True)
np.savez_compressed("gpurun_out/orders.npz", **{f"s{i}":o for i,o in enumerate(orders)})
g=w.get_bodies(); print("KE", float((g["vx"]**2+g["vy"]**2).sum()))
