import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
from tmkit_b200 import amino as A, scenes as S, sussman as X
r = X.run(seed=0)
sc = S.baxter_sussman(); sg = A.SceneGraph.from_scene(sc); q = S.q0_vector(sc)
sg.tf(q)
T = {}
def tm(k, f):
    t=time.perf_counter(); v=f(); T.setdefault(k,[]).append(time.perf_counter()-t); return v
for it in range(5):
    _, ab = tm('tf', lambda: sg.tf(q))
    ssg = tm('ssg', lambda: A.SubSceneGraph(sg, "", X.FRAME))
    mp = tm('mp_create', lambda: A.MotionPlanner(ssg))
    tm('set_start', lambda: mp.set_start(q))
    goal = ab[sg.frame_id('block_c')]
    n = tm('set_wsgoal', lambda: mp.set_wsgoal(goal))
    p = tm('plan', lambda: mp.plan(3.0))
    sg2 = tm('copy', lambda: sg.copy())
    rel = sg.rel_tf(sg.frame_id(X.FRAME), sg.frame_id('block_c'), ab)
    tm('reparent', lambda: sg2.reparent_name(X.FRAME, 'block_c', rel))
    tm('init', lambda: sg2.init())
    tm('cl_init', lambda: sg2.cl_init())
    tm('tf2', lambda: sg2.tf(q))
for k,v in T.items(): print(f"{k:12s} " + " ".join(f"{x*1e3:7.2f}" for x in v))
