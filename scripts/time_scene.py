import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
from tmkit_b200 import amino as A, scenes as S, sussman as SU
sc=S.baxter_sussman()
t=time.perf_counter(); sg=A.SceneGraph.from_scene(sc); print('from_scene',time.perf_counter()-t)
q=S.q0_vector(sc)
t=time.perf_counter(); sg.tf(q); print('first tf (cuda init)',time.perf_counter()-t)
for rep in range(3):
    t=time.perf_counter(); sg.tf(q); print('tf',time.perf_counter()-t)
    t=time.perf_counter(); ssg=A.SubSceneGraph(sg,"",SU.FRAME); print('chain',time.perf_counter()-t)
    t=time.perf_counter(); mp=A.MotionPlanner(ssg); print('mp create',time.perf_counter()-t)
    t=time.perf_counter(); mp.set_start(q); print('set_start',time.perf_counter()-t)
    _,ab=sg.tf(q)
    t=time.perf_counter(); n=mp.set_wsgoal(ab[sg.frame_id('block_c')]); print('wsgoal',time.perf_counter()-t, n)
    t=time.perf_counter(); p=mp.plan(3.0); print('plan',time.perf_counter()-t)
    t=time.perf_counter(); sg2=sg.copy(); sg2.init(); sg2.cl_init(); print('copy+init',time.perf_counter()-t)
    t=time.perf_counter(); cl=A.Collision(sg); print('cl create',time.perf_counter()-t)
    t=time.perf_counter(); del cl; print('cl destroy',time.perf_counter()-t)
