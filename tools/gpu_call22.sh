set -x
python -m pytest tests/test_gpu_solver_parity.py -q -m gpu -k "diph_3d" 2>&1 | tail -5
timeout 300 python bench.py --no-cpu --no-profile 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e'])"
python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, penguin_b200 as pb
from oracle import geom, penguin_oracle as po
from helpers import rel_l2
pb.init()
# device-built capacities, 3-D diphasic at 24^3, 6 steps: GPU vs oracle extremes
nx=24
mo, mg = po.Mesh((nx,)*3,(4.0,)*3), pb.Mesh((nx,)*3,(4.0,)*3)
body=pb.Sphere((2.0,2.0,2.0),1.0)
c1,c2=pb.Capacity(body,mg),pb.Capacity(-body,mg)
p1,p2=pb.Phase(c1,pb.DiffusionOps(c1),0.0,1.0),pb.Phase(c2,pb.DiffusionOps(c2),0.0,1.0)
n=c1.nloc; h=4.0/nx; dt=0.5*h*h
u0=np.concatenate([np.ones(2*n),np.zeros(2*n)])
ic=pb.InterfaceConditions(pb.ScalarJump(1.0,2.0,0.0),pb.FluxJump(1.0,1.0,0.0))
s=pb.DiffusionUnsteadyDiph(p1,p2,pb.BorderConditions(),ic,dt,u0,"BE")
pb.solve_DiffusionUnsteadyDiph_(s,p1,p2,dt,5.5*dt,pb.BorderConditions(),ic,"BE",reltol=1e-12)
ls=geom.LevelSet.ball((2.0,2.0,2.0),1.0)
o1,o2=geom.capacity(mo,ls),geom.capacity(mo,ls.flipped())
f=lambda x,y,z,t:0.0*x
q1,q2=po.Phase(o1,po.DiffusionOps(o1),f,1.0),po.Phase(o2,po.DiffusionOps(o2),f,1.0)
ico=po.InterfaceConditions(po.ScalarJump(1.0,2.0,0.0),po.FluxJump(1.0,1.0,0.0))
so=po.DiffusionUnsteadyDiph(q1,q2,po.BorderConditions(),ico,dt,u0,"BE")
po.solve_DiffusionUnsteadyDiph(so,q1,q2,dt,5.5*dt,po.BorderConditions(),ico,"BE")
for k,(a,b) in enumerate(zip(s.states,so.states)):
    print(k,"rel",rel_l2(a,b),"gpu max",np.abs(a).max(),"oracle max",np.abs(b).max())
PY
