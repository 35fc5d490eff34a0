import sys, os, random
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'path-tracing__ray-tracer_b200'))
import numpy as np
from b200rt import renderer, packer
from b200rt.cornell import CustomSceneBuilder
from oracle import cpu_oracle as O
random.seed(0); b=CustomSceneBuilder(texture_dir=False); scene=b.build_scene(); cam=b.create_camera(4/3)
W,H=320,240
exp=O.cpu_export(scene,cam)
ref_ids,ref_t=O.cpu_primary_ids_bruteforce(exp,W,H)
obj,t,pid=renderer.primary_hits(scene,cam,W,H,'cpu','f64')
d=np.argwhere(obj!=ref_ids)
print('id mismatches',len(d),'t equal',np.array_equal(t,ref_t))
for y,x in d[:10]: print(y,x,'gpu',obj[y,x],type(scene.objects[obj[y,x]]).__name__,'ref',ref_ids[y,x],type(scene.objects[ref_ids[y,x]]).__name__,t[y,x],ref_t[y,x])
D=4
ref=O.cpu_whitted(exp,W,H,D)
r=renderer.B200WhittedRenderer(precision='f64',jitter_seed=None)
rgb=r.trace(scene,cam,W,H,D)
err=np.abs(rgb-ref['rgb']).max(axis=2)
print('whitted max err',err.max(),'n>1e-4',(err>1e-4).sum(),'of',err.size)
bad=np.argwhere(err>1e-4)
import collections
print('bad by primary obj', collections.Counter([(int(ref_ids[y,x]), type(scene.objects[ref_ids[y,x]]).__name__ if ref_ids[y,x]>=0 else None) for y,x in bad]).most_common(12))
print('all by primary obj', collections.Counter([int(v) for v in ref_ids.reshape(-1)]).most_common(12))
for y,x in bad[:6]: print(y,x,rgb[y,x],ref['rgb'][y,x])
# depth 0 (no children)
ref0=O.cpu_whitted(exp,W,H,0); rgb0=r.trace(scene,cam,W,H,0); e0=np.abs(rgb0-ref0['rgb']).max(axis=2); print('depth0 err',e0.max(),(e0>1e-4).sum())
