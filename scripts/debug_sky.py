import sys, os, random
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'path-tracing__ray-tracer_b200'))
import numpy as np
from b200rt import renderer
from b200rt.cornell import CustomSceneBuilder
from b200rt.scene_api import RenderSettings
from oracle import cpu_oracle as O
random.seed(0); b=CustomSceneBuilder(texture_dir=False); scene=b.build_scene(); cam=b.create_camera(16/9)
W,H,D,N=160,90,8,1024
pk=O.nb_pack(scene,cam,with_textures=False)
ids,_=O.nb_primary_hits(pk,W,H)
res={}
for prec in ('f32','f64'):
    r=renderer.B200PathTracer(precision=prec,rng='pcg',seed=7)
    acc,cnt=r.render_accum(scene,cam,RenderSettings(W,H,N,D))
    res[prec]=acc[...,:3].astype(np.float64)/N
    print(prec,'rays/path',cnt[1]/cnt[0],cnt[2]/cnt[0],'unshadowed/path',cnt[3]/cnt[0])
from scipy.ndimage import minimum_filter
sky=minimum_filter((ids<0).astype(np.uint8),size=5).astype(bool)
for prec in res:
    dev=np.abs(res[prec]-0.1).max(axis=2)*sky
    bad=np.argwhere(dev>1e-5)
    print(prec,'sky px',sky.sum(),'bad',len(bad),'max',dev.max())
    for y,x in bad[:12]: print('   ',y,x,res[prec][y,x], 'other', res['f64' if prec=='f32' else 'f32'][y,x] if len(res)==2 else None)
d=np.abs(res['f32']-res['f64']).max(axis=2)
print('f32 vs f64 same-seed: max',d.max(),'n>1e-3',(d>1e-3).sum(),'of',d.size)
