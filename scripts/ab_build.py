"""Build experimental variants of libb200rt.so for A/B timing on the GPU box (they travel with the snapshot):
    python scripts/ab_build.py name1:-DFOO=1,-DBAR=0 name2:-DFOO=0 ...
-> path-tracing__ray-tracer_b200/build/variants/libb200rt_<name>.so"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))
from b200rt import build as B
out_dir = os.path.join(B.BUILD, "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defines = tuple(d for d in defs.split(",") if d)
    out = os.path.join(out_dir, f"libb200rt_{name}.so")
    B.build(force=True, defines=defines, out=out)
    print(out)
