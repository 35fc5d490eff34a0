"""Copy the reference's data assets (7 JPEG textures + output_RayTracer.png) into the git-ignored
tests/golden/_local/ so that an optional GPU test can check the CUDA float64 path against the
reference's golden render on the GPU box (the directory travels with gpurun, never with git).

    python oracle/stage_local_assets.py
"""
import os
import shutil

SRC = os.environ.get("B200RT_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "_local")

if __name__ == "__main__":
    os.makedirs(os.path.join(DST, "textures"), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(SRC, "textures"))):
        if f.endswith(".jpg"):
            shutil.copy(os.path.join(SRC, "textures", f), os.path.join(DST, "textures", f))
    shutil.copy(os.path.join(SRC, "output_RayTracer.png"), os.path.join(DST, "output_RayTracer.png"))
    print("staged into", DST)
