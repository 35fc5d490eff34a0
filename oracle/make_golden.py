"""Generate tests/golden/*.npz by running the REAL reference in this container.

    python oracle/make_golden.py            # needs /root/reference (read-only is fine)

Every fixture is the output of the reference's own code (see oracle/ref_harness.py for how it
is run) on the Cornell scene built by the reference's ``CustomSceneBuilder`` with
``random.seed(0)``.  The only substitution is the texture *files*: the reference's JPEGs cannot
be shipped, so ``textures/<name>.jpg`` resolve to the deterministic synthetic images of
``b200rt.cornell.synthetic_texture`` (same dimensions), written as lossless PNG bytes.  The
real-texture golden vector (``output_RayTracer.png``) is checked by the container-only test
``tests/test_oracle_reference_pin.py`` instead.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.path and os.path.abspath(sys.path[0]) == os.path.dirname(os.path.abspath(__file__)):
    sys.path.pop(0)                      # keep "oracle" resolving to the package, not this directory
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "path-tracing__ray-tracer_b200"))

from oracle import ref_harness as RH  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    assert RH.available(), "reference not found"
    os.makedirs(OUT, exist_ok=True)
    tex_root = RH.write_synthetic_texture_dir(tempfile.mkdtemp(prefix="b200rt_syn_"))
    t0 = time.time()

    # ---- packed scene (the reference's own host packers) -------------------------------
    scene, cam169 = RH.build_reference_scene(0, 16 / 9, texture_root=tex_root)
    _, cam43 = RH.build_reference_scene(0, 4 / 3, texture_root=tex_root)      # same seed -> same scene
    pk = RH.reference_pack(scene, cam169, "path", texture_root=tex_root)
    pk43 = RH.reference_pack(scene, cam43, "path", texture_root=tex_root)
    pk_tex = RH.reference_pack(scene, cam169, "texture", texture_root=tex_root)
    assert all(np.array_equal(pk[k], pk_tex[k]) for k in pk)                  # both renderers pack alike
    kinds = np.array([type(o).__name__ for o in scene.objects])
    np.savez_compressed(os.path.join(OUT, "packed_scene_seed0.npz"),
                        scene=pk["scene"], camera_16x9=pk["camera"], camera_4x3=pk43["camera"],
                        lights=pk["lights"], tex_info=pk["tex_info"],
                        tex_sha256=np.array(hashlib.sha256(pk["tex"].tobytes()).hexdigest()),
                        object_kinds=kinds)
    print("packed scene", pk["scene"].shape, f"{time.time() - t0:.1f}s")

    # ---- cuda_xorshift / cuda_random / cuda_tonemap known answers ---------------------
    mod = RH.njit_path_tracer()
    seeds = np.array([0, 1, 12345, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 40 + 17, 1103515245 * 2073599 + 12345,
                      -5, -2 ** 40, 987654321987], dtype=np.int64)
    xs = np.array([mod.cuda_xorshift(int(s)) for s in seeds], dtype=np.int64)
    rnd = np.array([mod.cuda_random(int(s)) for s in seeds])
    chain = [int(seeds[6])]
    for _ in range(32):
        chain.append(int(mod.cuda_xorshift(chain[-1])))
    tm_in = np.linspace(0, 4, 33)
    tm = np.array([mod.cuda_tonemap(float(x)) for x in tm_in])
    np.savez_compressed(os.path.join(OUT, "nb_rng_tonemap.npz"), seeds=seeds, xorshift=xs, random=rnd,
                        chain=np.array(chain, dtype=np.int64), tonemap_in=tm_in, tonemap=tm)

    # ---- cuda_scene_hit on explicit rays ----------------------------------------------
    rng = np.random.default_rng(1)
    n = 3000
    o = np.concatenate([np.tile([0, 0, 50.0], (n // 2, 1)), rng.uniform(-14, 14, (n // 2, 3))])
    d = rng.normal(size=(n, 3))
    d[: n // 2, 2] = -np.abs(d[: n // 2, 2]) * 4
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    hit, rec = RH.run_scene_hit(pk, o, d)
    np.savez_compressed(os.path.join(OUT, "nb_scene_hit_rays.npz"), o=o, d=d, hit=hit, rec=rec)
    print("scene_hit rays", hit.mean(), f"{time.time() - t0:.1f}s")

    # ---- path tracer: the kernel's uint8 + float sums via the reference's device functions
    W, H, SPP, D = 64, 36, 8, 8
    for fc in (0, 1):
        u8 = RH.run_path_kernel(pk, W, H, SPP, D, fc)
        s1, s2 = RH.run_path_float(pk, W, H, SPP, D, fc)
        np.savez_compressed(os.path.join(OUT, f"nb_path_{W}x{H}_spp{SPP}_d{D}_f{fc}.npz"),
                            u8=u8.reshape(H, W, 3), sum=s1.reshape(H, W, 3), sumsq=s2.reshape(H, W, 3),
                            params=np.array([W, H, SPP, D, fc]))
    print("path tracer", f"{time.time() - t0:.1f}s")

    # ---- textured Whitted kernel ----------------------------------------------------------
    W, H, SPP, D = 96, 54, 4, 6
    u8 = RH.run_texture_kernel(pk, W, H, SPP, D)
    np.savez_compressed(os.path.join(OUT, f"nb_texture_{W}x{H}_spp{SPP}_d{D}.npz"), u8=u8.reshape(H, W, 3),
                        params=np.array([W, H, SPP, D]))
    W, H, SPP, D = 64, 48, 9, 16            # 4:3, the golden-PNG setting's depth
    u8 = RH.run_texture_kernel(pk43, W, H, SPP, D)
    np.savez_compressed(os.path.join(OUT, f"nb_texture_{W}x{H}_spp{SPP}_d{D}.npz"), u8=u8.reshape(H, W, 3),
                        params=np.array([W, H, SPP, D]))
    print("texture whitted", f"{time.time() - t0:.1f}s")

    # ---- cpu_raytracer: CPURenderer._trace on pixel-centre rays + primary hits ------------
    sys.path.insert(0, RH.REF_ROOT)
    from core.material import HitRecord  # type: ignore
    R = RH.reference_cpu_renderer()
    W, H, D = 64, 48, 4
    rgb = np.zeros((H, W, 3)); tt = np.full((H, W), -1.0); ids = np.full((H, W), -1, dtype=np.int32)
    for j in range(H):
        for i in range(W):
            ray = cam43.get_ray((i + 0.5) / W, (j + 0.5) / H)
            c = R._trace(ray, scene, 0, D)
            rgb[j, i] = (c.x, c.y, c.z)
            rec = HitRecord()
            if scene.hit(ray, 1e-3, float("inf"), rec):
                tt[j, i] = rec.t
            best, bt = -1, float("inf")            # each object alone, first arg-min
            for k, ob in enumerate(scene.objects):
                r2 = HitRecord()
                if ob.hit(ray, 1e-3, float("inf"), r2) and r2.t < bt:
                    best, bt = k, r2.t
            ids[j, i] = best
    np.savez_compressed(os.path.join(OUT, f"cpu_whitted_{W}x{H}_d{D}.npz"), rgb=rgb, t=tt, ids=ids,
                        params=np.array([W, H, D]))
    print("cpu whitted", f"{time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
