"""Generates the committed golden fixtures from the reference's own code.

Runs only where /root/reference exists (this container): it loads
oracle/_ref/libref_sceneprep.so, which `make ref` compiles IN PLACE from the
reference's src/resources/meshlets.cpp, src/resources/bounding_sphere.h,
src/shared_cpu_gpu.h and the vendored meshoptimizer v0.20.  Outputs:

  reference_layout.json      sizeof() of the nine contract structs
  torus_32x16_ref.npz        a 1 024-triangle torus meshletised by the reference
  torus_125x80_ref.npz       BASELINE config 1's 20 000-triangle torus, ditto
                             (214 meshlets expected)

Each .npz holds the quantised vertex streams, the triangle list fed to the
reference, and the reference's outputs: Meshlet[] (56 B records), micro
indices, vertex indices, per-meshlet bounds and the mesh bounding sphere.
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def torus(su, sv, R=1.0, r=0.35):
    """Quantised torus streams + triangle list (row-major grid order)."""
    i, j = np.meshgrid(np.arange(su), np.arange(sv), indexing="xy")
    u = (i / su * 2 * np.pi).astype(np.float32).reshape(-1)
    v = (j / sv * 2 * np.pi).astype(np.float32).reshape(-1)
    pos = np.stack([(R + r * np.cos(v)) * np.cos(u), r * np.sin(v), (R + r * np.cos(v)) * np.sin(u)], 1).astype(np.float32)
    nrm = np.stack([np.cos(v) * np.cos(u), np.sin(v), np.cos(v) * np.sin(u)], 1).astype(np.float32)
    uv = np.stack([i.reshape(-1) / su, j.reshape(-1) / sv], 1).astype(np.float32)
    lo = pos.min(0)
    ext = float((pos.max(0) - lo).max())
    q = np.clip(np.rint((pos - lo) / ext * 65535.0), 0, 65535).astype(np.uint16)
    nq = np.clip(np.rint(nrm * 127.0), -127, 127).astype(np.int8)
    uq = np.clip(np.rint(uv * 65535.0), 0, 65535).astype(np.uint16)
    tris = []
    for jj in range(sv):
        for ii in range(su):
            a = jj * su + ii
            b = jj * su + (ii + 1) % su
            c = ((jj + 1) % sv) * su + (ii + 1) % su
            d = ((jj + 1) % sv) * su + ii
            tris += [a, d, c, a, c, b]  # outward-facing for the reference's CCW-front convention
    return q, nq, uq, np.array(tris, np.uint32), np.float32(ext / 65535.0), lo.astype(np.float32)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_sceneprep.so"))
    sizes = (C.c_uint32 * 9)()
    lib.ref_struct_sizes(sizes)
    names = ["Instance", "MeshInfo", "Meshlet", "MeshletReference", "PrefixSumValue", "DrawIndirectCommand",
             "DispatchIndirectCommand", "MiscStorage", "Uniforms"]
    with open(os.path.join(HERE, "reference_layout.json"), "w") as f:
        json.dump({"source": "/root/reference/src/shared_cpu_gpu.h compiled with vendored glm",
                   "sizes": dict(zip(names, [int(x) for x in sizes]))}, f, indent=1)

    vp = C.c_void_p
    lib.ref_build_meshlets.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), vp, vp, vp]
    lib.ref_compute_bounding_sphere.argtypes = [vp, vp, C.c_uint64]
    lib.ref_compute_meshlet_bounds.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp]
    for su, sv in ((32, 16), (125, 80)):
        q, nq, uq, tris, dq_scale, dq_offset = torus(su, sv)
        raw = np.ascontiguousarray(q.astype(np.float32))  # raw integer values as floats (mesh_loading.cpp:584-597)
        n_m, n_micro, n_idx = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = lib.ref_build_meshlets(tris.ctypes.data, len(tris), raw.ctypes.data, len(q), C.byref(n_m), C.byref(n_micro),
                                    C.byref(n_idx), None, None, None)
        assert rc == 0
        meshlets = np.zeros((n_m.value, 56), np.uint8)
        micro = np.zeros(n_micro.value, np.uint8)
        idx = np.zeros(n_idx.value, np.uint16)
        lib.ref_build_meshlets(tris.ctypes.data, len(tris), raw.ctypes.data, len(q), C.byref(n_m), C.byref(n_micro),
                               C.byref(n_idx), meshlets.ctypes.data, micro.ctypes.data, idx.ctypes.data)
        sphere = np.zeros(4, np.float32)
        lib.ref_compute_bounding_sphere(sphere.ctypes.data, raw.ctypes.data, len(q))
        # bounds recomputed one meshlet at a time through meshopt_computeMeshletBounds (what meshlets.cpp:69-86 calls)
        bounds = np.zeros((n_m.value, 11), np.float32)
        tri_off = meshlets[:, 44:48].copy().view(np.uint32).reshape(-1)
        idx_off = meshlets[:, 48:52].copy().view(np.uint32).reshape(-1)
        for m in range(n_m.value):
            tc, ic = int(meshlets[m, 52]), int(meshlets[m, 53])
            mv = np.ascontiguousarray(idx[idx_off[m]:idx_off[m] + ic].astype(np.uint32))
            mt = np.ascontiguousarray(micro[tri_off[m]:tri_off[m] + 3 * tc])
            lib.ref_compute_meshlet_bounds(mv.ctypes.data, mt.ctypes.data, tc, raw.ctypes.data, len(q), bounds[m].ctypes.data)
        print(f"torus {su}x{sv}: {len(tris) // 3} triangles, {len(q)} vertices -> {n_m.value} meshlets, "
              f"{n_micro.value} micro-index bytes, {n_idx.value} vertex indices")
        np.savez_compressed(os.path.join(HERE, f"torus_{su}x{sv}_ref.npz"), positions=q, normals=nq, uvs=uq, triangles=tris,
                            meshlets=meshlets, micro_indices=micro, indices=idx, bounding_sphere=sphere, bounds=bounds,
                            dequant_scale=dq_scale, dequant_offset=dq_offset)


if __name__ == "__main__":
    main()
