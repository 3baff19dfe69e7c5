"""ctypes mirrors of include/lh_shared_cpu_gpu.h, include/lighthugger_b200.h and
include/lh_scene.h (which restate /root/reference/src/shared_cpu_gpu.h:8-166).
Sizes are asserted against the numbers the reference header compiles to
(SURVEY.md 8a-L)."""
import ctypes as C

u8, u16, u32, u64, i32, f32 = C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int32, C.c_float


class LhMat4(C.Structure):
    _fields_ = [("m", (f32 * 4) * 4)]  # m[column][row]


class LhMat3(C.Structure):
    _fields_ = [("m", (f32 * 3) * 3)]


class LhMeshInfo(C.Structure):
    _fields_ = [
        ("positions", u64), ("indices", u64), ("normals", u64), ("uvs", u64), ("micro_indices", u64), ("meshlets", u64),
        ("num_meshlets", u16), ("flags", u8), ("_pad0", u8),
        ("bounding_sphere", f32 * 4), ("texture_scale", f32 * 2), ("texture_offset", f32 * 2),
        ("base_color_texture_index", u16), ("metallic_roughness_texture_index", u16), ("normal_texture_index", u16),
        ("_pad1", u16), ("base_color_factor", f32 * 3),
    ]


class LhMiscStorage(C.Structure):
    _fields_ = [
        ("shadow_matrices", LhMat4 * 4), ("uv_space_shadow_matrices", LhMat4 * 4), ("shadow_view_matrices", LhMat4 * 4),
        ("shadow_sphere_radii", f32 * 4), ("min_depth", u32), ("max_depth", u32),
    ]


class LhUniforms(C.Structure):
    _fields_ = [
        ("combined_perspective_view", LhMat4), ("inv_perspective_view", LhMat4), ("view", LhMat4), ("initial_view", LhMat4),
        ("perspective", LhMat4), ("perspective_inverse", LhMat4), ("view_inverse", LhMat4),
        ("meshlet_references", u64), ("instances", u64), ("draw_calls", u64), ("misc_storage", u64),
        ("num_meshlets_prefix_sum", u64), ("dispatches", u64),
        ("camera_pos", f32 * 3), ("sun_dir", f32 * 3), ("sun_intensity", f32 * 3),
        ("num_instances", u32), ("window_size", u32 * 2), ("shadow_cam_distance", f32), ("cascade_split_pow", f32),
        ("debug", i32), ("debug_shadowmaps", u32),
    ]


class LhDrawIndirectCommand(C.Structure):
    _fields_ = [("vertexCount", u32), ("instanceCount", u32), ("firstVertex", u32), ("firstInstance", u32)]


class LhDispatchIndirectCommand(C.Structure):
    _fields_ = [("x", u32), ("y", u32), ("z", u32)]


class LhInstance(C.Structure):
    _fields_ = [("transform", LhMat4), ("mesh_info_address", u64), ("normal_transform", LhMat3), ("_pad0", u32)]


class LhMeshlet(C.Structure):
    _fields_ = [
        ("cone_apex", f32 * 3), ("cone_axis", f32 * 3), ("cone_cutoff", f32), ("bounding_sphere", f32 * 4),
        ("triangle_offset", u32), ("index_offset", u32), ("triangle_count", u8), ("index_count", u8), ("_pad0", u16),
    ]


class LhMeshletReference(C.Structure):
    _fields_ = [("instance_index", u32), ("meshlet_index", u16), ("_pad0", u16)]


class LhPrefixSumValue(C.Structure):
    _fields_ = [("index", u32), ("sum", u32)]


for _t, _n in ((LhMeshInfo, 104), (LhMiscStorage, 792), (LhUniforms, 560), (LhDrawIndirectCommand, 16),
               (LhDispatchIndirectCommand, 12), (LhInstance, 112), (LhMeshlet, 56), (LhMeshletReference, 8),
               (LhPrefixSumValue, 8)):
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)

# ---- include/lighthugger_b200.h ----
LH_PASS_NAMES = ["reset", "cull_instances", "cull_meshlets", "visbuffer", "composite", "depth_reduction", "shadow_matrices",
                 "cull_instances_shadows", "shadowmaps", "render_geometry", "display_transform", "shadow_exchange", "frame_barrier"]
LH_PASS_COUNT = len(LH_PASS_NAMES)
LH_CONFIG_WRITE_DRAW_COMMANDS = 1
LH_CONFIG_PASS_TIMINGS = 2
LH_CONFIG_COUNT_SAMPLES = 4
LH_TARGET_VIS64, LH_TARGET_DEPTH, LH_TARGET_SCENE_RGBA16F, LH_TARGET_SHADOW, LH_TARGET_DISPLAY_BGRA8 = range(5)
LH_TEXTURE_RGBA8_SRGB, LH_TEXTURE_RGBA8_UNORM, LH_TEXTURE_BC7_SRGB, LH_TEXTURE_BC1_SRGB = range(4)
LH_ERR_CAPACITY = -4
LH_EXCHANGE_PEER, LH_EXCHANGE_NCCL = 0, 1
# the reference's compile-time capacities (shared_cpu_gpu.h:119-125,159): LhConfig / LhCapacities defaults
LH_MAX_INSTANCES, LH_MAX_OPAQUE_DRAWS, LH_MAX_ALPHA_CLIP_DRAWS = 1024, 200000, 200000
LH_VIS_CLEAR = 0xFFFFFFFFFFFFFFFF
LH_VIS_TRI_BITS = 7
LH_VIS_MAX_SLOTS = 1 << 25
LH_UNUSED_TEXTURE_INDEX = 0xFFFF
LH_MESH_INFO_FLAGS_32_BIT_INDICES = 1
LH_MESH_INFO_FLAGS_ALPHA_CLIP = 2


class LhConfig(C.Structure):
    _fields_ = [
        ("struct_size", u32), ("device", i32), ("max_instances", u32), ("max_opaque_draws", u32),
        ("max_alpha_clip_draws", u32), ("width", u32), ("height", u32), ("shadowmap_size", u32), ("flags", u32),
        ("max_large_triangles", u32),
    ]


class LhBufferSizes(C.Structure):
    _fields_ = [("meshlet_references", u64), ("num_meshlets_prefix_sum", u64), ("draw_calls", u64), ("dispatches", u64),
                ("misc_storage", u64), ("uniforms", u64)]


class LhFrameStats(C.Structure):
    _fields_ = [
        ("num_visible_instances", u32), ("num_candidate_meshlets", u32), ("num_opaque", u32), ("num_alpha_clip", u32),
        ("min_depth", u32), ("max_depth", u32), ("shadow_visible_instances", u32 * 4),
        ("shadow_candidate_meshlets", u32 * 4), ("shadow_opaque", u32 * 4), ("shadow_alpha_clip", u32 * 4),
        ("num_large_triangles", u32), ("overflow", u32), ("rasterized_triangles", u64),
        ("meshlets_skipped_no_sample", u32 * 5), ("meshlets_skipped_occluded", u32 * 5),
        ("meshlets_queued", u32 * 5), ("_pad1", u32), ("emitted_samples", C.c_uint64 * 5),
    ]


class LhPassTimes(C.Structure):
    _fields_ = [("ms", f32 * LH_PASS_COUNT), ("frame_ms", f32), ("raster_meshlets_kernel_ms", f32),
                ("raster_meshlets_kernel_launches", u32)]


# ---- include/lh_scene.h ----
LH_MESH_TORUS, LH_MESH_SPHERE, LH_MESH_PLANE, LH_MESH_BOX = range(4)


class LhMeshDesc(C.Structure):
    _fields_ = [
        ("kind", i32), ("segments_u", u32), ("segments_v", u32), ("flags", u32), ("base_color_texture_index", u16),
        ("metallic_roughness_texture_index", u16), ("normal_texture_index", u16), ("_pad", u16),
        ("base_color_factor", f32 * 3), ("uv_repeat", f32),
    ]


class LhMeshView(C.Structure):
    _fields_ = [
        ("positions", C.c_void_p), ("num_vertices", u64), ("normals", C.c_void_p), ("uvs", C.c_void_p),
        ("indices", C.c_void_p), ("num_indices", u64), ("index_bytes", u32), ("micro_indices", C.c_void_p),
        ("micro_indices_bytes", u64), ("meshlets", C.c_void_p), ("num_meshlets", u64), ("num_triangles", u64),
        ("info", LhMeshInfo), ("dequant_scale", f32 * 3), ("dequant_offset", f32 * 3),
    ]


class LhCamera(C.Structure):
    _fields_ = [("position", f32 * 3), ("yaw", f32), ("pitch", f32), ("fov_degrees", f32), ("sun_latitude", f32),
                ("sun_longitude", f32), ("shadow_cam_distance", f32), ("cascade_split_pow", f32)]


class LhCapacities(C.Structure):
    _fields_ = [("max_instances", u32), ("max_opaque_draws", u32), ("max_alpha_clip_draws", u32)]


class LhPresetInfo(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("num_texture_sets", u32), ("alpha_cutout_set", u32), ("camera", LhCamera)]
