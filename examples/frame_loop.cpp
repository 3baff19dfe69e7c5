// A C++20 host driving the CUDA layer the way the reference's main() drives Vulkan (src/main.cpp:41-832): create the
// context and the buffers, upload a scene, then per frame update the uniforms from the camera, render, run the display
// transform and wait.  Everything goes through the two C ABIs (include/lighthugger_b200.h, include/lh_scene.h); see
// INTEGRATION.md for the line-by-line correspondence.
//
//   lh_frame_loop [preset: tiny|c1|c2|c4|c5] [frames] [width height]
//
// Prints one line per frame (GPU milliseconds, visible instances, rasterised meshlets) and a checksum of the last
// BGRA8 image.  Exits with 3 and the library's message when no CUDA device is usable (there is no CPU fallback).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lh_scene.h"
#include "lighthugger_b200.h"

namespace {

[[noreturn]] void die(LhContext* ctx, const char* what, int rc) {   // check_vk_result of src/util.cpp:4-11, without the abort
    fprintf(stderr, "lh_frame_loop: %s failed (%d): %s\n", what, rc, lh_last_error(ctx));
    exit(rc == LH_ERR_NO_DEVICE ? 3 : 1);
}
#define CHECK(ctx, call)                        \
    do {                                        \
        int rc_ = (call);                       \
        if (rc_ < 0) die(ctx, #call, rc_);      \
    } while (0)

int preset_id(const std::string& name) {
    if (name == "tiny") return LH_PRESET_TINY;
    if (name == "c1") return LH_PRESET_C1;
    if (name == "c2") return LH_PRESET_C2;
    if (name == "c4") return LH_PRESET_C4;
    if (name == "c5") return LH_PRESET_C5;
    fprintf(stderr, "unknown preset %s\n", name.c_str());
    exit(2);
}

}  // namespace

int main(int argc, char** argv) {
    const std::string preset = argc > 1 ? argv[1] : "tiny";
    const int frames = argc > 2 ? atoi(argv[2]) : 8;

    // ---- scene (stands in for load_gltf + the instance list, src/main.cpp:373-452) ----
    LhScene* scene = nullptr;
    if (lh_scene_create(&scene) != 0) return 1;
    LhPresetInfo info{};
    if (int rc = lh_scene_build_preset(scene, preset_id(preset), 0, 0, &info)) {
        fprintf(stderr, "lh_scene_build_preset failed: %d\n", rc);
        return 1;
    }
    const uint32_t width = argc > 4 ? (uint32_t)atoi(argv[3]) : info.width, height = argc > 4 ? (uint32_t)atoi(argv[4]) : info.height;

    // ---- context with capacities for this scene (the reference's are compile-time constants, shared_cpu_gpu.h:119-125) ----
    LhConfig cfg;
    lh_config_default(&cfg);
    cfg.width = width;
    cfg.height = height;
    cfg.max_instances = lh_scene_num_instances(scene) ? lh_scene_num_instances(scene) : 1;
    const uint64_t meshlets = lh_scene_total_meshlets(scene);
    cfg.max_opaque_draws = (uint32_t)(meshlets + 1024);
    cfg.max_alpha_clip_draws = (uint32_t)(meshlets / 4 + 1024);
    cfg.flags |= LH_CONFIG_PASS_TIMINGS;
    LhContext* ctx = nullptr;
    CHECK(nullptr, lh_context_create(&cfg, &ctx));

    // ---- upload: every stream, MeshInfo, the instances and the five pipeline buffers (main.cpp:442-532,646-670) ----
    LhCapacities caps{cfg.max_instances, cfg.max_opaque_draws, cfg.max_alpha_clip_draws};
    LhUniforms addresses{};
    LhBinding* binding = nullptr;
    if (int rc = lh_scene_bind(scene, &caps, (LhAllocFn)lh_buffer_alloc, (LhUploadFn)lh_upload, (LhFreeFn)lh_buffer_free, ctx, &addresses, &binding)) {
        fprintf(stderr, "lh_scene_bind failed: %d (%s)\n", rc, lh_last_error(ctx));
        return 1;
    }

    // ---- textures: the preset's procedural sets, BC7-encoded like the reference's DDS assets; a procedural sky ----
    for (uint32_t kind = 0; kind < 3; kind++)
        for (uint32_t v = 0; v < info.num_texture_sets; v++) {
            const uint32_t size = 64;
            uint32_t levels = 0;
            const uint64_t bytes = lh_scene_make_texture((int)kind, size, v, kind == 0 && v == info.alpha_cutout_set, nullptr, &levels);
            std::vector<uint8_t> rgba(bytes), bc7;
            lh_scene_make_texture((int)kind, size, v, kind == 0 && v == info.alpha_cutout_set, rgba.data(), &levels);
            uint64_t offset = 0;
            for (uint32_t l = 0; l < levels; l++) {
                const uint32_t s = size >> l ? size >> l : 1;
                const uint64_t n = lh_bc7_encode(rgba.data() + offset, s, s, nullptr);
                const size_t at = bc7.size();
                bc7.resize(at + n);
                lh_bc7_encode(rgba.data() + offset, s, s, bc7.data() + at);
                offset += (uint64_t)s * s * 4;
            }
            uint16_t slot = 0;
            CHECK(ctx, lh_texture_create(ctx, LH_TEXTURE_BC7_SRGB, size, size, levels, bc7.data(), bc7.size(), &slot));
        }
    {
        const uint32_t sky = 16;
        std::vector<float> faces((size_t)6 * sky * sky * 4);
        lh_scene_make_skybox(sky, faces.data());
        CHECK(ctx, lh_skybox_create(ctx, sky, faces.data()));
    }

    // ---- frame loop (main.cpp:674-832): the camera turns a little every frame, as it would under the mouse ----
    LhCamera camera = info.camera;
    double total_ms = 0.0;
    for (int frame = 0; frame < frames; frame++) {
        camera.yaw += 0.002f;
        LhUniforms u{};
        if (lh_scene_uniforms(scene, &camera, width, height, &u) != 0) return 1;   // main.cpp:734-760
        u.meshlet_references = addresses.meshlet_references;
        u.instances = addresses.instances;
        u.draw_calls = addresses.draw_calls;
        u.misc_storage = addresses.misc_storage;
        u.num_meshlets_prefix_sum = addresses.num_meshlets_prefix_sum;
        u.dispatches = addresses.dispatches;
        CHECK(ctx, lh_timer_start(ctx));
        CHECK(ctx, lh_frame_begin(ctx, &u));          // uniform_buffer.flush, main.cpp:784
        CHECK(ctx, lh_render_frame(ctx));             // render(...), rendering.cpp:28-474
        CHECK(ctx, lh_display_transform(ctx));        // rendering.cpp:491-502
        float ms = 0.0f;
        CHECK(ctx, lh_timer_stop(ctx, &ms));          // the fence wait, main.cpp:766-773
        LhFrameStats stats{};
        CHECK(ctx, lh_get_frame_stats(ctx, &stats));
        total_ms += ms;
        printf("frame %d: %.3f ms, %u visible instances, %u + %u meshlets after culling, depth bits [%08x, %08x]\n", frame, ms,
               stats.num_visible_instances, stats.num_opaque, stats.num_alpha_clip, stats.min_depth, stats.max_depth);
    }
    std::vector<uint32_t> image((size_t)width * height);
    CHECK(ctx, lh_get_target(ctx, LH_TARGET_DISPLAY_BGRA8, image.data(), image.size() * 4));
    uint64_t checksum = 1469598103934665603ull;   // FNV-1a over the presented image
    for (uint32_t p : image) checksum = (checksum ^ p) * 1099511628211ull;
    printf("%d frames of %s at %ux%u: %.3f ms/frame, image checksum %016llx\n", frames, preset.c_str(), width, height, total_ms / (frames ? frames : 1),
           (unsigned long long)checksum);

    lh_binding_destroy(binding);
    lh_context_destroy(ctx);
    lh_scene_destroy(scene);
    return 0;
}
