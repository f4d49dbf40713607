// C ABI entry points of the decode engine (include/av1r.h).  Thin: everything lives in Engine.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/av1r.h"
#include "engine.h"

using namespace av1r;

struct av1r_ctx {
    Engine* eng;
    std::string err;
};

extern "C" int av1r_open(const av1r_config* cfg, av1r_ctx** out) {
    if (!out) return AV1R_EINVAL;
    *out = nullptr;
    av1r_config c;
    av1r_default_config(&c);
    if (cfg) {
        size_t n = cfg->struct_size && cfg->struct_size < sizeof(c) ? cfg->struct_size : sizeof(c);
        memcpy(&c, cfg, n);
        c.struct_size = sizeof(c);
    }
    av1r_ctx* ctx = new av1r_ctx();
    ctx->eng = new Engine();
    int rc = ctx->eng->open(c);
    if (rc) {
        // keep the message reachable: callers get the code, stderr gets the text
        fprintf(stderr, "av1r_open: %s\n", ctx->eng->error().c_str());
        delete ctx->eng;
        delete ctx;
        return rc;
    }
    *out = ctx;
    return 0;
}

extern "C" void av1r_close(av1r_ctx* ctx) {
    if (!ctx) return;
    delete ctx->eng;
    delete ctx;
}

extern "C" int av1r_submit_tu(av1r_ctx* ctx, const uint8_t* data, size_t len, int64_t pts) {
    if (!ctx || !data) return AV1R_EINVAL;
    return ctx->eng->submit_tu(data, len, pts);
}

extern "C" int av1r_collect(av1r_ctx* ctx, av1r_frame_result* out, int cap, int* n) {
    if (!ctx || !n) return AV1R_EINVAL;
    return ctx->eng->collect(out, cap, n);
}

extern "C" int av1r_flush(av1r_ctx* ctx) {
    if (!ctx) return AV1R_EINVAL;
    return ctx->eng->flush();
}

extern "C" const char* av1r_last_error(const av1r_ctx* ctx) {
    if (!ctx) return "null context";
    return ctx->eng->error().c_str();
}

extern "C" int av1r_copy_frame(av1r_ctx* ctx, int64_t handle, int plane, void* dst, size_t dst_stride) {
    if (!ctx || !dst) return AV1R_EINVAL;
    return ctx->eng->copy_frame(handle, plane, dst, dst_stride);
}

extern "C" int av1r_release_frame(av1r_ctx* ctx, int64_t handle) {
    if (!ctx) return AV1R_EINVAL;
    return ctx->eng->release_frame(handle);
}

extern "C" int av1r_verify_file(const char* path, const av1r_config* cfg, av1r_report* out) {
    if (!path || !out) return AV1R_EINVAL;
    return Engine::verify_file(path, cfg, out);
}
extern "C" int av1r_verify_buffer(const uint8_t* data, size_t len, const av1r_config* cfg, av1r_report* out, uint64_t* digests,
                                  int64_t cap_frames) {
    if (!data || !out) return AV1R_EINVAL;
    return Engine::verify_buffer(data, len, cfg, out, digests, cap_frames);
}

extern "C" int av1r_clip_load(av1r_ctx* ctx, const uint8_t* const* tus, const size_t* lens, int n_tus, av1r_clip** out) {
    if (!ctx || !tus || !lens || !out) return AV1R_EINVAL;
    return ctx->eng->clip_load(tus, lens, n_tus, out);
}
extern "C" int av1r_clip_decode(av1r_ctx* ctx, av1r_clip* clip, uint64_t* checksums, int cap_frames, int* n_frames, float* device_ms) {
    if (!ctx || !clip) return AV1R_EINVAL;
    return ctx->eng->clip_decode(clip, checksums, cap_frames, n_frames, device_ms);
}
extern "C" int av1r_clip_decode_passes(av1r_ctx* ctx, av1r_clip* clip, int passes, uint64_t* checksums, int cap_frames, int* n_frames, float* device_ms) {
    if (!ctx || !clip) return AV1R_EINVAL;
    return ctx->eng->clip_decode_passes(clip, passes, checksums, cap_frames, n_frames, device_ms);
}
extern "C" int av1r_clip_profile(av1r_ctx* ctx, av1r_clip* clip, av1r_stage_times* out) {
    if (!ctx || !clip || !out) return AV1R_EINVAL;
    return ctx->eng->clip_profile(clip, out);
}

// clip info / free need the full type: provided by engine.cu

extern "C" int av1r_ctx_verify_buffer(av1r_ctx* ctx, const uint8_t* data, size_t len, av1r_report* out, uint64_t* digests, int64_t cap_frames) {
    if (!ctx || !data || !out) return AV1R_EINVAL;
    return ctx->eng->verify(data, len, out, digests, cap_frames);
}
