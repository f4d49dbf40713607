// Engine skeleton (device plumbing); the frame pipeline is filled in as stages land.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <deque>
#include <vector>

#include "demux.h"
#include "engine.h"
#include "obu.h"

namespace av1r {

struct EngineImpl {
    av1r_config cfg;
    std::string err;
    HeaderParser hp;
    bool opened = false;
    std::deque<av1r_frame_result> done;
};

Engine::Engine() : impl_(new EngineImpl()) {}
Engine::~Engine() { delete impl_; }
const std::string& Engine::error() const { return impl_->err; }

int Engine::open(const av1r_config& cfg) {
    impl_->cfg = cfg;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        impl_->err = "no CUDA device available (this engine has no CPU fallback)";
        return AV1R_EIO;
    }
    if (cfg.device < 0 || cfg.device >= ndev) {
        impl_->err = "bad device ordinal";
        return AV1R_EINVAL;
    }
    e = cudaSetDevice(cfg.device);
    if (e != cudaSuccess) {
        impl_->err = cudaGetErrorString(e);
        return AV1R_EIO;
    }
    impl_->opened = true;
    return 0;
}

int Engine::submit_tu(const uint8_t* data, size_t len, int64_t pts) {
    (void)data; (void)len; (void)pts;
    impl_->err = "tile reconstruction not built yet";
    return AV1R_ENOSYS;
}

int Engine::collect(av1r_frame_result* out, int cap, int* n) {
    int k = 0;
    while (k < cap && !impl_->done.empty()) {
        out[k++] = impl_->done.front();
        impl_->done.pop_front();
    }
    *n = k;
    return 0;
}

int Engine::flush() { return cudaDeviceSynchronize() == cudaSuccess ? 0 : AV1R_EIO; }
int Engine::copy_frame(int64_t, int, void*, size_t) { return AV1R_EINVAL; }
int Engine::release_frame(int64_t) { return AV1R_EINVAL; }

int Engine::verify_file(const char* path, const av1r_config* cfg, av1r_report* out) {
    (void)path; (void)cfg;
    memset(out, 0, sizeof(*out));
    out->struct_size = sizeof(*out);
    out->status = AV1R_ENOSYS;
    snprintf(out->message, sizeof(out->message), "tile reconstruction not built yet");
    return AV1R_ENOSYS;
}

}  // namespace av1r
