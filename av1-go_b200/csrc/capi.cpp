// C ABI glue that needs no GPU (version / defaults).
#include <cstring>

#include "../../include/av1r.h"

extern "C" uint32_t av1r_abi_version(void) { return AV1R_ABI_VERSION; }

extern "C" void av1r_default_config(av1r_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->device = 0;
    cfg->streams = 2;
    cfg->frames_in_flight = 8;
    cfg->parity_md5 = 0;
    cfg->apply_grain = 1;
    cfg->inloop_filters = 7;
    cfg->keep_frames = 0;
    cfg->host_threads = 0;
}
