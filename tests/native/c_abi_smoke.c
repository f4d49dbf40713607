/* Plain C99 consumer of include/av1r.h: proves the header is C-clean (no C++ leaks) and that the link line a cgo package uses
 * (-lav1r -lcudart) resolves.  Built by tests/test_abi.py with `gcc -std=c99 -pedantic -Wall -Werror`.
 *   c_abi_smoke FILE          host-only calls: version, defaults, probe, parse (runs on the GPU-less build box)
 *   c_abi_smoke FILE gpu      additionally av1r_open + av1r_verify_file + av1r_verify_batch on device 0
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "av1r.h"

static unsigned char* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    long sz;
    unsigned char* buf;
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf = (unsigned char*)malloc(sz > 0 ? (size_t)sz : 1);
    if (buf && sz > 0 && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { free(buf); buf = NULL; }
    fclose(f);
    *n = sz > 0 ? (size_t)sz : 0;
    return buf;
}

int main(int argc, char** argv) {
    av1r_config cfg;
    av1r_stream_info si;
    av1r_report rep;
    unsigned char* data;
    size_t n = 0;
    int rc;
    if (argc < 2) { fprintf(stderr, "usage: %s FILE [gpu]\n", argv[0]); return 2; }
    if (av1r_abi_version() != AV1R_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    av1r_default_config(&cfg);
    if (cfg.struct_size != sizeof(cfg) || cfg.apply_grain != 1 || cfg.inloop_filters != 7) { fprintf(stderr, "bad defaults\n"); return 1; }
    data = slurp(argv[1], &n);
    if (!data) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    memset(&si, 0, sizeof(si));
    si.struct_size = sizeof(si);
    rc = av1r_probe_buffer(data, n, &si);
    if (rc || !si.is_av1) { fprintf(stderr, "probe failed: %d\n", rc); return 1; }
    rc = av1r_parse_buffer(data, n, 1, 0, &rep);
    if (rc || rep.frames <= 0) { fprintf(stderr, "parse failed: %d %s\n", rc, rep.message); return 1; }
    printf("probe %dx%d %d-bit, %lld temporal units; host parse: %lld frames\n", si.width, si.height, si.bit_depth,
           (long long)si.temporal_units, (long long)rep.frames);
    if (argc > 2 && strcmp(argv[2], "gpu") == 0) {
        av1r_ctx* ctx = NULL;
        av1r_report one, total;
        const char* paths[1];
        int dev = 0;
        long long parsed = rep.frames;
        rc = av1r_open(&cfg, &ctx);
        if (rc) { fprintf(stderr, "av1r_open: %d\n", rc); return 1; }
        rc = av1r_ctx_verify_buffer(ctx, data, n, &rep, NULL, 0);
        if (rc || rep.frames != parsed) { fprintf(stderr, "verify: %d %s (%s)\n", rc, rep.message, av1r_last_error(ctx)); return 1; }
        av1r_close(ctx);
        paths[0] = argv[1];
        rc = av1r_verify_batch(paths, 1, &dev, 1, NULL, &one, &total);
        if (rc || total.frames != parsed || one.frames != parsed) { fprintf(stderr, "batch: %d %s\n", rc, total.message); return 1; }
        printf("gpu verify: %lld frames, %.1f frames/s; batch: %s\n", (long long)rep.frames, rep.frames_per_sec, total.message);
    }
    free(data);
    return 0;
}
