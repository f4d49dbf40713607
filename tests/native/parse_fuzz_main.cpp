// Host-only harness for the sanitizer run (tests/test_robustness.py): demux + probe + full symbol parse of every file named on
// the command line through the same C entry points the daemon binding uses.  Built with -fsanitize=address,undefined from the
// product's host sources (no CUDA objects); any out-of-bounds access aborts the process, which is what the test looks for.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/av1r.h"

int main(int argc, char** argv) {
    int parsed_ok = 0, rejected = 0;
    for (int i = 1; i < argc; i++) {
        FILE* f = fopen(argv[i], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[i]); return 2; }
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        // exact-size heap buffer: one byte past the end is a sanitizer report
        std::vector<uint8_t> buf((size_t)(n > 0 ? n : 0));
        if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); return 2; }
        fclose(f);
        av1r_stream_info si;
        av1r_probe_buffer(buf.data(), buf.size(), &si);
        av1r_report rep;
        const int rc = av1r_parse_buffer(buf.data(), buf.size(), 1, 0, &rep);
        if (rc > 0) { fprintf(stderr, "%s: positive return code %d\n", argv[i], rc); return 3; }
        if (rc == 0) parsed_ok++; else rejected++;
    }
    printf("ok=%d rejected=%d\n", parsed_ok, rejected);
    return 0;
}
