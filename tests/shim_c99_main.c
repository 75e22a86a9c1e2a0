/* tests/shim_c99_main.c -- TEST INFRASTRUCTURE: compiles the binding INTEGRATION.md section 2 shows (extracted
 * verbatim into shim_block.c by tools/extract_shim.py) as C99 against the reference's own headers, links it with
 * the reference's unmodified objects (oracle/_ref/obj) and libr433b.so, and replays capture files through
 * replay_files_on_gpu(): the reference's decoders and its JSON output run behind the GPU path.
 *
 *   shim_c99 RATE FILE.cu8 [FILE.cu8 ...]     -> one JSON line per decoded message on stdout
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shim_block.c"

#include "r_api.h"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s SAMPLE_RATE FILE.cu8 [FILE.cu8 ...]\n", argv[0]);
        return 2;
    }
    r_cfg_t *cfg = r_create_cfg();
    cfg->samp_rate = (uint32_t)atoi(argv[1]);
    cfg->center_frequency = 433920000;
    cfg->report_time = REPORT_TIME_OFF;
    add_json_output(cfg, NULL); /* stdout */
    register_all_protocols(cfg, 0);

    uint32_t n_files = (uint32_t)(argc - 2);
    uint64_t *offsets = calloc(n_files + 1, sizeof(uint64_t));
    uint64_t *lengths = calloc(n_files, sizeof(uint64_t));
    size_t cap = 0;
    uint8_t *all = NULL;
    for (uint32_t f = 0; f < n_files; ++f) {
        FILE *fp = fopen(argv[2 + f], "rb");
        if (!fp) {
            perror(argv[2 + f]);
            return 2;
        }
        fseek(fp, 0, SEEK_END);
        long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        size_t padded = ((size_t)n + 15) / 16 * 16; /* 16-byte aligned starts */
        all = realloc(all, cap + padded);
        memset(all + cap, 0, padded);
        if (fread(all + cap, 1, (size_t)n, fp) != (size_t)n) return 2;
        fclose(fp);
        offsets[f] = cap;
        lengths[f] = (uint64_t)n;
        cap += padded;
    }
    offsets[n_files] = cap;
    int rc = replay_files_on_gpu(cfg, all, offsets, lengths, n_files, 2);
    fflush(stdout);
    free(all);
    free(offsets);
    free(lengths);
    r_free_cfg(cfg);
    return rc ? 1 : 0;
}
