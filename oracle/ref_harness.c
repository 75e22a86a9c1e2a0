/* ref_harness.c -- TEST INFRASTRUCTURE (oracle side), never part of the product path.
 *
 * Drives the UNMODIFIED merbanan/rtl_433 C sources (compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libr433ref.so) through the
 * reference's own flow driver and records everything the hot path produces:
 *   - per block: the AM (am_buf) and FM (buf.fm) stage arrays       src/r_flow.c:151-208
 *   - every package pulse_detect_package() returns                  src/r_flow.c:241-243
 *   - every (device, bitbuffer) event a slicer hands to decode_fn   src/pulse_slicer.c:26-31
 * The block loop restates only the file-replay loop of main()       src/rtl_433.c:1797-1854
 * and the few assignments of process_sdr_frame()                    src/rtl_433.c:1084-1123
 * (both live in the CLI translation unit and cannot be linked).
 *
 * Packages are captured with two "sentinel" r_devices placed first in the device list:
 * a PWM slicer configuration that emits exactly one event per package, at its last
 * pulse (reset limit far beyond any gap).  Real decoders can be chained behind the
 * capture stub so the oracle also yields the reference's decoded JSON.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtl_433.h"
#include "r_private.h"
#include "r_device.h"
#include "r_api.h"
#include "r_flow.h"
#include "baseband.h"
#include "bitbuffer.h"
#include "pulse_data.h"
#include "pulse_detect.h"
#include "pulse_slicer.h"
#include "data.h"
#include "list.h"
#include "logger.h"
#include "fileformat.h"

#define REFH_EXPORT __attribute__((visibility("default")))

typedef struct {
    int32_t type;  /* 1 = OOK, 2 = FSK (enum package_types) */
    int32_t block; /* index of the push that returned it; == n_blocks for the flush */
    uint64_t offset;
    uint32_t sample_rate, depth_bits, start_ago, end_ago, num_pulses;
    int32_t ook_low_estimate, ook_high_estimate, fsk_f1_est, fsk_f2_est;
    float freq1_hz, freq2_hz, centerfreq_hz, range_db, rssi_db, snr_db, noise_db;
    float sample_file_pos;
    uint32_t pulse_off;   /* first entry in the pulse/gap pools */
    uint32_t pulse_count; /* entries stored (num_pulses + 1, capped at 1200) */
    uint32_t first_event, num_events;
} refh_package;

typedef struct {
    uint32_t package;
    uint32_t dev;    /* registration index (sentinels excluded) */
    int32_t ret;     /* real decoder's return value when chained, else 0 */
    uint32_t bb_idx; /* index into stored bitbuffers, UINT32_MAX when only hashed */
    uint64_t hash;   /* FNV-1a over the whole bitbuffer_t */
} refh_event;

typedef struct {
    uint32_t protocol_num;
    uint32_t modulation;
    float short_width, long_width, reset_limit, gap_limit, sync_width, tolerance;
    uint32_t priority, disabled;
    char name[96];
} refh_devinfo;

typedef struct refh {
    r_cfg_t cfg;
    int n_sentinels;
    /* registration-order side tables */
    int n_devs;
    r_device **devs;
    int (**orig_fn)(r_device *, bitbuffer_t *);
    unsigned *proto_num; /* DEVICES index + 1 (create_fn devices do not carry it) */
    /* capture options */
    int chain_decoders, store_bitbuffers, store_stages;
    /* results of the current stream */
    refh_package *pkgs; size_t n_pkgs, cap_pkgs;
    refh_event *evts; size_t n_evts, cap_evts;
    bitbuffer_t *bbs; size_t n_bbs, cap_bbs;
    int32_t *pulse_pool, *gap_pool; size_t n_pool, cap_pool, cap_gpool;
    int16_t *am, *fm; size_t n_stage, cap_stage, cap_fstage;
    char *json; size_t n_json, cap_json;
    int cur_block;
    uint64_t decoded_msgs;
} refh_t;

static refh_t *g_active; /* the reference is single-threaded and so is this harness */

static void *grow(void *p, size_t *cap, size_t need, size_t elem)
{
    if (need <= *cap) return p;
    size_t c = *cap ? *cap : 64;
    while (c < need) c *= 2;
    p = realloc(p, c * elem);
    if (!p) { fprintf(stderr, "refh: out of memory\n"); abort(); }
    *cap = c;
    return p;
}

static uint64_t fnv1a(void const *p, size_t n)
{
    uint8_t const *b = p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

static void quiet_log(log_level_t level, char const *src, char const *msg, void *ud)
{
    (void)level; (void)src; (void)msg; (void)ud;
}

static void sink_log(r_device *d, int level, data_t *data)
{
    (void)d; (void)level;
    data_free(data);
}

static void sink_output(r_device *d, data_t *data)
{
    /* register_protocol() stores the r_cfg_t in output_ctx and the cfg is the first member of refh_t: with several
       harness instances dispatching at once (threaded dispatch tests) every decoder reports to its own instance */
    refh_t *h = d->output_ctx ? (refh_t *)d->output_ctx : g_active;
    if (h) {
        h->decoded_msgs++;
        char buf[4096];
        size_t n = data_print_jsons(data, buf, sizeof(buf));
        (void)n;
        size_t len = strlen(buf);
        h->json = grow(h->json, &h->cap_json, h->n_json + len + 2, 1);
        memcpy(h->json + h->n_json, buf, len);
        h->n_json += len;
        h->json[h->n_json++] = '\n';
        h->json[h->n_json] = 0;
    }
    data_free(data);
}

static int dev_index(refh_t *h, r_device *d)
{
    for (int i = 0; i < h->n_devs; ++i)
        if (h->devs[i] == d) return i;
    return -1;
}

static int sentinel_cb(r_device *d, bitbuffer_t *bits)
{
    (void)bits;
    refh_t *h = g_active;
    struct dm_state *dm = h->cfg.demod;
    int fsk = d->modulation >= FSK_DEMOD_MIN_VAL;
    pulse_data_t const *pd = fsk ? &dm->fsk_pulse_data : &dm->pulse_data;

    h->pkgs = grow(h->pkgs, &h->cap_pkgs, h->n_pkgs + 1, sizeof(*h->pkgs));
    refh_package *p = &h->pkgs[h->n_pkgs++];
    memset(p, 0, sizeof(*p));
    p->type = fsk ? PULSE_DATA_FSK : PULSE_DATA_OOK;
    p->block = h->cur_block;
    p->offset = pd->offset;
    p->sample_rate = pd->sample_rate;
    p->depth_bits = pd->depth_bits;
    p->start_ago = pd->start_ago;
    p->end_ago = pd->end_ago;
    p->num_pulses = pd->num_pulses;
    p->ook_low_estimate = pd->ook_low_estimate;
    p->ook_high_estimate = pd->ook_high_estimate;
    p->fsk_f1_est = pd->fsk_f1_est;
    p->fsk_f2_est = pd->fsk_f2_est;
    p->freq1_hz = pd->freq1_hz;
    p->freq2_hz = pd->freq2_hz;
    p->centerfreq_hz = pd->centerfreq_hz;
    p->range_db = pd->range_db;
    p->rssi_db = pd->rssi_db;
    p->snr_db = pd->snr_db;
    p->noise_db = pd->noise_db;
    p->sample_file_pos = dm->sample_file_pos;
    uint32_t cnt = pd->num_pulses + 1;
    if (cnt > PD_MAX_PULSES) cnt = PD_MAX_PULSES;
    h->pulse_pool = grow(h->pulse_pool, &h->cap_pool, h->n_pool + cnt, sizeof(int32_t));
    h->gap_pool = grow(h->gap_pool, &h->cap_gpool, h->n_pool + cnt, sizeof(int32_t));
    memcpy(h->pulse_pool + h->n_pool, pd->pulse, cnt * sizeof(int32_t));
    memcpy(h->gap_pool + h->n_pool, pd->gap, cnt * sizeof(int32_t));
    p->pulse_off = (uint32_t)h->n_pool;
    p->pulse_count = cnt;
    h->n_pool += cnt;
    p->first_event = (uint32_t)h->n_evts;
    return 0;
}

static int capture_cb(r_device *d, bitbuffer_t *bits)
{
    refh_t *h = g_active;
    int idx = dev_index(h, d);
    h->evts = grow(h->evts, &h->cap_evts, h->n_evts + 1, sizeof(*h->evts));
    refh_event *e = &h->evts[h->n_evts++];
    e->package = h->n_pkgs ? (uint32_t)(h->n_pkgs - 1) : UINT32_MAX;
    e->dev = (uint32_t)idx;
    e->ret = 0;
    e->hash = fnv1a(bits, sizeof(*bits));
    e->bb_idx = UINT32_MAX;
    if (h->store_bitbuffers) {
        h->bbs = grow(h->bbs, &h->cap_bbs, h->n_bbs + 1, sizeof(*h->bbs));
        e->bb_idx = (uint32_t)h->n_bbs;
        h->bbs[h->n_bbs++] = *bits;
    }
    if (h->n_pkgs) h->pkgs[h->n_pkgs - 1].num_events++;
    int ret = 0;
    if (h->chain_decoders && idx >= 0 && h->orig_fn[idx]) {
        ret = h->orig_fn[idx](d, bits);
        /* h->evts may have moved if a decoder re-entered; it does not, but stay safe */
        h->evts[h->n_evts - 1].ret = ret;
    }
    return ret;
}

REFH_EXPORT refh_t *refh_create(void)
{
    refh_t *h = calloc(1, sizeof(*h));
    if (!h) return NULL;
    r_init_cfg(&h->cfg);
    r_logger_set_log_handler(quiet_log, NULL);
    h->cfg.report_time = REPORT_TIME_OFF;
    h->chain_decoders = 0;
    h->store_bitbuffers = 1;
    h->store_stages = 0;
    return h;
}

REFH_EXPORT void refh_destroy(refh_t *h)
{
    if (!h) return;
    if (g_active == h) g_active = NULL;
    free(h->pkgs); free(h->evts); free(h->bbs); free(h->pulse_pool); free(h->gap_pool);
    free(h->am); free(h->fm); free(h->json); free(h->devs); free(h->orig_fn); free(h->proto_num);
    /* the cfg is left to the process (r_free_cfg tears down outputs we never created) */
    free(h);
}

REFH_EXPORT void refh_set_capture(refh_t *h, int chain_decoders, int store_bitbuffers, int store_stages)
{
    h->chain_decoders = chain_decoders;
    h->store_bitbuffers = store_bitbuffers;
    h->store_stages = store_stages;
}

REFH_EXPORT void refh_set_levels(refh_t *h, int use_mag_est, float level_limit, float min_level, float min_snr)
{
    struct dm_state *dm = h->cfg.demod;
    dm->use_mag_est = use_mag_est;
    dm->level_limit = level_limit;
    dm->min_level = min_level;
    dm->min_snr = min_snr;
}

REFH_EXPORT void refh_set_fm_low_pass(refh_t *h, float fm_low_pass)
{
    h->cfg.demod->fm_low_pass = fm_low_pass;
}

REFH_EXPORT int refh_num_protocols(refh_t *h) { return h->cfg.num_r_devices; }

static void fill_info(refh_devinfo *o, r_device const *d)
{
    memset(o, 0, sizeof(*o));
    o->protocol_num = d->protocol_num;
    o->modulation = d->modulation;
    o->short_width = d->short_width;
    o->long_width = d->long_width;
    o->reset_limit = d->reset_limit;
    o->gap_limit = d->gap_limit;
    o->sync_width = d->sync_width;
    o->tolerance = d->tolerance;
    o->priority = d->priority;
    o->disabled = d->disabled;
    if (d->name) snprintf(o->name, sizeof(o->name), "%s", d->name);
}

REFH_EXPORT int refh_get_protocol(refh_t *h, int idx, refh_devinfo *out)
{
    if (idx < 0 || idx >= h->cfg.num_r_devices) return -1;
    fill_info(out, &h->cfg.devices[idx]);
    return 0;
}

static void add_sentinel(refh_t *h, unsigned modulation, char const *name)
{
    r_device *p = calloc(1, sizeof(*p));
    p->name = name;
    p->modulation = modulation;
    p->short_width = 100.0f;
    p->long_width = 200.0f;
    p->reset_limit = 1.0e9f; /* s_reset = 2.5e8 samples at 250 kHz: never reached */
    p->decode_fn = sentinel_cb;
    p->log_fn = sink_log;
    p->output_fn = sink_output;
    list_push(&h->cfg.demod->r_devs, p);
    h->n_sentinels++;
}

static void ensure_sentinels(refh_t *h)
{
    if (h->n_sentinels) return;
    add_sentinel(h, OOK_PULSE_PWM, "refh-sentinel-ook");
    add_sentinel(h, FSK_PULSE_PWM, "refh-sentinel-fsk");
}

static int track(refh_t *h, r_device *p)
{
    h->devs = realloc(h->devs, (h->n_devs + 1) * sizeof(*h->devs));
    h->orig_fn = realloc(h->orig_fn, (h->n_devs + 1) * sizeof(*h->orig_fn));
    h->proto_num = realloc(h->proto_num, (h->n_devs + 1) * sizeof(*h->proto_num));
    h->devs[h->n_devs] = p;
    h->proto_num[h->n_devs] = p->protocol_num;
    h->orig_fn[h->n_devs] = p->decode_fn;
    p->decode_fn = capture_cb;
    p->log_fn = sink_log;
    p->output_fn = sink_output;
    return h->n_devs++;
}

/* register_protocol() (src/r_api.c:235) on one entry of the DEVICES table; idx is 0-based */
REFH_EXPORT int refh_register(refh_t *h, int idx)
{
    if (idx < 0 || idx >= h->cfg.num_r_devices) return -1;
    ensure_sentinels(h);
    size_t before = h->cfg.demod->r_devs.len;
    register_protocol(&h->cfg, &h->cfg.devices[idx], NULL);
    if (h->cfg.demod->r_devs.len != before + 1) return -1;
    int t = track(h, h->cfg.demod->r_devs.elems[before]);
    h->proto_num[t] = (unsigned)idx + 1;
    return t;
}

/* register_all_protocols(cfg, 0) (src/r_api.c:294): everything with disabled == 0 */
REFH_EXPORT int refh_register_defaults(refh_t *h)
{
    int n = 0;
    for (int i = 0; i < h->cfg.num_r_devices; ++i) {
        if (h->cfg.devices[i].disabled <= 0) {
            if (refh_register(h, i) < 0) return -1;
            n++;
        }
    }
    return n;
}

/* a decoder-less device with explicit slicer parameters (slicer parity tests) */
REFH_EXPORT int refh_register_custom(refh_t *h, unsigned modulation, float short_width, float long_width,
        float reset_limit, float gap_limit, float sync_width, float tolerance, unsigned priority)
{
    ensure_sentinels(h);
    r_device *p = calloc(1, sizeof(*p));
    p->name = "refh-custom";
    p->protocol_num = 9000 + h->n_devs;
    p->modulation = modulation;
    p->short_width = short_width;
    p->long_width = long_width;
    p->reset_limit = reset_limit;
    p->gap_limit = gap_limit;
    p->sync_width = sync_width;
    p->tolerance = tolerance;
    p->priority = priority;
    list_push(&h->cfg.demod->r_devs, p);
    return track(h, p);
}

static int null_cb(r_device *d, bitbuffer_t *bits)
{
    (void)d; (void)bits;
    return 0;
}

/* timing mode: every device (and the sentinels) gets a do-nothing decode_fn, so a timed
   run measures demod + pulse detect + slicers and nothing of this harness's capture */
REFH_EXPORT void refh_set_timing_mode(refh_t *h, int on)
{
    ensure_sentinels(h);
    for (int i = 0; i < h->n_devs; ++i)
        h->devs[i]->decode_fn = on ? null_cb : capture_cb;
    void **el = h->cfg.demod->r_devs.elems;
    for (int i = 0; i < h->n_sentinels; ++i)
        ((r_device *)el[i])->decode_fn = on ? null_cb : sentinel_cb;
}

REFH_EXPORT int refh_num_registered(refh_t *h) { return h->n_devs; }

REFH_EXPORT int refh_get_registered(refh_t *h, int idx, refh_devinfo *out)
{
    if (idx < 0 || idx >= h->n_devs) return -1;
    fill_info(out, h->devs[idx]);
    out->protocol_num = h->proto_num[idx];
    return 0;
}

static void clear_results(refh_t *h)
{
    h->n_pkgs = h->n_evts = h->n_bbs = h->n_pool = h->n_stage = h->n_json = 0;
    h->decoded_msgs = 0;
    if (h->json) h->json[0] = 0;
}

static void frame(refh_t *h, unsigned char *buf, uint32_t len, int fpdm_mode)
{
    r_cfg_t *cfg = &h->cfg;
    struct dm_state *dm = cfg->demod;
    /* process_sdr_frame(), src/rtl_433.c:1094-1123 */
    unsigned fpdm = fpdm_mode;
    if (fpdm_mode == FSK_PULSE_DETECT_AUTO)
        fpdm = cfg->center_frequency > FSK_PULSE_DETECTOR_LIMIT ? FSK_PULSE_DETECT_NEW : FSK_PULSE_DETECT_OLD;
    dm->raw_handler = &cfg->raw_handler;
    dm->fsk_pulse_detect_mode = fpdm;
    dm->report_noise = 0;
    dm->verbosity = cfg->verbosity;
    dm->raw_mode = 0;
    dm->grab_mode = 0;
    if (dm->center_frequency != cfg->center_frequency || dm->samp_rate != cfg->samp_rate)
        flush_sdr_flow(cfg);
    dm->center_frequency = cfg->center_frequency;
    dm->samp_rate = cfg->samp_rate;
    push_sdr_flow(cfg, buf, len);
}

/* One input "file": block loop, flush, reset (src/rtl_433.c:1797-1854).
   sample_size: 2 = cu8, 4 = cs16.  fpdm_mode: 0 classic, 1 minmax, 2 auto.
   block_bytes: 0 = DEFAULT_BUF_LENGTH (262144). */
REFH_EXPORT int refh_run_stream(refh_t *h, void const *iq, size_t bytes, int sample_size,
        uint32_t samp_rate, uint32_t center_freq, int fpdm_mode, uint32_t block_bytes)
{
    r_cfg_t *cfg = &h->cfg;
    struct dm_state *dm = cfg->demod;
    ensure_sentinels(h);
    clear_results(h);
    g_active = h;
    if (!block_bytes) block_bytes = DEFAULT_BUF_LENGTH;

    pulse_detect_set_levels(dm->pulse_detect, dm->use_mag_est, dm->level_limit, dm->min_level, dm->min_snr, dm->detect_verbosity);
    dm->enable_FM_demod = 0;
    for (int i = 0; i < h->n_devs; ++i)
        if (h->devs[i]->modulation >= FSK_DEMOD_MIN_VAL) dm->enable_FM_demod = 1;

    cfg->samp_rate = samp_rate;
    cfg->center_frequency = center_freq;
    dm->sample_size = sample_size;
    dm->load_info.format = sample_size == 2 ? CU8_IQ : CS16_IQ;
    dm->sample_file_pos = 0.0f;
    /* the CLI never rewinds input_pos between files (src/r_flow.c:491); every stream here
       is treated as the first file of a fresh process, so offsets count from its start */
    dm->input_pos = 0;

    unsigned char *blk = malloc(block_bytes);
    int n_blocks = 0;
    size_t pos = 0;
    while (pos < bytes) {
        size_t n_read = bytes - pos < block_bytes ? bytes - pos : block_bytes;
        memcpy(blk, (unsigned char const *)iq + pos, n_read);
        pos += n_read;
        dm->sample_file_pos = ((float)n_blocks * block_bytes + n_read) / cfg->samp_rate / dm->sample_size;
        h->cur_block = n_blocks;
        n_blocks++;
        frame(h, blk, (uint32_t)n_read, fpdm_mode);
        if (h->store_stages) {
            size_t n = n_read / sample_size;
            h->am = grow(h->am, &h->cap_stage, h->n_stage + n, sizeof(int16_t));
            h->fm = grow(h->fm, &h->cap_fstage, h->n_stage + n, sizeof(int16_t));
            memcpy(h->am + h->n_stage, dm->am_buf, n * sizeof(int16_t));
            memcpy(h->fm + h->n_stage, dm->buf.fm, n * sizeof(int16_t));
            h->n_stage += n;
        }
    }
    free(blk);
    h->cur_block = n_blocks;
    flush_sdr_flow(cfg);
    reset_sdr_flow(cfg);
    g_active = NULL;
    return (int)h->n_pkgs;
}

REFH_EXPORT uint64_t refh_input_pos(refh_t *h) { return h->cfg.demod->input_pos; }
REFH_EXPORT size_t refh_num_packages(refh_t *h) { return h->n_pkgs; }
REFH_EXPORT size_t refh_num_events(refh_t *h) { return h->n_evts; }
REFH_EXPORT size_t refh_num_bitbuffers(refh_t *h) { return h->n_bbs; }
REFH_EXPORT size_t refh_num_stage(refh_t *h) { return h->n_stage; }
REFH_EXPORT uint64_t refh_num_decoded(refh_t *h) { return h->decoded_msgs; }
REFH_EXPORT refh_package const *refh_packages(refh_t *h) { return h->pkgs; }
REFH_EXPORT refh_event const *refh_events(refh_t *h) { return h->evts; }
REFH_EXPORT bitbuffer_t const *refh_bitbuffers(refh_t *h) { return h->bbs; }
REFH_EXPORT int32_t const *refh_pulse_pool(refh_t *h) { return h->pulse_pool; }
REFH_EXPORT int32_t const *refh_gap_pool(refh_t *h) { return h->gap_pool; }
REFH_EXPORT int16_t const *refh_am(refh_t *h) { return h->am; }
REFH_EXPORT int16_t const *refh_fm(refh_t *h) { return h->fm; }
REFH_EXPORT char const *refh_json(refh_t *h) { return h->json ? h->json : ""; }

REFH_EXPORT void refh_device_stats(refh_t *h, int idx, uint32_t out[8])
{
    r_device *d = h->devs[idx];
    out[0] = d->decode_events;
    out[1] = d->decode_ok;
    out[2] = d->decode_messages;
    for (int i = 0; i < 5; ++i) out[3 + i] = d->decode_fails[i];
}

/* ABI facts the product's own headers must match (checked by tests) */
REFH_EXPORT void refh_abi_facts(uint32_t out[16])
{
    out[0] = sizeof(r_device);
    out[1] = sizeof(bitbuffer_t);
    out[2] = sizeof(pulse_data_t);
    out[3] = offsetof(r_device, modulation);
    out[4] = offsetof(r_device, short_width);
    out[5] = offsetof(r_device, decode_fn);
    out[6] = offsetof(r_device, priority);
    out[7] = offsetof(r_device, decode_events);
    out[8] = offsetof(r_device, decode_ctx);
    out[9] = offsetof(bitbuffer_t, bits_per_row);
    out[10] = offsetof(bitbuffer_t, syncs_before_row);
    out[11] = offsetof(bitbuffer_t, bb);
    out[12] = offsetof(pulse_data_t, pulse);
    out[13] = offsetof(pulse_data_t, gap);
    out[14] = offsetof(pulse_data_t, ook_low_estimate);
    out[15] = offsetof(pulse_data_t, freq1_hz);
}

/* -------- direct entry points to single reference functions (function-level oracle) -------- */

REFH_EXPORT float refh_envelope_detect(uint8_t const *iq, uint16_t *y, uint32_t n) { baseband_init(); return envelope_detect(iq, y, n); }
REFH_EXPORT float refh_magnitude_est_cu8(uint8_t const *iq, uint16_t *y, uint32_t n) { return magnitude_est_cu8(iq, y, n); }
REFH_EXPORT float refh_magnitude_est_cs16(int16_t const *iq, uint16_t *y, uint32_t n) { return magnitude_est_cs16(iq, y, n); }

REFH_EXPORT void refh_low_pass(int16_t state[2], uint16_t const *x, int16_t *y, uint32_t n)
{
    filter_state_t st;
    st.y[0] = state[0];
    st.x[0] = state[1];
    baseband_low_pass_filter(&st, x, y, n);
    state[0] = st.y[0];
    state[1] = st.x[0];
}

REFH_EXPORT void refh_demod_fm(int cs16, void const *iq, int16_t *y, unsigned long n, uint32_t rate, float low_pass, int32_t coef_out[2])
{
    demodfm_state_t st;
    baseband_demod_FM_reset(&st);
    r_logger_set_log_handler(quiet_log, NULL);
    if (cs16) {
        baseband_demod_FM_cs16(&st, iq, y, n, rate, low_pass);
        coef_out[0] = (int32_t)st.alp_32[1];
        coef_out[1] = (int32_t)st.blp_32[0];
    } else {
        baseband_demod_FM(&st, iq, y, n, rate, low_pass);
        coef_out[0] = st.alp_16[1];
        coef_out[1] = st.blp_16[0];
    }
}

/* run one slicer on a caller-built pulse train; returns number of captured events */
REFH_EXPORT int refh_slice(refh_t *h, int dev_idx, int fsk, uint32_t sample_rate, uint32_t num_pulses,
        int32_t const *pulse, int32_t const *gap)
{
    static pulse_data_t pd;
    memset(&pd, 0, sizeof(pd));
    pd.sample_rate = sample_rate;
    pd.num_pulses = num_pulses;
    memcpy(pd.pulse, pulse, num_pulses * sizeof(int32_t));
    memcpy(pd.gap, gap, num_pulses * sizeof(int32_t));
    clear_results(h);
    g_active = h;
    /* a package record so events have something to attach to */
    h->pkgs = grow(h->pkgs, &h->cap_pkgs, 1, sizeof(*h->pkgs));
    memset(&h->pkgs[0], 0, sizeof(h->pkgs[0]));
    h->n_pkgs = 1;
    list_t one = {0};
    list_ensure_size(&one, 2);
    list_push(&one, h->devs[dev_idx]);
    if (fsk)
        run_fsk_demods(&one, &pd);
    else
        run_ook_demods(&one, &pd);
    free(one.elems);
    g_active = NULL;
    return (int)h->n_evts;
}

/* run_ook_demods()/run_fsk_demods() over ALL registered devices on a caller-built pulse train */
REFH_EXPORT int refh_slice_all(refh_t *h, int fsk, uint32_t sample_rate, uint32_t num_pulses,
        int32_t const *pulse, int32_t const *gap)
{
    static pulse_data_t pd;
    memset(&pd, 0, sizeof(pd));
    pd.sample_rate = sample_rate;
    pd.num_pulses = num_pulses;
    memcpy(pd.pulse, pulse, num_pulses * sizeof(int32_t));
    memcpy(pd.gap, gap, num_pulses * sizeof(int32_t));
    clear_results(h);
    g_active = h;
    h->pkgs = grow(h->pkgs, &h->cap_pkgs, 1, sizeof(*h->pkgs));
    memset(&h->pkgs[0], 0, sizeof(h->pkgs[0]));
    h->n_pkgs = 1;
    list_t all = {0};
    list_ensure_size(&all, (size_t)h->n_devs + 1);
    for (int i = 0; i < h->n_devs; ++i) list_push(&all, h->devs[i]);
    if (fsk)
        run_fsk_demods(&all, &pd);
    else
        run_ook_demods(&all, &pd);
    free(all.elems);
    g_active = NULL;
    return (int)h->n_evts;
}

/* Hand the registered r_device structs to an EXTERNAL dispatcher (the product's
   r433b_dispatch_r_devices): real decode_fn restored, output collected by this harness. */
REFH_EXPORT r_device **refh_begin_external_dispatch(refh_t *h)
{
    clear_results(h);
    for (int i = 0; i < h->n_devs; ++i) {
        r_device *d = h->devs[i];
        d->decode_fn = h->orig_fn[i];
        d->decode_events = d->decode_ok = d->decode_messages = 0;
        memset(d->decode_fails, 0, sizeof(d->decode_fails));
    }
    g_active = h;
    return h->devs;
}

REFH_EXPORT void refh_end_external_dispatch(refh_t *h)
{
    for (int i = 0; i < h->n_devs; ++i) h->devs[i]->decode_fn = capture_cb;
    g_active = NULL;
}

REFH_EXPORT void refh_reset_stats(refh_t *h)
{
    for (int i = 0; i < h->n_devs; ++i) {
        r_device *d = h->devs[i];
        d->decode_events = d->decode_ok = d->decode_messages = 0;
        memset(d->decode_fails, 0, sizeof(d->decode_fails));
    }
}

/* file_info_parse_filename(), src/fileformat.c:298: out = {format, sample_rate, center_frequency} */
REFH_EXPORT void refh_parse_filename(char const *name, uint32_t out[3])
{
    file_info_t info;
    memset(&info, 0, sizeof(info));
    file_info_parse_filename(&info, name);
    out[0] = info.format;
    out[1] = info.sample_rate;
    out[2] = info.center_frequency;
}

/* -------- pulse-level I/O of the reference (src/pulse_data.c, src/rfraw.c), SURVEY 8(f4) -------- */
#include "rfraw.h"

/* pulse_data_load() until a package comes back empty (the .ook loop of src/rtl_433.c:1755-1760); the loaded
   pulse_data_t are stored back to back in `out` (cap entries).  Returns the number loaded. */
REFH_EXPORT int refh_load_ook(char const *text, size_t len, uint32_t samp_rate, pulse_data_t *out, int cap)
{
    FILE *f = fmemopen((void *)text, len, "r");
    if (!f) return -1;
    struct timeval now = {0, 0};
    r_logger_set_log_handler(quiet_log, NULL);
    int n = 0;
    while (n < cap) {
        pulse_data_load(f, &now, &out[n], samp_rate);
        if (!out[n].num_pulses) break;
        n++;
    }
    fclose(f);
    return n;
}

REFH_EXPORT int refh_rfraw(char const *line, pulse_data_t *out)
{
    memset(out, 0, sizeof(*out));
    if (!rfraw_check(line)) return 0;
    rfraw_parse(out, line);
    return 1;
}

/* pulse_data_dump() / pulse_data_print_vcd() / pulse_data_dump_raw() output into a caller buffer */
REFH_EXPORT size_t refh_dump_ook(pulse_data_t const *pd, char *buf, size_t cap)
{
    char *mem = NULL;
    size_t n = 0;
    FILE *f = open_memstream(&mem, &n);
    pulse_data_dump(f, pd);
    fclose(f);
    if (n < cap) memcpy(buf, mem, n + 1);
    free(mem);
    return n;
}

REFH_EXPORT size_t refh_dump_vcd(pulse_data_t const *pd, int ch_id, int with_header, char *buf, size_t cap)
{
    char *mem = NULL;
    size_t n = 0;
    FILE *f = open_memstream(&mem, &n);
    if (with_header) pulse_data_print_vcd_header(f, pd->sample_rate);
    pulse_data_print_vcd(f, pd, ch_id);
    fclose(f);
    if (n < cap) memcpy(buf, mem, n + 1);
    free(mem);
    return n;
}

REFH_EXPORT void refh_dump_raw(uint8_t *buf, unsigned len, uint64_t buf_offset, pulse_data_t const *pd, uint8_t bits)
{
    pulse_data_dump_raw(buf, len, buf_offset, pd, bits);
}

/* run_ook_demods() / run_fsk_demods() (by fsk_f2_est, src/rtl_433.c:1774-1778) over all registered devices on a
   loaded pulse_data_t; events are captured like those of refh_run_stream() */
REFH_EXPORT int refh_slice_pulse_data(refh_t *h, pulse_data_t *pd)
{
    clear_results(h);
    g_active = h;
    h->pkgs = grow(h->pkgs, &h->cap_pkgs, 1, sizeof(*h->pkgs));
    memset(&h->pkgs[0], 0, sizeof(h->pkgs[0]));
    h->n_pkgs = 1;
    list_t all = {0};
    list_ensure_size(&all, (size_t)h->n_devs + 1);
    for (int i = 0; i < h->n_devs; ++i) list_push(&all, h->devs[i]);
    if (pd->fsk_f2_est)
        run_fsk_demods(&all, pd);
    else
        run_ook_demods(&all, pd);
    free(all.elems);
    g_active = NULL;
    return (int)h->n_evts;
}

/* -------- decoder length gates: what does decode_fn do with events that are too short? --------
   For device idx find the largest T <= max_bits such that EVERY probed bitbuffer with at least one row whose rows all
   hold fewer than T bits makes the decoder return a constant code <= 0 without producing output -- code[0] for buffers
   of one row, code[1] for buffers of several (many decoders test num_rows first): all 2^L contents of a single row
   of L bits (exhaustive, L < T), and n_random more bitbuffers per L (random row count, lengths <= L with one row of
   exactly L, random bytes -- also beyond the row lengths --, random sync counts).  Returns T. */
static uint64_t probe_rng(uint64_t *s)
{
    *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
    return *s;
}

static int probe_call(refh_t *h, int idx, bitbuffer_t const *bb, int *out_delta)
{
    static bitbuffer_t work;
    work = *bb;
    uint64_t before = h->decoded_msgs;
    int ret = h->orig_fn[idx] ? h->orig_fn[idx](h->devs[idx], &work) : 0;
    *out_delta = (int)(h->decoded_msgs - before);
    return ret;
}

REFH_EXPORT int refh_probe_gate(refh_t *h, int idx, int max_bits, int n_random, uint64_t seed, int code[2])
{
    code[0] = code[1] = 0;
    if (idx < 0 || idx >= h->n_devs || !h->orig_fn[idx]) return 0;
    refh_t *saved = g_active;
    g_active = h;
    static bitbuffer_t bb;
    int delta = 0;
    /* c1: one empty row; cN: two empty rows.  (A buffer without any row -- an nrzs event can be one -- is never gated.) */
    memset(&bb, 0, sizeof(bb));
    bb.num_rows = bb.free_row = 1;
    int const c1 = probe_call(h, idx, &bb, &delta);
    int bad = delta;
    bb.num_rows = bb.free_row = 2;
    int const cN = probe_call(h, idx, &bb, &delta);
    bad |= delta;
    code[0] = c1;
    code[1] = cN;
    int T = 0;
    if (bad || c1 > 0 || c1 < -4 || cN > 0 || cN < -4) { g_active = saved; return 0; }
    uint64_t rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull + (uint64_t)idx;
    for (int L = 0; L < max_bits; ++L) {
        int ok = 1;
        /* one row of exactly L bits, every content */
        for (uint32_t v = 0; ok && v < (1u << L); ++v) {
            memset(&bb, 0, 204 + 8);
            bb.num_rows = 1;
            bb.free_row = 1;
            bb.bits_per_row[0] = (uint16_t)L;
            uint32_t left = L ? v << (32 - L) : 0;
            bb.bb[0][0] = left >> 24; bb.bb[0][1] = left >> 16; bb.bb[0][2] = left >> 8; bb.bb[0][3] = left;
            if (probe_call(h, idx, &bb, &delta) != c1 || delta) ok = 0;
        }
        /* one row with dirt beyond its length and sync counts; two or more rows, none longer than L */
        for (int k = 0; ok && k < n_random; ++k) {
            memset(&bb, 0, sizeof(bb));
            uint64_t r = probe_rng(&rng);
            int rows = k % 8 == 0 ? 1 : 2 + (int)(r % ((r >> 8) % 4 == 0 ? 49 : 5));
            int longest = (int)((r >> 16) % rows);
            int mode = (int)((r >> 24) % 5); /* random / zeros / ones / 0xaa / 0x55 */
            bb.num_rows = (uint16_t)rows;
            bb.free_row = (uint16_t)rows;
            for (int row = 0; row < rows; ++row) {
                uint64_t q = probe_rng(&rng);
                int bits = row == longest ? L : (int)(q % (L + 1));
                bb.bits_per_row[row] = (uint16_t)bits;
                bb.syncs_before_row[row] = (q >> 20) % 7 == 0 ? (uint16_t)((q >> 24) % 4) : 0;
                for (int b = 0; b < 4; ++b) {
                    uint8_t byte = mode == 0 ? (uint8_t)(q >> (32 + 8 * b)) : mode == 1 ? 0 : mode == 2 ? 0xff : mode == 3 ? 0xaa : 0x55;
                    /* half of the random buffers are clean beyond the row length, as the bit writer leaves them */
                    if ((r >> 40) & 1) {
                        int keep = bits - 8 * b;
                        byte = keep <= 0 ? 0 : keep >= 8 ? byte : (uint8_t)(byte & (0xff00 >> keep));
                    }
                    bb.bb[row][b] = byte;
                }
            }
            if (probe_call(h, idx, &bb, &delta) != (rows == 1 ? c1 : cN) || delta) ok = 0;
        }
        if (!ok) break;
        T = L + 1;
    }
    g_active = saved;
    return T;
}

/* decode_fn of device idx on one bitbuffer (validation of a gate table with independent inputs) */
REFH_EXPORT int refh_call_decoder(refh_t *h, int idx, bitbuffer_t const *bb, int *outputs)
{
    refh_t *saved = g_active;
    g_active = h;
    int ret = probe_call(h, idx, bb, outputs);
    g_active = saved;
    return ret;
}

/* -------- pulse analyzer of the reference (src/pulse_analyzer.c), SURVEY 8(f3) -------- */
#include <unistd.h>
#include <fcntl.h>
#include "pulse_analyzer.h"

/* pulse_analyzer() on a copy of `pd`; everything it prints to stderr goes to `buf`, the events of its trial
   demodulation are captured like any others (device index -1).  Returns the text length. */
REFH_EXPORT size_t refh_analyze(refh_t *h, pulse_data_t const *pd, int package_type, char *buf, size_t cap)
{
    static pulse_data_t work;
    work = *pd;
    clear_results(h);
    g_active = h;
    h->pkgs = grow(h->pkgs, &h->cap_pkgs, 1, sizeof(*h->pkgs));
    memset(&h->pkgs[0], 0, sizeof(h->pkgs[0]));
    h->n_pkgs = 1;
    r_device device;
    memset(&device, 0, sizeof(device));
    device.log_fn = sink_log;
    device.decode_fn = capture_cb;
    char path[] = "/tmp/refh_analyze_XXXXXX";
    int tmp = mkstemp(path);
    if (tmp < 0) return 0;
    fflush(stderr);
    int saved = dup(2);
    dup2(tmp, 2);
    pulse_analyzer(&work, package_type, &device);
    fflush(stderr);
    dup2(saved, 2);
    close(saved);
    off_t len = lseek(tmp, 0, SEEK_END);
    lseek(tmp, 0, SEEK_SET);
    size_t n = (size_t)len < cap - 1 ? (size_t)len : cap - 1;
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(tmp, buf + got, n - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    buf[got] = 0;
    close(tmp);
    unlink(path);
    g_active = NULL;
    return (size_t)len;
}

/* -------- SigMF container of the reference (src/sigmf.c), SURVEY 8(f2) -------- */
#include "sigmf.h"

/* sigmf_reader_open(): out = {rc, sample_rate, first_frequency, position of the stream after opening} -- the file
   loop of src/rtl_433.c:1712-1723 then reads cu8 blocks from that position to the end of the archive */
REFH_EXPORT void refh_sigmf_open(char const *path, uint64_t out[4])
{
    sigmf_t s;
    memset(&s, 0, sizeof(s));
    r_logger_set_log_handler(quiet_log, NULL);
    fflush(stdout);
    int saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); /* json_parse() printf()s what it reads */
    int rc = sigmf_reader_open(&s, path);
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    close(nul);
    out[0] = (uint64_t)(int64_t)rc;
    out[1] = s.sample_rate;
    out[2] = s.first_frequency;
    out[3] = s.mtar.stream ? (uint64_t)ftell(s.mtar.stream) : 0;
    if (s.mtar.stream) fclose(s.mtar.stream);
    sigmf_free_items(&s);
}
