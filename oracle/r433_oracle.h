/* r433_oracle.h -- TEST INFRASTRUCTURE: CPU restatement of the rtl_433 IQ->bitbuffer hot path.
 *
 * Not a product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * reference legs may load this.  Pinned against the compiled reference (oracle/_ref) and
 * the committed fixtures under tests/golden/ by tests/test_oracle_*.py.
 */
#ifndef R433_ORACLE_H_
#define R433_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_PULSES 1200 /* include/pulse_data.h:21 */
#define ORC_BB_ROWS 50      /* include/bitbuffer.h:28 */
#define ORC_BB_COLS 128     /* include/bitbuffer.h:27 */

/* same layout as refh_package / refh_event in oracle/ref_harness.c so one reader serves both */
typedef struct {
    int32_t type;
    int32_t block;
    uint64_t offset;
    uint32_t sample_rate, depth_bits, start_ago, end_ago, num_pulses;
    int32_t ook_low_estimate, ook_high_estimate, fsk_f1_est, fsk_f2_est;
    float freq1_hz, freq2_hz, centerfreq_hz, range_db, rssi_db, snr_db, noise_db;
    float sample_file_pos;
    uint32_t pulse_off, pulse_count;
    uint32_t first_event, num_events;
} orc_package;

typedef struct {
    uint32_t package;
    uint32_t dev;
    int32_t ret;
    uint32_t bb_idx;
    uint64_t hash;
} orc_event;

/* include/bitbuffer.h:34-40, byte for byte (6604 bytes) */
typedef struct {
    uint16_t num_rows;
    uint16_t free_row;
    uint16_t bits_per_row[ORC_BB_ROWS];
    uint16_t syncs_before_row[ORC_BB_ROWS];
    uint8_t bb[ORC_BB_ROWS][ORC_BB_COLS];
} orc_bitbuffer;

typedef struct {
    uint32_t modulation; /* enum modulation_types, include/r_device.h:24-40 */
    float short_width, long_width, reset_limit, gap_limit, sync_width, tolerance;
    uint32_t priority;
} orc_device;

typedef struct orc orc_t;

orc_t *orc_create(void);
void orc_destroy(orc_t *o);
void orc_set_capture(orc_t *o, int store_bitbuffers, int store_stages);
void orc_set_levels(orc_t *o, int use_mag_est, float level_limit, float min_level, float min_snr);
void orc_set_fm_low_pass(orc_t *o, float fm_low_pass);
int orc_add_device(orc_t *o, orc_device const *d);
int orc_num_devices(orc_t *o);

/* one capture file: block loop + flush + reset, as rtl_433 -r does */
/* src/rtl_433.c:1811-1825 */
void orc_cf32_to_cs16(float const *in, int16_t *out, size_t n);

int orc_run_stream(orc_t *o, void const *iq, size_t bytes, int sample_size, uint32_t samp_rate,
        uint32_t center_freq, int fpdm_mode, uint32_t block_bytes);

size_t orc_num_packages(orc_t *o);
size_t orc_num_events(orc_t *o);
size_t orc_num_bitbuffers(orc_t *o);
size_t orc_num_stage(orc_t *o);
orc_package const *orc_packages(orc_t *o);
orc_event const *orc_events(orc_t *o);
orc_bitbuffer const *orc_bitbuffers(orc_t *o);
int32_t const *orc_pulse_pool(orc_t *o);
int32_t const *orc_gap_pool(orc_t *o);
int16_t const *orc_am(orc_t *o);
int16_t const *orc_fm(orc_t *o);

/* function-level entry points */
void orc_envelope_cu8(uint8_t const *iq, uint16_t *y, uint32_t n);
void orc_magnitude_cu8(uint8_t const *iq, uint16_t *y, uint32_t n);
void orc_magnitude_cs16(int16_t const *iq, uint16_t *y, uint32_t n);
void orc_low_pass(int16_t state[2], uint16_t const *x, int16_t *y, uint32_t n);
void orc_fm_coeffs(int cs16, uint32_t rate, float low_pass, int32_t coef_out[2]);
void orc_demod_fm(int cs16, void const *iq, int16_t *y, unsigned long n, uint32_t rate, float low_pass, int32_t coef_out[2]);
void orc_detector_levels(int use_mag_est, float level_limit, float min_level, float min_snr, int32_t out[4]);
int orc_slice(orc_t *o, int dev_idx, uint32_t sample_rate, uint32_t num_pulses, int32_t const *pulse, int32_t const *gap);

#ifdef __cplusplus
}
#endif
#endif
