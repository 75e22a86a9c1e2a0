/* r433b_abi.h -- the three reference structs the host-side replay has to speak, re-declared so
 * libr433b builds without the reference tree.  Layouts must stay byte-identical to
 *   r_device     include/r_device.h:59-92      (152 bytes on LP64)
 *   bitbuffer_t  include/bitbuffer.h:34-40     (6604 bytes)
 *   pulse_data_t include/pulse_data.h:30-50    (9672 bytes)
 * of merbanan/rtl_433 25.12; tests/test_abi.py checks every size and offset below against the
 * compiled reference (oracle/_ref).  A program that includes the reference's own headers must
 * NOT include this file; the struct tags are the same on purpose, so pointers are compatible.
 */
#ifndef R433B_ABI_H_
#define R433B_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R433B_BITBUF_COLS 128
#define R433B_BITBUF_ROWS 50
#define R433B_PD_MAX_PULSES 1200

struct data;
struct bitbuffer;

struct r_device {
    unsigned protocol_num;
    char const *name;
    unsigned modulation;
    float short_width;
    float long_width;
    float reset_limit;
    float gap_limit;
    float sync_width;
    float tolerance;
    int (*decode_fn)(struct r_device *decoder, struct bitbuffer *bitbuffer);
    struct r_device *(*create_fn)(char const *args);
    unsigned priority;
    unsigned disabled;
    char const *const *fields;
    int verbose;
    int verbose_bits;
    void (*log_fn)(struct r_device *decoder, int level, struct data *data);
    void (*output_fn)(struct r_device *decoder, struct data *data);
    unsigned decode_events;
    unsigned decode_ok;
    unsigned decode_messages;
    unsigned decode_fails[5];
    void *decode_ctx;
    void *output_ctx;
};

struct bitbuffer {
    uint16_t num_rows;
    uint16_t free_row;
    uint16_t bits_per_row[R433B_BITBUF_ROWS];
    uint16_t syncs_before_row[R433B_BITBUF_ROWS];
    uint8_t bb[R433B_BITBUF_ROWS][R433B_BITBUF_COLS];
};

struct pulse_data {
    uint64_t offset;
    uint32_t sample_rate;
    unsigned depth_bits;
    unsigned start_ago;
    unsigned end_ago;
    unsigned int num_pulses;
    int pulse[R433B_PD_MAX_PULSES];
    int gap[R433B_PD_MAX_PULSES];
    int ook_low_estimate;
    int ook_high_estimate;
    int fsk_f1_est;
    int fsk_f2_est;
    float freq1_hz;
    float freq2_hz;
    float centerfreq_hz;
    float range_db;
    float rssi_db;
    float snr_db;
    float noise_db;
};

#ifdef __cplusplus
}
#endif
#endif /* R433B_ABI_H_ */
