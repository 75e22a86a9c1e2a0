/* r433_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the rtl_433 IQ->bitbuffer hot path.
 *
 * Plain scalar C, one stream at a time, written for clarity.  Each function cites the
 * reference file:line (relative to the merbanan/rtl_433 tree, release 25.12) whose behaviour it
 * restates.  This file is the checker for the CUDA path and is itself pinned against the
 * compiled reference (oracle/_ref/libr433ref.so) and tests/golden/ -- see tests/test_oracle_*.py.
 * It must never be linked into, or called from, the product (rtl_433_b200/).
 *
 * Build: gcc -O2 -ffp-contract=off (the float expressions below must not be fused).
 */
#include "r433_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define DEFAULT_BLOCK_BYTES (16 * 32 * 512) /* include/rtl_433.h:17 */

/* ---------------------------------------------------------------- envelope / magnitude -- */

/* src/baseband.c:22-45: (127-I)^2 + (127-Q)^2, note 127 not 128 */
void orc_envelope_cu8(uint8_t const *iq, uint16_t *y, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) {
        int di = 127 - (int)iq[2 * k];
        int dq = 127 - (int)iq[2 * k + 1];
        y[k] = (uint16_t)(di * di + dq * dq);
    }
}

/* src/baseband.c:65-79: 122*max + 51*min of |v-128| */
void orc_magnitude_cu8(uint8_t const *iq, uint16_t *y, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) {
        int a = abs((int)iq[2 * k] - 128);
        int b = abs((int)iq[2 * k + 1] - 128);
        int hi = a > b ? a : b, lo = a > b ? b : a;
        y[k] = (uint16_t)(122 * hi + 51 * lo);
    }
}

/* src/baseband.c:96-110: (122*max + 51*min) >> 8 of |v| */
void orc_magnitude_cs16(int16_t const *iq, uint16_t *y, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t a = (uint32_t)abs((int)iq[2 * k]);
        uint32_t b = (uint32_t)abs((int)iq[2 * k + 1]);
        uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
        y[k] = (uint16_t)((122 * hi + 51 * lo) >> 8);
    }
}

/* ------------------------------------------------------------------- AM low-pass (IIR) -- */

/* src/baseband.c:151-152: FIX(0.85408)>>1 and FIX(0.07296)>>1, FIX(x) = (int)(x*32768) */
static int lpf_a1(void) { return ((int)(0.85408 * 32768)) >> 1; }
static int lpf_b0(void) { return ((int)(0.07296 * 32768)) >> 1; }

/* src/baseband.c:145-169.  state[0] = y[-1], state[1] = x[-1] (an int16: a raw envelope of
   32768 re-enters the next block as -32768, :167) */
void orc_low_pass(int16_t state[2], uint16_t const *x, int16_t *y, uint32_t n)
{
    int const a1 = lpf_a1(), b0 = lpf_b0();
    if (n < 1) return;
    int yp = state[0];
    int xp = state[1];
    for (uint32_t k = 0; k < n; ++k) {
        int xc = x[k];
        yp = (int16_t)((a1 * yp + b0 * (xc + xp)) >> 14);
        y[k] = (int16_t)yp;
        xp = xc;
    }
    state[0] = (int16_t)yp;
    state[1] = (int16_t)x[n - 1];
}

/* ------------------------------------------------------------------------ FM demod ------ */

/* src/baseband.c:217-231 (cu8, Q0.15 halved gain) and :310-324 (cs16, Q0.30) */
void orc_fm_coeffs(int cs16, uint32_t rate, float low_pass, int32_t coef_out[2])
{
    if (low_pass > 1e4f)
        low_pass = low_pass / rate;
    else if (low_pass >= 1.0f)
        low_pass = 1e6f / low_pass / rate;
    double ita = 1.0 / tan(M_PI_2 * low_pass);
    if (!cs16) {
        double gain = 1.0 / (1.0 + ita) / 2;
        coef_out[0] = (int)(((ita - 1.0) * gain) * 32768);
        coef_out[1] = (int)(gain * 32768);
    } else {
        double gain = 1.0 / (1.0 + ita);
        coef_out[0] = (int)(((ita - 1.0) * gain) * (1 << 30));
        coef_out[1] = (int)(gain * (1 << 30));
    }
}

/* src/baseband.c:181-202; pi == 32767; C division truncates toward zero */
static int16_t angle16(int32_t y, int32_t x)
{
    int32_t const q = INT16_MAX / 4, q3 = 3 * INT16_MAX / 4;
    int32_t ay = y < 0 ? -y : y;
    int32_t ang;
    if (x == 0 && y == 0) return 0;
    if (x >= 0) {
        int32_t d = ay + x;
        if (d == 0) d = 1;
        ang = q - q * (x - ay) / d;
    } else {
        int32_t d = ay - x;
        if (d == 0) d = 1;
        ang = q3 - q * (x + ay) / d;
    }
    return (int16_t)(y < 0 ? -ang : ang);
}

/* src/baseband.c:281-300; no (0,0) special case; arguments arrive narrowed to int32 (:352).
   |y| of INT32_MIN is undefined in the reference; two's complement wrap is used here. */
static int32_t angle32(int32_t y, int32_t x)
{
    int64_t const q = INT32_MAX / 4, q3 = 3ll * INT32_MAX / 4;
    int64_t ay = (int64_t)(int32_t)(y < 0 ? (int32_t)(0u - (uint32_t)y) : y);
    int64_t ang;
    if (x >= 0) {
        int64_t d = ay + x;
        if (d == 0) d = 1;
        ang = q - q * (x - ay) / d;
    } else {
        int64_t d = ay - x;
        if (d == 0) d = 1;
        ang = q3 - q * (x + ay) / d;
    }
    return (int32_t)(y < 0 ? -ang : ang);
}

typedef struct {
    int32_t xr, xi, xf, yf; /* include/baseband.h:97-107 */
} fm_state;

/* src/baseband.c:236-271 */
static void fm_cu8(fm_state *st, int32_t const coef[2], uint8_t const *iq, int16_t *out, size_t n)
{
    int16_t xr = (int16_t)st->xr, xi = (int16_t)st->xi, xf = (int16_t)st->xf, yf = (int16_t)st->yf;
    for (size_t k = 0; k < n; ++k) {
        int16_t pr_ = xr, pi_ = xi, pf = xf, py = yf;
        xr = (int16_t)(iq[2 * k] - 128);
        xi = (int16_t)(iq[2 * k + 1] - 128);
        int32_t re = xr * pr_ + xi * pi_;
        int32_t im = xi * pr_ - xr * pi_;
        xf = angle16(im, re);
        yf = (int16_t)((coef[0] * py + coef[1] * (xf + pf)) >> 14);
        out[k] = yf;
    }
    st->xr = xr; st->xi = xi; st->xf = xf; st->yf = yf;
}

/* src/baseband.c:329-365 */
static void fm_cs16(fm_state *st, int32_t const coef[2], int16_t const *iq, int16_t *out, size_t n)
{
    int32_t xr = st->xr, xi = st->xi, xf = st->xf, yf = st->yf;
    int64_t const a1 = coef[0], b0 = coef[1];
    for (size_t k = 0; k < n; ++k) {
        int32_t pr_ = xr, pi_ = xi, pf = xf, py = yf;
        xr = iq[2 * k];
        xi = iq[2 * k + 1];
        int64_t re = (int64_t)xr * pr_ + (int64_t)xi * pi_;
        int64_t im = (int64_t)xi * pr_ - (int64_t)xr * pi_;
        xf = angle32((int32_t)(uint32_t)(uint64_t)im, (int32_t)(uint32_t)(uint64_t)re);
        yf = (int32_t)((a1 * py + b0 * ((int64_t)xf + pf)) >> 30);
        out[k] = (int16_t)(yf >> 16);
    }
    st->xr = xr; st->xi = xi; st->xf = xf; st->yf = yf;
}

void orc_demod_fm(int cs16, void const *iq, int16_t *y, unsigned long n, uint32_t rate, float low_pass, int32_t coef_out[2])
{
    fm_state st = {0, 0, 0, 0};
    orc_fm_coeffs(cs16, rate, low_pass, coef_out);
    if (cs16)
        fm_cs16(&st, coef_out, iq, y, n);
    else
        fm_cu8(&st, coef_out, iq, y, n);
}

/* --------------------------------------------------------------------- detector levels -- */

/* include/baseband.h:43-46 with _exp10f(x) == powf(10, x) */
static int db_to_amp(float db) { return (int)(powf(10, (db + 42.1442f) / 10.0f)); }
static int db_to_mag(float db) { return (int)(powf(10, (db + 84.2884f) / 20.0f)); }
static int db_to_amp_f(float db) { return (int)(0.5 + powf(10, db / 10.0f)); }
static int db_to_mag_f(float db) { return (int)(0.5 + powf(10, db / 20.0f)); }

/* src/pulse_detect.c:86-98 and :24.  out = {fixed_high, min_high, high_low_ratio, max_high} */
void orc_detector_levels(int use_mag_est, float level_limit, float min_level, float min_snr, int32_t out[4])
{
    if (use_mag_est) {
        out[0] = level_limit < 0.0 ? db_to_mag(level_limit) : 0;
        out[1] = db_to_mag(min_level);
        out[2] = db_to_mag_f(min_snr);
    } else {
        out[0] = level_limit < 0.0 ? db_to_amp(level_limit) : 0;
        out[1] = db_to_amp(min_level);
        out[2] = db_to_amp_f(min_snr);
    }
    out[3] = db_to_amp(0); /* OOK_MAX_HIGH_LEVEL is DB_TO_AMP(0) in either mode */
}

/* --------------------------------------------------------------------------- pulse train -- */

typedef struct {
    uint64_t offset;
    uint32_t sample_rate;
    uint32_t start_ago, end_ago;
    uint32_t n;
    int pulse[ORC_MAX_PULSES];
    int gap[ORC_MAX_PULSES];
    int low_est, high_est, f1_est, f2_est;
} train_t; /* the integer part of pulse_data_t, include/pulse_data.h:30-50 */

static void train_clear(train_t *t) { memset(t, 0, sizeof(*t)); } /* src/pulse_data.c:22 */

/* src/pulse_data.c:27-34: drop the oldest 600; `offset` grows by the COUNT (kept quirk) */
static void train_shift(train_t *t)
{
    int const half = ORC_MAX_PULSES / 2;
    memmove(t->pulse, t->pulse + half, (ORC_MAX_PULSES - half) * sizeof(int));
    memmove(t->gap, t->gap + half, (ORC_MAX_PULSES - half) * sizeof(int));
    t->n -= half;
    t->offset += half;
}

/* ----------------------------------------------------------------------- FSK sub-detector -- */

enum { FK_INIT = 0, FK_HIGH = 1, FK_LOW = 2, FK_ERR = 3 };

typedef struct {
    unsigned len;
    int st;
    int f1, f2;
    int16_t vmax, vmin;
    int skip;
} fsk_t; /* include/pulse_detect_fsk.h:23-41 */

static void fsk_reset(fsk_t *f) /* src/pulse_detect_fsk.c:26-32 */
{
    memset(f, 0, sizeof(*f));
    f->vmax = INT16_MIN;
    f->vmin = INT16_MAX;
    f->skip = 40;
}

/* src/pulse_detect_fsk.c:34-141 ("classic"): two frequency trackers, short runs folded back */
static void fsk_classic(fsk_t *f, int16_t v, train_t *t)
{
    int d1 = abs(v - f->f1);
    int d2 = abs(v - f->f2);
    f->len += 1;
    if (f->st == FK_INIT) {
        if (f->len < 10) {
            f->f1 = f->f1 / 2 + v / 2;
        } else if (d1 > 6000 / 2) {
            if (v > f->f1) { /* started low: a gap came first */
                f->st = FK_HIGH;
                f->f2 = f->f1;
                f->f1 = v;
                t->pulse[0] = 0;
                t->gap[0] = (int)f->len;
                t->n += 1;
                f->len = 0;
            } else { /* started high */
                f->st = FK_LOW;
                f->f2 = v;
                t->pulse[0] = (int)f->len;
                f->len = 0;
            }
        } else {
            f->f1 += v / 16 - f->f1 / 16;
        }
    } else if (f->st == FK_HIGH) {
        if (d1 > d2) {
            f->st = FK_LOW;
            if (f->len >= 10) {
                t->pulse[t->n] = (int)f->len;
                f->len = 0;
            } else {
                f->len += (unsigned)t->gap[t->n - 1];
                t->n -= 1;
                if (t->n == 0 && t->pulse[0] == 0) {
                    f->f1 = f->f2;
                    f->st = FK_INIT;
                }
            }
        } else {
            int div = v > f->f1 ? 16 : 64;
            f->f1 += v / div - f->f1 / div;
        }
    } else if (f->st == FK_LOW) {
        if (d2 > d1) {
            f->st = FK_HIGH;
            if (f->len >= 10) {
                t->gap[t->n] = (int)f->len;
                t->n += 1;
                f->len = 0;
                if (t->n >= ORC_MAX_PULSES) train_shift(t);
            } else {
                f->len += (unsigned)t->pulse[t->n];
                if (t->n == 0) f->st = FK_INIT;
            }
        } else {
            int div = v < f->f2 ? 16 : 64;
            f->f2 += v / div - f->f2 / div;
        }
    }
}

/* src/pulse_detect_fsk.c:143-156 */
static void fsk_finish(fsk_t *f, train_t *t)
{
    if (t->n >= ORC_MAX_PULSES) return;
    f->len += 1;
    if (f->st == FK_HIGH) {
        t->pulse[t->n] = (int)f->len;
        t->gap[t->n] = 0;
    } else {
        t->gap[t->n] = (int)f->len;
    }
    t->n += 1;
}

/* src/pulse_detect_fsk.c:158-221 ("minmax"): decaying envelope, mid-point comparator.
   The f1/f2 names are swapped relative to the states on purpose (:192, :208). */
static void fsk_minmax(fsk_t *f, int16_t v, train_t *t)
{
    if (f->skip == 0) {
        if (v > f->vmax) f->vmax = v;
        if (v < f->vmin) f->vmin = v;
        int16_t mid = (int16_t)((f->vmax + f->vmin) / 2);
        if (v > mid) f->vmax = (int16_t)(f->vmax - 10);
        if (v < mid) f->vmin = (int16_t)(f->vmin + 10);
        f->len += 1;
        if (f->st == FK_INIT) {
            f->st = v > mid ? FK_HIGH : FK_LOW;
        } else if (f->st == FK_HIGH) {
            if (v < mid) {
                f->st = FK_LOW;
                t->pulse[t->n] = (int)f->len;
                f->len = 0;
            }
            f->f2 += v / 64 - f->f2 / 64;
        } else if (f->st == FK_LOW) {
            if (v > mid) {
                f->st = FK_HIGH;
                t->gap[t->n] = (int)f->len;
                t->n += 1;
                f->len = 0;
                if (t->n >= ORC_MAX_PULSES) train_shift(t);
            }
            f->f1 += v / 64 - f->f1 / 64;
        }
    }
    if (f->skip > 0) f->skip -= 1;
}

/* -------------------------------------------------------------------- OOK package detector -- */

enum { ST_IDLE = 0, ST_PULSE = 1, ST_GAP_START = 2, ST_GAP = 3 };

typedef struct {
    /* levels, src/pulse_detect.c:31-34 */
    int fixed_high, min_high, ratio, max_high;
    /* running state, :36-53 */
    int st;
    int run;       /* pulse_length */
    int longest;   /* max_pulse */
    int pos;       /* data_counter */
    int lead_in;
    int low, high;
    fsk_t fsk;
} detector_t;

static void detector_reset(detector_t *d) /* src/pulse_detect.c:74-84 */
{
    d->st = ST_IDLE;
    d->run = d->longest = d->pos = d->lead_in = d->low = d->high = 0;
    fsk_reset(&d->fsk);
}

static void feed_fsk(detector_t *d, int16_t v, train_t *fsk, unsigned fpdm)
{
    if (fpdm == 0)
        fsk_classic(&d->fsk, v, fsk);
    else
        fsk_minmax(&d->fsk, v, fsk);
}

/* the FSK package hand-over, src/pulse_detect.c:239-253 and :387-410 */
static void close_fsk(detector_t *d, train_t *ook, train_t *fsk, unsigned fpdm, int len)
{
    if (fpdm == 0) fsk_finish(&d->fsk, fsk);
    fsk->f1_est = d->fsk.f1;
    fsk->f2_est = d->fsk.f2;
    fsk->low_est = d->low;
    fsk->high_est = d->high;
    ook->end_ago = (uint32_t)(len - d->pos);
    fsk->end_ago = (uint32_t)(len - d->pos);
    d->st = ST_IDLE;
}

/* the OOK package hand-over, src/pulse_detect.c:263-272, :431-439, :451-468 */
static void close_ook(detector_t *d, train_t *ook, int len)
{
    d->st = ST_IDLE;
    ook->low_est = d->low;
    ook->high_est = d->high;
    ook->end_ago = (uint32_t)(len - d->pos);
}

/* src/pulse_detect.c:199-483.  Returns 0 (block consumed), 1 (OOK package), 2 (FSK package).
   On 1/2 the position is NOT advanced: the same sample is looked at again, in IDLE. */
static int detect(detector_t *d, int16_t const *am, int16_t const *fm, int len, uint32_t rate,
        uint64_t base, train_t *ook, train_t *fsk, unsigned fpdm)
{
    if (len == 0) { /* end-of-input flush, :204-278 */
        int st = d->st;
        if (st == ST_PULSE) {
            if (d->run < 10) {
                if (ook->n <= 1) {
                    d->st = ST_IDLE;
                    st = ST_IDLE;
                } else {
                    st = ST_GAP_START;
                }
            } else {
                ook->pulse[ook->n] = d->run;
                if (d->run > d->longest) d->longest = d->run;
                d->run = 0;
                st = ST_GAP_START;
            }
        }
        if (st == ST_GAP_START) {
            if (fsk->n > 16) {
                d->st = ST_GAP;
                close_fsk(d, ook, fsk, fpdm, len);
                return 2;
            }
            st = ST_GAP;
        }
        if (st == ST_GAP) {
            ook->gap[ook->n] = d->run;
            ook->n += 1;
            close_ook(d, ook, len);
            return 1;
        }
    }

    int const per_ms = (int)(rate / 1000);
    if (d->high < d->min_high) d->high = d->min_high; /* :283 */
    if (d->pos == 0) {                                /* :285-289 */
        ook->start_ago += (uint32_t)len;
        fsk->start_ago += (uint32_t)len;
    }
    int spurious_eop = 0; /* :291, a local: forgotten at every call boundary */

    for (; d->pos < len; d->pos++) {
        int16_t const a = am[d->pos];
        int16_t thr = (int16_t)((d->low + (d->high < d->max_high ? d->high : d->max_high)) / 2);
        if (d->fixed_high != 0) thr = (int16_t)d->fixed_high;
        int16_t const hys = (int16_t)(thr / 8);
        int const above = a > (thr + hys);
        int const below = a < (thr - hys);

        switch (d->st) {
        case ST_IDLE:
            if (above && d->lead_in > 1024) { /* :309-324 */
                train_clear(ook);
                train_clear(fsk);
                ook->sample_rate = fsk->sample_rate = rate;
                ook->offset = fsk->offset = base + (uint64_t)d->pos;
                ook->start_ago = fsk->start_ago = (uint32_t)(len - d->pos);
                d->run = 0;
                d->longest = 0;
                fsk_reset(&d->fsk);
                d->st = ST_PULSE;
            } else { /* :325-334 noise floor tracker */
                int delta = a - d->low;
                d->low += delta / 1024;
                d->low += delta > 0 ? 1 : -1;
                d->high = d->ratio * d->low;
                if (d->high < d->min_high) d->high = d->min_high;
                if (d->lead_in <= 1024) d->lead_in += 1;
            }
            break;
        case ST_PULSE: /* :336-375 */
            d->run += 1;
            if (below) {
                if (d->run < 10) {
                    if (ook->n <= 1) {
                        d->st = ST_IDLE;
                    } else {
                        spurious_eop = 1;
                        d->st = ST_GAP;
                    }
                } else {
                    ook->pulse[ook->n] = d->run;
                    if (d->run > d->longest) d->longest = d->run;
                    d->run = 0;
                    d->st = ST_GAP_START;
                }
            } else {
                d->high += a / 64 - d->high / 64;
                if (d->high < d->min_high) d->high = d->min_high;
                ook->f1_est += fm[d->pos] / 64 - ook->f1_est / 64;
            }
            if (ook->n == 0) feed_fsk(d, fm[d->pos], fsk, fpdm);
            break;
        case ST_GAP_START: /* :376-421 */
            d->run += 1;
            if (above) {
                d->run += ook->pulse[ook->n];
                d->st = ST_PULSE;
            } else if (d->run >= 10) {
                d->st = ST_GAP;
                if (fsk->n > 16) {
                    close_fsk(d, ook, fsk, fpdm, len);
                    return 2;
                }
            }
            if (ook->n == 0) feed_fsk(d, fm[d->pos], fsk, fpdm);
            break;
        case ST_GAP: /* :422-470 */
            d->run += 1;
            if (above) {
                ook->gap[ook->n] = d->run;
                ook->n += 1;
                if (ook->n >= ORC_MAX_PULSES) {
                    close_ook(d, ook, len);
                    return 1;
                }
                d->run = 0;
                d->st = ST_PULSE;
            }
            if (spurious_eop
                    || (d->run > 10 * d->longest && d->run > 10 * per_ms)
                    || d->run > 100 * per_ms) {
                ook->gap[ook->n] = d->run;
                ook->n += 1;
                close_ook(d, ook, len);
                return 1;
            }
            break;
        }
    }
    d->pos = 0;
    return 0;
}

/* -------------------------------------------------------------------------- bit buffer ---- */

static void bb_clear(orc_bitbuffer *b) { memset(b, 0, sizeof(*b)); } /* src/bitbuffer.c:17 */

static void bb_first_row(orc_bitbuffer *b)
{
    if (b->num_rows == 0) b->free_row = b->num_rows = 1;
}

/* src/bitbuffer.c:22-56: MSB first; every 1024 bits the row silently spills into the next
   physical row and takes it (free_row++) */
static void bb_bit(orc_bitbuffer *b, int bit)
{
    bb_first_row(b);
    uint16_t *len = &b->bits_per_row[b->num_rows - 1];
    if (*len == UINT16_MAX) return;
    unsigned byte = *len / 8, shift = 7 - *len % 8;
    if (*len > 0 && *len % (ORC_BB_COLS * 8) == 0) {
        if (b->free_row < ORC_BB_ROWS)
            b->free_row++;
        else
            return;
    }
    uint8_t *row = &b->bb[0][0] + (size_t)(b->num_rows - 1) * ORC_BB_COLS; /* spilled bytes run on into the following rows */
    row[byte] |= (uint8_t)(bit << shift);
    (*len)++;
}

/* src/bitbuffer.c:106-122 */
static void bb_row(orc_bitbuffer *b)
{
    bb_first_row(b);
    if (b->free_row < ORC_BB_ROWS) {
        b->free_row++;
        b->num_rows = b->free_row;
    } else {
        b->bits_per_row[b->num_rows - 1] = 0;
    }
}

/* src/bitbuffer.c:124-133 */
static void bb_sync(orc_bitbuffer *b)
{
    bb_first_row(b);
    if (b->bits_per_row[b->num_rows - 1]) bb_row(b);
    b->syncs_before_row[b->num_rows - 1]++;
}

/* -------------------------------------------------------------------------------- oracle -- */

struct orc {
    /* configuration */
    int use_mag_est;
    float level_limit, min_level, min_snr, fm_low_pass;
    orc_device *devs;
    int n_devs;
    int store_bitbuffers, store_stages;
    /* per-stream state */
    detector_t det;
    train_t ook, fsk;
    /* results */
    orc_package *pkgs; size_t n_pkgs, cap_pkgs;
    orc_event *evts; size_t n_evts, cap_evts;
    orc_bitbuffer *bbs; size_t n_bbs, cap_bbs;
    int32_t *ppool, *gpool; size_t n_pool, cap_ppool, cap_gpool;
    int16_t *am, *fm; size_t n_stage, cap_am, cap_fm;
    /* context of the package being sliced */
    uint32_t cur_dev;
};

static void *grow(void *p, size_t *cap, size_t need, size_t elem)
{
    if (need <= *cap) return p;
    size_t c = *cap ? *cap : 64;
    while (c < need) c *= 2;
    p = realloc(p, c * elem);
    if (!p) abort();
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

/* what account_event() hands to decode_fn, then clears: src/pulse_slicer.c:26-66 */
static int emit(orc_t *o, orc_bitbuffer *b)
{
    o->evts = grow(o->evts, &o->cap_evts, o->n_evts + 1, sizeof(*o->evts));
    orc_event *e = &o->evts[o->n_evts++];
    e->package = o->n_pkgs ? (uint32_t)(o->n_pkgs - 1) : UINT32_MAX;
    e->dev = o->cur_dev;
    e->ret = 0;
    e->hash = fnv1a(b, sizeof(*b));
    e->bb_idx = UINT32_MAX;
    if (o->store_bitbuffers) {
        o->bbs = grow(o->bbs, &o->cap_bbs, o->n_bbs + 1, sizeof(*o->bbs));
        e->bb_idx = (uint32_t)o->n_bbs;
        o->bbs[o->n_bbs++] = *b;
    }
    if (o->n_pkgs) o->pkgs[o->n_pkgs - 1].num_events++;
    bb_clear(b);
    return 0; /* no decoder behind the oracle */
}

/* ------------------------------------------------------------------------------ slicers ---- */

typedef struct {
    int s_short, s_long, s_reset, s_gap, s_sync, s_tol;
    float per_us;
    int ok;
} timing_t;

/* the preamble every slicer shares, e.g. src/pulse_slicer.c:341-359: float product, int truncation */
static timing_t timing_of(orc_device const *d, uint32_t rate)
{
    timing_t t;
    t.per_us = rate / 1.0e6f;
    t.s_short = d->short_width * t.per_us;
    t.s_long = d->long_width * t.per_us;
    t.s_reset = d->reset_limit * t.per_us;
    t.s_gap = d->gap_limit * t.per_us;
    t.s_sync = d->sync_width * t.per_us;
    t.s_tol = d->tolerance * t.per_us;
    t.ok = !((d->short_width > 0 && t.s_short <= 0) || (d->long_width > 0 && t.s_long <= 0)
            || (d->reset_limit > 0 && t.s_reset <= 0) || (d->gap_limit > 0 && t.s_gap <= 0)
            || (d->sync_width > 0 && t.s_sync <= 0) || (d->tolerance > 0 && t.s_tol <= 0));
    return t;
}

static int in_tol(int v, int nominal, int tol) { return v >= nominal - tol && v <= nominal + tol; }

/* src/pulse_slicer.c:68-259 */
static int slice_pcm(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    float f_short = d->short_width > 0.0f ? 1.0f / (d->short_width * t.per_us) : 0;
    float f_long = d->long_width > 0.0f ? 1.0f / (d->long_width * t.per_us) : 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    int const gap_limit = t.s_gap ? t.s_gap : t.s_reset;
    if (t.s_long == 0) return 0; /* the reference divides by zero here (:97) */
    int const max_zeros = gap_limit / t.s_long;
    int tol = t.s_tol;
    if (tol <= 0) tol = t.s_long / 4;
    int const rz = t.s_short != t.s_long;
    int need = rz ? 4 : 12;
    int preamble = 0;
    unsigned const N = p->n;

    if (rz) {
        /* :105-132 longest run of in-tolerance RZ bits re-tunes the bit period */
        for (unsigned n = 0; n < N; ++n) {
            int sw = 0, lw = 0, cnt = 0;
            while (n < N && in_tol(p->pulse[n], t.s_short, tol) && in_tol(p->pulse[n] + p->gap[n], t.s_long, tol)) {
                sw += p->pulse[n];
                lw += p->pulse[n] + p->gap[n];
                cnt++;
                n++;
            }
            if (cnt >= need) {
                f_long = (float)cnt / lw;
                f_short = (float)cnt / sw;
                need = cnt;
                preamble = cnt;
            }
        }
        /* :134-157 otherwise any in-tolerance bits */
        int sw = 0, lw = 0, cnt = 0;
        for (unsigned n = 0; preamble == 0 && n < N; ++n) {
            if (in_tol(p->pulse[n], t.s_short, tol) && in_tol(p->pulse[n] + p->gap[n], t.s_long, tol)) {
                sw += p->pulse[n];
                lw += p->pulse[n] + p->gap[n];
                cnt++;
            }
        }
        if (cnt > 8) {
            f_long = (float)cnt / lw;
            f_short = (float)cnt / sw;
        }
    } else {
        /* :159-180 NRZ 1010.. preamble; float product, DOUBLE add of 0.5, truncation */
        for (unsigned n = 0; n < N; ++n) {
            int w = 0, cnt = 0;
            while (n < N && (int)(p->pulse[n] * f_short + 0.5) == 1 && (int)(p->gap[n] * f_long + 0.5) == 1) {
                w += p->pulse[n] + p->gap[n];
                cnt += 2;
                n++;
            }
            if (cnt >= need) {
                f_short = f_long = (float)cnt / w;
                need = cnt;
                preamble = cnt;
            }
        }
        /* :182-214 otherwise single and double width symbols anywhere */
        int w = 0, cnt = 0;
        for (unsigned n = 0; preamble == 0 && n < N; ++n) {
            if (in_tol(p->pulse[n], t.s_short, tol)) { w += p->pulse[n]; cnt += 1; }
            if (in_tol(p->pulse[n], 2 * t.s_short, tol)) { w += p->pulse[n]; cnt += 2; }
            if (in_tol(p->gap[n], t.s_long, tol)) { w += p->gap[n]; cnt += 1; }
            if (in_tol(p->gap[n], 2 * t.s_long, tol)) { w += p->gap[n]; cnt += 2; }
        }
        if (cnt > 20) f_short = f_long = (float)cnt / w;
    }

    for (unsigned n = 0; n < N; ++n) { /* :216-257 */
        int highs = p->pulse[n] * f_short + 0.5f;
        int lows = (p->gap[n] + t.s_short - t.s_long) * f_long + 0.5f;
        for (int i = 0; i < highs; ++i) bb_bit(&bits, 1);
        if (lows > max_zeros) lows = max_zeros;
        for (int i = 0; i < lows; ++i) bb_bit(&bits, 0);
        if (rz && abs(p->pulse[n] - t.s_short) > tol)
            bb_clear(&bits);
        else if (p->gap[n] > gap_limit && p->gap[n] <= t.s_reset)
            bb_row(&bits);
        if ((n == N - 1 || p->gap[n] > t.s_reset) && (bits.bits_per_row[0] > 0 || bits.num_rows > 1))
            events += emit(o, &bits);
    }
    return events;
}

/* src/pulse_slicer.c:261-337: the GAP carries the bit */
static int slice_ppm(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    int z_lo, z_hi, o_lo, o_hi, s_lo = 0, s_hi = 0;
    if (t.s_tol > 0) {
        z_lo = t.s_short - t.s_tol; z_hi = t.s_short + t.s_tol;
        o_lo = t.s_long - t.s_tol;  o_hi = t.s_long + t.s_tol;
        if (t.s_sync > 0) { s_lo = t.s_sync - t.s_tol; s_hi = t.s_sync + t.s_tol; }
    } else {
        z_lo = 0;
        z_hi = (t.s_short + t.s_long) / 2 + 1;
        o_lo = z_hi - 1;
        o_hi = t.s_gap ? t.s_gap : t.s_reset;
    }
    for (unsigned n = 0; n < p->n; ++n) {
        int g = p->gap[n];
        if (g > z_lo && g < z_hi) bb_bit(&bits, 0);
        else if (g > o_lo && g < o_hi) bb_bit(&bits, 1);
        else if (g > s_lo && g < s_hi) bb_sync(&bits);
        else if (g < t.s_reset) bb_row(&bits);
        if ((n == p->n - 1 || g >= t.s_reset) && (bits.bits_per_row[0] > 0 || bits.num_rows > 1))
            events += emit(o, &bits);
    }
    return events;
}

/* src/pulse_slicer.c:339-449: the PULSE carries the bit */
static int slice_pwm(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    int o_lo, o_hi, z_lo, z_hi, s_lo = 0, s_hi = 0;
    if (t.s_tol > 0) {
        o_lo = t.s_short - t.s_tol; o_hi = t.s_short + t.s_tol;
        z_lo = t.s_long - t.s_tol;  z_hi = t.s_long + t.s_tol;
        if (t.s_sync > 0) { s_lo = t.s_sync - t.s_tol; s_hi = t.s_sync + t.s_tol; }
    } else if (t.s_sync <= 0) {
        o_lo = 0; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = INT_MAX;
    } else if (t.s_sync < t.s_short) {
        s_lo = 0; s_hi = (t.s_sync + t.s_short) / 2 + 1;
        o_lo = s_hi - 1; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = INT_MAX;
    } else if (t.s_sync < t.s_long) {
        o_lo = 0; o_hi = (t.s_short + t.s_sync) / 2 + 1;
        s_lo = o_hi - 1; s_hi = (t.s_sync + t.s_long) / 2 + 1;
        z_lo = s_hi - 1; z_hi = INT_MAX;
    } else {
        o_lo = 0; o_hi = (t.s_short + t.s_long) / 2 + 1;
        z_lo = o_hi - 1; z_hi = (t.s_long + t.s_sync) / 2 + 1;
        s_lo = z_hi - 1; s_hi = INT_MAX;
    }
    for (unsigned n = 0; n < p->n; ++n) {
        int w = p->pulse[n];
        if (w > o_lo && w < o_hi) bb_bit(&bits, 1);
        else if (w > z_lo && w < z_hi) bb_bit(&bits, 0);
        else if (w > s_lo && w < s_hi) bb_sync(&bits);
        else if (w <= o_lo) { /* too short: ignored */ }
        else bb_row(&bits);
        if ((n == p->n - 1 || p->gap[n] > t.s_reset) && bits.num_rows > 0)
            events += emit(o, &bits);
        else if (t.s_gap > 0 && p->gap[n] > t.s_gap && bits.num_rows > 0 && bits.bits_per_row[bits.num_rows - 1] > 0)
            bb_row(&bits);
    }
    return events;
}

/* src/pulse_slicer.c:451-527 */
static int slice_manchester(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0, since = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    bb_bit(&bits, 0);
    double const edge = t.s_short * 1.5;
    for (unsigned n = 0; n < p->n; ++n) {
        int w = p->pulse[n], g = p->gap[n];
        int lo = t.s_short - t.s_tol, hi = t.s_short * 2 + t.s_tol;
        if (t.s_tol > 0 && (w < lo || w > hi || g < lo || g > hi)) {
            if (w > edge && w <= hi) bb_bit(&bits, 1);
            bb_row(&bits);
            bb_bit(&bits, 0);
            since = 0;
        } else if (w + since > edge) {
            bb_bit(&bits, 1);
            since = 0;
        } else {
            since += w;
        }
        if ((n == p->n - 1 || g > t.s_reset) && bits.num_rows > 0) {
            events += emit(o, &bits);
            bb_bit(&bits, 0);
            since = 0;
        } else if (g + since > edge) {
            bb_bit(&bits, 0);
            since = 0;
        } else {
            since += g;
        }
    }
    return events;
}

/* src/pulse_slicer.c:529-535: pulses and gaps as one alternating symbol stream */
static int symbol_at(train_t const *p, unsigned k) { return k % 2 == 0 ? p->pulse[k / 2] : p->gap[k / 2]; }

/* src/pulse_slicer.c:537-595 */
static int slice_dmc(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    unsigned const total = p->n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        if (abs(s - t.s_short) < t.s_tol) {
            bb_bit(&bits, 1);
            s = k + 1 < total ? symbol_at(p, ++k) : 0;
            if (abs(s - t.s_short) > t.s_tol) {
                if (s >= t.s_reset - t.s_tol)
                    k--;
                else if (bits.num_rows > 0 && bits.bits_per_row[bits.num_rows - 1] > 0)
                    bb_row(&bits);
            }
        } else if (abs(s - t.s_long) < t.s_tol) {
            bb_bit(&bits, 0);
        } else if (s >= t.s_reset - t.s_tol && bits.num_rows > 0) {
            events += emit(o, &bits);
        }
    }
    return events;
}

/* src/pulse_slicer.c:597-657 */
static int slice_piwm_raw(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    float f_short = d->short_width > 0.0f ? 1.0f / (d->short_width * t.per_us) : 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    unsigned const total = p->n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        int w = s * f_short + 0.5; /* float product, double add */
        if (s > t.s_long) {
            bb_row(&bits);
        } else if (abs(s - w * t.s_short) < t.s_tol) {
            for (; w > 0; --w) bb_bit(&bits, 1 - k % 2);
        } else if (s < t.s_reset && bits.num_rows > 0 && bits.bits_per_row[bits.num_rows - 1] > 0) {
            bb_row(&bits);
        }
        if ((k == total - 1 || s > t.s_reset) && bits.num_rows > 0) events += emit(o, &bits);
    }
    return events;
}

/* src/pulse_slicer.c:659-713 */
static int slice_piwm_dc(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    unsigned const total = p->n * 2;
    for (unsigned k = 0; k < total; ++k) {
        int s = symbol_at(p, k);
        if (abs(s - t.s_short) < t.s_tol) bb_bit(&bits, 1);
        else if (abs(s - t.s_long) < t.s_tol) bb_bit(&bits, 0);
        else if (s < t.s_reset && bits.num_rows > 0 && bits.bits_per_row[bits.num_rows - 1] > 0) bb_row(&bits);
        if ((k == total - 1 || s > t.s_reset) && bits.num_rows > 0) events += emit(o, &bits);
    }
    return events;
}

/* src/pulse_slicer.c:715-759 */
static int slice_nrzs(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    int const limit = t.s_short;
    for (unsigned n = 0; n < p->n; ++n) {
        if (p->pulse[n] > limit) {
            for (int i = 0; i < p->pulse[n] / limit; ++i) bb_bit(&bits, 1);
            bb_bit(&bits, 0);
        } else if (p->pulse[n] < limit) {
            bb_bit(&bits, 0);
        }
        if (n == p->n - 1 || p->gap[n] >= t.s_reset) events += emit(o, &bits);
    }
    return events;
}

/* src/pulse_slicer.c:775-864 (Oregon Scientific v1) */
static int slice_osv1(orc_t *o, train_t const *p, orc_device const *d)
{
    timing_t t = timing_of(d, p->sample_rate);
    if (!t.ok) return 0;
    int events = 0, pre = 0, man = 0;
    orc_bitbuffer bits;
    bb_clear(&bits);
    int const half_lo = t.s_short / 2, half_hi = t.s_short * 3 / 2, sync_lo = 2 * half_hi;
    unsigned n;
    for (n = 0; n < p->n; ++n) {
        if (p->pulse[n] > half_lo && p->gap[n] > half_lo) {
            pre++;
            if (p->gap[n] > half_hi) break;
        } else {
            return events;
        }
    }
    if (pre != 12) return events;
    ++n;
    if (n >= ORC_MAX_PULSES) return events; /* the reference reads past the array here */
    if (p->pulse[n] < sync_lo || p->gap[n] < sync_lo) return events;
    if (p->gap[n] > p->pulse[n]) {
        man ^= 1;
        if (man) bb_bit(&bits, 0);
    }
    for (n++; n < p->n; ++n) {
        man ^= 1;
        if (man) bb_bit(&bits, 1);
        if (p->pulse[n] > half_hi) {
            man ^= 1;
            if (man) bb_bit(&bits, 1);
        }
        if ((n == p->n - 1 || p->gap[n] > t.s_reset) && bits.num_rows > 0) {
            events += emit(o, &bits);
            return events;
        }
        man ^= 1;
        if (man) bb_bit(&bits, 0);
        if (p->gap[n] > half_hi) {
            man ^= 1;
            if (man) bb_bit(&bits, 0);
        }
    }
    return events;
}

/* src/pulse_slicer.c:866-918; only short/long/reset take part in the rate check */
static int slice_rzi(orc_t *o, train_t const *p, orc_device const *d)
{
    float per_us = p->sample_rate / 1.0e6f;
    int s_short = d->short_width * per_us;
    int s_long = d->long_width * per_us;
    int s_reset = d->reset_limit * per_us;
    int s_base = s_long - s_short;
    if ((d->short_width > 0 && s_short <= 0) || (d->long_width > 0 && s_long <= 0) || (d->reset_limit > 0 && s_reset <= 0))
        return 0;
    if (s_long == 0) return 0; /* the reference divides by zero */
    int events = 0, fresh = 1;
    orc_bitbuffer bits;
    bb_clear(&bits);
    for (unsigned n = 0; n < p->n; ++n) {
        int high = p->pulse[n];
        int ones = fresh ? (high + s_long / 2) / s_long : (high - s_base + s_long / 2) / s_long;
        fresh = 0;
        for (int k = 0; k < ones; ++k) bb_bit(&bits, 1);
        if (p->gap[n] > s_reset || n == p->n - 1) {
            if (bits.bits_per_row[0] > 0) events += emit(o, &bits);
            bb_clear(&bits);
            fresh = 1;
            continue;
        }
        bb_bit(&bits, 0);
    }
    return events;
}

enum { /* include/r_device.h:24-40 */
    M_OOK_MC = 3, M_OOK_PCM = 4, M_OOK_PPM = 5, M_OOK_PWM = 6, M_OOK_PIWM_RAW = 8, M_OOK_DMC = 9,
    M_OOK_OSV1 = 10, M_OOK_PIWM_DC = 11, M_OOK_NRZS = 12, M_OOK_RZI = 13,
    M_FSK_PCM = 16, M_FSK_PWM = 17, M_FSK_MC = 18
};

static int slice_one(orc_t *o, train_t const *p, int fsk, orc_device const *d)
{
    if (!fsk) { /* src/r_api.c:456-495 */
        switch (d->modulation) {
        case M_OOK_PCM: return slice_pcm(o, p, d);
        case M_OOK_PPM: return slice_ppm(o, p, d);
        case M_OOK_PWM: return slice_pwm(o, p, d);
        case M_OOK_MC: return slice_manchester(o, p, d);
        case M_OOK_PIWM_RAW: return slice_piwm_raw(o, p, d);
        case M_OOK_PIWM_DC: return slice_piwm_dc(o, p, d);
        case M_OOK_DMC: return slice_dmc(o, p, d);
        case M_OOK_OSV1: return slice_osv1(o, p, d);
        case M_OOK_NRZS: return slice_nrzs(o, p, d);
        case M_OOK_RZI: return slice_rzi(o, p, d);
        default: return 0;
        }
    } else { /* src/r_api.c:520-544 */
        switch (d->modulation) {
        case M_FSK_PCM: return slice_pcm(o, p, d);
        case M_FSK_PWM: return slice_pwm(o, p, d);
        case M_FSK_MC: return slice_manchester(o, p, d);
        default: return 0;
        }
    }
}

/* run_ook_demods / run_fsk_demods, src/r_api.c:438-550: ascending priority classes, registration
   order inside a class, later classes only while no event decoded (never, without decoders) */
static int slice_all(orc_t *o, train_t const *p, int fsk)
{
    int decoded = 0;
    unsigned next = 0;
    for (unsigned prio = 0; !decoded && prio < UINT_MAX; prio = next) {
        next = UINT_MAX;
        for (int i = 0; i < o->n_devs; ++i) {
            orc_device const *d = &o->devs[i];
            if (d->priority > prio && d->priority < next) next = d->priority;
            if (d->priority != prio) continue;
            o->cur_dev = (uint32_t)i;
            decoded += slice_one(o, p, fsk, d);
        }
    }
    return decoded;
}

/* calc_rssi_snr, src/r_flow.c:35-64 */
static void fill_levels(orc_package *k, train_t const *t, uint32_t rate, uint32_t center, int sample_size, int use_mag_est, int max_high)
{
    float hi = t->high_est > 0 ? t->high_est : 1;
    float lo = t->low_est > 0 ? t->low_est : 1;
    float top = hi < max_high ? hi : max_high;
    float asnr = top / lo;
    float off1 = (float)t->f1_est / INT16_MAX * rate / 2.0f;
    float off2 = (float)t->f2_est / INT16_MAX * rate / 2.0f;
    k->freq1_hz = off1 + center;
    k->freq2_hz = off2 + center;
    k->centerfreq_hz = center;
    k->depth_bits = (uint32_t)sample_size * 4;
    if (sample_size == 2 && !use_mag_est) {
        k->range_db = 42.1442f;
        k->rssi_db = 10.0f * log10f(hi) - 42.1442f;
        k->noise_db = 10.0f * log10f(lo) - 42.1442f;
        k->snr_db = 10.0f * log10f(asnr);
    } else {
        k->range_db = 84.2884f;
        k->rssi_db = 20.0f * log10f(hi) - 84.2884f;
        k->noise_db = 20.0f * log10f(lo) - 84.2884f;
        k->snr_db = 20.0f * log10f(asnr);
    }
}

static void record_package(orc_t *o, int type, int block, float file_pos, train_t const *t, uint32_t rate,
        uint32_t center, int sample_size)
{
    o->pkgs = grow(o->pkgs, &o->cap_pkgs, o->n_pkgs + 1, sizeof(*o->pkgs));
    orc_package *k = &o->pkgs[o->n_pkgs++];
    memset(k, 0, sizeof(*k));
    k->type = type;
    k->block = block;
    k->offset = t->offset;
    k->sample_rate = t->sample_rate;
    k->start_ago = t->start_ago;
    k->end_ago = t->end_ago;
    k->num_pulses = t->n;
    k->ook_low_estimate = t->low_est;
    k->ook_high_estimate = t->high_est;
    k->fsk_f1_est = t->f1_est;
    k->fsk_f2_est = t->f2_est;
    k->sample_file_pos = file_pos;
    fill_levels(k, t, rate, center, sample_size, o->use_mag_est, o->det.max_high);
    uint32_t cnt = t->n + 1;
    if (cnt > ORC_MAX_PULSES) cnt = ORC_MAX_PULSES;
    o->ppool = grow(o->ppool, &o->cap_ppool, o->n_pool + cnt, sizeof(int32_t));
    o->gpool = grow(o->gpool, &o->cap_gpool, o->n_pool + cnt, sizeof(int32_t));
    memcpy(o->ppool + o->n_pool, t->pulse, cnt * sizeof(int32_t));
    memcpy(o->gpool + o->n_pool, t->gap, cnt * sizeof(int32_t));
    k->pulse_off = (uint32_t)o->n_pool;
    k->pulse_count = cnt;
    o->n_pool += cnt;
    k->first_event = (uint32_t)o->n_evts;
}

/* ------------------------------------------------------------------------------- driver ---- */

orc_t *orc_create(void)
{
    orc_t *o = calloc(1, sizeof(*o));
    if (!o) return NULL;
    o->level_limit = 0.0f;      /* src/r_api.c:153-155 */
    o->min_level = -12.1442f;
    o->min_snr = 9.0f;
    o->store_bitbuffers = 1;
    return o;
}

// src/rtl_433.c:1811-1825: a cf32 capture is turned into cs16 while it is read, block by block:
// "clamp float to [-1,1] and scale to Q0.15".  The float -> int conversion of an out-of-range
// or NaN product is undefined in C; on the reference's x86-64 builds (cvttss2si) it yields
// INT_MIN, which the clamp turns into -INT16_MAX.  Restated with that behaviour made explicit.
void orc_cf32_to_cs16(float const *in, int16_t *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        float v = in[i] * INT16_MAX;
        int s_tmp = (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : INT32_MIN;
        if (s_tmp < -INT16_MAX)
            s_tmp = -INT16_MAX;
        else if (s_tmp > INT16_MAX)
            s_tmp = INT16_MAX;
        out[i] = (int16_t)s_tmp;
    }
}

void orc_destroy(orc_t *o)
{
    if (!o) return;
    free(o->devs); free(o->pkgs); free(o->evts); free(o->bbs); free(o->ppool); free(o->gpool);
    free(o->am); free(o->fm);
    free(o);
}

void orc_set_capture(orc_t *o, int store_bitbuffers, int store_stages)
{
    o->store_bitbuffers = store_bitbuffers;
    o->store_stages = store_stages;
}

void orc_set_levels(orc_t *o, int use_mag_est, float level_limit, float min_level, float min_snr)
{
    o->use_mag_est = use_mag_est;
    o->level_limit = level_limit;
    o->min_level = min_level;
    o->min_snr = min_snr;
}

void orc_set_fm_low_pass(orc_t *o, float v) { o->fm_low_pass = v; }

int orc_add_device(orc_t *o, orc_device const *d)
{
    o->devs = realloc(o->devs, (size_t)(o->n_devs + 1) * sizeof(*o->devs));
    o->devs[o->n_devs] = *d;
    return o->n_devs++;
}

int orc_num_devices(orc_t *o) { return o->n_devs; }

static void clear_results(orc_t *o) { o->n_pkgs = o->n_evts = o->n_bbs = o->n_pool = o->n_stage = 0; }

/* rtl_433 -r FILE: src/rtl_433.c:1797-1854 (block loop, flush, reset) around
   push_sdr_flow(), src/r_flow.c:104-263 */
int orc_run_stream(orc_t *o, void const *iq, size_t bytes, int sample_size, uint32_t rate,
        uint32_t center_freq, int fpdm_mode, uint32_t block_bytes)
{
    clear_results(o);
    if (!block_bytes) block_bytes = DEFAULT_BLOCK_BYTES;
    size_t const max_n = block_bytes / (size_t)sample_size;
    uint16_t *env = malloc(max_n * sizeof(uint16_t));
    int16_t *am = malloc(max_n * sizeof(int16_t));
    int16_t *fm = malloc(max_n * sizeof(int16_t));

    /* src/rtl_433.c:1094-1102 */
    unsigned fpdm = (unsigned)fpdm_mode;
    if (fpdm_mode == 2) fpdm = center_freq > 800000000u ? 1 : 0;
    /* src/rtl_433.c:1515-1522 */
    int enable_fm = 0;
    for (int i = 0; i < o->n_devs; ++i)
        if (o->devs[i].modulation >= 16) enable_fm = 1;

    int32_t lv[4];
    orc_detector_levels(o->use_mag_est, o->level_limit, o->min_level, o->min_snr, lv);
    detector_t *d = &o->det;
    d->fixed_high = lv[0];
    d->min_high = lv[1];
    d->ratio = lv[2];
    d->max_high = lv[3];
    detector_reset(d); /* src/r_flow.c:79-97 at the end of the previous file */
    train_clear(&o->ook);
    train_clear(&o->fsk);
    int16_t lpf_state[2] = {0, 0};
    fm_state fms = {0, 0, 0, 0};
    int32_t coef[2] = {0, 0};
    if (enable_fm) { /* src/r_flow.c:204 */
        float lp = o->fm_low_pass != 0.0f ? o->fm_low_pass : fpdm ? 0.2f : 0.1f;
        orc_fm_coeffs(sample_size == 4, rate, lp, coef);
    }

    uint64_t input_pos = 0;
    int n_blocks = 0;
    size_t pos = 0;
    float file_pos = 0.0f; /* demod->sample_file_pos, src/rtl_433.c:1752 */
    for (;;) {
        size_t n_read = bytes - pos < block_bytes ? bytes - pos : block_bytes;
        int const n = (int)(n_read / (size_t)sample_size); /* 0 => flush_sdr_flow(), len == 0 */
        uint8_t const *blk = (uint8_t const *)iq + pos;
        pos += n_read;
        int const block_idx = n_blocks;
        if (n_read) {
            /* src/rtl_433.c:1839, float arithmetic */
            file_pos = ((float)n_blocks * block_bytes + n_read) / rate / sample_size;
            n_blocks++;
            /* src/r_flow.c:151-162 */
            if (sample_size == 2) {
                if (o->use_mag_est)
                    orc_magnitude_cu8(blk, env, (uint32_t)n);
                else
                    orc_envelope_cu8(blk, env, (uint32_t)n);
            } else {
                orc_magnitude_cs16((int16_t const *)blk, env, (uint32_t)n);
            }
            orc_low_pass(lpf_state, env, am, (uint32_t)n); /* :198 */
            if (enable_fm) {                                /* :202-210 */
                if (sample_size == 2)
                    fm_cu8(&fms, coef, blk, fm, (size_t)n);
                else
                    fm_cs16(&fms, coef, (int16_t const *)blk, fm, (size_t)n);
            } else {
                /* buf.fm aliases buf.temp (include/r_private.h:32-36): with FM off the detector
                   reads the raw envelope, reinterpreted as int16 */
                for (int k = 0; k < n; ++k) fm[k] = (int16_t)env[k];
            }
            if (o->store_stages) {
                o->am = grow(o->am, &o->cap_am, o->n_stage + (size_t)n, sizeof(int16_t));
                o->fm = grow(o->fm, &o->cap_fm, o->n_stage + (size_t)n, sizeof(int16_t));
                memcpy(o->am + o->n_stage, am, (size_t)n * sizeof(int16_t));
                memcpy(o->fm + o->n_stage, fm, (size_t)n * sizeof(int16_t));
                o->n_stage += (size_t)n;
            }
        }
        /* src/r_flow.c:241-334 */
        int type;
        while ((type = detect(d, am, fm, n, rate, input_pos, &o->ook, &o->fsk, fpdm)) != 0) {
            train_t const *t = type == 1 ? &o->ook : &o->fsk;
            record_package(o, type, block_idx, file_pos, t, rate, center_freq, sample_size);
            slice_all(o, t, type == 2);
        }
        input_pos += (uint64_t)n; /* :491 */
        if (!n_read) break;
    }
    free(env);
    free(am);
    free(fm);
    return (int)o->n_pkgs;
}

size_t orc_num_packages(orc_t *o) { return o->n_pkgs; }
size_t orc_num_events(orc_t *o) { return o->n_evts; }
size_t orc_num_bitbuffers(orc_t *o) { return o->n_bbs; }
size_t orc_num_stage(orc_t *o) { return o->n_stage; }
orc_package const *orc_packages(orc_t *o) { return o->pkgs; }
orc_event const *orc_events(orc_t *o) { return o->evts; }
orc_bitbuffer const *orc_bitbuffers(orc_t *o) { return o->bbs; }
int32_t const *orc_pulse_pool(orc_t *o) { return o->ppool; }
int32_t const *orc_gap_pool(orc_t *o) { return o->gpool; }
int16_t const *orc_am(orc_t *o) { return o->am; }
int16_t const *orc_fm(orc_t *o) { return o->fm; }

int orc_slice(orc_t *o, int dev_idx, uint32_t sample_rate, uint32_t num_pulses, int32_t const *pulse, int32_t const *gap)
{
    static train_t t;
    train_clear(&t);
    t.sample_rate = sample_rate;
    t.n = num_pulses;
    memcpy(t.pulse, pulse, num_pulses * sizeof(int32_t));
    memcpy(t.gap, gap, num_pulses * sizeof(int32_t));
    clear_results(o);
    o->pkgs = grow(o->pkgs, &o->cap_pkgs, 1, sizeof(*o->pkgs));
    memset(&o->pkgs[0], 0, sizeof(o->pkgs[0]));
    o->n_pkgs = 1;
    o->cur_dev = (uint32_t)dev_idx;
    slice_one(o, &t, o->devs[dev_idx].modulation >= 16, &o->devs[dev_idx]);
    return (int)o->n_evts;
}
