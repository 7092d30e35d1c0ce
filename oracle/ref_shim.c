/* ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin flat-array wrapper around the UNMODIFIED reference functions so that
 * tests and bench.py's cpu_baseline leg can call them through ctypes:
 *
 *   signal_in_picoamps()   /root/reference/src/misc.c:15-32   (sigtk.h:124)
 *   getevents()            /root/reference/src/events.c:553-573 (sigtk.h:134)
 *   meanf/stdvf/medianf/meani16/stdvi16/mediani16  src/stat.h:17-73
 *
 * This file is ours; it is compiled together with the reference sources where
 * they lie under /root/reference (see oracle/Makefile, target _ref) into
 * oracle/_ref/libsigtk_ref.so.  No reference source is copied into the repo.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sigtk.h" /* reference header, -I/root/reference/src */
#include "stat.h"  /* reference inline stat functions */

/* pA conversion of one read through the reference's own function. */
int ref_pa(const int16_t *raw, uint64_t n, double digitisation, double offset,
           double range, float *out) {
    slow5_rec_t rec;
    memset(&rec, 0, sizeof rec);
    rec.raw_signal = (int16_t *)raw;
    rec.len_raw_signal = n;
    rec.digitisation = digitisation;
    rec.offset = offset;
    rec.range = range;
    float *pa = signal_in_picoamps(&rec);
    if (!pa) return -1;
    memcpy(out, pa, n * sizeof(float));
    free(pa);
    return 0;
}

/* Event detection of one read. Returns the number of events, or -(needed) if
 * cap is too small. Aborts (like the reference) on degenerate inputs. */
int64_t ref_getevents(uint64_t n, const float *pa, int rna, uint64_t cap,
                      uint64_t *start, float *length, float *mean, float *stdv) {
    event_table et = getevents((size_t)n, (float *)pa, (int8_t)rna);
    int64_t ne = (int64_t)et.n;
    if ((uint64_t)ne > cap) {
        free(et.event);
        return -ne;
    }
    for (int64_t i = 0; i < ne; i++) {
        start[i] = et.event[i].start;
        length[i] = et.event[i].length;
        mean[i] = et.event[i].mean;
        stdv[i] = et.event[i].stdv;
    }
    free(et.event);
    return ne;
}

/* raw int16 -> pA -> events, the per-record work of event_func (cfunc.c:72-83)
 * without the printf. Used for parity and for CPU timing. */
int64_t ref_event_read(const int16_t *raw, uint64_t n, double digitisation,
                       double offset, double range, int rna, uint64_t cap,
                       uint64_t *start, float *length, float *mean, float *stdv) {
    slow5_rec_t rec;
    memset(&rec, 0, sizeof rec);
    rec.raw_signal = (int16_t *)raw;
    rec.len_raw_signal = n;
    rec.digitisation = digitisation;
    rec.offset = offset;
    rec.range = range;
    float *pa = signal_in_picoamps(&rec);
    int64_t ne = ref_getevents(n, pa, rna, cap, start, length, mean, stdv);
    free(pa);
    return ne;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Time signal_in_picoamps + getevents over a batch of reads laid out as a flat
 * int16 array with read_off[n_reads+1] (in samples). Single thread. Returns
 * seconds; *total_events receives the event count (so the work is not dead). */
double ref_time_events(const int16_t *samples, const uint64_t *read_off,
                       uint64_t n_reads, const double *digitisation,
                       const double *offset, const double *range, int rna,
                       uint64_t *total_events) {
    uint64_t nev = 0;
    double t0 = now_s();
    for (uint64_t r = 0; r < n_reads; r++) {
        slow5_rec_t rec;
        memset(&rec, 0, sizeof rec);
        rec.raw_signal = (int16_t *)(samples + read_off[r]);
        rec.len_raw_signal = read_off[r + 1] - read_off[r];
        rec.digitisation = digitisation[r];
        rec.offset = offset[r];
        rec.range = range[r];
        float *pa = signal_in_picoamps(&rec);
        event_table et = getevents((size_t)rec.len_raw_signal, pa, (int8_t)rna);
        nev += et.n;
        free(pa);
        free(et.event);
    }
    double t1 = now_s();
    if (total_events) *total_events = nev;
    return t1 - t0;
}

/* The six numbers of stat_func (cfunc.c:126-159), in its print order:
 * raw_mean, pa_mean, raw_std, pa_std, raw_median, pa_median. */
int ref_stat(const int16_t *raw, uint64_t n, double digitisation, double offset,
             double range, float *out6) {
    slow5_rec_t rec;
    memset(&rec, 0, sizeof rec);
    rec.raw_signal = (int16_t *)raw;
    rec.len_raw_signal = n;
    rec.digitisation = digitisation;
    rec.offset = offset;
    rec.range = range;
    float m1 = meani16(rec.raw_signal, n);
    float s1 = stdvi16(rec.raw_signal, n);
    int16_t k1 = mediani16(rec.raw_signal, n);
    float *pa = signal_in_picoamps(&rec);
    float m2 = meanf(pa, n);
    float s2 = stdvf(pa, n);
    float k2 = medianf(pa, n);
    free(pa);
    out6[0] = m1; out6[1] = m2; out6[2] = s1; out6[3] = s2;
    out6[4] = (float)k1; out6[5] = k2;
    return 0;
}

/* svb-zd through the reference's own slow5lib (slow5_press.c:1055-1150). encode: returns the stream length or -1;
 * decode: returns the number of samples or -1. */
#include <slow5/slow5_press.h>
int64_t ref_svbzd_encode(const int16_t *raw, uint64_t n, uint8_t *out, uint64_t cap) {
    size_t bytes = 0;
    void *p = slow5_ptr_compress_solo(SLOW5_COMPRESS_SVB_ZD, raw, (size_t)n * sizeof *raw, &bytes);
    if (!p) return -1;
    if (bytes > cap) { free(p); return -1; }
    memcpy(out, p, bytes);
    free(p);
    return (int64_t)bytes;
}
int64_t ref_svbzd_decode(const uint8_t *in, uint64_t n_bytes, int16_t *out, uint64_t cap) {
    size_t bytes = 0;
    void *p = slow5_ptr_depress_solo(SLOW5_COMPRESS_SVB_ZD, in, (size_t)n_bytes, &bytes);
    if (!p) return -1;
    if (bytes / 2 > cap) { free(p); return -1; }
    memcpy(out, p, bytes);
    free(p);
    return (int64_t)(bytes / 2);
}
