/*
 * oracle/ref_fft.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of kopek's FFT frequency-analysis path, used only as the
 * parity checker (tests/, __graft_entry__.smoke()) and as the reported CPU
 * baseline (bench.py cpu_baseline / --impl reference).  Nothing under
 * kopek_b200/ may link, import or call this file.
 *
 * PARITY UNPINNED: the reference (Rust, /root/reference) cannot be compiled in
 * this image (no rustc/cargo, crate `num = "0.4"` is un-vendored and has no
 * lockfile) and its own tests hold no golden vector for src/fft.rs.  This file
 * follows the reference line by line instead; it is cross-checked against an
 * O(N^2) f64 DFT and numpy.fft in tests/test_oracle.py.
 *
 * Third-party arithmetic restated here (crate num-complex 0.4.x, published
 * semantics): Complex*scalar and Complex/scalar act component-wise;
 * Complex::exp() = from_polar(re.exp(), im) = (r*cos(im), r*sin(im));
 * Mul = (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re); norm() = hypot(re, im).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile) so no
 * FMA contraction changes the rounding sequence of the Rust original.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* usize::next_power_of_two (src/fft.rs:32): 0 -> 1, 1 -> 1, 3 -> 4 ... */
ORACLE_API size_t oracle_next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

/* ------------------------------------------------------------------ f64 --- */

/* fft_inner, src/fft.rs:7-28.  buf_a/buf_b are (re,im) interleaved; the Rust
 * slices `&mut buf_b[step..]` become pointer offsets, `split_at_mut(n/2)`
 * becomes the index n/2 + i/2 relative to the (possibly offset) slice start. */
static void fft_inner_f64(double *buf_a, double *buf_b, size_t n, size_t step) {
    if (step >= n) return;                                   /* :14-16 */
    fft_inner_f64(buf_b, buf_a, n, step * 2);                /* :18 */
    fft_inner_f64(buf_b + 2 * step, buf_a + 2 * step, n, step * 2); /* :19 */
    double *left = buf_a, *right = buf_a + 2 * (n / 2);      /* :21 */
    for (size_t i = 0; i < n; i += step * 2) {               /* :23 */
        /* :24  (-I * PI * (i as f64) / (n as f64)).exp()
         *   -I            = (-0.0, -1.0)
         *   * PI          = (-0.0, -PI)          component-wise
         *   * (i as f64)  = (-0.0*i, (-PI)*i)    one rounding on im
         *   / (n as f64)  = (.., ((-PI)*i)/n)    second rounding on im
         *   .exp()        = exp(re) * (cos im, sin im), exp(-0.0) == 1.0 exactly */
        double fi = (double)i, fn = (double)n;
        double z_re = (-0.0 * M_PI) * fi / fn;
        double z_im = (-1.0 * M_PI) * fi / fn;
        double r = exp(z_re);
        double w_re = r * cos(z_im), w_im = r * sin(z_im);
        double b_re = buf_b[2 * (i + step)], b_im = buf_b[2 * (i + step) + 1];
        double t_re = w_re * b_re - w_im * b_im;             /* Complex Mul */
        double t_im = w_re * b_im + w_im * b_re;
        double a_re = buf_b[2 * i], a_im = buf_b[2 * i + 1];
        left[2 * (i / 2)] = a_re + t_re;                     /* :25 */
        left[2 * (i / 2) + 1] = a_im + t_im;
        right[2 * (i / 2)] = a_re - t_re;                    /* :26 */
        right[2 * (i / 2) + 1] = a_im - t_im;
    }
}

/* pub fn fft, src/fft.rs:6-41.  `in` holds n_in complex (2*n_in doubles);
 * `out` must hold oracle_next_pow2(n_in) complex.  Returns that length. */
ORACLE_API size_t oracle_fft_f64(const double *in, size_t n_in, double *out) {
    size_t n = oracle_next_pow2(n_in);                       /* :31-32 */
    double *buf_a = out;                                     /* :34 (result IS buf_a, :40) */
    memcpy(buf_a, in, n_in * 2 * sizeof(double));
    for (size_t i = 2 * n_in; i < 2 * n; ++i) buf_a[i] = 0.0;/* :36 */
    double *buf_b = (double *)malloc(n * 2 * sizeof(double));/* :38 */
    memcpy(buf_b, buf_a, n * 2 * sizeof(double));
    fft_inner_f64(buf_a, buf_b, n, 1);                       /* :39 */
    free(buf_b);
    return n;
}

/* ------------------------------------------------------------------ f32 --- */
/* The same recursion instantiated at f32 ("the f32 reference" of north_star;
 * the reference itself only has the f64 form -- SURVEY.md D1). */
static void fft_inner_f32(float *buf_a, float *buf_b, size_t n, size_t step) {
    if (step >= n) return;
    fft_inner_f32(buf_b, buf_a, n, step * 2);
    fft_inner_f32(buf_b + 2 * step, buf_a + 2 * step, n, step * 2);
    float *left = buf_a, *right = buf_a + 2 * (n / 2);
    for (size_t i = 0; i < n; i += step * 2) {
        float fi = (float)i, fn = (float)n;
        float z_im = (-1.0f * (float)M_PI) * fi / fn;
        float w_re = cosf(z_im), w_im = sinf(z_im);
        float b_re = buf_b[2 * (i + step)], b_im = buf_b[2 * (i + step) + 1];
        float t_re = w_re * b_re - w_im * b_im;
        float t_im = w_re * b_im + w_im * b_re;
        float a_re = buf_b[2 * i], a_im = buf_b[2 * i + 1];
        left[2 * (i / 2)] = a_re + t_re;
        left[2 * (i / 2) + 1] = a_im + t_im;
        right[2 * (i / 2)] = a_re - t_re;
        right[2 * (i / 2) + 1] = a_im - t_im;
    }
}

ORACLE_API size_t oracle_fft_f32(const float *in, size_t n_in, float *out) {
    size_t n = oracle_next_pow2(n_in);
    memcpy(out, in, n_in * 2 * sizeof(float));
    for (size_t i = 2 * n_in; i < 2 * n; ++i) out[i] = 0.0f;
    float *buf_b = (float *)malloc(n * 2 * sizeof(float));
    memcpy(buf_b, out, n * 2 * sizeof(float));
    fft_inner_f32(out, buf_b, n, 1);
    free(buf_b);
    return n;
}

/* Independent check: O(N^2) f64 DFT, X[k] = sum x[j] e^{-2 pi i jk/n}. */
ORACLE_API void oracle_dft_naive_f64(const double *in, size_t n, double *out) {
    for (size_t k = 0; k < n; ++k) {
        long double sr = 0, si = 0;
        for (size_t j = 0; j < n; ++j) {
            size_t m = (j * k) % n;
            long double th = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
            long double c = cosl(th), s = sinl(th);
            sr += in[2 * j] * c - in[2 * j + 1] * s;
            si += in[2 * j] * s + in[2 * j + 1] * c;
        }
        out[2 * k] = (double)sr;
        out[2 * k + 1] = (double)si;
    }
}

/* ----------------------------------------------------------------- show --- */
/* pub fn show, src/fft.rs:43-52: label line, then "{:.4}{:+.4}i" joined by
 * ", " and a newline.  Rust prints NaN as "NaN" (never signed) and infinities
 * as "inf"/"-inf" ("+inf" under {:+}).  Writes at most cap-1 bytes + NUL and
 * returns the full length that was needed (snprintf convention). */
typedef struct { char *dst; size_t cap, need; } sbuf;
static void app(sbuf *b, const char *piece) {
    size_t len = strlen(piece);
    if (b->need < b->cap) {
        size_t room = b->cap - b->need - 1;
        size_t cp = len < room ? len : room;
        memcpy(b->dst + b->need, piece, cp);
        b->dst[b->need + cp] = 0;
    }
    b->need += len;
}
static void app_f(sbuf *b, double v, int plus) {
    char tmp[400];
    if (isnan(v)) snprintf(tmp, sizeof tmp, "NaN");
    else snprintf(tmp, sizeof tmp, plus ? "%+.4f" : "%.4f", v);
    app(b, tmp);
}
ORACLE_API size_t oracle_show(const char *label, const double *buf, size_t n, char *dst, size_t cap) {
    sbuf b = {dst, cap, 0};
    if (cap) dst[0] = 0;
    app(&b, label);
    app(&b, "\n");
    for (size_t i = 0; i < n; ++i) {
        if (i) app(&b, ", ");
        app_f(&b, buf[2 * i], 0);
        app_f(&b, buf[2 * i + 1], 1);
        app(&b, "i");
    }
    app(&b, "\n");
    return b.need;
}

/* ------------------------------------------------------ signal sources --- */
/* Oscillator::sine, src/oscillator.rs:47-53 -- f32 phase accumulator:
 *   value = phase.sin(); inc = 2.0*PI*freq/sample_rate; phase = (phase+inc) % (2.0*PI) */
ORACLE_API void oracle_osc_sine(float sample_rate, float frequency, float *phase_io, float *out, size_t count) {
    float phase = *phase_io;
    const float pi = 3.14159265358979323846f;
    for (size_t i = 0; i < count; ++i) {
        float value = sinf(phase);
        float inc = 2.0f * pi * frequency / sample_rate;
        phase = fmodf(phase + inc, 2.0f * pi);
        out[i] = value;
    }
    *phase_io = phase;
}
/* fake_sine / sawtooth / square / triangle, src/oscillator.rs:56-91 */
ORACLE_API void oracle_osc_wave(int wave /*1 fake_sine, 2 saw, 3 square, 4 triangle*/, float duty,
                                float sample_rate, float frequency, float *phase_io, float *out, size_t count) {
    float phase = *phase_io;
    const float pi = 3.14159265358979323846f;
    for (size_t i = 0; i < count; ++i) {
        float value;
        if (wave == 1) {
            float x = (phase / pi) - 1.0f;
            value = 4.0f * x * (1.0f - fabsf(x));
            phase = fmodf(phase + 2.0f * pi * frequency / sample_rate, 2.0f * pi);
        } else if (wave == 2) {
            float np_ = phase / (2.0f * pi);
            value = 2.0f * np_ - 1.0f;
            phase = fmodf(phase + 2.0f * pi * frequency / sample_rate, 2.0f * pi);
        } else if (wave == 3) {
            float d = duty < 0.0f ? 0.0f : (duty > 1.0f ? 1.0f : duty);
            value = phase < d ? 1.0f : -1.0f;
            float p = phase + frequency / sample_rate;
            phase = p - truncf(p);
        } else {
            value = 1.0f - 4.0f * fabsf(phase - 0.5f);
            float p = phase + frequency / sample_rate;
            phase = p - truncf(p);
        }
        out[i] = value;
    }
    *phase_io = phase;
}

/* -------------------------------------------------- caller-side pieces --- */
/* Marshalling, examples/beep/view.rs:140-144, graph/player.rs:56-59 (scale =
 * i16::MAX) and pong/main.rs:192-195 (scale = 1): f32 -> f64 -> Complex{re,0}. */
ORACLE_API void oracle_marshal(const float *samples, size_t n, double scale_div, double *out_cplx) {
    for (size_t i = 0; i < n; ++i) {
        out_cplx[2 * i] = (double)samples[i] / scale_div;
        out_cplx[2 * i + 1] = 0.0;
    }
}
/* Magnitude A, examples/beep/view.rs:155: complex.norm() == hypot (f64). */
ORACLE_API void oracle_mag_norm_f64(const double *cplx, size_t n, double *out) {
    for (size_t i = 0; i < n; ++i) out[i] = hypot(cplx[2 * i], cplx[2 * i + 1]);
}
/* Magnitude B, examples/graph/utils.rs:53, examples/pong/utils.rs:40:
 * ((re as f32).powf(2.0) + (im as f32).powf(2.0)).sqrt()  (display scale left to caller). */
ORACLE_API void oracle_mag_sqrt_f32(const double *cplx, size_t n, float *out) {
    for (size_t i = 0; i < n; ++i) {
        float re = (float)cplx[2 * i], im = (float)cplx[2 * i + 1];
        out[i] = sqrtf(re * re + im * im);
    }
}
/* get_narrow_bar_spectrum, examples/graph/utils.rs:85-116 / pong/utils.rs:62-94:
 * means of y over consecutive bands of sizes [4,4,8,16,32,64,128,256] from index 0. */
ORACLE_API void oracle_bands_log(const float *y, float *avg8) {
    static const int sizes[8] = {4, 4, 8, 16, 32, 64, 128, 256};
    int start = 0;
    for (int b = 0; b < 8; ++b) {
        float sum = 0.0f;
        for (int i = start; i < start + sizes[b]; ++i) sum += y[i];
        avg8[b] = sum / (float)sizes[b];
        start += sizes[b];
    }
}
/* get_narrow_bar_spectrum_low, examples/pong/utils.rs:96-114: 8 bands of 3 bins from bin 20. */
ORACLE_API void oracle_bands_low(const float *y, float *avg8) {
    int start = 20;
    for (int b = 0; b < 8; ++b) {
        float sum = 0.0f;
        for (int i = start; i < start + 3; ++i) sum += y[i];
        avg8[b] = sum / 3.0f;
        start += 3;
    }
}
/* argmax + threshold, examples/pong/main.rs:200-213 (strict '>' from max = 0). */
ORACLE_API int oracle_peak_band(const float *avg, int nb, float *max_out) {
    float max = 0.0f;
    int index = 0;
    for (int i = 0; i < nb; ++i)
        if (avg[i] > max) { max = avg[i]; index = i; }
    *max_out = max;
    return index;
}

/* ------------------------------------- composition used by STFT parity --- */
/* The reference has no window / hop / dB.  Parity for the
 * fused GPU path is defined by composition:  w (f32 table) * frame (f32
 * product) -> widen -> oracle_fft -> norm() -> 20*log10 (the dB of the
 * comment at examples/beep/view.rs:155). */
enum { ORACLE_WIN_RECT = 0, ORACLE_WIN_HANN = 1, ORACLE_WIN_HAMMING = 2 };
/* Periodic (DFT-even) windows, computed in f64 and rounded once to f32:
 * hann w[i] = 0.5 - 0.5 cos(2 pi i/n); hamming w[i] = 0.54 - 0.46 cos(2 pi i/n). */
ORACLE_API void oracle_window(int kind, size_t n, float *w) {
    for (size_t i = 0; i < n; ++i) {
        double c = cos(2.0 * M_PI * (double)i / (double)n);
        double v = kind == ORACLE_WIN_HANN ? 0.5 - 0.5 * c : kind == ORACLE_WIN_HAMMING ? 0.54 - 0.46 * c : 1.0;
        w[i] = (float)v;
    }
}

/* One STFT frame through the f64 recursion: out_cplx gets n/2+1 (half) or n
 * (full) complex doubles. */
static void stft_frame_f64(const float *x, const float *w, size_t n, double *scratch_in, double *scratch_out) {
    for (size_t i = 0; i < n; ++i) {
        float xw = w ? x[i] * w[i] : x[i];
        scratch_in[2 * i] = (double)xw;
        scratch_in[2 * i + 1] = 0.0;
    }
    oracle_fft_f64(scratch_in, n, scratch_out);
}
static void stft_frame_f32(const float *x, const float *w, size_t n, float *scratch_in, float *scratch_out) {
    for (size_t i = 0; i < n; ++i) {
        scratch_in[2 * i] = w ? x[i] * w[i] : x[i];
        scratch_in[2 * i + 1] = 0.0f;
    }
    oracle_fft_f32(scratch_in, n, scratch_out);
}

/* Batched drivers: frames are independent, so they are split in contiguous
 * ranges over POSIX threads (the reference itself is single-threaded; threads
 * only matter for the "all host cores" CPU baseline). */
#include <pthread.h>
#include <unistd.h>

ORACLE_API int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
ORACLE_API size_t oracle_stft_frames(size_t n_samples, size_t n, size_t hop) {
    return n_samples < n ? 0 : 1 + (n_samples - n) / hop;
}

typedef struct {
    const float *samples; const float *w;
    size_t n, hop, bins, f0, f1;
    int precision, output;   /* output 0: complex f64; 1 |X|, 2 |X|^2, 3 dB (f32) */
    void *out;
} stft_job;

static void *stft_worker(void *arg) {
    stft_job *j = (stft_job *)arg;
    size_t n = j->n;
    double *si = (double *)malloc(2 * n * sizeof(double)), *so = (double *)malloc(2 * n * sizeof(double));
    float *fi = (float *)malloc(2 * n * sizeof(float)), *fo = (float *)malloc(2 * n * sizeof(float));
    for (size_t f = j->f0; f < j->f1; ++f) {
        const float *x = j->samples + f * j->hop;
        if (j->output == 0) {
            double *o = (double *)j->out + f * j->bins * 2;
            if (j->precision == 64) {
                stft_frame_f64(x, j->w, n, si, so);
                memcpy(o, so, j->bins * 2 * sizeof(double));
            } else {
                stft_frame_f32(x, j->w, n, fi, fo);
                for (size_t k = 0; k < 2 * j->bins; ++k) o[k] = (double)fo[k];
            }
        } else {
            float *o = (float *)j->out + f * j->bins;
            stft_frame_f64(x, j->w, n, si, so);
            for (size_t k = 0; k < j->bins; ++k) {
                double m = hypot(so[2 * k], so[2 * k + 1]);   /* norm(), beep/view.rs:155 */
                o[k] = j->output == 2 ? (float)(m * m) : j->output == 3 ? (float)(20.0 * log10(m)) : (float)m;
            }
        }
    }
    free(si); free(so); free(fi); free(fo);
    return NULL;
}

static void stft_run(const float *samples, size_t n_samples, size_t n, size_t hop, int window, size_t bins,
                     int precision, int output, int threads, void *out) {
    size_t frames = oracle_stft_frames(n_samples, n, hop);
    float *w = NULL;
    if (window != ORACLE_WIN_RECT) {
        w = (float *)malloc(n * sizeof(float));
        oracle_window(window, n, w);
    }
    if (threads <= 0) threads = oracle_max_threads();
    if ((size_t)threads > frames) threads = frames ? (int)frames : 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    stft_job *jobs = (stft_job *)malloc(sizeof(stft_job) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        stft_job j = {samples, w, n, hop, bins, frames * (size_t)t / (size_t)threads,
                      frames * (size_t)(t + 1) / (size_t)threads, precision, output, out};
        jobs[t] = j;
        if (threads == 1) stft_worker(&jobs[t]);
        else pthread_create(&tid[t], NULL, stft_worker, &jobs[t]);
    }
    if (threads > 1)
        for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(tid); free(jobs); free(w);
}

/* Batched STFT, complex f64 result [frames][bins][2].  precision: 64 -> f64
 * recursion, 32 -> f32 recursion widened on output.  threads<=0 -> all cores. */
ORACLE_API void oracle_stft_complex(const float *samples, size_t n_samples, size_t n, size_t hop, int window,
                                    size_t bins, int precision, int threads, double *out) {
    stft_run(samples, n_samples, n, hop, window, bins, precision, 0, threads, out);
}
/* Magnitude-only f32 result [frames][bins] (norm() of the f64 recursion,
 * narrowed) -- the shape bench.py's CPU baseline times. output: 1 |X|, 2 |X|^2, 3 dB. */
ORACLE_API void oracle_stft_mag(const float *samples, size_t n_samples, size_t n, size_t hop, int window,
                                size_t bins, int output, int threads, float *out) {
    stft_run(samples, n_samples, n, hop, window, bins, 64, output, threads, out);
}

/* ------------------------------------------------- decoder-side callers --- */

/* i16 frame -> the f32 sample the f32 pipeline transforms.  examples/graph/player.rs:56-59:
 *   frame[0] as f64 / std::i16::MAX as f64            (f64 division, then our f32 path narrows once)
 * pcm: n_frames x channels interleaved; out[i] = (float)((double)pcm[i*channels + channel] / divisor). */
ORACLE_API void oracle_pcm16_to_f32(const int16_t *pcm, size_t n_frames, size_t channels, size_t channel,
                                    double divisor, float *out) {
    for (size_t i = 0; i < n_frames; ++i) out[i] = (float)((double)pcm[i * channels + channel] / divisor);
}

/* load_track_at_path, examples/graph/player.rs:81-102: frames.iter().map(|frame| [volume_factor * frame[0] as f32,
 * volume_factor * frame[1] as f32]).flatten() -- one f32 product per i16 value (volume_factor = 0.00002 there). */
ORACLE_API void oracle_pcm16_scale_f32(const int16_t *pcm, size_t n_values, float factor, float *out) {
    for (size_t i = 0; i < n_values; ++i) out[i] = factor * (float)pcm[i];
}

/* the capture callback of examples/pong/audio.rs:57-70: for frame in data.chunks(channel_count) { average = 0.0; for
 * sample in frame.iter().take(2) { average += sample; } average = average / frame.len() as f32; push(average) } */
ORACLE_API void oracle_f32_downmix(const float *frames, size_t n_frames, size_t channels, float *out) {
    for (size_t i = 0; i < n_frames; ++i) {
        float average = 0.0f;
        for (size_t c = 0; c < channels && c < 2; ++c) average += frames[i * channels + c];
        out[i] = average / (float)channels;
    }
}

/* detect_bpm, src/decoder.rs:39-71, statement by statement.
 *   f_frames = frames.map(|f| [f[0] as f32 / i16::MAX as f32, f[1] as f32 / i16::MAX as f32])     :41-44
 *   duration = (frames.len() as f32 / sample_rate) as usize                                        :46 (:35-37)
 *   for i in 0..duration: start = i*44100, end = (i+1)*44100      (44100 is hard coded)            :48-50
 *     average_e = sum_{start..end} (f[0].powf(2.0) + f[1].powf(2.0))   sequential f32 fold from 0  :51-53
 *     average_e *= 1024.0 / 44100.0                                                                :54
 *     for j in 0..43: e = the same sum over the 1024 frames of block j; if e > C*average_e: beats  :57-66
 *   beats * (60 / duration as u32)                                    u32 division                 :70
 * powf(x, 2.0) is x*x correctly rounded.  The Rust original panics when duration == 0 (division by zero) or when
 * duration*44100 exceeds the slice (sample_rate < 44100): both return -1 here.  block_e (duration*43) and avg_e
 * (duration, already scaled) are optional outputs for the parity tests.  Returns bpm, *beats_out = beats. */
ORACLE_API int64_t oracle_detect_bpm(const int16_t *frames, size_t n_frames, float sample_rate, uint32_t *beats_out,
                                     float *block_e, float *avg_e) {
    const float C = 5.5f, imax = (float)INT16_MAX;
    const size_t duration = (size_t)((float)n_frames / sample_rate);
    if (duration == 0 || duration * 44100 > n_frames) return -1;
    uint32_t beats = 0;
    for (size_t i = 0; i < duration; ++i) {
        const int16_t *sec = frames + 2 * i * 44100;
        float average_e = 0.0f;
        for (size_t k = 0; k < 44100; ++k) {
            float a = (float)sec[2 * k] / imax, b = (float)sec[2 * k + 1] / imax;
            float a2 = a * a, b2 = b * b;
            float s = a2 + b2;
            average_e = average_e + s;
        }
        average_e = average_e * (1024.0f / 44100.0f);
        if (avg_e) avg_e[i] = average_e;
        for (size_t j = 0; j < 43; ++j) {
            float e = 0.0f;
            for (size_t k = j * 1024; k < (j + 1) * 1024; ++k) {
                float a = (float)sec[2 * k] / imax, b = (float)sec[2 * k + 1] / imax;
                float a2 = a * a, b2 = b * b;
                float s = a2 + b2;
                e = e + s;
            }
            if (block_e) block_e[i * 43 + j] = e;
            if (e > C * average_e) beats += 1;
        }
    }
    if (beats_out) *beats_out = beats;
    return (int64_t)(beats * (60u / (uint32_t)duration));
}
