// libkopek_cuda.so -- implementation of the C ABI in include/kopek_cuda.h.
// Host-side plan/stream management around the sm_100a kernels K1 (fft_batched.cuh)
// and K2 (fft_c2c_f64.cuh).  No CPU compute path exists here: without a CUDA
// device every call fails with KOPEK_ERR_CUDA.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "fft_c2c_f64.cuh"
#include "k1_launch.cuh"
#include "k1w_launch.cuh"

namespace kopek {

// ---------------------------------------------------------------- errors ----
char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
kopek_status fail(kopek_status code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

static int ilog2(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) ++l;
    return l;
}

// ------------------------------------------------------- K2 host context ----
// The drop-in fft() is a pure function in the reference (src/fft.rs:6); here it
// owns a lazily created per-process context (device buffers, pinned staging,
// one stream, twiddle tables keyed by n) behind a mutex so that any thread may
// call it, like the original.
struct TwiddleTable {
    size_t n = 0;
    int device = -1;
    double2 *d = nullptr;
};

struct K2Context {
    std::mutex mu;
    int device = 0;
    bool init = false;
    cudaStream_t stream = nullptr;
    void *pinned = nullptr, *pinned_dev = nullptr;   // mapped: the device reads/writes it directly on the latency path
    uint32_t *flag = nullptr, *flag_dev = nullptr;   // mapped completion word of the latency path
    uint32_t token = 0;
    size_t pinned_bytes = 0;
    double2 *d_in = nullptr, *d_out = nullptr;
    size_t d_in_elems = 0, d_out_elems = 0;
};
static K2Context g_k2;
static std::mutex g_tw_mu;
static std::vector<TwiddleTable> g_tw_tables;   // one table per (n, device), kept for the life of the process

constexpr int K2_CHUNK_LOG_MAX = 12;       // 4096 complex f64 = 64 KiB of shared memory
constexpr size_t K2_PINNED_MAX = 8u << 20; // staging through pinned memory up to 8 MiB
constexpr uint64_t K1_HOST_CHUNK_MB = 16;     // samples per in-flight chunk of kopek_stft_exec_host
constexpr size_t K2_ZERO_COPY_MAX = 512u << 10;   // in + out of a call served without DMA (mapped pinned memory)

// tw[j] = exp(-i pi (2j)/n) exactly as src/fft.rs:24 builds it:
//   im = ((-PI) * (i as f64)) / (n as f64), i = 2j;  w = (cos im, sin im)
static kopek_status k2_get_twiddles(size_t n, int device, const double2 **out) {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    for (size_t t = 0; t < g_tw_tables.size(); ++t)
        if (g_tw_tables[t].n == n && g_tw_tables[t].device == device) {
            *out = g_tw_tables[t].d;
            return KOPEK_OK;
        }
    size_t cnt = n / 2 ? n / 2 : 1;
    std::vector<double2> h(cnt);
    const double fn = (double)n;
    for (size_t j = 0; j < cnt; ++j) {
        double fi = (double)(2 * j);
        double th = (-1.0 * M_PI) * fi / fn;
        h[j] = make_double2(cos(th), sin(th));
    }
    TwiddleTable tb;
    tb.n = n;
    tb.device = device;
    KOPEK_CUDA(cudaMalloc(&tb.d, cnt * sizeof(double2)));
    cudaError_t e = cudaMemcpy(tb.d, h.data(), cnt * sizeof(double2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(tb.d);
        return fail(KOPEK_ERR_CUDA, "twiddle upload failed: %s", cudaGetErrorString(e));
    }
    // never evicted: a caller may still be about to launch on a table handed out earlier, and there are at most
    // 31 power-of-two lengths per device (2^30 points = 8 GiB of twiddles only if someone really transforms 2^30)
    g_tw_tables.push_back(tb);
    *out = tb.d;
    return KOPEK_OK;
}

// Device-resident batched c2c f64.  n = next_pow2(n_in) per transform.  `flag`/`token`: see k2_c2c_f64_chunk; only
// honoured (and *flag_used set) when the whole call is one single-CTA launch.
static kopek_status k2_run_device(const double2 *d_in, size_t n_in, size_t batch, double2 *d_out, int device,
                                  cudaStream_t stream, volatile uint32_t *flag = nullptr, uint32_t token = 0,
                                  bool *flag_used = nullptr) {
    const size_t n = kopek_next_pow2(n_in);
    if (n > ((size_t)1 << 30)) return fail(KOPEK_ERR_UNSUPPORTED, "fft length %zu exceeds 2^30", n);
    const int logn = ilog2(n);
    const double2 *tw = nullptr;
    kopek_status st = k2_get_twiddles(n, device, &tw);
    if (st != KOPEK_OK) return st;
    const int chunk_log = std::min(logn, K2_CHUNK_LOG_MAX);
    const size_t smem = sizeof(double2) << chunk_log;
    // function attributes belong to the device's context: one flag per device (idempotent, so a race is benign)
    static std::atomic<bool> attr_done[64];
    if (device < 0 || device >= 64 || !attr_done[device].load(std::memory_order_acquire)) {
        KOPEK_CUDA(cudaFuncSetAttribute(k2_c2c_f64_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double2) << K2_CHUNK_LOG_MAX)));
        if (device >= 0 && device < 64) attr_done[device].store(true, std::memory_order_release);
    }
    const size_t chunks = n >> chunk_log;
    const int threads = (int)std::max<size_t>(32, ((size_t)1 << chunk_log) / 8);   // one 8-element item per thread
    const bool use_flag = flag && chunks == 1 && batch == 1;
    if (flag_used) *flag_used = use_flag;
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {
        size_t nb = std::min<size_t>(65535, batch - b0);
        dim3 grid((unsigned)chunks, (unsigned)nb);
        k2_c2c_f64_chunk<<<grid, threads, smem, stream>>>(d_in + b0 * n_in, n_in, n_in, d_out + b0 * n, n, tw, logn,
                                                         chunk_log, use_flag ? flag : nullptr, token);
        KOPEK_CUDA(cudaGetLastError());
        // the levels above the chunk size: global passes of up to three levels each
        for (int lv = chunk_log + 1; lv <= logn;) {
            const int r = (logn - lv + 1) % 3 ? (logn - lv + 1) % 3 : 3;
            const size_t items = n >> r;
            for (size_t b = 0; b < nb; ++b) {
                double2 *d = d_out + (b0 + b) * n;
                const unsigned g = (unsigned)((items + 255) / 256);
                if (r == 1) k2_c2c_f64_levels<1><<<g, 256, 0, stream>>>(d, tw, logn, lv);
                else if (r == 2) k2_c2c_f64_levels<2><<<g, 256, 0, stream>>>(d, tw, logn, lv);
                else k2_c2c_f64_levels<3><<<g, 256, 0, stream>>>(d, tw, logn, lv);
                KOPEK_CUDA(cudaGetLastError());
            }
            lv += r;
        }
    }
    return KOPEK_OK;
}

// Wait for `token` to land in the mapped flag word: the kernel's last action.  Polling host memory costs a cache miss
// per look; a stream synchronisation costs several microseconds after the kernel has ended.  The stream is queried
// now and then so that a failed launch cannot spin forever.
static kopek_status k2_wait_flag(const volatile uint32_t *flag, uint32_t token, cudaStream_t stream) {
    for (uint32_t spins = 1;; ++spins) {
        if (*flag == token) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return KOPEK_OK;
        }
        if ((spins & 0xFFFu) == 0) {
            cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaSuccess) {           // the kernel has ended: its token is on its way or there already
                KOPEK_CUDA(cudaStreamSynchronize(stream));
                return KOPEK_OK;
            }
            if (q != cudaErrorNotReady) return fail(KOPEK_ERR_CUDA, "fft kernel failed: %s", cudaGetErrorString(q));
        }
    }
}

static kopek_status k2_ensure(K2Context &c, size_t in_elems, size_t out_elems, size_t pinned_bytes) {
    if (!c.init) {
        const char *env = getenv("KOPEK_DEVICE");
        c.device = env ? atoi(env) : 0;
        int cnt = 0;
        KOPEK_CUDA(cudaGetDeviceCount(&cnt));
        if (cnt <= 0) return fail(KOPEK_ERR_CUDA, "no CUDA device visible (no CPU fallback exists)");
        if (c.device < 0 || c.device >= cnt) return fail(KOPEK_ERR_INVALID, "KOPEK_DEVICE=%d out of range", c.device);
    }
    DeviceGuard g(c.device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", c.device);
    if (!c.init) {
        KOPEK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        KOPEK_CUDA(cudaHostAlloc(&c.flag, 64, cudaHostAllocMapped));
        KOPEK_CUDA(cudaHostGetDevicePointer(&c.flag_dev, c.flag, 0));
        *c.flag = 0;
        c.init = true;
    }
    if (in_elems > c.d_in_elems) {
        if (c.d_in) cudaFree(c.d_in);
        c.d_in = nullptr; c.d_in_elems = 0;
        KOPEK_CUDA(cudaMalloc(&c.d_in, in_elems * sizeof(double2)));
        c.d_in_elems = in_elems;
    }
    if (out_elems > c.d_out_elems) {
        if (c.d_out) cudaFree(c.d_out);
        c.d_out = nullptr; c.d_out_elems = 0;
        KOPEK_CUDA(cudaMalloc(&c.d_out, out_elems * sizeof(double2)));
        c.d_out_elems = out_elems;
    }
    if (pinned_bytes > c.pinned_bytes) {
        if (c.pinned) cudaFreeHost(c.pinned);
        c.pinned = nullptr; c.pinned_bytes = 0;
        KOPEK_CUDA(cudaHostAlloc(&c.pinned, pinned_bytes, cudaHostAllocMapped));
        c.pinned_bytes = pinned_bytes;
        KOPEK_CUDA(cudaHostGetDevicePointer(&c.pinned_dev, c.pinned, 0));
    }
    return KOPEK_OK;
}

static kopek_status k2_run_host(const double *in, size_t n_in, size_t batch, double *out) {
    const size_t n = kopek_next_pow2(n_in);
    if (batch == 0) return KOPEK_OK;
    if (n_in > 0 && !in) return fail(KOPEK_ERR_INVALID, "in is NULL");
    if (!out) return fail(KOPEK_ERR_INVALID, "out is NULL");
    K2Context &c = g_k2;
    std::lock_guard<std::mutex> lk(c.mu);
    const size_t in_elems = std::max<size_t>(1, n_in * batch), out_elems = n * batch;
    const size_t in_bytes = n_in * batch * sizeof(double2), out_bytes = out_elems * sizeof(double2);
    const bool stage = std::max(in_bytes, out_bytes) <= K2_PINNED_MAX;
    // Latency path (one UI-frame call = 16 KB in, 16 KB out): no DMA at all.  The pinned staging buffer is mapped into
    // the device's address space, the single-CTA kernel gathers its bit-reversed input straight from it over PCIe and
    // writes the spectrum back into it, so a call is one launch + one stream synchronisation instead of two copy
    // launches around the kernel (46 -> ~25 us per 1024-point call).  Same kernel, same bits.
    const bool zero_copy = stage && n <= ((size_t)1 << K2_CHUNK_LOG_MAX) && in_bytes + out_bytes <= K2_ZERO_COPY_MAX;
    const size_t out_off = (in_bytes + 255) & ~(size_t)255;
    kopek_status st = k2_ensure(c, zero_copy ? 1 : in_elems, zero_copy ? 1 : out_elems,
                                zero_copy ? out_off + out_bytes : stage ? std::max(in_bytes, out_bytes) : 0);
    if (st != KOPEK_OK) return st;
    DeviceGuard g(c.device);
    if (zero_copy) {
        double2 *pin_in = reinterpret_cast<double2 *>(c.pinned);
        double2 *pin_out = reinterpret_cast<double2 *>(reinterpret_cast<char *>(c.pinned) + out_off);
        if (in_bytes) memcpy(pin_in, in, in_bytes);
        bool flagged = false;
        if (++c.token == 0) c.token = 1;
        st = k2_run_device(reinterpret_cast<const double2 *>(c.pinned_dev), n_in, batch,
                           reinterpret_cast<double2 *>(reinterpret_cast<char *>(c.pinned_dev) + out_off), c.device, c.stream,
                           c.flag_dev, c.token, &flagged);
        if (st != KOPEK_OK) return st;
        if (flagged) st = k2_wait_flag(c.flag, c.token, c.stream);
        else KOPEK_CUDA(cudaStreamSynchronize(c.stream));
        if (st != KOPEK_OK) return st;
        memcpy(out, pin_out, out_bytes);
        return KOPEK_OK;
    }
    if (in_bytes) {
        if (stage) {
            memcpy(c.pinned, in, in_bytes);
            KOPEK_CUDA(cudaMemcpyAsync(c.d_in, c.pinned, in_bytes, cudaMemcpyHostToDevice, c.stream));
        } else {
            KOPEK_CUDA(cudaMemcpyAsync(c.d_in, in, in_bytes, cudaMemcpyHostToDevice, c.stream));
        }
    }
    st = k2_run_device(c.d_in, n_in, batch, c.d_out, c.device, c.stream);
    if (st != KOPEK_OK) return st;
    KOPEK_CUDA(cudaMemcpyAsync(stage ? c.pinned : (void *)out, c.d_out, out_bytes, cudaMemcpyDeviceToHost, c.stream));
    KOPEK_CUDA(cudaStreamSynchronize(c.stream));
    if (stage) memcpy(out, c.pinned, out_bytes);
    return KOPEK_OK;
}

// --------------------------------------------------------------- K1 plans ----
typedef cudaError_t (*k1_launch_fn)(const K1Params &, int, bool, bool, bool, int, cudaStream_t);
typedef int (*k1_occ_fn)(int, bool);
typedef K1Info (*k1_info_fn)();

struct HostSlot {            // one in-flight chunk of kopek_stft_exec_host
    cudaStream_t stream = nullptr;
    float *d_in = nullptr;
    void *d_out = nullptr;
};

}  // namespace kopek

using namespace kopek;

struct kopek_plan {
    uint32_t n, hop, window, output, bins_mode;
    uint64_t bins, pitch, elem_bytes;
    int device;
    int sm_count, occupancy;
    K1Info info;
    k1_launch_fn launch;
    float *d_win;      // 0.5 * w
    float2 *d_tw, *d_pw;
    float4 *d_k1w;     // tables of the sub-warp kernels (n == 256 / 512); n == 1024: the three-pass kernel in tuning builds
    float4 *d_k1b;     // tables of the two-exchange 1024-point kernel
    int k1b_warps_per_sm;
    float4 *d_k1m;     // tables of the team-per-frame kernel (n == 2048 / 4096)
    int k1m_teams_per_sm, k1m_wink;
    int k1w_variant, k1w_warps_per_sm;
    float band_scale;
    int bulk_store;
    mutable uint32_t last_launches;
    // exec_host resources (lazy)
    std::mutex mu;
    HostSlot slots[3];
    uint64_t slot_frames;
};

namespace kopek {

static void host_window(uint32_t kind, uint32_t n, std::vector<float> &w) {
    w.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        double c = cos(2.0 * M_PI * (double)i / (double)n);
        double v = kind == KOPEK_WINDOW_HANN ? 0.5 - 0.5 * c : kind == KOPEK_WINDOW_HAMMING ? 0.54 - 0.46 * c : 1.0;
        w[i] = (float)v;
    }
}

static kopek_status upload_window(kopek_plan *p, const float *w) {
    // the kernel's window table carries the 1/2 of the real-input split (exact scaling)
    std::vector<float> half(p->n);
    for (uint32_t i = 0; i < p->n; ++i) half[i] = 0.5f * w[i];
    if (!p->d_win) KOPEK_CUDA(cudaMalloc(&p->d_win, p->n * sizeof(float)));
    KOPEK_CUDA(cudaMemcpy(p->d_win, half.data(), p->n * sizeof(float), cudaMemcpyHostToDevice));
    if (p->d_k1m) {
        std::vector<float4> t(k1m_table_f4((int)p->n));
        const double cs_a = p->window == KOPEK_WINDOW_HANN ? 0.5 : 0.54, cs_b = p->window == KOPEK_WINDOW_HANN ? 0.5 : 0.46;
        k1m_build_tables((int)p->n, half.data(), cs_a, cs_b, t.data());
        KOPEK_CUDA(cudaMemcpy(p->d_k1m, t.data(), sizeof(float4) * t.size(), cudaMemcpyHostToDevice));
    }
    if (p->d_k1b) {
        std::vector<float4> tb(K1B_TABLE_F4);
        k1b_build_tables(half.data(), tb.data());
        KOPEK_CUDA(cudaMemcpy(p->d_k1b, tb.data(), sizeof(float4) * K1B_TABLE_F4, cudaMemcpyHostToDevice));
    }
    if (p->d_k1w && p->n == 256) {
        std::vector<float4> t(K1W256_TABLE_F4);
        k1w256_build_tables(half.data(), t.data());
        KOPEK_CUDA(cudaMemcpy(p->d_k1w, t.data(), sizeof(float4) * K1W256_TABLE_F4, cudaMemcpyHostToDevice));
    } else if (p->d_k1w && p->n == 512) {
        std::vector<float4> t(K1W512_TABLE_F4);
        k1w512_build_tables(half.data(), t.data());
        KOPEK_CUDA(cudaMemcpy(p->d_k1w, t.data(), sizeof(float4) * K1W512_TABLE_F4, cudaMemcpyHostToDevice));
    }
    return KOPEK_OK;
}

static void plan_free_device(kopek_plan *p) {
    if (p->d_win) cudaFree(p->d_win);
    if (p->d_tw) cudaFree(p->d_tw);
    if (p->d_pw) cudaFree(p->d_pw);
    if (p->d_k1w) cudaFree(p->d_k1w);
    if (p->d_k1b) cudaFree(p->d_k1b);
    if (p->d_k1m) cudaFree(p->d_k1m);
    for (HostSlot &s : p->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
}

}  // namespace kopek

// ================================================================ C ABI =====
extern "C" {

const char *kopek_last_error(void) { return err_buf(); }
const char *kopek_version(void) { return "kopek_b200 0.1.0 (sm_100a)"; }

int32_t kopek_device_count(void) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return cnt;
}

size_t kopek_next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

kopek_status kopek_fft_c2c_f64(const double *in, size_t n_in, double *out, size_t n_out) {
    if (n_out != kopek_next_pow2(n_in))
        return fail(KOPEK_ERR_INVALID, "n_out=%zu must be next_power_of_two(n_in=%zu)=%zu", n_out, n_in,
                    kopek_next_pow2(n_in));
    return k2_run_host(in, n_in, 1, out);
}

// Inverse through the forward transform: x = conj(fft(conj(X))) / n.  The reference has no inverse (SURVEY.md 8f row
// 4); this is the one a caller of its fft() would write, and it makes fft -> ifft round trips a property test of the
// forward path at any size.  n must be a power of two (an inverse cannot zero-pad); 1/n is exact.
kopek_status kopek_ifft_c2c_f64(const double *in, size_t n, double *out) {
    if (n == 0 || (n & (n - 1))) return fail(KOPEK_ERR_INVALID, "inverse transform: n=%zu must be a power of two >= 1", n);
    if (!in || !out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    std::vector<double> tmp;
    try { tmp.resize(2 * n); } catch (...) { return fail(KOPEK_ERR_NOMEM, "out of host memory"); }
    for (size_t i = 0; i < n; ++i) { tmp[2 * i] = in[2 * i]; tmp[2 * i + 1] = -in[2 * i + 1]; }
    kopek_status st = k2_run_host(tmp.data(), n, 1, out);
    if (st != KOPEK_OK) return st;
    const double inv = 1.0 / (double)n;
    for (size_t i = 0; i < n; ++i) { out[2 * i] = out[2 * i] * inv; out[2 * i + 1] = -out[2 * i + 1] * inv; }
    return KOPEK_OK;
}

kopek_status kopek_fft_c2c_f64_batched(const double *in, size_t n_in, size_t batch, double *out) {
    return k2_run_host(in, n_in, batch, out);
}

kopek_status kopek_fft_c2c_f64_device(const double *d_in, size_t n_in, size_t batch, double *d_out, int device,
                                      void *cuda_stream) {
    if (batch == 0) return KOPEK_OK;
    if ((n_in > 0 && !d_in) || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if (d_in == d_out) return fail(KOPEK_ERR_INVALID, "in-place transform is not supported");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    return k2_run_device(reinterpret_cast<const double2 *>(d_in), n_in, batch, reinterpret_cast<double2 *>(d_out),
                         device, (cudaStream_t)cuda_stream);
}

// "{:.4}{:+.4}i" -- Rust prints NaN unsigned ("NaN"), infinities as "inf"/"-inf"/"+inf".
static void fmt_bin(std::string &s, double v, bool plus) {
    char tmp[400];
    if (v != v) snprintf(tmp, sizeof tmp, "NaN");
    else snprintf(tmp, sizeof tmp, plus ? "%+.4f" : "%.4f", v);
    s += tmp;
}
size_t kopek_show_format(const char *label, const double *buf, size_t n, char *dst, size_t cap) {
    std::string s(label ? label : "");
    s += '\n';
    for (size_t i = 0; i < n; ++i) {
        if (i) s += ", ";
        fmt_bin(s, buf[2 * i], false);
        fmt_bin(s, buf[2 * i + 1], true);
        s += 'i';
    }
    s += '\n';
    if (dst && cap) {
        size_t cp = std::min(cap - 1, s.size());
        memcpy(dst, s.data(), cp);
        dst[cp] = 0;
    }
    return s.size();
}
kopek_status kopek_show(const char *label, const double *buf, size_t n) {
    if (n && !buf) return fail(KOPEK_ERR_INVALID, "buf is NULL");
    size_t need = kopek_show_format(label, buf, n, nullptr, 0);
    std::string s(need + 1, '\0');
    kopek_show_format(label, buf, n, &s[0], need + 1);
    fwrite(s.data(), 1, need, stdout);
    fflush(stdout);
    return KOPEK_OK;
}

// ------------------------------------------------------------------ plans ----
kopek_status kopek_plan_create(kopek_plan **plan, uint32_t n, uint32_t hop, uint32_t window, uint32_t output,
                               uint32_t bins, uint64_t out_pitch_elems, int device) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    *plan = nullptr;
    k1_launch_fn launch = nullptr;
    k1_occ_fn occ = nullptr;
    k1_info_fn info = nullptr;
    switch (n) {
#define KOPEK_CASE(N_) case N_: launch = k1_launch_##N_; occ = k1_occupancy_##N_; info = k1_info_##N_; break;
        KOPEK_CASE(256) KOPEK_CASE(512) KOPEK_CASE(1024) KOPEK_CASE(2048) KOPEK_CASE(4096)
#undef KOPEK_CASE
        default:
            return fail(KOPEK_ERR_UNSUPPORTED, "frame size %u: must be a power of two in [256, 4096]", n);
    }
    if (hop == 0) return fail(KOPEK_ERR_INVALID, "hop must be >= 1");
    if (window > KOPEK_WINDOW_CUSTOM) return fail(KOPEK_ERR_INVALID, "unknown window %u", window);
    if (output > KOPEK_OUT_BANDS_LOW) return fail(KOPEK_ERR_INVALID, "unknown output %u", output);
    const bool bands = output >= KOPEK_OUT_BANDS_LOG;
    if (bands && (n != 1024 || bins != KOPEK_BINS_HALF))
        return fail(KOPEK_ERR_UNSUPPORTED, "band epilogues need n = 1024 and KOPEK_BINS_HALF (the reference's callers "
                                           "analyse 1024-sample frames)");
    if (bins > KOPEK_BINS_FULL) return fail(KOPEK_ERR_INVALID, "unknown bins mode %u", bins);
    const uint64_t nb = bands ? KOPEK_BAND_ROW : bins == KOPEK_BINS_FULL ? n : n / 2 + 1;
    if (out_pitch_elems == 0) out_pitch_elems = nb;
    if (out_pitch_elems < nb) return fail(KOPEK_ERR_INVALID, "out_pitch_elems %llu < bins per frame %llu",
                                          (unsigned long long)out_pitch_elems, (unsigned long long)nb);
    int cnt = 0;
    KOPEK_CUDA(cudaGetDeviceCount(&cnt));
    if (cnt <= 0) return fail(KOPEK_ERR_CUDA, "no CUDA device visible (no CPU fallback exists)");
    if (device < 0 || device >= cnt) return fail(KOPEK_ERR_INVALID, "device %d out of range [0,%d)", device, cnt);
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);

    kopek_plan *p = new (std::nothrow) kopek_plan();
    if (!p) return fail(KOPEK_ERR_NOMEM, "out of host memory");
    p->n = n; p->hop = hop; p->window = window; p->output = output; p->bins_mode = bins;
    p->bins = nb; p->pitch = out_pitch_elems; p->elem_bytes = output == KOPEK_OUT_COMPLEX ? 8 : 4;
    p->device = device; p->launch = launch; p->info = info();
    p->d_win = nullptr; p->d_tw = nullptr; p->d_pw = nullptr; p->last_launches = 0; p->slot_frames = 0;
    p->d_k1b = nullptr; p->k1b_warps_per_sm = 0; p->d_k1m = nullptr; p->k1m_teams_per_sm = 0;
    p->d_k1w = nullptr; p->k1w_variant = 0; p->k1w_warps_per_sm = 0; p->band_scale = 1.0f; p->bulk_store = 0;

    kopek_status st = KOPEK_OK;
    do {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) { st = fail(KOPEK_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); break; }
        if (prop.major != 10) {
            st = fail(KOPEK_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library holds sm_100a code only", device,
                      prop.major, prop.minor);
            break;
        }
        p->sm_count = prop.multiProcessorCount;
        p->occupancy = occ(bands ? (int)KOPEK_OUT_MAG : (int)output, window != KOPEK_WINDOW_RECT);
        if (p->occupancy <= 0) { st = fail(KOPEK_ERR_CUDA, "kernel for n=%u cannot be resident: %s", n,
                                           cudaGetErrorString(cudaGetLastError())); break; }
        const uint32_t M = n / 2;
        std::vector<float2> tw(M), pw(M / 2 + 1);
        switch (n) {
            case 256: k1_build_stage_twiddles<256>(tw.data()); break;
            case 512: k1_build_stage_twiddles<512>(tw.data()); break;
            case 1024: k1_build_stage_twiddles<1024>(tw.data()); break;
            case 2048: k1_build_stage_twiddles<2048>(tw.data()); break;
            default: k1_build_stage_twiddles<4096>(tw.data()); break;
        }
        for (uint32_t k = 0; k <= M / 2; ++k) {
            double th = 2.0 * M_PI * (double)k / (double)n;
            pw[k] = make_float2((float)-sin(th), (float)-cos(th));
        }
        pw[M / 2] = make_float2(-1.0f, 0.0f);   // -i W_N^{N/4} = -1 exactly
        if ((e = cudaMalloc(&p->d_tw, M * sizeof(float2))) != cudaSuccess ||
            (e = cudaMalloc(&p->d_pw, (M / 2 + 1) * sizeof(float2))) != cudaSuccess ||
            (e = cudaMemcpy(p->d_tw, tw.data(), M * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(p->d_pw, pw.data(), (M / 2 + 1) * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
            st = fail(KOPEK_ERR_CUDA, "plan table upload failed: %s", cudaGetErrorString(e));
            break;
        }
        if (n == 1024) {
            // headline size: the two-exchange warp-per-frame kernel (fft_warp1024b.cuh), every output; K1 stays for
            // the generic K1 is kept as a cross-check (KOPEK_K1W_VARIANT < 0)
            const char *env = getenv("KOPEK_K1W_VARIANT");
            p->k1w_variant = env ? atoi(env) : 200;
            if (p->k1w_variant >= 0) p->k1w_variant = std::max(p->k1w_variant, 200);
            if (p->k1w_variant < 0 && bands) p->k1w_variant = 200;      // the generic kernel has no band epilogue
            K1WParams dummy{};
            if (p->k1w_variant >= 200) {
                std::vector<float4> tb(K1B_TABLE_F4);
                k1b_build_tables(nullptr, tb.data());
                if ((e = cudaMalloc(&p->d_k1b, sizeof(float4) * K1B_TABLE_F4)) != cudaSuccess ||
                    (e = cudaMemcpy(p->d_k1b, tb.data(), sizeof(float4) * K1B_TABLE_F4, cudaMemcpyHostToDevice)) !=
                        cudaSuccess ||
                    (e = k1b_launch(dummy, (int)output, window != KOPEK_WINDOW_RECT, false, 0, nullptr,
                                    &p->k1b_warps_per_sm)) != cudaSuccess ||
                    p->k1b_warps_per_sm <= 0) {
                    st = fail(KOPEK_ERR_CUDA, "K1W-1024b setup failed: %s", cudaGetErrorString(e));
                    break;
                }
            }
        }
        if ((n == 256 || n == 512) && output <= KOPEK_OUT_DB) {
            // sub-warp frames: 8 (fft_warp256.cuh) / 16 (fft_warp512.cuh) lanes per frame; K1 stays for FULL bins / odd hops
            const char *env = getenv("KOPEK_K1W_VARIANT");
            if (!(env && atoi(env) < 0)) {
                const int nf4 = n == 256 ? K1W256_TABLE_F4 : K1W512_TABLE_F4;
                std::vector<float4> t(nf4);
                if (n == 256) k1w256_build_tables(nullptr, t.data());
                else k1w512_build_tables(nullptr, t.data());
                if ((e = cudaMalloc(&p->d_k1w, sizeof(float4) * nf4)) != cudaSuccess ||
                    (e = cudaMemcpy(p->d_k1w, t.data(), sizeof(float4) * nf4, cudaMemcpyHostToDevice)) != cudaSuccess) {
                    st = fail(KOPEK_ERR_CUDA, "K1W-%u table upload failed: %s", n, cudaGetErrorString(e));
                    break;
                }
                K1WParams dummy{};
                e = n == 256 ? k1w256_launch(dummy, (int)output, window != KOPEK_WINDOW_RECT, 0, 0, nullptr, &p->k1w_warps_per_sm)
                             : k1w512_launch(dummy, (int)output, window != KOPEK_WINDOW_RECT, 0, 0, nullptr, &p->k1w_warps_per_sm);
                if (e != cudaSuccess || p->k1w_warps_per_sm <= 0) {
                    st = fail(KOPEK_ERR_CUDA, "K1W-%u kernel cannot be resident: %s", n, cudaGetErrorString(e));
                    break;
                }
            }
        }
        {
            // team-per-frame kernel (fft_team.cuh): 2048 / 4096 points; K1 stays for FULL bins / odd hops
            const char *env = getenv("KOPEK_K1W_VARIANT");
            const int v = env ? atoi(env) : 0;
            if (output <= KOPEK_OUT_DB && v >= 0 && (n == 2048 || n == 4096)) {
                std::vector<float4> t(k1m_table_f4((int)n));
                k1m_build_tables((int)n, nullptr, 0.0, 0.0, t.data());
                K1WParams dummy{};
                p->k1m_wink = window == KOPEK_WINDOW_RECT ? 0 : window == KOPEK_WINDOW_CUSTOM ? 3 : 2;
                if (const char *wk = getenv("KOPEK_K1M_WINK"))          // tuning: 3 = shared-memory table for Hann / Hamming too
                    if (window != KOPEK_WINDOW_RECT && atoi(wk) == 3) p->k1m_wink = 3;
                if ((e = cudaMalloc(&p->d_k1m, sizeof(float4) * t.size())) != cudaSuccess ||
                    (e = cudaMemcpy(p->d_k1m, t.data(), sizeof(float4) * t.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
                    (e = k1m_launch((int)n, dummy, (int)output, p->k1m_wink, 0, 0, nullptr,
                                    &p->k1m_teams_per_sm)) != cudaSuccess ||
                    p->k1m_teams_per_sm <= 0) {
                    st = fail(KOPEK_ERR_CUDA, "K1M-%u setup failed: %s", n, cudaGetErrorString(e));
                    break;
                }
                if (getenv("KOPEK_DEBUG")) fprintf(stderr, "[kopek] K1M-%u: %d teams per SM\n", n, p->k1m_teams_per_sm);
            }
        }
        if (window == KOPEK_WINDOW_HANN || window == KOPEK_WINDOW_HAMMING) {
            std::vector<float> w;
            host_window(window, n, w);
            st = upload_window(p, w.data());
        } else if (window == KOPEK_WINDOW_CUSTOM) {
            std::vector<float> w(n, 1.0f);   // until kopek_plan_set_window is called
            st = upload_window(p, w.data());
        }
    } while (0);
    if (st != KOPEK_OK) {
        plan_free_device(p);
        delete p;
        return st;
    }
    *plan = p;
    return KOPEK_OK;
}

kopek_status kopek_plan_set_bulk_store(kopek_plan *plan, int enable) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    if (enable && (plan->n != 1024 || plan->pitch % 4 != 0 || plan->output < KOPEK_OUT_MAG || plan->output > KOPEK_OUT_DB ||
                   plan->bins_mode != KOPEK_BINS_HALF))
        return fail(KOPEK_ERR_UNSUPPORTED, "bulk store needs n = 1024, HALF bins, an f32 output (MAG/POWER/DB) and "
                                           "out_pitch_elems %% 4 == 0 (e.g. 516)");
    plan->bulk_store = enable ? 1 : 0;
    return KOPEK_OK;
}

kopek_status kopek_plan_set_band_scale(kopek_plan *plan, float scale) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    plan->band_scale = scale;
    return KOPEK_OK;
}

kopek_status kopek_plan_set_window(kopek_plan *plan, const float *window_host) {
    if (!plan || !window_host) return fail(KOPEK_ERR_INVALID, "NULL argument");
    if (plan->window != KOPEK_WINDOW_CUSTOM)
        return fail(KOPEK_ERR_INVALID, "plan was not created with KOPEK_WINDOW_CUSTOM");
    DeviceGuard g(plan->device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", plan->device);
    KOPEK_CUDA(cudaDeviceSynchronize());
    return upload_window(plan, window_host);
}

kopek_status kopek_plan_destroy(kopek_plan *plan) {
    if (!plan) return KOPEK_OK;
    DeviceGuard g(plan->device);
    cudaDeviceSynchronize();
    plan_free_device(plan);
    delete plan;
    return KOPEK_OK;
}

uint64_t kopek_plan_frames(const kopek_plan *plan, uint64_t n_samples) {
    if (!plan || n_samples < plan->n) return 0;
    return 1 + (n_samples - plan->n) / plan->hop;
}
uint64_t kopek_plan_bins(const kopek_plan *plan) { return plan ? plan->bins : 0; }
uint64_t kopek_plan_out_pitch(const kopek_plan *plan) { return plan ? plan->pitch : 0; }
uint64_t kopek_plan_out_elem_bytes(const kopek_plan *plan) { return plan ? plan->elem_bytes : 0; }
uint32_t kopek_plan_last_launches(const kopek_plan *plan) { return plan ? plan->last_launches : 0; }

static kopek_status stft_launch(const kopek_plan *plan, const float *d_samples, uint64_t frames, void *d_out,
                                cudaStream_t stream) {
    if (plan->d_k1w && (plan->n == 256 || plan->n == 512)) {
        // sub-warp kernels: HALF or FULL rows, any hop, any 4-byte aligned base
        const uint32_t need = plan->n == 256 ? 4 : 2;          // 128-bit (256) / 64-bit (512) vector loads
        const bool vecl = plan->hop % need == 0 && (uintptr_t)d_samples % (4 * need) == 0;
        const int mode = (plan->bins_mode == KOPEK_BINS_FULL ? 1 : 0) + (vecl ? 0 : 2);
        K1WParams wp{};
        wp.x = d_samples;
        wp.out = d_out;
        wp.tables = plan->d_k1w;
        wp.n_frames = frames;
        wp.hop = plan->hop;
        wp.pitch = plan->pitch;
        wp.band_scale = 1.0f;
        const uint64_t fpw = plan->n == 256 ? 4 : 2;              // frames per warp
        const uint64_t groups = (frames + fpw - 1) / fpw;
        const uint64_t warps = (uint64_t)plan->sm_count * plan->k1w_warps_per_sm;
        const int grid = (int)std::min<uint64_t>((groups + 3) / 4, warps / 4);
        const bool win = plan->window != KOPEK_WINDOW_RECT;
        cudaError_t e = plan->n == 256 ? k1w256_launch(wp, (int)plan->output, win, mode, grid, stream, nullptr)
                                       : k1w512_launch(wp, (int)plan->output, win, mode, grid, stream, nullptr);
        if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "K1W-%u launch failed: %s", plan->n, cudaGetErrorString(e));
        return KOPEK_OK;
    }
    {
        // team-per-frame kernels (2048 / 4096 points): HALF or FULL rows, any hop, any 4-byte aligned base, any window
        const bool vec = plan->hop % 2 == 0 && (uintptr_t)d_samples % 8 == 0;
        const int mode = (plan->bins_mode == KOPEK_BINS_FULL ? 1 : 0) + (vec ? 0 : 2);
        if (plan->d_k1m) {
            K1WParams wp{};
            wp.x = d_samples;
            wp.out = d_out;
            wp.tables = plan->d_k1m;
            wp.n_frames = frames;
            wp.hop = plan->hop;
            wp.pitch = plan->pitch;
            wp.band_scale = 1.0f;
            wp.vec = vec ? 1u : 0u;
            const int grid = (int)std::min<uint64_t>(frames, (uint64_t)plan->sm_count * plan->k1m_teams_per_sm);
            cudaError_t e = k1m_launch((int)plan->n, wp, (int)plan->output, plan->k1m_wink, mode, grid, stream, nullptr);
            if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "K1M-%u launch failed: %s", plan->n, cudaGetErrorString(e));
            return KOPEK_OK;
        }
    }
    if (plan->d_k1b && plan->n == 1024) {
        // every 1024-point mode runs on the two-exchange kernel: HALF or FULL rows, any hop, any 4-byte aligned base
        K1WParams wp{};
        wp.x = d_samples;
        wp.out = d_out;
        wp.tables = plan->d_k1b;
        wp.n_frames = frames;
        wp.hop = plan->hop;
        wp.pitch = plan->pitch;
        wp.band_scale = plan->band_scale;
        wp.vec = (plan->hop % 2 == 0 && (uintptr_t)d_samples % 8 == 0) ? 1u : 0u;
        const uint64_t warps_b = (uint64_t)plan->sm_count * plan->k1b_warps_per_sm;
        const int grid_b = (int)std::min<uint64_t>((frames + 3) / 4, warps_b / 4);
        const bool win = plan->window != KOPEK_WINDOW_RECT;
        cudaError_t eb;
        if (plan->bins_mode == KOPEK_BINS_FULL) {
            eb = k1b_launch_full(wp, (int)plan->output, win, grid_b, stream);
        } else {
            // bulk-store epilogue: opt-in, f32 spectra, 16-byte aligned rows (else the direct 32-bit stores)
            const bool bulk = plan->bulk_store && wp.vec && plan->output >= KOPEK_OUT_MAG && plan->output <= KOPEK_OUT_DB &&
                              plan->pitch % 4 == 0 && (uintptr_t)d_out % 16 == 0;
            eb = k1b_launch(wp, (int)plan->output, win, bulk, grid_b, stream, nullptr);
        }
        if (eb != cudaSuccess) return fail(KOPEK_ERR_CUDA, "K1W-1024b launch failed: %s", cudaGetErrorString(eb));
        return KOPEK_OK;
    }
    K1Params kp;
    kp.x = d_samples;
    kp.out = d_out;
    kp.win = plan->d_win;
    kp.tw = plan->d_tw;
    kp.pw = plan->d_pw;
    kp.n_frames = frames;
    kp.hop = plan->hop;
    kp.pitch = plan->pitch;
    kp.full = 0;
    kp.vec = 0;
    const bool vec = (plan->hop % 2 == 0) && ((uintptr_t)d_samples % 8 == 0);
    const uint64_t groups = (frames + plan->info.fpb - 1) / plan->info.fpb;
    const uint64_t resident = (uint64_t)plan->sm_count * plan->occupancy;
    const int grid = (int)std::min<uint64_t>(groups, resident);
    cudaError_t e = plan->launch(kp, (int)plan->output, plan->window != KOPEK_WINDOW_RECT,
                                 plan->bins_mode == KOPEK_BINS_FULL, vec, grid, stream);
    if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "K1 launch (n=%u) failed: %s", plan->n, cudaGetErrorString(e));
    return KOPEK_OK;
}

kopek_status kopek_stft_exec(const kopek_plan *plan, const float *d_samples, uint64_t n_samples, void *d_out,
                             uint64_t *n_frames_out, void *cuda_stream) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    const uint64_t frames = kopek_plan_frames(plan, n_samples);
    if (n_frames_out) *n_frames_out = frames;
    plan->last_launches = 0;
    if (frames == 0) return KOPEK_OK;
    if (!d_samples || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if ((uintptr_t)d_samples % 4 != 0 || (uintptr_t)d_out % plan->elem_bytes != 0)
        return fail(KOPEK_ERR_INVALID, "misaligned device pointer");
    DeviceGuard g(plan->device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", plan->device);
    kopek_status st = stft_launch(plan, d_samples, frames, d_out, (cudaStream_t)cuda_stream);
    if (st == KOPEK_OK) plan->last_launches = 1;
    return st;
}

// decoded PCM in, spectra out: interleaved stereo i16 frames -> frame[channel] / i16::MAX -> plan.
// n = 1024 (the frame size of the reference's callers) fuses the marshalling into the kernel's load stage; other
// geometries convert into a stream-ordered scratch buffer first (kopek_pcm16_to_f32) and run the plan on it.
kopek_status kopek_stft_exec_pcm16(const kopek_plan *plan, const int16_t *d_frames, uint64_t n_frames, uint32_t channel,
                                   void *d_out, uint64_t *n_frames_out, void *cuda_stream) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    const uint64_t frames = kopek_plan_frames(plan, n_frames);       // one sample per PCM frame
    if (n_frames_out) *n_frames_out = frames;
    plan->last_launches = 0;
    if (frames == 0) return KOPEK_OK;
    if (!d_frames || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if (channel > 1) return fail(KOPEK_ERR_INVALID, "channel %u of a stereo frame", channel);
    if ((uintptr_t)d_frames % 4 != 0 || (uintptr_t)d_out % plan->elem_bytes != 0)
        return fail(KOPEK_ERR_INVALID, "misaligned device pointer");
    DeviceGuard g(plan->device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", plan->device);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    if (plan->d_k1b && plan->n == 1024 && plan->bins_mode == KOPEK_BINS_HALF && plan->hop % 4 == 0 &&
        (uintptr_t)d_frames % 16 == 0) {
        K1WParams wp{};
        wp.x = reinterpret_cast<const float *>(d_frames);              // one 32-bit word per stereo frame
        wp.out = d_out;
        wp.tables = plan->d_k1b;
        wp.n_frames = frames;
        wp.hop = plan->hop;
        wp.pitch = plan->pitch;
        wp.band_scale = plan->band_scale;
        wp.pcm_shift = channel ? 0u : 16u;
        const uint64_t warps_b = (uint64_t)plan->sm_count * plan->k1b_warps_per_sm;
        const int grid_b = (int)std::min<uint64_t>((frames + 3) / 4, warps_b / 4);
        cudaError_t e = k1b_launch_pcm(wp, (int)plan->output, plan->window != KOPEK_WINDOW_RECT, grid_b, stream);
        if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "K1W-1024b (PCM) launch failed: %s", cudaGetErrorString(e));
        plan->last_launches = 1;
        return KOPEK_OK;
    }
    const uint64_t n_used = (frames - 1) * plan->hop + plan->n;
    float *scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, n_used * sizeof(float), stream);
    if (e != cudaSuccess) return fail(KOPEK_ERR_NOMEM, "PCM scratch of %llu samples: %s", (unsigned long long)n_used,
                                      cudaGetErrorString(e));
    kopek_status st = kopek_pcm16_to_f32(d_frames, n_used, 2, channel, 32767.0, scratch, plan->device, cuda_stream);
    if (st == KOPEK_OK) st = stft_launch(plan, scratch, frames, d_out, stream);
    cudaFreeAsync(scratch, stream);
    if (st == KOPEK_OK) plan->last_launches = 2;
    return st;
}

kopek_status kopek_stft_exec_host(kopek_plan *plan, const float *h_samples, uint64_t n_samples, void *h_out,
                                  uint64_t *n_frames_out) {
    if (!plan) return fail(KOPEK_ERR_INVALID, "plan is NULL");
    const uint64_t frames = kopek_plan_frames(plan, n_samples);
    if (n_frames_out) *n_frames_out = frames;
    plan->last_launches = 0;
    if (frames == 0) return KOPEK_OK;
    if (!h_samples || !h_out) return fail(KOPEK_ERR_INVALID, "NULL host pointer");
    std::lock_guard<std::mutex> lk(plan->mu);
    DeviceGuard g(plan->device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", plan->device);

    // chunk so that one slot moves ~KOPEK_HOST_CHUNK_MB MiB over its busier direction (samples in, or spectra out when
    // frames overlap: hop = 1, n = 4096 writes 8 KB per 4 bytes read); three slots keep the H2D engine, the SMs and
    // the D2H engine busy at the same time.  The first upload and the last download are not overlapped with anything,
    // so the chunk is kept small against the call (measured on B200, tools/sweep_host_chunk.py).
    const uint64_t n = plan->n, hop = plan->hop;
    uint64_t chunk_mb = K1_HOST_CHUNK_MB;
    if (const char *env = getenv("KOPEK_HOST_CHUNK_MB")) chunk_mb = std::max(1, atoi(env));
    const uint64_t per_frame = std::max<uint64_t>(hop * sizeof(float), plan->pitch * plan->elem_bytes);
    uint64_t cf = std::max<uint64_t>(1, (chunk_mb << 20) / per_frame);
    cf = std::min(cf, frames);
    const uint64_t in_elems = (cf - 1) * hop + n;
    const uint64_t out_bytes = cf * plan->pitch * plan->elem_bytes;
    if (plan->slot_frames < cf) {
        cudaError_t e = cudaSuccess;
        for (HostSlot &s : plan->slots) {
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            s.d_in = nullptr; s.d_out = nullptr;
        }
        plan->slot_frames = 0;
        for (HostSlot &s : plan->slots) {
            if (!s.stream && (e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) break;
            if ((e = cudaMalloc(&s.d_in, in_elems * sizeof(float))) != cudaSuccess) break;
            if ((e = cudaMalloc(&s.d_out, out_bytes)) != cudaSuccess) break;
        }
        if (e != cudaSuccess) {          // all or nothing: a later call must not find half-sized slots
            for (HostSlot &s : plan->slots) {
                if (s.d_in) cudaFree(s.d_in);
                if (s.d_out) cudaFree(s.d_out);
                s.d_in = nullptr; s.d_out = nullptr;
            }
            cudaGetLastError();
            return fail(KOPEK_ERR_NOMEM, "chunk slots (%llu frames each): %s", (unsigned long long)cf, cudaGetErrorString(e));
        }
        plan->slot_frames = cf;
    }
    // fault injection for tests/test_gpu_guards.py: fail the enqueue of chunk K (0-based) as a launch failure would
    long fail_chunk = -1;
    if (const char *env = getenv("KOPEK_TEST_FAIL_CHUNK")) fail_chunk = atol(env);
    uint32_t launches = 0;
    uint64_t c = 0;
    kopek_status st = KOPEK_OK;
    for (uint64_t f0 = 0; f0 < frames && st == KOPEK_OK; f0 += cf, ++c) {
        HostSlot &s = plan->slots[c % 3];
        const uint64_t nf = std::min(cf, frames - f0);
        const uint64_t ne = (nf - 1) * hop + n;
        cudaError_t e = cudaMemcpyAsync(s.d_in, h_samples + f0 * hop, ne * sizeof(float), cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) { st = fail(KOPEK_ERR_CUDA, "chunk %llu upload: %s", (unsigned long long)c, cudaGetErrorString(e)); break; }
        if ((long)c == fail_chunk) { st = fail(KOPEK_ERR_CUDA, "chunk %llu: injected launch failure", (unsigned long long)c); break; }
        st = stft_launch(plan, s.d_in, nf, s.d_out, s.stream);
        if (st != KOPEK_OK) break;
        ++launches;
        // rows are contiguous (pitch) so one copy returns the chunk; the tail of the
        // last row beyond `bins` is unspecified padding when pitch > bins
        const uint64_t ob = ((nf - 1) * plan->pitch + plan->bins) * plan->elem_bytes;
        e = cudaMemcpyAsync((char *)h_out + f0 * plan->pitch * plan->elem_bytes, s.d_out, ob, cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) st = fail(KOPEK_ERR_CUDA, "chunk %llu download: %s", (unsigned long long)c, cudaGetErrorString(e));
    }
    // drain EVERY slot stream, also on the error path: chunks enqueued before the failure are still copying into the
    // caller's h_out, which the caller is free to release as soon as this call returns
    for (HostSlot &s : plan->slots) {
        if (!s.stream) continue;
        cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && st == KOPEK_OK) st = fail(KOPEK_ERR_CUDA, "chunk pipeline: %s", cudaGetErrorString(e));
    }
    plan->last_launches = launches;
    return st;
}

// ------------------------------------------------------------- live stream ----
// The beep view's sliding VecDeque of the last n samples (examples/beep/view.rs:44,57-60).  A push is a latency
// path (a few KB per UI frame), so nothing is copied by DMA: the window and the spectrum live in MAPPED pinned
// host memory, the plan's kernel reads the n samples over PCIe (in order) and stores the spectrum straight back --
// one launch and one stream synchronisation per push.  The ring itself is ordinary host memory; every push
// linearises the latest n samples into the mapped, 256-byte aligned window (two memcpys, <= 16 KB), so the kernel
// always sees an aligned frame: the same (fast) kernel, the same rounding, whatever the push sizes add up to.
struct kopek_stream {
    kopek_plan *plan;
    uint32_t n;
    int device;
    uint64_t written;          // total samples pushed so far
    float *h_ring;             // host ring of n floats: sample i lives at i % n
    float *h_win;              // mapped pinned, n floats: the latest n samples in order (rebuilt by every push)
    float *d_win;              // the device's address of h_win
    void *h_out;               // mapped pinned, one spectrum
    void *d_out;               // the device's address of h_out
    cudaStream_t cs;
    std::mutex mu;
};

kopek_status kopek_stream_create(kopek_stream **stream, uint32_t n, uint32_t window, uint32_t output,
                                 float input_scale, int device) {
    if (!stream) return fail(KOPEK_ERR_INVALID, "stream is NULL");
    *stream = nullptr;
    if (window > KOPEK_WINDOW_HAMMING) return fail(KOPEK_ERR_INVALID, "stream window must be RECT, HANN or HAMMING");
    // the scale rides on the window table (exactly one f32 product per sample, like w[i] * x[i])
    const bool scaled = input_scale != 1.0f;
    kopek_plan *plan = nullptr;
    kopek_status st = kopek_plan_create(&plan, n, n, scaled ? (uint32_t)KOPEK_WINDOW_CUSTOM : window, output,
                                        KOPEK_BINS_HALF, 0, device);
    if (st != KOPEK_OK) return st;
    if (scaled) {
        std::vector<float> w;
        host_window(window, n, w);
        for (float &v : w) v = (float)((double)v * (double)input_scale);
        st = kopek_plan_set_window(plan, w.data());
        if (st != KOPEK_OK) { kopek_plan_destroy(plan); return st; }
    }
    kopek_stream *s = new (std::nothrow) kopek_stream();
    if (!s) { kopek_plan_destroy(plan); return fail(KOPEK_ERR_NOMEM, "out of host memory"); }
    s->plan = plan; s->n = n; s->device = device; s->written = 0;
    s->h_ring = nullptr; s->h_win = nullptr; s->d_win = nullptr; s->h_out = nullptr; s->d_out = nullptr; s->cs = nullptr;
    DeviceGuard g(device);
    const size_t ob = plan->bins * plan->elem_bytes;
    cudaError_t e;
    s->h_ring = static_cast<float *>(calloc(n, sizeof(float)));      // the VecDeque starts as n zeros
    if (!s->h_ring) { kopek_stream_destroy(s); return fail(KOPEK_ERR_NOMEM, "out of host memory"); }
    if ((e = cudaHostAlloc(&s->h_win, n * sizeof(float), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(&s->d_win, s->h_win, 0)) != cudaSuccess ||
        (e = cudaHostAlloc(&s->h_out, ob, cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(&s->d_out, s->h_out, 0)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&s->cs, cudaStreamNonBlocking)) != cudaSuccess) {
        kopek_stream_destroy(s);
        return fail(KOPEK_ERR_CUDA, "stream setup failed: %s", cudaGetErrorString(e));
    }
    *stream = s;
    return KOPEK_OK;
}

uint64_t kopek_stream_bins(const kopek_stream *stream) { return stream ? stream->plan->bins : 0; }

kopek_status kopek_stream_push(kopek_stream *s, const float *h_samples, uint32_t count, void *h_out) {
    if (!s || !h_out || (count && !h_samples)) return fail(KOPEK_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard g(s->device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", s->device);
    const uint32_t n = s->n;
    if (count > n) {                       // only the last n samples can survive in the window
        s->written += count - n;
        h_samples += count - n;
        count = n;
    }
    if (count) {
        uint32_t pos = (uint32_t)(s->written % n), done = 0;
        while (done < count) {             // at most two segments (wrap)
            uint32_t seg = std::min(count - done, n - pos);
            memcpy(s->h_ring + pos, h_samples + done, seg * sizeof(float));
            done += seg;
            pos = (pos + seg) % n;
        }
        s->written += count;
    }
    // the most recent n samples start at (written % n) in the ring: linearise them into the mapped window; the launch
    // orders the host's writes before the kernel's reads, the synchronisation the kernel's stores before the host's reads
    const uint32_t head = (uint32_t)(s->written % n);
    memcpy(s->h_win, s->h_ring + head, (n - head) * sizeof(float));
    if (head) memcpy(s->h_win + (n - head), s->h_ring, head * sizeof(float));
    const float *frame = s->d_win;
    uint64_t nf = 0;
    kopek_status st = kopek_stft_exec(s->plan, frame, n, s->d_out, &nf, s->cs);
    if (st != KOPEK_OK) return st;
    KOPEK_CUDA(cudaStreamSynchronize(s->cs));
    memcpy(h_out, s->h_out, s->plan->bins * s->plan->elem_bytes);
    return KOPEK_OK;
}

kopek_status kopek_stream_destroy(kopek_stream *s) {
    if (!s) return KOPEK_OK;
    DeviceGuard g(s->device);
    if (s->cs) { cudaStreamSynchronize(s->cs); cudaStreamDestroy(s->cs); }
    free(s->h_ring);
    if (s->h_win) cudaFreeHost(s->h_win);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->plan) kopek_plan_destroy(s->plan);
    delete s;
    return KOPEK_OK;
}

// ------------------------------------------------------- peer-direct output ----
kopek_status kopek_device_alloc(void **d_ptr, size_t bytes, int device) {
    if (!d_ptr) return fail(KOPEK_ERR_INVALID, "d_ptr is NULL");
    *d_ptr = nullptr;
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(KOPEK_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return KOPEK_OK;
}
kopek_status kopek_device_free(void *d_ptr, int device) {
    if (!d_ptr) return KOPEK_OK;
    DeviceGuard g(device);
    KOPEK_CUDA(cudaFree(d_ptr));
    return KOPEK_OK;
}
kopek_status kopek_copy(void *dst, const void *src, size_t bytes, int device) {
    if (!bytes) return KOPEK_OK;
    if (!dst || !src) return fail(KOPEK_ERR_INVALID, "NULL pointer");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    KOPEK_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return KOPEK_OK;
}
kopek_status kopek_ipc_export(const void *d_ptr, unsigned char *handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == KOPEK_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_ptr || !handle_out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    KOPEK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle_out, &h, sizeof h);
    return KOPEK_OK;
}
kopek_status kopek_ipc_open(const unsigned char *handle, int device, void **d_ptr) {
    if (!handle || !d_ptr) return fail(KOPEK_ERR_INVALID, "NULL argument");
    *d_ptr = nullptr;
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    KOPEK_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KOPEK_OK;
}
kopek_status kopek_ipc_close(void *d_ptr, int device) {
    if (!d_ptr) return KOPEK_OK;
    DeviceGuard g(device);
    KOPEK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return KOPEK_OK;
}

kopek_status kopek_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(KOPEK_ERR_INVALID, "ptr is NULL");
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(KOPEK_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return KOPEK_OK;
}
kopek_status kopek_host_free(void *ptr) {
    if (!ptr) return KOPEK_OK;
    KOPEK_CUDA(cudaFreeHost(ptr));
    return KOPEK_OK;
}

}  // extern "C"
