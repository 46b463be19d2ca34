// Multi-GPU entry points of the C ABI (include/kopek_cuda.h section 4): ONE process drives
// R devices.  Frames are independent (the reference's fft is a pure function, src/fft.rs:6), so the transform has no
// data-path collective: device r owns the contiguous frame range [floor(r F / R), floor((r+1) F / R)) and reads the
// sample range that covers it (the (n - hop)-sample halo between neighbours is simply read twice).  Spectra stay
// sharded unless the caller asks for ONE spectrogram on devices[0]; then there are three ways to get it there:
//   kopek_mgpu_exec_gathered   every device's kernel stores its rows straight into devices[0]'s buffer through the
//                              peer mapping -- transform and transfer are one kernel (the 2048-byte bulk-store epilogue
//                              when the plan allows it: n = 1024, f32 spectra, pitch % 4 == 0)
//   kopek_mgpu_gather (COPY)   copy-engine peer copies of finished shards, pushed by the source device's stream
//   kopek_mgpu_gather (NCCL)   one grouped ncclSend/ncclRecv (NCCL has no native gatherv); libnccl.so.2 is opened at
//                              run time so that the library itself depends on nothing but the CUDA driver
// No torch, no Python: this is what a Rust host (rust/kopek/src/fft.rs::stft_magnitude_multi) links against.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace kopek {

// ---- NCCL through dlopen: the seven entry points a grouped send/recv gather needs (nccl.h 2.27: ncclUint8 = 1) ----
struct NcclApi {
    typedef int (*init_all_fn)(void **comms, int ndev, const int *devlist);
    typedef int (*destroy_fn)(void *comm);
    typedef int (*group_fn)();
    typedef int (*send_fn)(const void *buf, size_t count, int dtype, int peer, void *comm, cudaStream_t stream);
    typedef int (*recv_fn)(void *buf, size_t count, int dtype, int peer, void *comm, cudaStream_t stream);
    typedef const char *(*errstr_fn)(int);
    void *handle = nullptr;
    init_all_fn init_all = nullptr;
    destroy_fn destroy = nullptr;
    group_fn group_start = nullptr, group_end = nullptr;
    send_fn send = nullptr;
    recv_fn recv = nullptr;
    errstr_fn errstr = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("KOPEK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.init_all = (NcclApi::init_all_fn)dlsym(api.handle, "ncclCommInitAll");
        api.destroy = (NcclApi::destroy_fn)dlsym(api.handle, "ncclCommDestroy");
        api.group_start = (NcclApi::group_fn)dlsym(api.handle, "ncclGroupStart");
        api.group_end = (NcclApi::group_fn)dlsym(api.handle, "ncclGroupEnd");
        api.send = (NcclApi::send_fn)dlsym(api.handle, "ncclSend");
        api.recv = (NcclApi::recv_fn)dlsym(api.handle, "ncclRecv");
        api.errstr = (NcclApi::errstr_fn)dlsym(api.handle, "ncclGetErrorString");
        api.ok = api.init_all && api.destroy && api.group_start && api.group_end && api.send && api.recv && api.errstr;
    });
    return api;
}

struct MgpuDev {
    int device = -1;
    kopek_plan *plan = nullptr;          // the device's own plan (tightly packed or caller's pitch)
    kopek_plan *plan_bulk = nullptr;     // same plan with the bulk-store epilogue, when the geometry allows it
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool peer_to_root = false;           // this device can store into devices[0]'s memory
    void *nccl = nullptr;
};

}  // namespace kopek

using namespace kopek;

struct kopek_mgpu {
    std::vector<MgpuDev> devs;
    uint32_t n, hop, window, output, bins_mode;
    uint64_t bins, pitch, elem_bytes;
    bool window_open = false;            // ev_begin recorded, waiting for kopek_mgpu_sync
    bool nccl_ready = false;
    std::mutex mu;
};

namespace kopek {

static void frame_range(uint32_t rank, uint32_t world, uint64_t frames, uint64_t *f0, uint64_t *f1) {
    // 128-bit products: rank * frames overflows 64 bits only beyond 2^58 frames, but be exact anyway
    *f0 = (uint64_t)(((unsigned __int128)rank * frames) / world);
    *f1 = (uint64_t)(((unsigned __int128)(rank + 1) * frames) / world);
}

static uint64_t total_frames(const kopek_mgpu *c, uint64_t n_samples) {
    return n_samples < c->n ? 0 : 1 + (n_samples - c->n) / c->hop;
}

static kopek_status open_window(kopek_mgpu *c) {
    if (c->window_open) return KOPEK_OK;
    for (MgpuDev &d : c->devs) {
        DeviceGuard g(d.device);
        if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", d.device);
        KOPEK_CUDA(cudaEventRecord(d.ev_begin, d.stream));
    }
    c->window_open = true;
    return KOPEK_OK;
}

static kopek_status close_enqueue(kopek_mgpu *c) {
    for (MgpuDev &d : c->devs) {
        DeviceGuard g(d.device);
        if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", d.device);
        KOPEK_CUDA(cudaEventRecord(d.ev_end, d.stream));
    }
    return KOPEK_OK;
}

// An entry point failed after open_window: the devices that did get work keep running, so the end events are still
// recorded (best effort) -- a later kopek_mgpu_sync then drains the streams and finds both events of every device --
// and the first error is what the caller sees.
static kopek_status fail_in_window(kopek_mgpu *c, kopek_status st) {
    const std::string msg(kopek_last_error());
    close_enqueue(c);
    return fail(st, "%s", msg.c_str());
}

static void mgpu_free(kopek_mgpu *c) {
    NcclApi &nc = nccl_api();
    for (MgpuDev &d : c->devs) {
        if (d.device < 0) continue;
        DeviceGuard g(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.nccl && nc.ok) nc.destroy(d.nccl);
        if (d.plan) kopek_plan_destroy(d.plan);
        if (d.plan_bulk) kopek_plan_destroy(d.plan_bulk);
        if (d.ev_begin) cudaEventDestroy(d.ev_begin);
        if (d.ev_end) cudaEventDestroy(d.ev_end);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
}

static kopek_status ensure_nccl(kopek_mgpu *c) {
    if (c->nccl_ready) return KOPEK_OK;
    NcclApi &nc = nccl_api();
    if (!nc.ok)
        return fail(KOPEK_ERR_UNSUPPORTED, "NCCL gather: %s", nc.handle ? "libnccl.so.2 lacks the send/recv entry points"
                                                                       : "libnccl.so.2 could not be opened (set KOPEK_NCCL_LIB)");
    const int R = (int)c->devs.size();
    std::vector<void *> comms(R, nullptr);
    std::vector<int> list(R);
    for (int r = 0; r < R; ++r) list[r] = c->devs[r].device;
    int rc = nc.init_all(comms.data(), R, list.data());
    if (rc != 0) return fail(KOPEK_ERR_CUDA, "ncclCommInitAll over %d devices: %s", R, nc.errstr(rc));
    for (int r = 0; r < R; ++r) c->devs[r].nccl = comms[r];
    c->nccl_ready = true;
    return KOPEK_OK;
}

}  // namespace kopek

extern "C" {

kopek_status kopek_mgpu_create(kopek_mgpu **ctx, const int32_t *devices, uint32_t n_devices, uint32_t n, uint32_t hop,
                               uint32_t window, uint32_t output, uint32_t bins, uint64_t out_pitch_elems) {
    if (!ctx) return fail(KOPEK_ERR_INVALID, "ctx is NULL");
    *ctx = nullptr;
    if (!devices || n_devices == 0) return fail(KOPEK_ERR_INVALID, "empty device list");
    if (n_devices > 64) return fail(KOPEK_ERR_INVALID, "at most 64 devices");
    for (uint32_t a = 0; a < n_devices; ++a)
        for (uint32_t b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return fail(KOPEK_ERR_INVALID, "device %d listed twice", devices[a]);
    kopek_mgpu *c = new (std::nothrow) kopek_mgpu();
    if (!c) return fail(KOPEK_ERR_NOMEM, "out of host memory");
    c->n = n; c->hop = hop; c->window = window; c->output = output; c->bins_mode = bins;
    c->devs.resize(n_devices);
    kopek_status st = KOPEK_OK;
    for (uint32_t r = 0; r < n_devices && st == KOPEK_OK; ++r) {
        MgpuDev &d = c->devs[r];
        st = kopek_plan_create(&d.plan, n, hop, window, output, bins, out_pitch_elems, devices[r]);
        if (st != KOPEK_OK) break;
        d.device = devices[r];
        DeviceGuard g(d.device);
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreate(&d.ev_begin)) != cudaSuccess || (e = cudaEventCreate(&d.ev_end)) != cudaSuccess) {
            st = fail(KOPEK_ERR_CUDA, "device %d: stream/event setup failed: %s", d.device, cudaGetErrorString(e));
            break;
        }
        if (r == 0) {
            c->bins = kopek_plan_bins(d.plan);
            c->pitch = kopek_plan_out_pitch(d.plan);
            c->elem_bytes = kopek_plan_out_elem_bytes(d.plan);
            d.peer_to_root = true;
        } else {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d.device, devices[0]) == cudaSuccess && can) {
                e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) d.peer_to_root = true;
                cudaGetLastError();
            }
        }
        // the bulk-store twin: one 2048-byte TMA store per frame is what reaches link bandwidth over NVLink
        if (n == 1024 && bins == KOPEK_BINS_HALF && output >= KOPEK_OUT_MAG && output <= KOPEK_OUT_DB &&
            c->pitch % 4 == 0 && hop % 4 == 0) {
            st = kopek_plan_create(&d.plan_bulk, n, hop, window, output, bins, c->pitch, d.device);
            if (st == KOPEK_OK) st = kopek_plan_set_bulk_store(d.plan_bulk, 1);
        }
    }
    if (st != KOPEK_OK) {
        std::string msg(kopek_last_error());
        mgpu_free(c);
        return fail(st, "%s", msg.c_str());
    }
    *ctx = c;
    return KOPEK_OK;
}

kopek_status kopek_mgpu_destroy(kopek_mgpu *ctx) {
    if (ctx) mgpu_free(ctx);
    return KOPEK_OK;
}

uint32_t kopek_mgpu_device_count(const kopek_mgpu *ctx) { return ctx ? (uint32_t)ctx->devs.size() : 0; }
uint64_t kopek_mgpu_bins(const kopek_mgpu *ctx) { return ctx ? ctx->bins : 0; }
uint64_t kopek_mgpu_out_pitch(const kopek_mgpu *ctx) { return ctx ? ctx->pitch : 0; }
uint64_t kopek_mgpu_out_elem_bytes(const kopek_mgpu *ctx) { return ctx ? ctx->elem_bytes : 0; }

kopek_status kopek_mgpu_set_custom_window(kopek_mgpu *ctx, const float *window_host) {
    if (!ctx) return fail(KOPEK_ERR_INVALID, "ctx is NULL");
    for (MgpuDev &d : ctx->devs) {
        kopek_status st = kopek_plan_set_window(d.plan, window_host);
        if (st == KOPEK_OK && d.plan_bulk) st = kopek_plan_set_window(d.plan_bulk, window_host);
        if (st != KOPEK_OK) return st;
    }
    return KOPEK_OK;
}

kopek_status kopek_mgpu_shard(const kopek_mgpu *ctx, uint64_t n_samples, uint32_t rank, uint64_t *sample_begin,
                              uint64_t *sample_count, uint64_t *frame_begin, uint64_t *frame_count) {
    if (!ctx) return fail(KOPEK_ERR_INVALID, "ctx is NULL");
    if (rank >= ctx->devs.size()) return fail(KOPEK_ERR_INVALID, "rank %u of %zu devices", rank, ctx->devs.size());
    uint64_t f0, f1;
    frame_range(rank, (uint32_t)ctx->devs.size(), total_frames(ctx, n_samples), &f0, &f1);
    const bool empty = f1 <= f0;
    if (sample_begin) *sample_begin = empty ? 0 : f0 * ctx->hop;
    if (sample_count) *sample_count = empty ? 0 : (f1 - 1 - f0) * ctx->hop + ctx->n;
    if (frame_begin) *frame_begin = f0;
    if (frame_count) *frame_count = f1 - f0;
    return KOPEK_OK;
}

kopek_status kopek_mgpu_exec(kopek_mgpu *ctx, const float *const *d_samples, const uint64_t *n_samples,
                             void *const *d_out, uint64_t *n_frames_out) {
    if (!ctx || !d_samples || !n_samples || !d_out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    kopek_status st = open_window(ctx);
    if (st != KOPEK_OK) return st;
    for (size_t r = 0; r < ctx->devs.size(); ++r) {
        MgpuDev &d = ctx->devs[r];
        uint64_t nf = 0;
        st = kopek_stft_exec(d.plan, d_samples[r], n_samples[r], d_out[r], &nf, d.stream);
        if (n_frames_out) n_frames_out[r] = nf;
        if (st != KOPEK_OK) return fail_in_window(ctx, st);
    }
    return close_enqueue(ctx);
}

kopek_status kopek_mgpu_exec_gathered(kopek_mgpu *ctx, const float *const *d_samples, const uint64_t *n_samples,
                                      void *d_dst, uint64_t *n_frames_total) {
    if (!ctx || !d_samples || !n_samples || !d_dst) return fail(KOPEK_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (const MgpuDev &d : ctx->devs)
        if (!d.peer_to_root)
            return fail(KOPEK_ERR_UNSUPPORTED, "device %d has no peer access to device %d: use kopek_mgpu_exec + "
                                               "kopek_mgpu_gather", d.device, ctx->devs[0].device);
    kopek_status st = open_window(ctx);
    if (st != KOPEK_OK) return st;
    uint64_t row = 0;
    for (size_t r = 0; r < ctx->devs.size(); ++r) {
        MgpuDev &d = ctx->devs[r];
        char *dst = static_cast<char *>(d_dst) + row * ctx->pitch * ctx->elem_bytes;
        const bool bulk = d.plan_bulk && (uintptr_t)dst % 16 == 0;
        uint64_t nf = 0;
        st = kopek_stft_exec(bulk ? d.plan_bulk : d.plan, d_samples[r], n_samples[r], dst, &nf, d.stream);
        if (st != KOPEK_OK) return fail_in_window(ctx, st);
        row += nf;
    }
    if (n_frames_total) *n_frames_total = row;
    return close_enqueue(ctx);
}

kopek_status kopek_mgpu_gather(kopek_mgpu *ctx, void *const *d_out, const uint64_t *n_frames, void *d_dst,
                               uint32_t mode) {
    if (!ctx || !d_out || !n_frames || !d_dst) return fail(KOPEK_ERR_INVALID, "NULL argument");
    if (mode > KOPEK_GATHER_NCCL) return fail(KOPEK_ERR_INVALID, "unknown gather mode %u", mode);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t R = ctx->devs.size();
    const uint64_t row_bytes = ctx->pitch * ctx->elem_bytes;
    kopek_status st = KOPEK_OK;
    if (mode == KOPEK_GATHER_NCCL && R > 1 && (st = ensure_nccl(ctx)) != KOPEK_OK) return st;
    if ((st = open_window(ctx)) != KOPEK_OK) return st;
    std::vector<uint64_t> off(R + 1, 0);
    for (size_t r = 0; r < R; ++r) off[r + 1] = off[r] + n_frames[r] * row_bytes;
    MgpuDev &root = ctx->devs[0];
    if (n_frames[0] && d_out[0] != d_dst) {
        DeviceGuard g(root.device);
        cudaError_t e = cudaMemcpyAsync(d_dst, d_out[0], off[1], cudaMemcpyDeviceToDevice, root.stream);
        if (e != cudaSuccess)
            return fail_in_window(ctx, fail(KOPEK_ERR_CUDA, "gather: device %d's own rows: %s", root.device, cudaGetErrorString(e)));
    }
    if (mode == KOPEK_GATHER_COPY) {
        // pushed by the source device's stream: ordered after that device's transform, all links busy at once
        for (size_t r = 1; r < R; ++r) {
            if (!n_frames[r]) continue;
            MgpuDev &d = ctx->devs[r];
            DeviceGuard g(d.device);
            cudaError_t e = cudaMemcpyPeerAsync(static_cast<char *>(d_dst) + off[r], root.device, d_out[r], d.device,
                                                off[r + 1] - off[r], d.stream);
            if (e != cudaSuccess)
                return fail_in_window(ctx, fail(KOPEK_ERR_CUDA, "gather: peer copy %d -> %d: %s", d.device, root.device,
                                                cudaGetErrorString(e)));
        }
    } else if (R > 1) {
        NcclApi &nc = nccl_api();
        int rc = nc.group_start();
        for (size_t r = 1; r < R && rc == 0; ++r) {
            if (!n_frames[r]) continue;
            const size_t bytes = off[r + 1] - off[r];
            rc = nc.send(d_out[r], bytes, 1 /* ncclUint8 */, 0, ctx->devs[r].nccl, ctx->devs[r].stream);
            if (rc == 0) rc = nc.recv(static_cast<char *>(d_dst) + off[r], bytes, 1, (int)r, root.nccl, root.stream);
        }
        const int rc2 = nc.group_end();
        if (rc != 0 || rc2 != 0)
            return fail_in_window(ctx, fail(KOPEK_ERR_CUDA, "NCCL gather: %s", nc.errstr(rc ? rc : rc2)));
    }
    return close_enqueue(ctx);
}

kopek_status kopek_mgpu_sync(kopek_mgpu *ctx, float *elapsed_ms_max) {
    if (!ctx) return fail(KOPEK_ERR_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    float worst = 0.0f;
    for (MgpuDev &d : ctx->devs) {
        DeviceGuard g(d.device);
        if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", d.device);
        KOPEK_CUDA(cudaStreamSynchronize(d.stream));
        if (ctx->window_open) {
            float ms = 0.0f;
            // after a failed entry point an end event may be missing: no time for that device, not an error of sync
            if (cudaEventElapsedTime(&ms, d.ev_begin, d.ev_end) == cudaSuccess) worst = std::max(worst, ms);
            else cudaGetLastError();
        }
    }
    ctx->window_open = false;
    if (elapsed_ms_max) *elapsed_ms_max = worst;
    return KOPEK_OK;
}

// One host buffer in, one host spectrogram out: the library splits the frame range, one host thread per device runs
// that device's chunked H2D -> kernel -> D2H pipeline (kopek_stft_exec_host) on its slice; rows land at their final
// place in h_out, so there is nothing to gather.
kopek_status kopek_mgpu_exec_host(kopek_mgpu *ctx, const float *h_samples, uint64_t n_samples, void *h_out,
                                  uint64_t *n_frames_out) {
    if (!ctx) return fail(KOPEK_ERR_INVALID, "ctx is NULL");
    const uint64_t frames = total_frames(ctx, n_samples);
    if (n_frames_out) *n_frames_out = frames;
    if (frames == 0) return KOPEK_OK;
    if (!h_samples || !h_out) return fail(KOPEK_ERR_INVALID, "NULL host pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const uint32_t R = (uint32_t)ctx->devs.size();
    std::vector<kopek_status> status(R, KOPEK_OK);
    std::vector<std::string> message(R);
    auto work = [&](uint32_t r) {
        uint64_t f0, f1;
        frame_range(r, R, frames, &f0, &f1);
        if (f1 <= f0) return;
        const uint64_t s0 = f0 * ctx->hop, cnt = (f1 - 1 - f0) * ctx->hop + ctx->n;
        uint64_t got = 0;
        status[r] = kopek_stft_exec_host(ctx->devs[r].plan, h_samples + s0, cnt,
                                         static_cast<char *>(h_out) + f0 * ctx->pitch * ctx->elem_bytes, &got);
        if (status[r] == KOPEK_OK && got != f1 - f0) status[r] = KOPEK_ERR_INVALID;
        if (status[r] != KOPEK_OK) message[r] = kopek_last_error();      // thread-local: copy before the thread ends
    };
    std::vector<std::thread> pool;
    try {
        for (uint32_t r = 1; r < R; ++r) pool.emplace_back(work, r);
    } catch (...) {
        for (std::thread &t : pool) t.join();
        return fail(KOPEK_ERR_NOMEM, "cannot start %u host threads", R - 1);
    }
    work(0);
    for (std::thread &t : pool) t.join();
    for (uint32_t r = 0; r < R; ++r)
        if (status[r] != KOPEK_OK) return fail(status[r], "device %d: %s", ctx->devs[r].device, message[r].c_str());
    return KOPEK_OK;
}

}  // extern "C"
