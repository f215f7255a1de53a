// batch.cpp -- whole-file batches over one or several GPUs (include/flacb200_stream.h: flacb200_encode_batch,
// flacb200_decode_batch): the file-level fan-out of the reference's rayon examples (examples/flac2wav.rs:31-38,
// examples/flac-split.rs:84-87) behind the C ABI.
//
// One host thread + one engine per device.  A device's share of the streams is cut into sub-batches of about
// SUB_BYTES; for sub-batch k the thread queues the uploads of k + 1 on a copy stream, runs the frame engine on k
// (device buffers in, device buffers out: one flacb200_encode / flacb200_decode call for all its streams) and queues the
// downloads of k on a second copy stream, each stream's bytes going straight to their place in that stream's own output
// (an exclusive scan of the frame sizes per stream -- the only "exchange" step, on the host).  MD5s are computed meanwhile
// by the other host threads, eight streams per thread at a time (md5_mb.h).  No CPU fallback: without a device every
// stream fails with FLACB200_E_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/flacb200_stream.h"
#include "md5_mb.h"

extern "C" int flacb200_build_stream_header(const flacb200_writer_options* opt, uint32_t sample_rate, uint32_t bits_per_sample, uint32_t channels,
                                            uint64_t total_pcm_frames, int total_known_at_open, const uint32_t* frame_sizes, uint64_t n_frames,
                                            const uint8_t md5[16], uint8_t* out, size_t capacity, size_t* len);

namespace {

constexpr size_t SUB_BYTES = (size_t)768 << 20;   // PCM per sub-batch: a couple of launch groups of the frame kernels

// engines are expensive to create (streams, events, scratch that grows to the working set): one per device is kept for
// the batch calls of a process; a device's engine is used by one batch call at a time
struct DeviceSlot {
    std::mutex mu;
    flacb200_engine* engine = nullptr;
    cudaStream_t in = nullptr, out = nullptr;
    void* d_in[2] = {nullptr, nullptr};
    void* d_out[2] = {nullptr, nullptr};
    size_t cap_in[2] = {0, 0}, cap_out[2] = {0, 0};
    cudaEvent_t out_done[2] = {nullptr, nullptr};
};
std::mutex g_slots_mu;
std::map<int, DeviceSlot*> g_slots;

DeviceSlot* slot_for(int device, int* rc)
{
    std::lock_guard<std::mutex> g(g_slots_mu);
    auto it = g_slots.find(device);
    if (it != g_slots.end()) return it->second;
    DeviceSlot* s = new DeviceSlot();
    *rc = flacb200_engine_create(device, &s->engine);
    if (*rc) {
        delete s;
        return nullptr;
    }
    cudaSetDevice(device);
    cudaStreamCreateWithFlags(&s->in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s->out, cudaStreamNonBlocking);
    for (auto& e : s->out_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    flacb200_engine_set_keep_info(s->engine, 0);
    g_slots[device] = s;
    return s;
}

int grow(void** p, size_t* cap, size_t want)
{
    if (want <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t n = want + want / 8 + 4096;
    if (cudaMalloc(p, n) != cudaSuccess) {
        cudaGetLastError();
        return FLACB200_E_OUT_OF_MEMORY;
    }
    *cap = n;
    return 0;
}

inline size_t lcm16(size_t a)
{
    size_t x = a, y = 16;
    while (y) { const size_t t = x % y; x = y; y = t; }
    return a / x * 16;
}

struct Key {
    uint32_t rate, bps, ch;
    int32_t kind;
    bool operator<(const Key& o) const { return std::tie(rate, bps, ch, kind) < std::tie(o.rate, o.bps, o.ch, o.kind); }
};

// host threads this process may use for hashing: all of them, unless FLACB200_HOST_THREADS says otherwise (several rank
// processes on one host share its cores)
unsigned host_threads()
{
    static const unsigned n = [] {
        if (const char* v = getenv("FLACB200_HOST_THREADS")) {
            const unsigned long k = strtoul(v, nullptr, 0);
            if (k) return (unsigned)k;
        }
        return std::max(1u, std::thread::hardware_concurrency());
    }();
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------------------
struct EncJob {
    const flacb200_track* tracks;
    flacb200_file* files;
    const flacb200_writer_options* opt;
    std::vector<size_t> mine;                    // track indices of this device, grouped by stream parameters
    std::vector<std::vector<uint32_t>>* sizes;   // per track: frame sizes
    std::vector<size_t>* header_len;
};

void encode_on_device(int device, EncJob job)
{
    int rc = 0;
    DeviceSlot* s = slot_for(device, &rc);
    auto fail_all = [&](int code) {
        for (size_t t : job.mine)
            if (job.files[t].status == 0) job.files[t].status = code;
    };
    if (!s) return fail_all(rc);
    std::lock_guard<std::mutex> g(s->mu);
    cudaSetDevice(device);
    // sub-batches: consecutive tracks with the same stream parameters, about SUB_BYTES of PCM
    struct Sub { size_t a, b; };
    std::vector<Sub> subs;
    auto key = [&](size_t t) { const flacb200_track& k = job.tracks[t]; return Key{k.sample_rate, k.bits_per_sample, k.channels, k.pcm_kind}; };
    auto track_bytes = [&](size_t t) {
        const flacb200_track& k = job.tracks[t];
        return (size_t)k.n_pcm_frames * k.channels * (k.pcm_kind == FLACB200_PCM_I32_INTERLEAVED ? 4 : (k.bits_per_sample + 7) / 8);
    };
    for (size_t i = 0; i < job.mine.size();) {
        size_t j = i, bytes = 0;
        while (j < job.mine.size() && !(key(job.mine[j]) < key(job.mine[i])) && !(key(job.mine[i]) < key(job.mine[j])) && (j == i || bytes < SUB_BYTES))
            bytes += track_bytes(job.mine[j++]);
        subs.push_back(Sub{i, j});
        i = j;
    }
    struct Placed { std::vector<size_t> off; size_t total = 0; };
    std::vector<Placed> placed(subs.size());
    auto upload = [&](size_t k) -> int {   // queue the PCM of sub-batch k into buffer k & 1
        const Sub& sb = subs[k];
        const flacb200_track& k0 = job.tracks[job.mine[sb.a]];
        const size_t fb = (size_t)k0.channels * (k0.pcm_kind == FLACB200_PCM_I32_INTERLEAVED ? 4 : (k0.bits_per_sample + 7) / 8);
        const size_t align = lcm16(fb);   // every track starts on a PCM-frame boundary that is also 16-byte aligned (k_lpc3's cp.async)
        Placed& pl = placed[k];
        pl.off.clear();
        size_t at = 0;
        for (size_t i = sb.a; i < sb.b; i++) {
            pl.off.push_back(at);
            at += (track_bytes(job.mine[i]) + align - 1) / align * align;
        }
        pl.total = at;
        const int r = grow(&s->d_in[k & 1], &s->cap_in[k & 1], at + 64);
        if (r) return r;
        for (size_t i = sb.a; i < sb.b; i++) {
            const size_t n = track_bytes(job.mine[i]);
            if (n && cudaMemcpyAsync((uint8_t*)s->d_in[k & 1] + pl.off[i - sb.a], job.tracks[job.mine[i]].pcm, n, cudaMemcpyHostToDevice, s->in) != cudaSuccess)
                return FLACB200_E_CUDA_BASE - (int)cudaGetLastError();
        }
        return 0;
    };
    if (!subs.empty() && (rc = upload(0)) != 0) return fail_all(rc);
    std::vector<uint32_t> fsz;
    std::vector<flacb200_segment> segs;
    for (size_t k = 0; k < subs.size(); k++) {
        const Sub& sb = subs[k];
        const flacb200_track& k0 = job.tracks[job.mine[sb.a]];
        const size_t fb = (size_t)k0.channels * (k0.pcm_kind == FLACB200_PCM_I32_INTERLEAVED ? 4 : (k0.bits_per_sample + 7) / 8);
        cudaStreamSynchronize(s->in);   // sub-batch k has landed
        if (k + 1 < subs.size() && (rc = upload(k + 1)) != 0) return fail_all(rc);   // k + 1 travels while k is encoded
        flacb200_stream_params prm{};
        prm.sample_rate = k0.sample_rate;
        prm.bits_per_sample = k0.bits_per_sample;
        prm.channels = k0.channels;
        segs.clear();
        uint64_t nframes = 0;
        const uint32_t bs = job.opt->frame.block_size;
        for (size_t i = sb.a; i < sb.b; i++) {
            const flacb200_track& tk = job.tracks[job.mine[i]];
            segs.push_back(flacb200_segment{placed[k].off[i - sb.a] / fb, tk.n_pcm_frames, 0});
            nframes += (tk.n_pcm_frames + bs - 1) / bs;
        }
        const size_t bound = flacb200_encode_bound(&job.opt->frame, &prm, segs.data(), segs.size()) + 256;
        if (k >= 2) cudaEventSynchronize(s->out_done[k & 1]);   // the downloads of sub-batch k - 2 have left this buffer
        if ((rc = grow(&s->d_out[k & 1], &s->cap_out[k & 1], bound)) != 0) return fail_all(rc);
        fsz.resize(nframes);
        uint64_t nf = 0, total = 0;
        rc = flacb200_encode(s->engine, &job.opt->frame, &prm, s->d_in[k & 1], placed[k].total, k0.pcm_kind, FLACB200_DEVICE, 0, segs.data(), segs.size(),
                             s->d_out[k & 1], s->cap_out[k & 1], FLACB200_DEVICE, fsz.data(), fsz.size(), &nf, &total);
        if (rc) {
            for (size_t i = sb.a; i < sb.b; i++) job.files[job.mine[i]].status = rc;
            continue;
        }
        // ---- placement: an exclusive scan of the frame sizes, per track ----
        size_t f = 0, off = 0;
        for (size_t i = sb.a; i < sb.b; i++) {
            const size_t t = job.mine[i];
            const flacb200_track& tk = job.tracks[t];
            const size_t nfr = (size_t)((tk.n_pcm_frames + bs - 1) / bs);
            std::vector<uint32_t>& sz = (*job.sizes)[t];
            sz.assign(fsz.begin() + (long)f, fsz.begin() + (long)(f + nfr));
            size_t bytes = 0;
            for (uint32_t v : sz) bytes += v;
            flacb200_file& fl = job.files[t];
            const size_t hl = (*job.header_len)[t], need = hl + bytes;
            if (!fl.data) {
                fl.data = (uint8_t*)malloc(need);
                fl.capacity = fl.data ? need : 0;
                if (!fl.data) fl.status = FLACB200_E_OUT_OF_MEMORY;
            } else if (fl.capacity < need) {
                fl.status = FLACB200_E_OUTPUT_TOO_SMALL;
            }
            if (fl.status == 0) {
                fl.len = need;
                fl.frames = (uint32_t)nfr;
                if (bytes && cudaMemcpyAsync(fl.data + hl, (const uint8_t*)s->d_out[k & 1] + off, bytes, cudaMemcpyDeviceToHost, s->out) != cudaSuccess)
                    fl.status = FLACB200_E_CUDA_BASE - (int)cudaGetLastError();
            }
            f += nfr;
            off += bytes;
        }
        cudaEventRecord(s->out_done[k & 1], s->out);
    }
    cudaStreamSynchronize(s->out);
}

}   // namespace

extern "C" {

void flacb200_md5_many(const uint8_t* const* data, const size_t* len, size_t n, uint8_t* digests, unsigned threads)
{
    flacb200::md5_many(data, len, n, reinterpret_cast<uint8_t(*)[16]>(digests), threads ? threads : host_threads());
}

size_t flacb200_encode_batch_bound(const flacb200_track* t, const flacb200_writer_options* opt)
{
    if (!t || !opt) return 0;
    size_t hl = 0;
    if (flacb200_build_stream_header(opt, t->sample_rate, t->bits_per_sample, t->channels, t->n_pcm_frames, 1, nullptr, 0, nullptr, nullptr, 0, &hl)) return 0;
    flacb200_stream_params prm{};
    prm.sample_rate = t->sample_rate;
    prm.bits_per_sample = t->bits_per_sample;
    prm.channels = t->channels;
    flacb200_segment seg{0, t->n_pcm_frames, 0};
    return hl + flacb200_encode_bound(&opt->frame, &prm, &seg, 1);
}

int flacb200_encode_batch(const flacb200_track* tracks, size_t n_tracks, const flacb200_writer_options* opt, const int* devices, int n_devices,
                          flacb200_file* files)
{
    if (!opt || (!tracks && n_tracks) || (!files && n_tracks) || n_devices < 0) return FLACB200_E_BAD_ARGUMENT;
    if (n_tracks == 0) return 0;
    const int dev0 = 0;
    if (!devices || n_devices == 0) {
        devices = &dev0;
        n_devices = 1;
    }
    // ---- Encoder::new's checks and the initial metadata blocks of every track ----
    std::vector<size_t> header_len(n_tracks, 0);
    std::vector<std::vector<uint32_t>> sizes(n_tracks);
    for (size_t t = 0; t < n_tracks; t++) {
        flacb200_file& fl = files[t];
        fl.len = 0;
        fl.frames = 0;
        memset(fl.md5, 0, 16);
        const flacb200_track& tk = tracks[t];
        fl.status = 0;
        if (!tk.pcm && tk.n_pcm_frames) fl.status = FLACB200_E_BAD_ARGUMENT;
        else if (tk.pcm_kind != FLACB200_PCM_BYTES_LE && tk.pcm_kind != FLACB200_PCM_BYTES_BE && tk.pcm_kind != FLACB200_PCM_I32_INTERLEAVED)
            fl.status = FLACB200_E_BAD_ARGUMENT;
        else if (tk.n_pcm_frames == 0) fl.status = 58;   // NoSamples (finalize_inner, src/encode.rs:2037)
        else
            fl.status = flacb200_build_stream_header(opt, tk.sample_rate, tk.bits_per_sample, tk.channels, tk.n_pcm_frames, 1, nullptr, 0, nullptr,
                                                     nullptr, 0, &header_len[t]);
    }
    // ---- deal the tracks to the devices: largest first onto the least loaded device, grouped by stream parameters ----
    std::vector<size_t> order;
    for (size_t t = 0; t < n_tracks; t++)
        if (files[t].status == 0) order.push_back(t);
    auto bytes_of = [&](size_t t) { return (size_t)tracks[t].n_pcm_frames * tracks[t].channels * ((tracks[t].bits_per_sample + 7) / 8); };
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bytes_of(a) > bytes_of(b); });
    std::vector<EncJob> jobs((size_t)n_devices);
    std::vector<size_t> load((size_t)n_devices, 0);
    for (size_t t : order) {
        const size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        jobs[d].mine.push_back(t);
        load[d] += bytes_of(t);
    }
    for (EncJob& j : jobs) {
        j.tracks = tracks;
        j.files = files;
        j.opt = opt;
        j.sizes = &sizes;
        j.header_len = &header_len;
        std::stable_sort(j.mine.begin(), j.mine.end(), [&](size_t a, size_t b) {
            const Key ka{tracks[a].sample_rate, tracks[a].bits_per_sample, tracks[a].channels, tracks[a].pcm_kind};
            const Key kb{tracks[b].sample_rate, tracks[b].bits_per_sample, tracks[b].channels, tracks[b].pcm_kind};
            return ka < kb || (!(kb < ka) && a < b);
        });
    }
    // ---- MD5 of every track on the host threads the devices do not need (update_md5, src/encode.rs:1292-1318) ----
    std::vector<uint8_t> digests(16 * n_tracks, 0);
    std::thread md5_thread([&] {
        std::vector<const uint8_t*> ptr;
        std::vector<size_t> len, which;
        std::vector<std::vector<uint8_t>> converted;   // big-endian / int32 input is hashed in its little-endian packed form
        for (size_t t : order) {
            const flacb200_track& tk = tracks[t];
            const size_t B = (tk.bits_per_sample + 7) / 8, n = (size_t)tk.n_pcm_frames * tk.channels;
            which.push_back(t);
            len.push_back(n * B);
            if (tk.pcm_kind == FLACB200_PCM_BYTES_LE || (tk.pcm_kind == FLACB200_PCM_BYTES_BE && B == 1)) {
                ptr.push_back((const uint8_t*)tk.pcm);
                continue;
            }
            converted.emplace_back(n * B);
            uint8_t* d = converted.back().data();
            if (tk.pcm_kind == FLACB200_PCM_BYTES_BE) {
                const uint8_t* sp = (const uint8_t*)tk.pcm;
                for (size_t i = 0; i < n; i++)
                    for (size_t k = 0; k < B; k++) d[i * B + k] = sp[i * B + B - 1 - k];
            } else {
                const int32_t* sp = (const int32_t*)tk.pcm;
                for (size_t i = 0; i < n; i++)
                    for (size_t k = 0; k < B; k++) d[i * B + k] = (uint8_t)((uint32_t)sp[i] >> (8 * k));
            }
            ptr.push_back(d);
        }
        std::vector<uint8_t> out(16 * which.size());
        // every hardware thread hashes: the device threads spend their time blocked in CUDA calls, and with eight streams per
        // group a batch rarely has more groups than threads -- one thread short doubles the rounds
        const unsigned threads = host_threads();
        flacb200::md5_many(ptr.data(), len.data(), which.size(), reinterpret_cast<uint8_t(*)[16]>(out.data()), threads);
        for (size_t i = 0; i < which.size(); i++) memcpy(digests.data() + 16 * which[i], out.data() + 16 * i, 16);
    });
    std::vector<std::thread> workers;
    for (int d = 0; d < n_devices; d++)
        if (!jobs[(size_t)d].mine.empty()) workers.emplace_back(encode_on_device, devices[d], jobs[(size_t)d]);
    for (auto& w : workers) w.join();
    md5_thread.join();
    // ---- Encoder::finalize_inner: STREAMINFO (frame sizes, total, MD5), SEEKTABLE, PADDING in front of the frames ----
    int first = 0;
    for (size_t t = 0; t < n_tracks; t++) {
        flacb200_file& fl = files[t];
        if (fl.status == 0) {
            const flacb200_track& tk = tracks[t];
            memcpy(fl.md5, digests.data() + 16 * t, 16);
            size_t hl = 0;
            fl.status = flacb200_build_stream_header(opt, tk.sample_rate, tk.bits_per_sample, tk.channels, tk.n_pcm_frames, 1, sizes[t].data(),
                                                     sizes[t].size(), fl.md5, fl.data, header_len[t], &hl);
            if (fl.status == 0 && hl != header_len[t]) fl.status = FLACB200_E_BAD_ARGUMENT;
        }
        if (fl.status && !first) first = fl.status;
    }
    return first;
}

void flacb200_files_free(flacb200_file* files, size_t n)
{
    if (!files) return;
    for (size_t i = 0; i < n; i++) {
        free(files[i].data);
        files[i].data = nullptr;
        files[i].capacity = files[i].len = 0;
    }
}

}   // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct DecJob {
    const uint8_t* const* flac;
    const size_t* flac_len;
    flacb200_pcm* out;
    int pcm_kind;
    std::vector<size_t> mine;
};

void decode_on_device(int device, DecJob job)
{
    int rc = 0;
    DeviceSlot* s = slot_for(device, &rc);
    auto fail_all = [&](int code) {
        for (size_t t : job.mine)
            if (job.out[t].status == 0) job.out[t].status = code;
    };
    if (!s) return fail_all(rc);
    std::lock_guard<std::mutex> g(s->mu);
    cudaSetDevice(device);
    auto key = [&](size_t t) { const flacb200_streaminfo& i = job.out[t].info; return Key{i.sample_rate, i.bits_per_sample, i.channels, (int32_t)i.max_block_size}; };
    auto sample_bytes = [&](size_t t) { return (size_t)(job.pcm_kind == FLACB200_PCM_I32_INTERLEAVED ? 4 : (job.out[t].info.bits_per_sample + 7) / 8); };
    auto pcm_bytes = [&](size_t t) { return (size_t)job.out[t].info.total_samples * job.out[t].info.channels * sample_bytes(t); };
    struct Sub { size_t a, b; };
    std::vector<Sub> subs;
    for (size_t i = 0; i < job.mine.size();) {
        size_t j = i, bytes = 0;
        while (j < job.mine.size() && !(key(job.mine[j]) < key(job.mine[i])) && !(key(job.mine[i]) < key(job.mine[j])) && (j == i || bytes < SUB_BYTES))
            bytes += pcm_bytes(job.mine[j++]);
        subs.push_back(Sub{i, j});
        i = j;
    }
    struct Placed { std::vector<size_t> off; size_t total = 0; };
    std::vector<Placed> placed(subs.size());
    auto frames_of = [&](size_t t, const uint8_t** p) {
        *p = job.flac[t] + job.out[t].info.frames_start;
        return job.flac_len[t] - (size_t)job.out[t].info.frames_start;
    };
    auto upload = [&](size_t k) -> int {
        const Sub& sb = subs[k];
        Placed& pl = placed[k];
        pl.off.clear();
        size_t at = 0;
        for (size_t i = sb.a; i < sb.b; i++) {
            const uint8_t* p;
            pl.off.push_back(at);
            at += (frames_of(job.mine[i], &p) + 15) & ~(size_t)15;
        }
        pl.total = at;
        const int r = grow(&s->d_in[k & 1], &s->cap_in[k & 1], at + 64);
        if (r) return r;
        for (size_t i = sb.a; i < sb.b; i++) {
            const uint8_t* p;
            const size_t n = frames_of(job.mine[i], &p);
            if (n && cudaMemcpyAsync((uint8_t*)s->d_in[k & 1] + pl.off[i - sb.a], p, n, cudaMemcpyHostToDevice, s->in) != cudaSuccess)
                return FLACB200_E_CUDA_BASE - (int)cudaGetLastError();
        }
        return 0;
    };
    if (!subs.empty() && (rc = upload(0)) != 0) return fail_all(rc);
    std::vector<flacb200_decode_segment> segs;
    for (size_t k = 0; k < subs.size(); k++) {
        const Sub& sb = subs[k];
        const flacb200_streaminfo& i0 = job.out[job.mine[sb.a]].info;
        cudaStreamSynchronize(s->in);
        if (k + 1 < subs.size() && (rc = upload(k + 1)) != 0) return fail_all(rc);
        flacb200_stream_params prm{};
        prm.sample_rate = i0.sample_rate;
        prm.bits_per_sample = i0.bits_per_sample;
        prm.channels = i0.channels;
        prm.max_block_size = i0.max_block_size;
        const size_t fb = (size_t)i0.channels * sample_bytes(job.mine[sb.a]);
        segs.clear();
        uint64_t pcm_at = 0;
        for (size_t i = sb.a; i < sb.b; i++) {
            const size_t t = job.mine[i];
            const uint8_t* p;
            segs.push_back(flacb200_decode_segment{placed[k].off[i - sb.a], frames_of(t, &p), pcm_at, job.out[t].info.total_samples});
            pcm_at += job.out[t].info.total_samples;
        }
        if (k >= 2) cudaEventSynchronize(s->out_done[k & 1]);
        if ((rc = grow(&s->d_out[k & 1], &s->cap_out[k & 1], (size_t)pcm_at * fb + 256)) != 0) return fail_all(rc);
        uint64_t nf = 0, ns = 0, bad = 0;
        rc = flacb200_decode(s->engine, &prm, s->d_in[k & 1], placed[k].total, FLACB200_DEVICE, segs.data(), segs.size(), s->d_out[k & 1],
                             (size_t)pcm_at * fb, job.pcm_kind, FLACB200_DEVICE, 0, &nf, &ns, &bad);
        if (rc < 0) {
            for (size_t i = sb.a; i < sb.b; i++) job.out[job.mine[i]].status = rc;
            continue;
        }
        if (rc > 0) {
            // a stream of the sub-batch is damaged: which one is found by decoding them one at a time (the batch call reports
            // the first error in stream order only)
            for (size_t i = sb.a; i < sb.b; i++) {
                flacb200_decode_segment one = segs[i - sb.a];
                uint64_t a = 0, b = 0, c = 0;
                const int r1 = flacb200_decode(s->engine, &prm, s->d_in[k & 1], placed[k].total, FLACB200_DEVICE, &one, 1, s->d_out[k & 1],
                                               (size_t)pcm_at * fb, job.pcm_kind, FLACB200_DEVICE, 0, &a, &b, &c);
                if (r1) job.out[job.mine[i]].status = r1;
            }
        }
        uint64_t at = 0;
        for (size_t i = sb.a; i < sb.b; i++) {
            const size_t t = job.mine[i];
            flacb200_pcm& o = job.out[t];
            const size_t need = (size_t)o.info.total_samples * fb;
            if (o.status == 0) {
                if (!o.data) {
                    o.data = malloc(std::max<size_t>(need, 1));
                    o.capacity = o.data ? need : 0;
                    if (!o.data) o.status = FLACB200_E_OUT_OF_MEMORY;
                } else if (o.capacity < need) {
                    o.status = FLACB200_E_OUTPUT_TOO_SMALL;
                }
            }
            if (o.status == 0) {
                o.len = need;
                if (need && cudaMemcpyAsync(o.data, (const uint8_t*)s->d_out[k & 1] + (size_t)at * fb, need, cudaMemcpyDeviceToHost, s->out) != cudaSuccess)
                    o.status = FLACB200_E_CUDA_BASE - (int)cudaGetLastError();
            }
            at += o.info.total_samples;
        }
        cudaEventRecord(s->out_done[k & 1], s->out);
    }
    cudaStreamSynchronize(s->out);
}

}   // namespace

extern "C" {

int flacb200_decode_batch(const uint8_t* const* flac, const size_t* flac_len, size_t n_files, int pcm_kind, int verify, const int* devices,
                          int n_devices, flacb200_pcm* out)
{
    if ((!flac || !flac_len || !out) && n_files) return FLACB200_E_BAD_ARGUMENT;
    if (pcm_kind != FLACB200_PCM_BYTES_LE && pcm_kind != FLACB200_PCM_BYTES_BE && pcm_kind != FLACB200_PCM_I32_INTERLEAVED) return FLACB200_E_BAD_ARGUMENT;
    if (n_files == 0) return 0;
    const int dev0 = 0;
    if (!devices || n_devices <= 0) {
        devices = &dev0;
        n_devices = 1;
    }
    std::vector<size_t> order;
    for (size_t t = 0; t < n_files; t++) {
        flacb200_pcm& o = out[t];
        o.len = 0;
        o.verified = -1;
        o.status = flac[t] ? flacb200_read_streaminfo(flac[t], flac_len[t], &o.info) : FLACB200_E_BAD_ARGUMENT;
        // the batch path sizes its buffers from STREAMINFO: streams of unknown length go through a reader handle instead
        if (o.status == 0 && o.info.total_samples == 0) o.status = FLACB200_E_BAD_ARGUMENT;
        // (an untrusted total cannot exceed what the bytes could possibly hold: 16 samples per channel and byte at the very least)
        if (o.status == 0 && o.info.total_samples / 65536 > flac_len[t]) o.status = 59;   // SampleCountMismatch
        if (o.status == 0) order.push_back(t);
    }
    auto weight = [&](size_t t) { return (size_t)out[t].info.total_samples * out[t].info.channels; };
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return weight(a) > weight(b); });
    std::vector<DecJob> jobs((size_t)n_devices);
    std::vector<size_t> load((size_t)n_devices, 0);
    for (size_t t : order) {
        const size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        jobs[d].mine.push_back(t);
        load[d] += weight(t);
    }
    for (DecJob& j : jobs) {
        j.flac = flac;
        j.flac_len = flac_len;
        j.out = out;
        j.pcm_kind = pcm_kind;
        std::stable_sort(j.mine.begin(), j.mine.end(), [&](size_t a, size_t b) {
            const flacb200_streaminfo &x = out[a].info, &y = out[b].info;
            const Key ka{x.sample_rate, x.bits_per_sample, x.channels, (int32_t)x.max_block_size}, kb{y.sample_rate, y.bits_per_sample, y.channels, (int32_t)y.max_block_size};
            return ka < kb || (!(kb < ka) && a < b);
        });
    }
    std::vector<std::thread> workers;
    for (int d = 0; d < n_devices; d++)
        if (!jobs[(size_t)d].mine.empty()) workers.emplace_back(decode_on_device, devices[d], jobs[(size_t)d]);
    for (auto& w : workers) w.join();
    if (verify) {   // verify_reader (src/decode.rs:1291-1309): MD5 of the little-endian PCM
        std::vector<const uint8_t*> ptr;
        std::vector<size_t> len, which;
        std::vector<std::vector<uint8_t>> converted;
        for (size_t t : order) {
            flacb200_pcm& o = out[t];
            if (o.status) continue;
            const size_t B = (o.info.bits_per_sample + 7) / 8, n = (size_t)o.info.total_samples * o.info.channels;
            which.push_back(t);
            len.push_back(n * B);
            if (pcm_kind == FLACB200_PCM_BYTES_LE || (pcm_kind == FLACB200_PCM_BYTES_BE && B == 1)) {
                ptr.push_back((const uint8_t*)o.data);
                continue;
            }
            converted.emplace_back(n * B);
            uint8_t* d = converted.back().data();
            if (pcm_kind == FLACB200_PCM_BYTES_BE) {
                const uint8_t* sp = (const uint8_t*)o.data;
                for (size_t i = 0; i < n; i++)
                    for (size_t k = 0; k < B; k++) d[i * B + k] = sp[i * B + B - 1 - k];
            } else {
                const int32_t* sp = (const int32_t*)o.data;
                for (size_t i = 0; i < n; i++)
                    for (size_t k = 0; k < B; k++) d[i * B + k] = (uint8_t)((uint32_t)sp[i] >> (8 * k));
            }
            ptr.push_back(d);
        }
        std::vector<uint8_t> dig(16 * std::max<size_t>(which.size(), 1));
        flacb200::md5_many(ptr.data(), len.data(), which.size(), reinterpret_cast<uint8_t(*)[16]>(dig.data()), host_threads());
        static const uint8_t zero[16] = {0};
        for (size_t i = 0; i < which.size(); i++) {
            flacb200_pcm& o = out[which[i]];
            if (memcmp(o.info.md5, zero, 16) == 0) o.verified = 2;
            else o.verified = memcmp(o.info.md5, dig.data() + 16 * i, 16) == 0 ? 0 : 1;
        }
    }
    int first = 0;
    for (size_t t = 0; t < n_files; t++)
        if (out[t].status && !first) first = out[t].status;
    return first;
}

void flacb200_pcm_free(flacb200_pcm* out, size_t n)
{
    if (!out) return;
    for (size_t i = 0; i < n; i++) {
        free(out[i].data);
        out[i].data = nullptr;
        out[i].capacity = out[i].len = 0;
    }
}

}   // extern "C"
