// engine.cu -- host side of libflacb200.so: the C ABI of include/flacb200.h.
//
// One engine = one CUDA device + one stream + grow-only scratch.  The engine cuts segments into
// blocks, uploads PCM once, runs the kernel pipeline of encode_kernels.cu / decode_kernels.cu per
// launch group and returns frames (or PCM) plus per-frame sizes.  There is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <array>
#include <vector>

#include "../../include/flacb200.h"
#include "common.cuh"
#include "decode.cuh"
#include "glibc_log.cuh"

namespace flacb200 {
// encode_kernels.cu
void launch_planes(const EncCfg&, const FrameDesc*, const uint8_t*, int32_t*, uint32_t*, unsigned long long*, cudaStream_t);
void launch_lpc(const EncCfg&, const FrameDesc*, const int32_t*, const uint32_t*, const unsigned long long*, const double*, LpcRec*, cudaStream_t);
bool residual_uses_smem(const EncCfg&);
cudaError_t launch_residual(const EncCfg&, const FrameDesc*, const int32_t*, const uint32_t*, const unsigned long long*, const LpcRec*, CandRec*,
                            int32_t*, cudaStream_t);
void launch_decide_scan(const EncCfg&, const FrameDesc*, const CandRec*, const unsigned long long*, FrameRec*, uint32_t*, unsigned long long*,
                        unsigned long long*, uint8_t*, bool, cudaStream_t);
bool pack3_ok(const EncCfg&);
bool analyze3_ok(const EncCfg&);
bool lpc3_ok(const EncCfg&, bool);
cudaError_t launch_lpc3(const EncCfg&, const FrameDesc*, const uint8_t*, const double*, LpcRec*, uint32_t, cudaStream_t);
bool lpc4_ok(const EncCfg&, bool);
cudaError_t launch_lpc4(const EncCfg&, const FrameDesc*, const uint8_t*, const double*, LpcRec*, uint32_t, cudaStream_t);
cudaError_t launch_analyze3(const EncCfg&, const FrameDesc*, const uint8_t*, const LpcRec*, CandRec*, unsigned long long*, uint4*, cudaStream_t);
cudaError_t launch_pack3(const EncCfg&, const FrameDesc*, const uint8_t*, const CandRec*, const FrameRec*, uint8_t*, const uint4*, cudaStream_t);
cudaError_t launch_pack_crc(const EncCfg&, const FrameDesc*, const int32_t*, const CandRec*, const FrameRec*, uint8_t*, cudaStream_t);
bool analyze_fast_ok(const EncCfg&);
cudaError_t launch_pack2_crc(const EncCfg&, const FrameDesc*, const uint8_t*, const CandRec*, const FrameRec*, uint8_t*, cudaStream_t);
cudaError_t launch_lpc2(const EncCfg&, const FrameDesc*, const uint8_t*, const double*, LpcRec*, cudaStream_t);
cudaError_t launch_analyze(const EncCfg&, const FrameDesc*, const uint8_t*, const LpcRec*, CandRec*, unsigned long long*, cudaStream_t);
void init_encode_tables(cudaStream_t);   // encode_frame.cu
void init_fused_tables(cudaStream_t);    // encode_analyze.cu
bool frame4_ok(const EncCfg&);
cudaError_t launch_frame4(const EncCfg&, const FrameDesc*, const uint8_t*, const LpcRec*, CandRec*, FrameRec*, unsigned long long*, unsigned long long*,
                          const unsigned long long*, unsigned long long*, unsigned long long*, unsigned long long*, uint32_t*, uint8_t*, cudaStream_t);
void init_decode_tables(cudaStream_t);   // decode_kernels.cu
// synth.cu
cudaError_t launch_synth(uint8_t* pcm, unsigned long long first_track, unsigned long long n_tracks, unsigned long long n_pcm_frames,
                         uint32_t channels, uint32_t sample_rate, uint32_t bps, unsigned long long seed, const int32_t* lut, cudaStream_t st);
// decode_kernels.cu
uint32_t find_tiles(unsigned long long nbytes);
void launch_find_count(const DecCfg&, const uint8_t*, const DecSeg*, uint32_t*, uint32_t*, uint32_t*, cudaStream_t);
void launch_find_write(const DecCfg&, const uint8_t*, const DecSeg*, const uint32_t*, FrameCand*, cudaStream_t);
size_t find_slots_bytes(unsigned long long nbytes);
void launch_find_park(const DecCfg&, const uint8_t*, const DecSeg*, uint32_t*, uint32_t*, FrameCand*, uint32_t*, cudaStream_t);
void launch_find_compact(const DecCfg&, const uint32_t*, const uint32_t*, const FrameCand*, FrameCand*, cudaStream_t);
void launch_decode(const DecCfg&, const uint8_t*, const DecSeg*, const FrameCand*, uint32_t, int32_t*, DecRec*, bool, cudaStream_t);
// decode_parse.cu
void launch_parse(const DecCfg&, const uint8_t*, const DecSeg*, const FrameCand*, uint32_t, int32_t*, SubRec*, DecRec*, uint32_t*, cudaStream_t);
bool restore_emit_ok(const DecCfg&, const uint8_t*);
void launch_restore_emit(const DecCfg&, const FrameCand*, uint32_t, const SubRec*, const DecRec*, const unsigned long long*, const int32_t*, uint8_t*,
                         const uint32_t*, bool, cudaStream_t);
void launch_restore(const DecCfg&, const FrameCand*, uint32_t, const SubRec*, const DecRec*, int32_t*, cudaStream_t);
void launch_crc16f(const uint8_t*, const FrameCand*, uint32_t, DecRec*, cudaStream_t);
cudaError_t launch_chain(const DecCfg&, const uint8_t*, const DecSeg*, const FrameCand*, DecRec*, uint32_t, const FrameCand*, uint32_t,
                         unsigned long long*, ChainState*, cudaStream_t);
void launch_emit(const DecCfg&, const FrameCand*, const DecRec*, const unsigned long long*, const int32_t*, uint32_t, uint8_t*, cudaStream_t);
void launch_chain_fast(const DecCfg&, const DecSeg*, const FrameCand*, const DecRec*, uint32_t, uint32_t, uint32_t, unsigned long long*, ChainState*,
                       uint32_t*, cudaStream_t);
// md5.cu
struct Md5Seg {
    unsigned long long pcm_off, n_pcm;
};
struct Md5Cfg {
    uint32_t channels, bytes_per_sample, pcm_kind, nseg;
    unsigned long long planar_stride;
};
cudaError_t launch_md5(const Md5Cfg&, const uint8_t*, const Md5Seg*, uint32_t*, cudaStream_t);
}   // namespace flacb200

namespace flacb200 {
thread_local unsigned long long g_kernel_launches = 0;
}
using namespace flacb200;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct SegRec {
    unsigned long long pcm_offset, n_pcm_frames, first_fnum, first_frame;   // first_frame: index of the segment's first block in the call
    uint32_t win_full, win_tail;                                            // window pool offsets of a full block and of the last, shorter one
};

struct flacb200_engine {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host<->device copies overlapped with the kernels of other launch groups
    cudaStream_t aux = nullptr;                            // decode: CRC-16 and the frame walk beside the predictor restoration
    cudaStream_t lpc_stream = nullptr;                     // encode, option lpc_overlap: the persistent k_lpc4 grid of the next launch group (high priority)
    // small host<->device messages of the decode path (segment table up; candidate count, walk verdict, walk state down) go
    // through pinned mapped memory and a tiny copy kernel, not through the copy engines: behind a 350 MB upload or download
    // of another batch a 4-byte cudaMemcpyAsync waits for milliseconds
    uint8_t* mbox_h = nullptr;
    uint8_t* mbox_d = nullptr;
    size_t mbox_cap = 0;
    std::vector<cudaEvent_t> batch_ev;                     // decode with host buffers: upload of segment batch b landed
    std::vector<cudaEvent_t> pipe_ev;                      // per group: input landed, group done
    unsigned long long* h_totals = nullptr;                // pinned + mapped: cumulative output bytes after each group,
    unsigned long long* d_h_totals = nullptr;              // written by k_scan itself (no trip through the copy queue)
    size_t h_totals_cap = 0;
    uint32_t chunk_frames = 0;
    bool profiling = false, keep_info = true;
    // runtime knobs: defaults from the environment, read ONCE at engine creation;
    // flacb200_engine_set_option changes them afterwards
    unsigned legacy = 0;
    bool no_batch = false, debug = false;
    size_t batch_bytes = 0;   // 0 = default
    uint32_t plane_mb = 12288;   // decode: MB of scratch planes per launch group (two such buffers when the input needs several groups: they are pipelined)
    uint32_t lpc_overlap = 0;   // (experiment, off: measured no gain -- both sides are occupancy-bound) CTAs per SM of the persistent k_lpc3 that runs beside the previous group's integer kernels; 0 = off
    bool fused_frame = false;   // k_frame4 (analysis + decision + packing in one kernel): parity-green, measured SLOWER than the three
                                // kernels on C4 (45 vs 41 ms per step), so it is an option, not the default
    int sm_count = 148;
    std::vector<cudaEvent_t> lpc_ev;   // per group: LPC parameters ready, analysis done (the two LpcRec buffers alternate)
    DevBuf pcm, planes, masks, lpcs, lpcs_all, cands, frecs, descs, segs, out, fbytes, totals, winpool, scratch, lut, lookback, find_slots, res16, dec[12];
    std::map<uint32_t, uint32_t> win_off;   // block length -> offset in doubles
    std::vector<double> win_host;
    flacb200_options win_opt{};
    bool win_dirty = false;
    flacb200_timings tm{};
    cudaEvent_t ev[32] = {};
    std::vector<cudaEvent_t> evpool;   // per-kernel timing events (profiling only), grown on demand
    // last-call debug info
    std::vector<CandRec> info_cands;
    std::vector<FrameRec> info_frecs;
    EncCfg info_cfg{};
    uint64_t info_frames = 0;
    uint32_t last_ncand = 0;   // candidates of the most recent decode call (flacb200_decode_last_frames)
    void* host_stage = nullptr;
    size_t host_stage_cap = 0;
    std::vector<FrameDesc> descs_host;
    std::vector<SegRec> segs_host;
};

static int cuda_err(cudaError_t e) { return e == cudaSuccess ? 0 : FLACB200_E_CUDA_BASE - (int)e; }
#define CK(x)                                 \
    do {                                      \
        cudaError_t _e = (x);                 \
        if (_e != cudaSuccess) return cuda_err(_e); \
    } while (0)

static int ensure(DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes + 256);
        if (e != cudaSuccess) return FLACB200_E_OUT_OF_MEMORY;
        want = bytes + 256;
    }
    b.cap = want;
    return 0;
}
#define ENS(buf, bytes)                  \
    do {                                 \
        int _r = ensure(buf, bytes);     \
        if (_r) return _r;               \
    } while (0)

// ---- frame descriptors built on the device ---------------------------------------------------------------------------
// With the PCM resident on the device nothing on the host needs the per-frame table: the host describes the segments
// (one record each) and a thread per frame finds its segment by binary search over the segments' first frame indices.
// (A 270 000-frame call spent more than a millisecond building and uploading 6.5 MB of descriptors before its first kernel.)

__global__ void k_descs(const SegRec* __restrict__ segs, uint32_t nseg, uint32_t bs, FrameDesc* __restrict__ out, unsigned long long nframes)
{
    const unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    uint32_t lo = 0, hi = nseg - 1;   // the LAST segment whose first_frame <= f (empty segments share their successor's index)
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (segs[mid].first_frame <= f) lo = mid;
        else hi = mid - 1;
    }
    const SegRec sg = segs[lo];
    const unsigned long long k = f - sg.first_frame, done = k * bs;
    const unsigned long long left = sg.n_pcm_frames - done;
    FrameDesc d;
    d.pcm_off = sg.pcm_offset + done;
    d.fnum = sg.first_fnum + k;
    d.n = left < bs ? (uint32_t)left : bs;
    d.win_off = d.n == bs ? sg.win_full : sg.win_tail;
    out[f] = d;
}

extern "C" {

void flacb200_options_default(flacb200_options* o)
{
    memset(o, 0, sizeof(*o));
    o->block_size = 4096;
    o->max_lpc_order = 8;
    o->max_partition_order = 5;
    o->mid_side = 1;
    o->exhaustive_channel_correlation = 1;
    o->window_kind = 2;
    o->tukey_p = 0.5f;
}

void flacb200_options_fast(flacb200_options* o)
{
    flacb200_options_default(o);
    o->block_size = 1152;
    o->mid_side = 0;
    o->max_partition_order = 3;
    o->max_lpc_order = 0;
    o->exhaustive_channel_correlation = 0;
}

void flacb200_options_best(flacb200_options* o)
{
    flacb200_options_default(o);
    o->max_partition_order = 6;
    o->max_lpc_order = 12;
}

int flacb200_engine_create(int device, flacb200_engine** out)
{
    if (!out) return FLACB200_E_BAD_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return FLACB200_E_NO_DEVICE;
    }
    if (device < 0 || device >= count) return FLACB200_E_BAD_ARGUMENT;
    CK(cudaSetDevice(device));
    flacb200_engine* e = new flacb200_engine();
    e->device = device;
    cudaError_t err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
        delete e;
        return cuda_err(err);
    }
    e->stream = e->own_stream;
    cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking);
    {   // the persistent LPC grid of the overlap mode: its one small CTA per SM should be placed as soon as an SM has room
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        cudaStreamCreateWithPriority(&e->lpc_stream, cudaStreamNonBlocking, hi_p);
    }
    for (auto& ev : e->ev) cudaEventCreate(&ev);
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char* v = getenv("FLACB200_LPC_OVERLAP")) e->lpc_overlap = (uint32_t)strtoul(v, nullptr, 0);
    if (const char* v = getenv("FLACB200_PLANE_MB")) e->plane_mb = std::max<uint32_t>((uint32_t)strtoul(v, nullptr, 0), 1);
    if (const char* v = getenv("FLACB200_FUSED")) e->fused_frame = strtoul(v, nullptr, 0) != 0;
    if (const char* v = getenv("FLACB200_LEGACY")) e->legacy = (unsigned)strtoul(v, nullptr, 0);
    if (const char* v = getenv("FLACB200_BATCH_BYTES")) e->batch_bytes = std::max<size_t>((size_t)strtoull(v, nullptr, 0), 1);
    e->no_batch = getenv("FLACB200_NO_BATCH") != nullptr;
    e->debug = getenv("FLACB200_DEBUG") != nullptr;
    // the CRC tables in device memory are built once per device, before any engine on it can launch a kernel that reads them
    static std::once_flag tables_once[64];
    static cudaError_t tables_err[64];
    if (device < 64) {
        std::call_once(tables_once[device], [&] {
            init_encode_tables(e->own_stream);
            init_fused_tables(e->own_stream);
            init_decode_tables(e->own_stream);
            tables_err[device] = cudaStreamSynchronize(e->own_stream);
        });
        if (tables_err[device] != cudaSuccess) {
            const int rc = cuda_err(tables_err[device]);
            flacb200_engine_destroy(e);
            return rc;
        }
    }
    *out = e;
    return 0;
}

void flacb200_engine_destroy(flacb200_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->pcm, &e->planes, &e->masks, &e->lpcs, &e->lpcs_all, &e->cands, &e->frecs, &e->descs, &e->segs, &e->out, &e->fbytes, &e->totals, &e->winpool,
                      &e->scratch, &e->lut, &e->lookback, &e->find_slots, &e->res16};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto& b : e->dec)
        if (b.p) cudaFree(b.p);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->evpool) cudaEventDestroy(ev);
    if (e->host_stage) cudaFreeHost(e->host_stage);
    if (e->h_totals) cudaFreeHost(e->h_totals);
    if (e->mbox_h) cudaFreeHost(e->mbox_h);
    for (auto& ev : e->pipe_ev) cudaEventDestroy(ev);
    for (auto& ev : e->batch_ev) cudaEventDestroy(ev);
    for (auto& ev : e->lpc_ev) cudaEventDestroy(ev);
    cudaStreamDestroy(e->aux);
    cudaStreamDestroy(e->lpc_stream);
    cudaStreamDestroy(e->copy_in);
    cudaStreamDestroy(e->copy_out);
    cudaStreamDestroy(e->own_stream);
    delete e;
}

int flacb200_engine_set_stream(flacb200_engine* e, void* cuda_stream)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return 0;
}

int flacb200_engine_set_chunk_frames(flacb200_engine* e, uint32_t frames)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    e->chunk_frames = frames;
    return 0;
}

int flacb200_engine_set_option(flacb200_engine* e, const char* key, uint64_t value)
{
    if (!e || !key) return FLACB200_E_BAD_ARGUMENT;
    if (!strcmp(key, "legacy")) e->legacy = (unsigned)value;
    else if (!strcmp(key, "batch_bytes")) e->batch_bytes = (size_t)value;
    else if (!strcmp(key, "no_batch")) e->no_batch = value != 0;
    else if (!strcmp(key, "debug")) e->debug = value != 0;
    else if (!strcmp(key, "lpc_overlap")) e->lpc_overlap = (uint32_t)value;
    else if (!strcmp(key, "plane_mb")) e->plane_mb = std::max<uint32_t>((uint32_t)value, 1);
    else if (!strcmp(key, "fused")) e->fused_frame = value != 0;
    else return FLACB200_E_BAD_ARGUMENT;
    return 0;
}

int flacb200_engine_set_keep_info(flacb200_engine* e, int enable)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    e->keep_info = enable != 0;
    return 0;
}

int flacb200_set_profiling(flacb200_engine* e, int enable)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    e->profiling = enable != 0;
    return 0;
}

int flacb200_last_timings(flacb200_engine* e, flacb200_timings* t)
{
    if (!e || !t) return FLACB200_E_BAD_ARGUMENT;
    *t = e->tm;
    return 0;
}

static size_t frame_bound(uint32_t n, uint32_t channels, uint32_t bps)
{
    // header <= 16 bytes, per subframe 8 + 32 header bits + n * (bps + 1), CRC-16
    return 16 + (size_t)channels * (((size_t)n * (bps + 1) + 40 + 7) / 8) + 2 + 4;
}

size_t flacb200_encode_bound(const flacb200_options* opt, const flacb200_stream_params* params, const flacb200_segment* segments,
                             size_t n_segments)
{
    if (!opt || !params || !segments) return 0;
    size_t total = 0;
    const uint32_t bs = opt->block_size;
    for (size_t s = 0; s < n_segments; s++) {
        const uint64_t nf = bs ? (segments[s].n_pcm_frames + bs - 1) / bs : 0;
        total += nf * frame_bound(bs, params->channels, params->bits_per_sample);
    }
    return total + 64;
}

}   // extern "C"

// Window::generate (src/encode.rs:1725-1783).  Tables are built on the host with the C library's cos()
// (what the reference's f64::cos lowers to on Linux) and uploaded: the device never evaluates cos.
static void make_window(const flacb200_options& o, uint32_t n, double* w)
{
    auto fill1 = [&]() { for (uint32_t i = 0; i < n; i++) w[i] = 1.0; };
    auto hann = [&]() {
        const double np = (double)n - 1.0;
        for (uint32_t i = 0; i < n; i++) w[i] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)i / np);
    };
    if (o.window_kind == 0) { fill1(); return; }
    if (o.window_kind == 1) { hann(); return; }
    float p = o.tukey_p;
    if (p != p) p = 0.5f;   // NaN -> Tukey(0.5)  :1778
    if (p <= 0.0f) { fill1(); return; }
    if (p >= 1.0f) { hann(); return; }
    const double t = (double)p / 2.0 * (double)n;
    const uint64_t tt = (uint64_t)t;
    fill1();
    if (tt == 0) return;
    const uint64_t np = tt - 1;
    if (np > n || np > n - np) return;
    for (uint64_t k = 0; k < np; k++) {
        const double x = 0.5 - 0.5 * cos(M_PI * (double)k / (double)np);   // :1764
        w[k] = x;
        w[n - 1 - k] = x;
    }
}

static uint32_t window_offset(flacb200_engine* e, const flacb200_options& o, uint32_t n)
{
    if (memcmp(&e->win_opt, &o, sizeof(o)) != 0 &&
        (e->win_opt.window_kind != o.window_kind || e->win_opt.tukey_p != o.tukey_p)) {
        e->win_off.clear();
        e->win_host.clear();
    }
    e->win_opt = o;
    auto it = e->win_off.find(n);
    if (it != e->win_off.end()) return it->second;
    const uint32_t off = (uint32_t)e->win_host.size();
    e->win_host.resize(off + ((n + 3) & ~3u));
    make_window(o, n, e->win_host.data() + off);
    e->win_off[n] = off;
    e->win_dirty = true;
    return off;
}

// per-kernel timing: events are recorded back to back on the engine's stream and read after the call's
// final synchronisation, so profiling adds no host synchronisation inside the timed region
static void time_mark(flacb200_engine* e, size_t idx, cudaStream_t on = nullptr)
{
    if (!e->profiling) return;
    while (e->evpool.size() <= idx) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->evpool.push_back(ev);
    }
    cudaEventRecord(e->evpool[idx], on ? on : e->stream);
}

extern "C" int flacb200_encode(flacb200_engine* e, const flacb200_options* opt, const flacb200_stream_params* params, const void* pcm,
                               size_t pcm_bytes, int pcm_kind, int pcm_location, uint64_t planar_stride, const flacb200_segment* segments,
                               size_t n_segments, void* out, size_t out_capacity, int out_location, uint32_t* frame_bytes,
                               size_t frame_bytes_capacity, uint64_t* n_frames_out, uint64_t* total_bytes_out)
{
    if (!e || !opt || !params || !segments || (!pcm && pcm_bytes)) return FLACB200_E_BAD_ARGUMENT;
    if (params->channels < 1 || params->channels > 8) return 30;            // ExcessiveChannels (src/encode.rs:1904-1908)
    if (params->bits_per_sample < 1 || params->bits_per_sample > 32) return 33;   // InvalidBitsPerSample
    if (params->sample_rate >= (1u << 20)) return 26;                        // InvalidSampleRate (:1899-1902)
    if (opt->block_size == 0) return 24;                                     // InvalidBlockSize
    if (opt->max_lpc_order > 32 || opt->max_partition_order > 15) return FLACB200_E_BAD_ARGUMENT;
    if (pcm_kind < 0 || pcm_kind > 3) return FLACB200_E_BAD_ARGUMENT;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    memset(&e->tm, 0, sizeof(e->tm));

    EncCfg cfg{};
    cfg.channels = params->channels;
    cfg.bps = params->bits_per_sample;
    cfg.sample_rate = params->sample_rate;
    cfg.subset = params->subset;
    cfg.block_size = opt->block_size;
    cfg.bpad = (opt->block_size + 31u) & ~31u;
    cfg.max_lpc_order = opt->max_lpc_order;
    cfg.max_porder = std::min<uint32_t>(opt->max_partition_order, MAX_PORDER);
    cfg.use_rice2 = cfg.bps > 16;   // src/encode.rs:1965, :1115
    cfg.pcm_kind = (uint32_t)pcm_kind;
    cfg.bytes_per_sample = pcm_kind <= 1 ? (cfg.bps + 7) / 8 : 4;
    cfg.planar_stride = planar_stride;
    if (cfg.channels == 2 && cfg.bps < 32) {   // side needs bps + 1 <= 32 (:2715, :2473)
        if (opt->exhaustive_channel_correlation) cfg.mode = opt->mid_side ? MODE_EXH_MID_SIDE : MODE_EXH_SIDE;
        else cfg.mode = opt->mid_side ? MODE_FAST_MID_SIDE : MODE_FAST_SIDE;
        cfg.nslots = 4;
    } else {
        cfg.mode = MODE_INDEPENDENT;
        cfg.nslots = cfg.channels;
    }
    if (cfg.subset) {   // FlacStreamWriter::write validation (:1128-1137)
        const uint32_t b = cfg.bps;
        if (!(b == 8 || b == 12 || b == 16 || b == 20 || b == 24 || b == 32)) return 28;   // NonSubsetBitsPerSample
        const uint32_t r = cfg.sample_rate;
        static const uint32_t common[] = {88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
        bool ok = false;
        for (uint32_t c : common) ok |= (r == c);
        ok |= (r % 1000 == 0 && r / 1000 < 255) || (r % 10 == 0 && r / 10 < 65535) || r < 65535;
        if (!ok) return 27;   // NonSubsetSampleRate
    }

    // ---- cut segments into blocks ----
    const uint32_t bs = opt->block_size;
    const size_t sample_bytes = (size_t)cfg.bytes_per_sample;
    const uint64_t pcm_frame_bytes = (uint64_t)cfg.channels * cfg.bytes_per_sample;
    const bool dev_descs = pcm_location != FLACB200_HOST && n_segments <= 0x7FFFFFFFull;   // host PCM: the upload pipeline reads the table
    std::vector<FrameDesc>& descs = e->descs_host;   // kept between calls: no reallocation on the steady-state path
    std::vector<SegRec>& segrecs = e->segs_host;
    descs.clear();
    segrecs.clear();
    uint64_t max_index = 0, want = 0;
    uint64_t misalign = 0;           // OR of every block's byte offset: k_lpc3 / k_lpc4 copy 16-byte chunks
    uint32_t win_n = 0, win_o = 0;   // the window table is looked up once per distinct block length, not per frame
    if (dev_descs) {
        segrecs.reserve(n_segments);
        for (size_t s = 0; s < n_segments; s++) {
            const flacb200_segment& sg = segments[s];
            const uint64_t nb = (sg.n_pcm_frames + bs - 1) / bs;
            if (nb && sg.first_frame_number + (nb - 1) > 0xFFFFFFFFFull) return 38;   // ExcessiveFrameNumber
            SegRec r;
            r.pcm_offset = sg.pcm_offset; r.n_pcm_frames = sg.n_pcm_frames; r.first_fnum = sg.first_frame_number; r.first_frame = want;
            r.win_full = r.win_tail = 0;
            const uint32_t tail = (uint32_t)(sg.n_pcm_frames % bs);
            if (opt->max_lpc_order) {
                if (sg.n_pcm_frames >= bs) {
                    if (win_n != bs) { win_o = window_offset(e, *opt, bs); win_n = bs; }
                    r.win_full = win_o;
                }
                if (tail) {
                    if (win_n != tail) { win_o = window_offset(e, *opt, tail); win_n = tail; }
                    r.win_tail = win_o;
                }
            }
            for (uint64_t k = 0; k < std::min<uint64_t>(nb, 16); k++) misalign |= (sg.pcm_offset + k * bs) * pcm_frame_bytes;   // (the low bits repeat every 16 blocks)
            segrecs.push_back(r);
            want += nb;
            max_index = std::max<uint64_t>(max_index, sg.pcm_offset + sg.n_pcm_frames);
        }
    } else {
    for (size_t s = 0; s < n_segments; s++) want += (segments[s].n_pcm_frames + bs - 1) / bs;
    descs.reserve(want);
    for (size_t s = 0; s < n_segments; s++) {
        const flacb200_segment& sg = segments[s];
        uint64_t done = 0, fn = sg.first_frame_number;
        while (done < sg.n_pcm_frames) {
            const uint32_t n = (uint32_t)std::min<uint64_t>(bs, sg.n_pcm_frames - done);
            if (fn > 0xFFFFFFFFFull) return 38;   // ExcessiveFrameNumber
            if (opt->max_lpc_order && n != win_n) {
                win_o = window_offset(e, *opt, n);
                win_n = n;
            }
            FrameDesc d;
            d.pcm_off = sg.pcm_offset + done;
            d.fnum = fn++;
            d.n = n;
            d.win_off = win_o;
            misalign |= d.pcm_off * pcm_frame_bytes;
            descs.push_back(d);
            done += n;
        }
        max_index = std::max<uint64_t>(max_index, sg.pcm_offset + sg.n_pcm_frames);
    }
    }
    const uint64_t nframes = dev_descs ? want : descs.size();
    if (n_frames_out) *n_frames_out = nframes;
    if (total_bytes_out) *total_bytes_out = 0;
    if (nframes == 0) return 0;
    if (pcm_kind == FLACB200_PCM_I32_PLANAR) {
        if (max_index > planar_stride || (size_t)planar_stride * cfg.channels * 4 > pcm_bytes) return FLACB200_E_BAD_ARGUMENT;
    } else if (max_index * cfg.channels * sample_bytes > pcm_bytes) {
        return FLACB200_E_BAD_ARGUMENT;
    }
    const size_t bound = (size_t)nframes * frame_bound(bs, cfg.channels, cfg.bps) + 64;

    // ---- device buffers ----
    // kernel selection: the register-tiled kernels cover the common shapes; FLACB200_LEGACY (bit mask: 1 analyze,
    // 2 lpc, 4 pack) forces the generic kernels, which the parity tests use to cover both paths
    const unsigned legacy = e->legacy;
    const bool fast_analyze = analyze_fast_ok(cfg) && !(legacy & 1u);
    const bool fast_lpc = cfg.max_lpc_order >= 1 && !(legacy & 2u);
    const bool fast_pack = analyze_fast_ok(cfg) && !(legacy & 4u);
    const bool frame_analyze = fast_analyze && analyze3_ok(cfg) && !(legacy & 16u);   // k_analyze3: CTA per frame, shared unpack
    const bool frame_pack = fast_pack && pack3_ok(cfg) && !(legacy & 8u);   // k_pack3: whole frames, CRC fused, no pre-zeroed output
    // k_frame4: analysis, decision and packing of a stereo frame in one kernel (option "fused")
    const bool fused = e->fused_frame && frame_analyze && frame_pack && frame4_ok(cfg);
    const bool need_planes = !(fast_analyze && fast_pack && (fast_lpc || cfg.max_lpc_order == 0));
    // launch group: without the int32 planes a group costs ~250 bytes per candidate, so it can be large enough to fill
    // the GPU even for the warp-per-8-candidates LPC kernel; with planes it is sized to stay near the L2 capacity
    // with host buffers the groups are also the grain of the copy/compute pipeline: smaller groups shorten its fill and
    // drain (first upload before any kernel, last kernels + download after the last upload)
    const bool host_io = pcm_location == FLACB200_HOST || (out && out_location == FLACB200_HOST);
    uint32_t chunk = e->chunk_frames ? e->chunk_frames : (need_planes ? 2048 : (host_io ? 8192 : 32768));
    chunk = (uint32_t)std::min<uint64_t>(std::min<uint32_t>(chunk, 32768), nframes);
    const size_t ncand_chunk = (size_t)chunk * cfg.nslots;
    ENS(e->descs, nframes * sizeof(FrameDesc));
    if (need_planes) ENS(e->planes, ncand_chunk * cfg.bpad * sizeof(int32_t));
    ENS(e->masks, (size_t)chunk * (cfg.nslots * sizeof(uint32_t) + 4 * sizeof(unsigned long long)) + 64);
    ENS(e->lpcs, 2 * ncand_chunk * sizeof(LpcRec));   // two buffers: k_lpc3 of group g + 1 runs beside the integer kernels of group g
    ENS(e->cands, ncand_chunk * sizeof(CandRec));
    ENS(e->frecs, (size_t)chunk * sizeof(FrameRec));
    ENS(e->fbytes, nframes * sizeof(uint32_t));
    ENS(e->totals, 64);
    if (fused) ENS(e->lookback, (size_t)chunk * sizeof(unsigned long long));
    // k_analyze3 -> k_pack3: the int16 LPC residuals of every candidate of the launch group (8 KB each)
    const bool keep_res16 = frame_analyze && frame_pack && !fused && !(legacy & 1024u);
    if (keep_res16) ENS(e->res16, ncand_chunk * 8192);
    uint4* const d_res16 = keep_res16 ? (uint4*)e->res16.p : nullptr;
    if (!(out && out_location == FLACB200_DEVICE && out_capacity >= bound + 64 && ((uintptr_t)out & 15) == 0)) ENS(e->out, bound + 64);
    const bool need_scratch = !residual_uses_smem(cfg);
    if (need_scratch) ENS(e->scratch, ncand_chunk * 2 * cfg.bpad * sizeof(int32_t));
    if (e->win_dirty || (opt->max_lpc_order && e->winpool.cap < e->win_host.size() * sizeof(double))) {
        ENS(e->winpool, std::max<size_t>(e->win_host.size() * sizeof(double), 64));
        CK(cudaMemcpyAsync(e->winpool.p, e->win_host.data(), e->win_host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // win_host is pageable and may be reallocated by a later call
        e->win_dirty = false;
    }
    const uint8_t* d_pcm;
    // small uploads first: queued behind the bulk PCM copies they would hold the first launch group back
    if (dev_descs) {
        ENS(e->segs, segrecs.size() * sizeof(SegRec));
        CK(cudaMemcpyAsync(e->segs.p, segrecs.data(), segrecs.size() * sizeof(SegRec), cudaMemcpyHostToDevice, st));
        count_launch(), k_descs<<<(unsigned)((nframes + 255) / 256), 256, 0, st>>>((const SegRec*)e->segs.p, (uint32_t)segrecs.size(), bs,
                                                                                (FrameDesc*)e->descs.p, nframes);
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(e->descs.p, descs.data(), nframes * sizeof(FrameDesc), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemsetAsync(e->totals.p, 0, 64, st));
    if (e->profiling) cudaEventRecord(e->ev[20], st);
    const size_t ngroups = (size_t)((nframes + chunk - 1) / chunk);
    // host buffers: the copies run on their own streams, one launch group at a time, so that the PCM of group k + 1 is
    // uploaded and the frames of group k - 1 are downloaded while group k is being encoded
    const bool pipe_in = pcm_location == FLACB200_HOST && pcm_kind != FLACB200_PCM_I32_PLANAR && ngroups > 1;
    const bool pipe_out = out && out_location == FLACB200_HOST && ngroups > 1;
    if (pipe_in || pipe_out) {
        while (e->pipe_ev.size() < 2 * ngroups) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->pipe_ev.push_back(ev);
        }
        if (e->h_totals_cap < ngroups) {
            if (e->h_totals) cudaFreeHost(e->h_totals);
            e->h_totals = nullptr;
            CK(cudaHostAlloc((void**)&e->h_totals, (ngroups + 16) * sizeof(unsigned long long), cudaHostAllocMapped));
            CK(cudaHostGetDevicePointer((void**)&e->d_h_totals, e->h_totals, 0));
            e->h_totals_cap = ngroups + 16;
        }
    }
    if (pcm_location == FLACB200_HOST) {
        ENS(e->pcm, pcm_bytes + 16);
        d_pcm = (const uint8_t*)e->pcm.p;
        if (!pipe_in) CK(cudaMemcpyAsync(e->pcm.p, pcm, pcm_bytes, cudaMemcpyHostToDevice, st));
        else {
            const size_t fb = (size_t)cfg.channels * sample_bytes;   // bytes per inter-channel sample
            for (size_t g = 0; g < ngroups; g++) {
                const uint64_t b0 = g * (uint64_t)chunk, b1 = std::min<uint64_t>(b0 + chunk, nframes);
                uint64_t lo = ~0ull, hi = 0;
                for (uint64_t f = b0; f < b1; f++) {
                    lo = std::min<uint64_t>(lo, descs[f].pcm_off);
                    hi = std::max<uint64_t>(hi, descs[f].pcm_off + descs[f].n);
                }
                CK(cudaMemcpyAsync((uint8_t*)e->pcm.p + lo * fb, (const uint8_t*)pcm + lo * fb, (hi - lo) * fb, cudaMemcpyHostToDevice, e->copy_in));
                CK(cudaEventRecord(e->pipe_ev[2 * g], e->copy_in));
            }
        }
    } else {
        d_pcm = (const uint8_t*)pcm;
    }
    // frames go straight into the caller's device buffer when it can hold the worst case
    // k_lpc3: cp.async staging needs every block of the (packed, stereo) PCM on a 16-byte boundary
    const bool staged_lpc = fast_lpc && !(legacy & 32u) && lpc3_ok(cfg, ((misalign | (uint64_t)(uintptr_t)d_pcm) & 15u) == 0);
    uint8_t* d_out = (uint8_t*)e->out.p;
    const bool direct_out = out && out_location == FLACB200_DEVICE && out_capacity >= bound + 64 && ((uintptr_t)out & 15) == 0;
    if (direct_out) d_out = (uint8_t*)out;
    if (e->profiling) cudaEventRecord(e->ev[21], st);

    uint32_t* d_ormask = (uint32_t*)e->masks.p;
    unsigned long long* d_abssum = (unsigned long long*)((uint8_t*)e->masks.p + (((size_t)chunk * cfg.nslots * sizeof(uint32_t) + 15) & ~(size_t)15));
    const size_t masks_bytes = (((size_t)chunk * cfg.nslots * sizeof(uint32_t) + 15) & ~(size_t)15) + (size_t)chunk * 4 * sizeof(unsigned long long);
    const bool keep = e->keep_info && nframes <= 65536;   // flacb200_engine_set_keep_info
    if (keep) {
        e->info_cands.resize(nframes * cfg.nslots);
        e->info_frecs.resize(nframes);
        e->info_cfg = cfg;
        e->info_frames = nframes;
    } else {
        e->info_frames = 0;
    }
    const unsigned long long launches0 = g_kernel_launches;
    size_t nchunks = 0;
    bool pipe_overflow = false;
    // FP64 / integer overlap: k_lpc3 of group g + 1 is launched on a second stream as a PERSISTENT grid (a few CTAs per SM --
    // small enough to be resident next to the CTAs of k_analyze3 / k_pack3 of group g, which leave the FP64 pipe idle); a
    // full-size grid on a second stream would only start when the first kernel's last CTA has been dispatched
    const bool overlap = staged_lpc && frame_analyze && frame_pack && !fused && e->lpc_overlap != 0 && ngroups > 1;
    // k_lpc4 (a lane per candidate, eight frames per warp) where it applies; legacy bit 2048 keeps k_lpc3
    const bool lane_lpc = staged_lpc && !(legacy & 2048u) && lpc4_ok(cfg, true);
    // ... over ALL frames of the call in one launch when the PCM is on the device already: at eight frames per warp a 32768-frame
    // group is 1.7 waves of resident warps, and every group would end in its own half-empty wave
    const bool lpc_upfront = lane_lpc && !overlap && !pipe_in && ngroups > 1 && !(legacy & 4096u) && nframes <= 0xFFFFFFFFull &&
                             (size_t)nframes * cfg.nslots * sizeof(LpcRec) <= ((size_t)1 << 30);
    if (lpc_upfront) ENS(e->lpcs_all, (size_t)nframes * cfg.nslots * sizeof(LpcRec));
    LpcRec* lpc_buf[2] = {(LpcRec*)e->lpcs.p, (LpcRec*)e->lpcs.p + ncand_chunk};
    if (overlap)
        while (e->lpc_ev.size() < 2 * ngroups + 1) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->lpc_ev.push_back(ev);
        }
    auto group_cfg = [&](uint64_t base) {
        EncCfg c = cfg;
        c.nframes = (uint32_t)std::min<uint64_t>(chunk, nframes - base);
        return c;
    };
    cudaEventRecord(e->ev[22], st);
    const size_t ev_upfront = 6 * (size_t)ngroups + 8;
    if (lpc_upfront) {
        EncCfg call = cfg;
        call.nframes = (uint32_t)nframes;
        time_mark(e, ev_upfront);
        CK(launch_lpc4(call, (const FrameDesc*)e->descs.p, d_pcm, (const double*)e->winpool.p, (LpcRec*)e->lpcs_all.p, 0, st));
        time_mark(e, ev_upfront + 1);
    }
    for (uint64_t base = 0; base < nframes; base += chunk) {
        const EncCfg c = group_cfg(base);
        const FrameDesc* dd = (const FrameDesc*)e->descs.p + base;
        LpcRec* lp = lpc_upfront ? (LpcRec*)e->lpcs_all.p + base * cfg.nslots : (overlap ? lpc_buf[nchunks & 1] : lpc_buf[0]);
        if (pipe_in) CK(cudaStreamWaitEvent(st, e->pipe_ev[2 * nchunks], 0));
        CK(cudaMemsetAsync(e->masks.p, 0, masks_bytes, st));
        const size_t eb = nchunks * 6;
        time_mark(e, eb + 0);
        if (need_planes) launch_planes(c, dd, d_pcm, (int32_t*)e->planes.p, d_ormask, d_abssum, st);
        time_mark(e, eb + 1);
        if (overlap) {
            if (nchunks == 0) CK(lane_lpc ? launch_lpc4(c, dd, d_pcm, (const double*)e->winpool.p, lp, 0, st)
                                          : launch_lpc3(c, dd, d_pcm, (const double*)e->winpool.p, lp, 0, st));
            else CK(cudaStreamWaitEvent(st, e->lpc_ev[2 * nchunks], 0));   // launched beside the previous group
        } else if (lpc_upfront) {
        } else if (lane_lpc) CK(launch_lpc4(c, dd, d_pcm, (const double*)e->winpool.p, lp, 0, st));
        else if (staged_lpc) CK(launch_lpc3(c, dd, d_pcm, (const double*)e->winpool.p, lp, 0, st));
        else if (fast_lpc) CK(launch_lpc2(c, dd, d_pcm, (const double*)e->winpool.p, lp, st));
        else launch_lpc(c, dd, (const int32_t*)e->planes.p, d_ormask, d_abssum, (const double*)e->winpool.p, lp, st);
        time_mark(e, eb + 2);
        if (overlap && nchunks == 0) CK(cudaEventRecord(e->lpc_ev[2 * ngroups], st));
        if (fused) {
            // the running byte total alternates between two words of `totals`: a late look-back of this group must still find
            // the total the group started from
            unsigned long long* tt = (unsigned long long*)e->totals.p;
            CK(launch_frame4(c, dd, d_pcm, lp, keep ? (CandRec*)e->cands.p : nullptr, keep ? (FrameRec*)e->frecs.p : nullptr, d_abssum,
                             (unsigned long long*)e->lookback.p, tt + 4 + (nchunks & 1), tt + 4 + ((nchunks + 1) & 1),
                             pipe_out ? e->d_h_totals + nchunks : nullptr, tt + 3, (uint32_t*)e->fbytes.p + base, d_out, st));
            time_mark(e, eb + 3);
            time_mark(e, eb + 4);
        } else {
        if (frame_analyze) CK(launch_analyze3(c, dd, d_pcm, lp, (CandRec*)e->cands.p, d_abssum, d_res16, st));
        else if (fast_analyze) CK(launch_analyze(c, dd, d_pcm, lp, (CandRec*)e->cands.p, d_abssum, st));
        else
            CK(launch_residual(c, dd, (const int32_t*)e->planes.p, d_ormask, d_abssum, lp, (CandRec*)e->cands.p, (int32_t*)e->scratch.p, st));
        time_mark(e, eb + 3);
        if (overlap) {
            CK(cudaEventRecord(e->lpc_ev[2 * nchunks + 1], st));   // this group's LpcRec buffer may be reused by group g + 2
            const uint64_t nb = base + chunk;
            if (nb < nframes) {
                cudaStream_t sb = e->lpc_stream;
                if (nchunks >= 1) CK(cudaStreamWaitEvent(sb, e->lpc_ev[2 * (nchunks - 1) + 1], 0));   // buffer (g + 1) & 1 was read by group g - 1
                else CK(cudaStreamWaitEvent(sb, e->lpc_ev[2 * ngroups], 0));   // (orders stream B behind this call's uploads and the first group's own LPC launch on the main stream)
                if (pipe_in) CK(cudaStreamWaitEvent(sb, e->pipe_ev[2 * (nchunks + 1)], 0));
                const EncCfg cn = group_cfg(nb);
                if (lane_lpc)
                    CK(launch_lpc4(cn, (const FrameDesc*)e->descs.p + nb, d_pcm, (const double*)e->winpool.p, lpc_buf[(nchunks + 1) & 1],
                                   e->lpc_overlap * (uint32_t)e->sm_count, sb));
                else
                    CK(launch_lpc3(cn, (const FrameDesc*)e->descs.p + nb, d_pcm, (const double*)e->winpool.p, lpc_buf[(nchunks + 1) & 1],
                                   e->lpc_overlap * (uint32_t)e->sm_count, sb));
                CK(cudaEventRecord(e->lpc_ev[2 * (nchunks + 1)], sb));
            }
        }
        launch_decide_scan(c, dd, (const CandRec*)e->cands.p, d_abssum, (FrameRec*)e->frecs.p, (uint32_t*)e->fbytes.p + base,
                           (unsigned long long*)e->totals.p, pipe_out ? e->d_h_totals + nchunks : nullptr, d_out, !frame_pack, st);
        time_mark(e, eb + 4);
        if (frame_pack) CK(launch_pack3(c, dd, d_pcm, (const CandRec*)e->cands.p, (const FrameRec*)e->frecs.p, d_out, d_res16, st));
        else if (fast_pack) CK(launch_pack2_crc(c, dd, d_pcm, (const CandRec*)e->cands.p, (const FrameRec*)e->frecs.p, d_out, st));
        else CK(launch_pack_crc(c, dd, (const int32_t*)e->planes.p, (const CandRec*)e->cands.p, (const FrameRec*)e->frecs.p, d_out, st));
        }
        time_mark(e, eb + 5);
        if (pipe_out) {
            // frames of this group occupy [h_totals[g - 1], h_totals[g]); the host learns the bounds one group late and
            // queues the download on the output stream while the next group is already running
            CK(cudaEventRecord(e->pipe_ev[2 * nchunks + 1], st));
            if (nchunks >= 1) {
                const size_t g = nchunks - 1;
                CK(cudaEventSynchronize(e->pipe_ev[2 * g + 1]));
                const unsigned long long a = g ? e->h_totals[g - 1] : 0, b = e->h_totals[g];
                if (b > out_capacity) pipe_overflow = true;
                else if (b > a) {
                    CK(cudaStreamWaitEvent(e->copy_out, e->pipe_ev[2 * g + 1], 0));
                    CK(cudaMemcpyAsync((uint8_t*)out + a, d_out + a, b - a, cudaMemcpyDeviceToHost, e->copy_out));
                }
            }
        }
        nchunks++;
        if (keep) {
            CK(cudaMemcpyAsync(e->info_cands.data() + base * cfg.nslots, e->cands.p, (size_t)c.nframes * cfg.nslots * sizeof(CandRec),
                               cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(e->info_frecs.data() + base, e->frecs.p, (size_t)c.nframes * sizeof(FrameRec), cudaMemcpyDeviceToHost, st));
        }
    }
    cudaEventRecord(e->ev[23], st);
    CK(cudaGetLastError());

    // ---- results ----
    unsigned long long totals[4] = {0, 0, 0, 0};
    if (fused)   // the final byte total sits in the word the last group wrote
        CK(cudaMemcpyAsync(e->totals.p, (unsigned long long*)e->totals.p + 4 + (nchunks & 1), sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(totals, e->totals.p, sizeof(totals), cudaMemcpyDeviceToHost, st));
    if (frame_bytes) {
        const size_t cnt = std::min<size_t>(frame_bytes_capacity, nframes);
        CK(cudaMemcpyAsync(frame_bytes, e->fbytes.p, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    const uint64_t total = totals[0];
    if (total_bytes_out) *total_bytes_out = total;
    if (totals[3]) return FLACB200_E_BAD_ARGUMENT;   // a frame referenced a candidate that was never encoded
    if (out) {
        if (total > out_capacity || pipe_overflow) {
            if (pipe_out) cudaStreamSynchronize(e->copy_out);
            return FLACB200_E_OUTPUT_TOO_SMALL;
        }
        if (e->profiling) cudaEventRecord(e->ev[24], st);
        if (pipe_out) {   // the last group's frames; everything before is already on its way
            const unsigned long long a = nchunks >= 2 ? e->h_totals[nchunks - 2] : 0;
            if (total > a) CK(cudaMemcpyAsync((uint8_t*)out + a, d_out + a, total - a, cudaMemcpyDeviceToHost, e->copy_out));
            CK(cudaStreamSynchronize(e->copy_out));
        } else if (!direct_out) {
            CK(cudaMemcpyAsync(out, d_out, total, out_location == FLACB200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
        }
        if (e->profiling) cudaEventRecord(e->ev[25], st);
        CK(cudaStreamSynchronize(st));
    }
    cudaEventElapsedTime(&e->tm.total_ms, e->ev[22], e->ev[23]);
    e->tm.launches = (uint32_t)(g_kernel_launches - launches0);
    if (e->profiling) {
        const uint32_t per_chunk[5] = {need_planes ? 1u : 0u, 1, 1, frame_pack ? 2u : 3u, 2u};   // planes, lpc, analyze, decide+scan(+zero), pack(+crc16)
        for (size_t ck = 0; ck < nchunks; ck++)
            for (int k = 0; k < 5; k++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, e->evpool[ck * 6 + k], e->evpool[ck * 6 + k + 1]);
                e->tm.kernel_ms[k] += ms;
                e->tm.kernel_launches[k] += per_chunk[k];
            }
        if (lpc_upfront) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e->evpool[ev_upfront], e->evpool[ev_upfront + 1]);
            e->tm.kernel_ms[1] += ms;
            e->tm.kernel_launches[1] += 1 - (uint32_t)nchunks;   // one launch instead of one per group
        }
        cudaEventElapsedTime(&e->tm.h2d_ms, e->ev[20], e->ev[21]);
        if (out) cudaEventElapsedTime(&e->tm.d2h_ms, e->ev[24], e->ev[25]);
    }
    return 0;
}

extern "C" int flacb200_encode_last_info(flacb200_engine* e, flacb200_frame_info* infos, size_t capacity, uint64_t* n_frames)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    if (n_frames) *n_frames = e->info_frames;
    if (!infos) return 0;
    const EncCfg& cfg = e->info_cfg;
    for (uint64_t f = 0; f < e->info_frames && f < capacity; f++) {
        const FrameRec& fr = e->info_frecs[f];
        flacb200_frame_info& fi = infos[f];
        memset(&fi, 0, sizeof(fi));
        fi.channel_assignment = fr.assignment;
        fi.channels = fr.nsub;
        fi.frame_bytes = fr.frame_bytes;
        for (uint32_t k = 0; k < fr.nsub; k++) {
            const CandRec& c = e->info_cands[f * cfg.nslots + fr.slot[k]];
            flacb200_subframe_info& s = fi.sub[k];
            s.type = c.type;
            s.wasted = c.wasted;
            s.bps = c.bps;
            s.bits = c.bits;
            if (c.type >= 2) {
                s.order = c.order;
                s.coding_method = c.method;
                s.partition_order = c.porder_w;
                for (uint32_t j = 0; j < c.nparts && j < 64; j++) {
                    const uint8_t r = c.rice[j];
                    s.kind[j] = r < 0x40 ? 0 : ((r & 0x40) ? 1 : 2);
                    s.rice[j] = r < 0x40 ? r : (r & 31);
                }
            }
            if (c.type == 3) {
                s.precision = c.precision;
                s.shift = c.shift;
                for (uint32_t j = 0; j < c.order; j++) s.coefs[j] = c.q[j];
            }
        }
    }
    return 0;
}

__global__ void k_copy_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t nwords)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static int mbox_reserve(flacb200_engine* e, size_t bytes)
{
    if (bytes <= e->mbox_cap) return 0;
    if (e->mbox_h) {
        cudaStreamSynchronize(e->stream);
        cudaFreeHost(e->mbox_h);
        e->mbox_h = nullptr;
        e->mbox_cap = 0;
    }
    const size_t cap = std::max<size_t>((bytes + 4095) & ~(size_t)4095, 65536);
    CK(cudaHostAlloc((void**)&e->mbox_h, cap, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&e->mbox_d, e->mbox_h, 0));
    e->mbox_cap = cap;
    return 0;
}
// the first 256 bytes of the mailbox are the download slots, uploads start at 256
constexpr size_t MBOX_UP = 256;

// bytes (a multiple of 4) of device memory -> mailbox slot `off`, on stream st; read e->mbox_h + off after syncing st
static void mbox_post(flacb200_engine* e, size_t off, const void* d_src, size_t bytes, cudaStream_t st)
{
    count_launch(), k_copy_words<<<1, 64, 0, st>>>((uint32_t*)(e->mbox_d + off), (const uint32_t*)d_src, bytes / 4);
}

// Host frames -> host PCM for many streams: the segments (independent streams) are cut into batches by bytes; the upload
// of batch b + 1, the kernels of batch b and the download of batch b - 1 overlap (three streams, one whole-size device
// buffer each for frames and PCM; every batch is an ordinary flacb200_decode call on device-resident ranges).  Only for
// calls whose segments announce their sample counts and write disjoint, increasing PCM ranges; returns -9999 when the
// call does not qualify.
static int decode_batched(flacb200_engine* e, const flacb200_stream_params* params, const uint8_t* frames, size_t frames_bytes,
                          const flacb200_decode_segment* segments, size_t n_segments, uint8_t* pcm_out, size_t pcm_out_bytes, int pcm_kind,
                          uint64_t* n_frames_out, uint64_t* n_pcm_out, uint64_t* bad_frame)
{
    size_t MIN_BYTES = (size_t)64 << 20, BATCH_BYTES = (size_t)192 << 20;
    if (e->batch_bytes) {   // FLACB200_BATCH_BYTES / "batch_bytes" (tests: small batches)
        BATCH_BYTES = e->batch_bytes;
        MIN_BYTES = 0;
    }
    if (n_segments < 4 || frames_bytes < MIN_BYTES || pcm_kind == FLACB200_PCM_I32_PLANAR || e->no_batch) return -9999;
    const size_t fb = (size_t)params->channels * (pcm_kind <= 1 ? (params->bits_per_sample + 7) / 8 : 4);
    uint64_t prev_end = 0, prev_pcm = 0;
    for (size_t s = 0; s < n_segments; s++) {
        const flacb200_decode_segment& g = segments[s];
        if (g.n_pcm_frames == 0 || g.byte_offset < prev_end || g.byte_offset + g.byte_length > frames_bytes || g.pcm_offset < prev_pcm ||
            (g.pcm_offset + g.n_pcm_frames) * fb > pcm_out_bytes)
            return -9999;
        prev_end = g.byte_offset + g.byte_length;
        prev_pcm = g.pcm_offset + g.n_pcm_frames;
    }
    struct Batch { size_t s0, s1; uint64_t lo, hi; };
    std::vector<Batch> batches;
    for (size_t s = 0; s < n_segments;) {
        Batch b{s, s, segments[s].byte_offset & ~15ull, 0};
        uint64_t bytes = 0;
        while (b.s1 < n_segments && (bytes < BATCH_BYTES || b.s1 == b.s0)) bytes += segments[b.s1++].byte_length;
        b.hi = segments[b.s1 - 1].byte_offset + segments[b.s1 - 1].byte_length;
        batches.push_back(b);
        s = b.s1;
    }
    if (batches.size() < 2) return -9999;
    CK(cudaSetDevice(e->device));
    ENS(e->dec[0], frames_bytes + 64);
    ENS(e->dec[1], pcm_out_bytes + 64);
    uint8_t* d_frames = (uint8_t*)e->dec[0].p;
    uint8_t* d_pcm = (uint8_t*)e->dec[1].p;
    while (e->batch_ev.size() < batches.size()) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->batch_ev.push_back(ev);
    }
    for (size_t b = 0; b < batches.size(); b++) {   // all uploads are queued now; each batch's kernels wait for theirs
        CK(cudaMemcpyAsync(d_frames + batches[b].lo, frames + batches[b].lo, batches[b].hi - batches[b].lo, cudaMemcpyHostToDevice, e->copy_in));
        CK(cudaEventRecord(e->batch_ev[b], e->copy_in));
    }
    uint64_t frames_total = 0, pcm_total = 0;
    int rc = 0, first_err = 0;
    std::vector<flacb200_decode_segment> sub;
    for (size_t b = 0; b < batches.size(); b++) {
        const Batch& bt = batches[b];
        sub.assign(segments + bt.s0, segments + bt.s1);
        for (auto& g : sub) g.byte_offset -= bt.lo;
        CK(cudaStreamWaitEvent(e->stream, e->batch_ev[b], 0));
        uint64_t nf = 0, ns = 0, bad = 0;
        rc = flacb200_decode(e, params, d_frames + bt.lo, bt.hi - bt.lo, FLACB200_DEVICE, sub.data(), sub.size(), d_pcm, pcm_out_bytes, pcm_kind,
                             FLACB200_DEVICE, 0, &nf, &ns, &bad);
        if (rc < 0) break;
        if (rc > 0 && first_err == 0) {   // the first error in stream order; later streams are still decoded, as in one call
            first_err = rc;
            if (bad_frame) *bad_frame = frames_total + bad;
        }
        frames_total += nf;
        pcm_total += ns;
        // the batch's PCM (the kernels are done: the call above returns after its last readback)
        const size_t a = (size_t)segments[bt.s0].pcm_offset * fb, z = (size_t)(segments[bt.s1 - 1].pcm_offset + segments[bt.s1 - 1].n_pcm_frames) * fb;
        CK(cudaMemcpyAsync(pcm_out + a, d_pcm + a, z - a, cudaMemcpyDeviceToHost, e->copy_out));
    }
    cudaStreamSynchronize(e->copy_in);
    CK(cudaStreamSynchronize(e->copy_out));
    if (n_frames_out) *n_frames_out = frames_total;
    if (n_pcm_out) *n_pcm_out = pcm_total;
    return rc < 0 ? rc : first_err;
}

extern "C" int flacb200_decode(flacb200_engine* e, const flacb200_stream_params* params, const void* frames, size_t frames_bytes,
                               int frames_location, const flacb200_decode_segment* segments, size_t n_segments, void* pcm_out,
                               size_t pcm_out_bytes, int pcm_kind, int pcm_location, uint64_t planar_stride, uint64_t* n_frames_out,
                               uint64_t* n_pcm_out, uint64_t* bad_frame)
{
    if (!e || !params || !segments || (!frames && frames_bytes) || (!pcm_out && pcm_out_bytes)) return FLACB200_E_BAD_ARGUMENT;
    if (params->channels < 1 || params->channels > 8) return 30;
    if (params->bits_per_sample < 1 || params->bits_per_sample > 32) return 33;
    if (pcm_kind < 0 || pcm_kind > 3 || n_segments > 0xFFFFFFF0ull) return FLACB200_E_BAD_ARGUMENT;
    if (n_frames_out) *n_frames_out = 0;
    if (n_pcm_out) *n_pcm_out = 0;
    if (bad_frame) *bad_frame = 0;
    if (frames_location == FLACB200_HOST && pcm_location == FLACB200_HOST) {
        const int rb = decode_batched(e, params, (const uint8_t*)frames, frames_bytes, segments, n_segments, (uint8_t*)pcm_out, pcm_out_bytes, pcm_kind,
                                      n_frames_out, n_pcm_out, bad_frame);
        if (rb != -9999) return rb;
    }
    CK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    memset(&e->tm, 0, sizeof(e->tm));

    DecCfg cfg{};
    cfg.channels = params->channels;
    cfg.bps = params->bits_per_sample;
    cfg.sample_rate = params->sample_rate;
    cfg.subset = params->subset;
    cfg.max_block_size = params->max_block_size;
    cfg.pcm_kind = (uint32_t)pcm_kind;
    cfg.bytes_per_sample = pcm_kind <= 1 ? (cfg.bps + 7) / 8 : 4;
    cfg.planar_stride = planar_stride;
    cfg.nbytes = frames_bytes;
    cfg.nseg = (uint32_t)n_segments;
    cfg.nslots = cfg.channels + ((cfg.channels == 2 && cfg.bps == 32) ? 1 : 0);
    if (pcm_kind == FLACB200_PCM_I32_PLANAR) {
        if ((size_t)planar_stride * cfg.channels * 4 > pcm_out_bytes) return FLACB200_E_BAD_ARGUMENT;
        cfg.out_samples = planar_stride;
    } else {
        cfg.out_samples = pcm_out_bytes / ((size_t)cfg.channels * cfg.bytes_per_sample);
    }
    // segments: sorted by byte offset, disjoint, inside the buffer
    std::vector<DecSeg> segs(n_segments);
    uint64_t prev_end = 0, extent = 0;
    bool extent_known = true;
    for (size_t s = 0; s < n_segments; s++) {
        const flacb200_decode_segment& g = segments[s];
        if (g.byte_offset < prev_end || g.byte_offset + g.byte_length > frames_bytes || g.byte_offset + g.byte_length < g.byte_offset)
            return FLACB200_E_BAD_ARGUMENT;
        segs[s] = DecSeg{g.byte_offset, g.byte_offset + g.byte_length, g.pcm_offset, g.n_pcm_frames};
        prev_end = g.byte_offset + g.byte_length;
        if (g.n_pcm_frames) extent = std::max<uint64_t>(extent, g.pcm_offset + g.n_pcm_frames);
        else extent_known = false;
    }
    e->last_ncand = 0;
    if (n_segments == 0 || frames_bytes == 0) return 0;

    // ---- buffers ----
    const uint8_t* d_bytes;
    if (e->profiling) cudaEventRecord(e->ev[20], st);
    if (frames_location == FLACB200_HOST || ((uintptr_t)frames & 15)) {
        ENS(e->dec[0], frames_bytes + 64);
        CK(cudaMemcpyAsync(e->dec[0].p, frames, frames_bytes, frames_location == FLACB200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
        d_bytes = (const uint8_t*)e->dec[0].p;
    } else {
        d_bytes = (const uint8_t*)frames;
    }
    uint8_t* d_out;
    if (pcm_location == FLACB200_HOST) {
        ENS(e->dec[1], pcm_out_bytes + 64);
        d_out = (uint8_t*)e->dec[1].p;
    } else {
        d_out = (uint8_t*)pcm_out;
    }
    ENS(e->dec[2], n_segments * sizeof(DecSeg));
    {
        const int rm = mbox_reserve(e, MBOX_UP + n_segments * sizeof(DecSeg));
        if (rm) return rm;
        CK(cudaStreamSynchronize(st));   // (the previous call's kernel may still be reading the upload area)
        memcpy(e->mbox_h + MBOX_UP, segs.data(), n_segments * sizeof(DecSeg));
        count_launch(), k_copy_words<<<(unsigned)std::min<size_t>((n_segments * sizeof(DecSeg) / 4 + 255) / 256, 64), 256, 0, st>>>(
            (uint32_t*)e->dec[2].p, (const uint32_t*)(e->mbox_d + MBOX_UP), n_segments * sizeof(DecSeg) / 4);
    }
    if (e->profiling) cudaEventRecord(e->ev[21], st);
    const DecSeg* d_segs = (const DecSeg*)e->dec[2].p;
    const uint32_t tiles = find_tiles(frames_bytes);
    ENS(e->dec[3], (size_t)(tiles + 2) * 4);        // per-tile counts
    ENS(e->dec[4], (size_t)(tiles + 2) * 4);        // exclusive scan (+ total)
    ENS(e->dec[5], 256);                            // [0] max block size; ChainState at +64
    CK(cudaMemsetAsync(e->dec[5].p, 0, 256, st));
    uint32_t* d_maxbs = (uint32_t*)e->dec[5].p;
    ChainState* d_state = (ChainState*)((uint8_t*)e->dec[5].p + 64);

    cudaEventRecord(e->ev[22], st);
    size_t ev = 0;
    time_mark(e, ev++);
    // frame discovery reads the bytes ONCE: every tile counts its candidates and parks them in its slots; the second pass over
    // the bytes only runs when a tile had more candidates than slots (legacy bit 256 forces it, for the tests)
    const bool park = !(e->legacy & 256u);
    if (park) {
        ENS(e->find_slots, find_slots_bytes(frames_bytes) + 64);
        launch_find_park(cfg, d_bytes, d_segs, (uint32_t*)e->dec[3].p, (uint32_t*)e->dec[4].p, (FrameCand*)e->find_slots.p, d_maxbs, st);
    } else {
        launch_find_count(cfg, d_bytes, d_segs, (uint32_t*)e->dec[3].p, (uint32_t*)e->dec[4].p, d_maxbs, st);
    }
    uint32_t ncand = 0, max_bs = 0;
    mbox_post(e, 0, (uint32_t*)e->dec[4].p + tiles, 4, st);
    mbox_post(e, 4, d_maxbs, 8, st);   // max block size, slot overflow flag
    CK(cudaStreamSynchronize(st));
    ncand = ((const uint32_t*)e->mbox_h)[0];
    max_bs = ((const uint32_t*)e->mbox_h)[1];
    const bool parked = park && ((const uint32_t*)e->mbox_h)[2] == 0;
    const unsigned long long launches0 = g_kernel_launches;
    ENS(e->dec[6], (size_t)std::max<uint32_t>(ncand, 1) * sizeof(FrameCand));
    ENS(e->dec[7], (size_t)std::max<uint32_t>(ncand, 1) * sizeof(DecRec));
    ENS(e->dec[8], (size_t)std::max<uint32_t>(ncand, 1) * sizeof(unsigned long long));
    FrameCand* d_cands = (FrameCand*)e->dec[6].p;
    DecRec* d_recs = (DecRec*)e->dec[7].p;
    unsigned long long* d_pos = (unsigned long long*)e->dec[8].p;
    if (ncand) {
        if (parked) launch_find_compact(cfg, (const uint32_t*)e->dec[3].p, (const uint32_t*)e->dec[4].p, (const FrameCand*)e->find_slots.p, d_cands, st);
        else launch_find_write(cfg, d_bytes, d_segs, (const uint32_t*)e->dec[4].p, d_cands, st);
    }
    time_mark(e, ev++);
    cfg.bstride = (std::max<uint32_t>(max_bs, 4) + 3u) & ~3u;
    // one thread decodes one frame, so a group should hold enough frames to fill the GPU; the scratch budget bounds it
    const size_t per_frame = (size_t)cfg.nslots * cfg.bstride * sizeof(int32_t);
    size_t budget = (size_t)6 << 30;
    uint32_t group = (uint32_t)std::min<size_t>(std::max<size_t>(budget / per_frame, 1), std::max<uint32_t>(ncand, 1));
    if (e->chunk_frames) group = std::min<uint32_t>(group, e->chunk_frames);
    // FLACB200_LEGACY bit 64: the thread-per-frame decoder (k_decode) for everything; default: k_parse + k_restore, and
    // k_decode only for frames with a 33-bit side channel (32-bit stereo streams)
    const bool split_decode = !(e->legacy & 64u);
    const bool maybe_wide = cfg.channels == 2 && cfg.bps == 32;
    // stage spans of the profile: {event a, event b, stage} (stages overlap in the pipelined path)
    std::vector<std::array<size_t, 3>> spans;
    spans.push_back({0, 1, 0});
    size_t ngroups = 0;
    uint32_t g0 = 0;
    const bool fuse = split_decode && ncand && restore_emit_ok(cfg, d_out) && !(e->legacy & 512u);
    if (fuse) {
        // Packed 16/24-bit mono/stereo output.  Per launch group: k_parse; CRC-16 and the frame walk (they need only the end
        // offsets k_parse found); then ONE pass over the planes that restores the predictors and the stereo pair and writes the
        // packed PCM (k_restore_emit).  The groups are software-pipelined over two plane buffers: k_parse of group g + 1 runs
        // on the engine's stream beside the tail of group g on the second stream -- the single-CTA frame walk, the host's look
        // at its verdict and the drain of every kernel are hidden behind the other stream's work.  k_restore_emit is queued
        // before the host has read the walk's verdict and checks it itself; when k_chain_fast declined the group the general
        // walk runs and the kernel is launched again; predictors beyond its register budget: k_restore + k_emit.
        cudaStream_t aux = e->aux;
        // a group should fill the GPU's 148 x 16 x 64 k_parse lanes as nearly as its frames allow: the walk of a group takes about
        // as long half full as full (2.4 ms for 73 k frames, 2.7 ms for 98 k, profiles/r02_v2_dec_summary.csv)
        size_t buf_budget = (size_t)e->plane_mb << 20;
        uint32_t pg = (uint32_t)std::min<size_t>(std::max<size_t>(buf_budget / per_frame, 1), ncand);
        if (e->chunk_frames) pg = std::min<uint32_t>(pg, e->chunk_frames);
        if (pg < ncand && (size_t)pg * 2 >= ncand) pg = (ncand + 1) / 2;   // two groups: equal halves
        const uint32_t nbuf = pg < ncand ? 2u : 1u;
        const size_t plane_stride = (size_t)((pg + 31u) & ~31u) * cfg.nslots * cfg.bstride;   // int32 per buffer
        const size_t sub_stride = (size_t)pg * cfg.channels;
        ENS(e->dec[9], nbuf * plane_stride * sizeof(int32_t));
        ENS(e->dec[10], nbuf * sub_stride * sizeof(SubRec));
        const size_t G = (ncand + pg - 1) / pg;
        while (e->pipe_ev.size() < 3 * G) {
            cudaEvent_t pe;
            CK(cudaEventCreateWithFlags(&pe, cudaEventDisableTiming));
            e->pipe_ev.push_back(pe);
        }
        uint32_t* d_clean = (uint32_t*)((uint8_t*)e->dec[5].p + 192);   // [0] k_chain_fast's verdict, [1] a predictor longer than k_restore_emit's
        auto planes_of = [&](size_t g) { return (int32_t*)e->dec[9].p + (g & (nbuf - 1)) * plane_stride; };
        auto subs_of = [&](size_t g) { return (SubRec*)e->dec[10].p + (g & (nbuf - 1)) * sub_stride; };
        auto parse = [&](size_t g) {
            const uint32_t f0 = (uint32_t)(g * pg), n = std::min<uint32_t>(pg, ncand - f0);
            if (g >= 2) cudaStreamWaitEvent(st, e->pipe_ev[3 * (g - 2) + 2], 0);   // the buffer's previous group has been emitted
            const size_t a = ev++;
            time_mark(e, a);
            launch_parse(cfg, d_bytes, d_segs, d_cands + f0, n, planes_of(g), subs_of(g), d_recs + f0, d_clean + 1, st);
            time_mark(e, ev++);
            spans.push_back({a, a + 1, 1});
            cudaEventRecord(e->pipe_ev[3 * g], st);
        };
        parse(0);
        for (size_t g = 0; g < G; g++) {
            const uint32_t f0 = (uint32_t)(g * pg), n = std::min<uint32_t>(pg, ncand - f0);
            const FrameCand* after = f0 + n < ncand ? d_cands + f0 + n : nullptr;
            CK(cudaStreamWaitEvent(aux, e->pipe_ev[3 * g], 0));
            const size_t t0 = ev;
            ev += 4;
            time_mark(e, t0, aux);
            launch_crc16f(d_bytes, d_cands + f0, n, d_recs + f0, aux);
            launch_chain_fast(cfg, d_segs, d_cands + f0, d_recs + f0, n, ncand - f0, f0 == 0, d_pos + f0, d_state, d_clean, aux);
            mbox_post(e, 8, d_clean, 8, aux);
            CK(cudaEventRecord(e->pipe_ev[3 * g + 1], aux));
            time_mark(e, t0 + 1, aux);
            launch_restore_emit(cfg, d_cands + f0, n, subs_of(g), d_recs + f0, d_pos + f0, planes_of(g), d_out, d_clean, true, aux);
            time_mark(e, t0 + 2, aux);
            if (g + 1 < G) parse(g + 1);
            CK(cudaEventSynchronize(e->pipe_ev[3 * g + 1]));
            const uint32_t clean = ((const uint32_t*)e->mbox_h)[2], high = ((const uint32_t*)e->mbox_h)[3];
            if (clean != 1 && e->debug) fprintf(stderr, "flacb200: k_chain_fast declined group at %u (reason 0x%x)\n", f0, clean);
            if (clean != 1) CK(launch_chain(cfg, d_bytes, d_segs, d_cands + f0, d_recs + f0, n, after, f0 == 0, d_pos + f0, d_state, aux));
            if (high) {
                launch_restore(cfg, d_cands + f0, n, subs_of(g), d_recs + f0, planes_of(g), aux);
                launch_emit(cfg, d_cands + f0, d_recs + f0, d_pos + f0, planes_of(g), n, d_out, aux);
            } else if (clean != 1) {
                launch_restore_emit(cfg, d_cands + f0, n, subs_of(g), d_recs + f0, d_pos + f0, planes_of(g), d_out, d_clean, false, aux);
            }
            time_mark(e, t0 + 3, aux);
            spans.push_back({t0, t0 + 1, 2});
            spans.push_back({t0 + 1, t0 + 2, 3});
            spans.push_back({t0 + 2, t0 + 3, 4});
            CK(cudaEventRecord(e->pipe_ev[3 * g + 2], aux));
            ngroups++;
        }
        CK(cudaStreamWaitEvent(st, e->pipe_ev[3 * (G - 1) + 2], 0));
        g0 = ncand;
    } else {
    ENS(e->dec[9], (size_t)((group + 31u) & ~31u) * per_frame);   // whole bundles of 32 interleaved planes
    if (split_decode) ENS(e->dec[10], (size_t)group * cfg.channels * sizeof(SubRec));
    do {
        const uint32_t n = std::min<uint32_t>(group, ncand - g0);
        const FrameCand* after = g0 + n < ncand ? d_cands + g0 + n : nullptr;
        const size_t b0 = ev - 1;   // the four stage marks of this group follow the last one recorded
        if (n && split_decode) {
            // k_parse, then two independent tails that meet before k_emit: predictor restoration over the planes on the
            // engine's stream, and CRC-16 + the frame chain (which need only the end offsets k_parse found) on a second one
            cudaStream_t aux = e->aux;
            while (e->pipe_ev.size() < 3 * (ngroups + 1)) {
                cudaEvent_t pe;
                CK(cudaEventCreateWithFlags(&pe, cudaEventDisableTiming));
                e->pipe_ev.push_back(pe);
            }
            uint32_t* d_clean = (uint32_t*)((uint8_t*)e->dec[5].p + 192);   // [0] k_chain_fast's verdict, [1] a predictor longer than k_restore_emit's
            launch_parse(cfg, d_bytes, d_segs, d_cands + g0, n, (int32_t*)e->dec[9].p, (SubRec*)e->dec[10].p, d_recs + g0, d_clean + 1, st);
            if (maybe_wide) launch_decode(cfg, d_bytes, d_segs, d_cands + g0, n, (int32_t*)e->dec[9].p, d_recs + g0, true, st);
            time_mark(e, ev++);
            // The tails: predictor restoration on the engine's stream; CRC-16 and the frame walk on the second one.  The
            // walk is k_chain_fast when the candidates are exactly the frames (the host reads its verdict: a 4-byte copy and
            // a sync of the second stream, while k_restore keeps the GPU busy), else k_chain -- one CTA that wants a whole
            // SM's shared memory and therefore only starts once k_restore has drained.
            CK(cudaEventRecord(e->pipe_ev[3 * ngroups], st));
            launch_restore(cfg, d_cands + g0, n, (const SubRec*)e->dec[10].p, d_recs + g0, (int32_t*)e->dec[9].p, st);
            CK(cudaStreamWaitEvent(aux, e->pipe_ev[3 * ngroups], 0));
            launch_crc16f(d_bytes, d_cands + g0, n, d_recs + g0, aux);
            launch_chain_fast(cfg, d_segs, d_cands + g0, d_recs + g0, n, ncand - g0, g0 == 0, d_pos + g0, d_state, d_clean, aux);
            mbox_post(e, 8, d_clean, 4, aux);
            CK(cudaStreamSynchronize(aux));
            const uint32_t clean = ((const uint32_t*)e->mbox_h)[2];
            if (clean != 1 && e->debug) fprintf(stderr, "flacb200: k_chain_fast declined group at %u (reason 0x%x)\n", g0, clean);
            if (clean != 1) CK(launch_chain(cfg, d_bytes, d_segs, d_cands + g0, d_recs + g0, n, after, g0 == 0, d_pos + g0, d_state, aux));
            CK(cudaEventRecord(e->pipe_ev[3 * ngroups + 2], aux));
            CK(cudaStreamWaitEvent(st, e->pipe_ev[3 * ngroups + 2], 0));
            time_mark(e, ev++);
            time_mark(e, ev++);
            launch_emit(cfg, d_cands + g0, d_recs + g0, d_pos + g0, (const int32_t*)e->dec[9].p, n, d_out, st);
            time_mark(e, ev++);
            for (size_t k = 0; k < 4; k++) spans.push_back({b0 + k, b0 + k + 1, 1 + k});
            ngroups++;
        } else {
            if (n) {
                launch_decode(cfg, d_bytes, d_segs, d_cands + g0, n, (int32_t*)e->dec[9].p, d_recs + g0, false, st);
                time_mark(e, ev++);
                launch_crc16f(d_bytes, d_cands + g0, n, d_recs + g0, st);
                time_mark(e, ev++);
            }
            CK(launch_chain(cfg, d_bytes, d_segs, d_cands + g0, d_recs + g0, n, after, g0 == 0, d_pos + g0, d_state, st));
            if (n) {
                time_mark(e, ev++);
                launch_emit(cfg, d_cands + g0, d_recs + g0, d_pos + g0, (const int32_t*)e->dec[9].p, n, d_out, st);
                time_mark(e, ev++);
                for (size_t k = 0; k < 4; k++) spans.push_back({b0 + k, b0 + k + 1, 1 + k});
                ngroups++;
            }
        }
        g0 += n;
    } while (g0 < ncand);
    }
    cudaEventRecord(e->ev[23], st);
    CK(cudaGetLastError());
    ChainState state;
    static_assert(sizeof(ChainState) % 4 == 0 && sizeof(ChainState) <= 128, "mailbox slot");
    mbox_post(e, 64, d_state, sizeof(state), st);
    CK(cudaStreamSynchronize(st));
    memcpy(&state, e->mbox_h + 64, sizeof(state));
    if (n_frames_out) *n_frames_out = state.frames_total;
    if (n_pcm_out) *n_pcm_out = state.samples_total;
    e->last_ncand = ncand;
    cudaEventElapsedTime(&e->tm.total_ms, e->ev[22], e->ev[23]);
    e->tm.launches = (uint32_t)(g_kernel_launches - launches0);
    if (e->profiling) {
        for (const auto& sp : spans) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e->evpool[sp[0]], e->evpool[sp[1]]);
            e->tm.kernel_ms[sp[2]] += ms;
            e->tm.kernel_launches[sp[2]] += sp[2] == 0 ? (ncand ? 3 : 2) : 1;
        }
        cudaEventElapsedTime(&e->tm.h2d_ms, e->ev[20], e->ev[21]);
    }
    if (pcm_location == FLACB200_HOST) {
        size_t bytes = pcm_out_bytes;
        if (n_segments == 1 && pcm_kind != FLACB200_PCM_I32_PLANAR)   // one stream: nothing lies behind what the walk delivered
            bytes = std::min<size_t>(bytes, (size_t)(segs[0].pcm_off + state.samples_total) * cfg.channels * cfg.bytes_per_sample);
        else if (extent_known && pcm_kind != FLACB200_PCM_I32_PLANAR)   // (+ one block: a last frame may overshoot its announced total)
            bytes = std::min<size_t>(bytes, (size_t)(extent + 65536) * cfg.channels * cfg.bytes_per_sample);
        if (e->profiling) cudaEventRecord(e->ev[24], st);
        CK(cudaMemcpyAsync(pcm_out, d_out, bytes, cudaMemcpyDeviceToHost, st));
        if (e->profiling) cudaEventRecord(e->ev[25], st);
        CK(cudaStreamSynchronize(st));
        if (e->profiling) cudaEventElapsedTime(&e->tm.d2h_ms, e->ev[24], e->ev[25]);
    }
    if (state.err) {
        if (bad_frame) *bad_frame = state.err_frame;
        return state.err == 0x80000000u ? FLACB200_E_OUTPUT_TOO_SMALL : (int)state.err;
    }
    return 0;
}

extern "C" int flacb200_decode_last_frames(flacb200_engine* e, flacb200_frame_entry* table, size_t capacity, uint64_t* n_entries)
{
    if (!e || !n_entries || (!table && capacity)) return FLACB200_E_BAD_ARGUMENT;
    *n_entries = 0;
    const uint32_t n = e->last_ncand;
    if (n == 0) return 0;
    CK(cudaSetDevice(e->device));
    std::vector<FrameCand> cands(n);
    std::vector<DecRec> recs(n);
    std::vector<unsigned long long> pos(n);
    CK(cudaMemcpyAsync(cands.data(), e->dec[6].p, (size_t)n * sizeof(FrameCand), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(recs.data(), e->dec[7].p, (size_t)n * sizeof(DecRec), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(pos.data(), e->dec[8].p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    uint64_t k = 0;
    for (uint32_t c = 0; c < n; c++) {
        if (pos[c] == ~0ull || recs[c].err) continue;   // a false candidate, or a frame the walk never reached
        if (k < capacity) table[k] = flacb200_frame_entry{cands[c].off, pos[c], (uint32_t)(recs[c].end - cands[c].off), cands[c].block_size};
        k++;
    }
    *n_entries = k;
    return 0;
}

extern "C" int flacb200_md5_batch(flacb200_engine* e, const void* pcm, size_t pcm_bytes, int pcm_kind, int pcm_location, uint64_t planar_stride,
                            uint32_t channels, uint32_t bits_per_sample, const flacb200_segment* segments, size_t n_segments, uint8_t* digests)
{
    if (!e || !segments || !digests || (!pcm && pcm_bytes)) return FLACB200_E_BAD_ARGUMENT;
    if (channels < 1 || channels > 8) return 30;                 // ExcessiveChannels
    if (bits_per_sample < 1 || bits_per_sample > 32) return 33;  // InvalidBitsPerSample
    if (pcm_kind < 0 || pcm_kind > 3 || n_segments > 0xFFFFFFF0ull) return FLACB200_E_BAD_ARGUMENT;
    if (n_segments == 0) return 0;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    Md5Cfg cfg{};
    cfg.channels = channels;
    cfg.bytes_per_sample = (bits_per_sample + 7) / 8;   // of the MESSAGE (src/encode.rs:1292); int32 input is narrowed to it
    cfg.pcm_kind = (uint32_t)pcm_kind;
    cfg.nseg = (uint32_t)n_segments;
    cfg.planar_stride = planar_stride;
    const size_t in_sample_bytes = pcm_kind <= 1 ? cfg.bytes_per_sample : 4;
    std::vector<Md5Seg> segs(n_segments);
    for (size_t s = 0; s < n_segments; s++) {
        const uint64_t end = segments[s].pcm_offset + segments[s].n_pcm_frames;
        if (end < segments[s].pcm_offset) return FLACB200_E_BAD_ARGUMENT;
        if (pcm_kind == FLACB200_PCM_I32_PLANAR) {
            if (end > planar_stride || (size_t)planar_stride * channels * 4 > pcm_bytes) return FLACB200_E_BAD_ARGUMENT;
        } else if (end * channels * in_sample_bytes > pcm_bytes) {
            return FLACB200_E_BAD_ARGUMENT;
        }
        segs[s] = Md5Seg{segments[s].pcm_offset, segments[s].n_pcm_frames};
    }
    const uint8_t* d_pcm = (const uint8_t*)pcm;
    if (pcm_location == FLACB200_HOST) {
        ENS(e->pcm, pcm_bytes + 16);
        CK(cudaMemcpyAsync(e->pcm.p, pcm, pcm_bytes, cudaMemcpyHostToDevice, st));
        d_pcm = (const uint8_t*)e->pcm.p;
    }
    ENS(e->dec[11], n_segments * (sizeof(Md5Seg) + 16));
    Md5Seg* d_segs = (Md5Seg*)e->dec[11].p;
    uint32_t* d_dig = (uint32_t*)((uint8_t*)e->dec[11].p + n_segments * sizeof(Md5Seg));
    CK(cudaMemcpyAsync(d_segs, segs.data(), n_segments * sizeof(Md5Seg), cudaMemcpyHostToDevice, st));
    cudaEventRecord(e->ev[22], st);
    CK(launch_md5(cfg, d_pcm, d_segs, d_dig, st));
    cudaEventRecord(e->ev[23], st);
    CK(cudaMemcpyAsync(digests, d_dig, n_segments * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memset(&e->tm, 0, sizeof(e->tm));
    cudaEventElapsedTime(&e->tm.total_ms, e->ev[22], e->ev[23]);
    e->tm.launches = 1;
    return 0;
}

// the device build of glibc_log.cuh over an array (parity tooling: tests/test_gpu_libm.py compares it with the C library)
__global__ void k_debug_libm(int fn, const double* __restrict__ in, double* __restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fn == 0 ? glibc_log(in[i]) : glibc_log2(in[i]);
}

extern "C" int flacb200_debug_libm(flacb200_engine* e, int fn, const double* in, double* out, size_t n)
{
    if (!e || !in || !out || fn < 0 || fn > 1) return FLACB200_E_BAD_ARGUMENT;
    if (n == 0) return 0;
    CK(cudaSetDevice(e->device));
    ENS(e->scratch, 2 * n * sizeof(double));
    double* d_in = (double*)e->scratch.p;
    double* d_out = d_in + n;
    CK(cudaMemcpyAsync(d_in, in, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    count_launch(), k_debug_libm<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, e->stream>>>(fn, d_in, d_out, n);
    CK(cudaMemcpyAsync(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int flacb200_synth_pcm(flacb200_engine* e, void* pcm_device, uint64_t first_track, uint64_t n_tracks, uint64_t n_pcm_frames,
                                  uint32_t channels, uint32_t sample_rate, uint32_t bits_per_sample, uint64_t seed)
{
    if (!e || !pcm_device || channels < 1 || channels > 8 || bits_per_sample < 8 || bits_per_sample > 32) return FLACB200_E_BAD_ARGUMENT;
    CK(cudaSetDevice(e->device));
    if (!e->lut.p) {
        ENS(e->lut, 4096 * sizeof(int32_t));
        std::vector<int32_t> lut(4096);
        for (int k = 0; k < 4096; k++) lut[k] = (int32_t)llround(sin(2.0 * M_PI * (double)k / 4096.0) * (double)((1 << 30) - 1));
        CK(cudaMemcpy(e->lut.p, lut.data(), lut.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    CK(launch_synth((uint8_t*)pcm_device, first_track, n_tracks, n_pcm_frames, channels, sample_rate, bits_per_sample, seed,
                    (const int32_t*)e->lut.p, e->stream));
    return 0;
}

extern "C" void* flacb200_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void flacb200_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

extern "C" void* flacb200_device_alloc(flacb200_engine* e, size_t bytes)
{
    if (!e) return nullptr;
    cudaSetDevice(e->device);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void flacb200_device_free(flacb200_engine* e, void* p)
{
    if (e && p) {
        cudaSetDevice(e->device);
        cudaFree(p);
    }
}

extern "C" int flacb200_memcpy(flacb200_engine* e, void* dst, const void* src, size_t bytes, int kind)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    CK(cudaSetDevice(e->device));
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(dst, src, bytes, k, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int flacb200_synchronize(flacb200_engine* e)
{
    if (!e) return FLACB200_E_BAD_ARGUMENT;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" const char* flacb200_strerror(int code)
{
    static const char* names[] = {"Ok", "Io", "Utf8", "MissingFlacTag", "MissingStreaminfo", "MultipleStreaminfo", "MultipleSeekTable",
                                  "MultipleVorbisComment", "InvalidSeekTableSize", "InvalidSeekTablePoint", "Cuesheet", "InvalidPictureType",
                                  "MultiplePngIcon", "MultipleGeneralIcon", "ReservedMetadataBlock", "InvalidMetadataBlock",
                                  "InvalidMetadataBlockSize", "InsufficientApplicationBlock", "ExcessiveVorbisEntries", "ExcessiveStringLength",
                                  "ExcessivePictureSize", "ShortBlock", "ExcessiveBlockSize", "InvalidSyncCode", "InvalidBlockSize",
                                  "BlockSizeMismatch", "InvalidSampleRate", "NonSubsetSampleRate", "NonSubsetBitsPerSample", "SampleRateMismatch",
                                  "ExcessiveChannels", "InvalidChannels", "ChannelsMismatch", "InvalidBitsPerSample", "ExcessiveBps",
                                  "BitsPerSampleMismatch", "InvalidFrameNumber", "InvalidSeek", "ExcessiveFrameNumber", "Crc8Mismatch",
                                  "Crc16Mismatch", "InvalidSubframeHeader", "InvalidSubframeHeaderType", "ExcessiveWastedBits", "MissingResiduals",
                                  "InvalidCodingMethod", "InvalidPartitionOrder", "InvalidFixedOrder", "InvalidLpcOrder", "InvalidQlpPrecision",
                                  "NegativeLpcShift", "NoBestLpcOrder", "InsufficientLpcSamples", "ZeroLpCoefficients", "LpNegativeShiftError",
                                  "AccumulatorOverflow", "TooManySamples", "ExcessiveTotalSamples", "NoSamples", "SampleCountMismatch",
                                  "ResidualOverflow", "SamplesNotDivisibleByChannels", "InvalidTotalBytes", "InvalidTotalSamples",
                                  "ChannelCountMismatch", "ChannelLengthMismatch"};
    if (code >= 0 && code <= 65) return names[code];
    switch (code) {
    case FLACB200_E_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case FLACB200_E_BAD_ARGUMENT: return "bad argument";
    case FLACB200_E_OUT_OF_MEMORY: return "out of device memory";
    case FLACB200_E_OUTPUT_TOO_SMALL: return "output buffer too small";
    default: break;
    }
    if (code <= FLACB200_E_CUDA_BASE) return cudaGetErrorString((cudaError_t)(FLACB200_E_CUDA_BASE - code));
    return "unknown error";
}

extern "C" const char* flacb200_version(void) { return "flacb200 0.1.0 (sm_100a)"; }
