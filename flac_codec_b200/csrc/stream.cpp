// stream.cpp -- stream-level host layer of libflacb200.so (include/flacb200_stream.h).
//
// The reference's writer facades feed one block at a time into Encoder::encode (src/encode.rs:1997); here the
// same facade state (partial-block buffer, MD5, seek points, STREAMINFO bookkeeping) sits in front of the batch
// engine: blocks are collected in pinned memory and encoded `launch_frames` at a time by flacb200_encode, while
// the MD5 of the same bytes is computed on a host thread.  The reader parses the metadata blocks on the host and
// decodes all frames of the file image in one flacb200_decode call.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <mutex>
#include <vector>

#include "../../include/flacb200_stream.h"

namespace {

// flac_codec::Error ordinals used here (src/lib.rs:57-193, 1-based; see flacb200_strerror)
enum : int {
    E_IO = 1, E_MISSING_FLAC_TAG = 3, E_MISSING_STREAMINFO = 4, E_MULTIPLE_STREAMINFO = 5, E_MULTIPLE_SEEKTABLE = 6,
    E_INVALID_SEEKTABLE_SIZE = 8, E_INVALID_SEEKTABLE_POINT = 9, E_RESERVED_METADATA_BLOCK = 14, E_INVALID_METADATA_BLOCK = 15,
    E_INVALID_METADATA_BLOCK_SIZE = 16, E_INVALID_SAMPLE_RATE = 26, E_EXCESSIVE_CHANNELS = 30, E_INVALID_BPS = 33,
    E_INVALID_SEEK = 37, E_EXCESSIVE_TOTAL_SAMPLES = 57, E_NO_SAMPLES = 58, E_SAMPLE_COUNT_MISMATCH = 59,
    E_SAMPLES_NOT_DIVISIBLE = 61, E_INVALID_TOTAL_BYTES = 62, E_INVALID_TOTAL_SAMPLES = 63, E_CHANNEL_COUNT_MISMATCH = 64,
    E_CHANNEL_LENGTH_MISMATCH = 65,
};

constexpr uint64_t MAX_SAMPLES = 1ull << 36;             // Encoder::MAX_SAMPLES: STREAMINFO keeps 36 bits
constexpr size_t MAX_SEEK_POINTS = (1u << 24) / 18;      // SeekTable::MAX_POINTS (src/metadata/mod.rs:1989)
constexpr uint32_t MAX_FRAME_SIZE = (1u << 24) - 1;      // Streaminfo::MAX_FRAME_SIZE

// ---- MD5 (RFC 1321), the sum STREAMINFO stores over the little-endian PCM (src/encode.rs:1292-1318) ----
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    uint64_t total = 0;
    uint8_t tail[64];
    uint32_t ntail = 0;

    static inline uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

    void block(const uint8_t* p)
    {
        // fully unrolled rounds: the per-stream MD5 is the serial tail of a single-stream encode (the GPU work of a 60 s
        // track is shorter than hashing its 10 MB), so the constants and rotations are compile-time here
        uint32_t m[16];
        memcpy(m, p, 64);   // little-endian host
        uint32_t A = a, B = b, C = c, D = d;
#define FLACB200_MD5_STEP(f, w, x, y, z, k, t, sft) \
    w += f(x, y, z) + m[k] + t;                     \
    w = rol(w, sft) + x;
#define FLACB200_F(x, y, z) (z ^ (x & (y ^ z)))
#define FLACB200_G(x, y, z) (y ^ (z & (x ^ y)))
#define FLACB200_MD5_H(x, y, z) (x ^ y ^ z)
#define FLACB200_I(x, y, z) (y ^ (x | ~z))
        FLACB200_MD5_STEP(FLACB200_F, A, B, C, D, 0, 0xd76aa478u, 7)  FLACB200_MD5_STEP(FLACB200_F, D, A, B, C, 1, 0xe8c7b756u, 12)
        FLACB200_MD5_STEP(FLACB200_F, C, D, A, B, 2, 0x242070dbu, 17) FLACB200_MD5_STEP(FLACB200_F, B, C, D, A, 3, 0xc1bdceeeu, 22)
        FLACB200_MD5_STEP(FLACB200_F, A, B, C, D, 4, 0xf57c0fafu, 7)  FLACB200_MD5_STEP(FLACB200_F, D, A, B, C, 5, 0x4787c62au, 12)
        FLACB200_MD5_STEP(FLACB200_F, C, D, A, B, 6, 0xa8304613u, 17) FLACB200_MD5_STEP(FLACB200_F, B, C, D, A, 7, 0xfd469501u, 22)
        FLACB200_MD5_STEP(FLACB200_F, A, B, C, D, 8, 0x698098d8u, 7)  FLACB200_MD5_STEP(FLACB200_F, D, A, B, C, 9, 0x8b44f7afu, 12)
        FLACB200_MD5_STEP(FLACB200_F, C, D, A, B, 10, 0xffff5bb1u, 17) FLACB200_MD5_STEP(FLACB200_F, B, C, D, A, 11, 0x895cd7beu, 22)
        FLACB200_MD5_STEP(FLACB200_F, A, B, C, D, 12, 0x6b901122u, 7) FLACB200_MD5_STEP(FLACB200_F, D, A, B, C, 13, 0xfd987193u, 12)
        FLACB200_MD5_STEP(FLACB200_F, C, D, A, B, 14, 0xa679438eu, 17) FLACB200_MD5_STEP(FLACB200_F, B, C, D, A, 15, 0x49b40821u, 22)
        FLACB200_MD5_STEP(FLACB200_G, A, B, C, D, 1, 0xf61e2562u, 5)  FLACB200_MD5_STEP(FLACB200_G, D, A, B, C, 6, 0xc040b340u, 9)
        FLACB200_MD5_STEP(FLACB200_G, C, D, A, B, 11, 0x265e5a51u, 14) FLACB200_MD5_STEP(FLACB200_G, B, C, D, A, 0, 0xe9b6c7aau, 20)
        FLACB200_MD5_STEP(FLACB200_G, A, B, C, D, 5, 0xd62f105du, 5)  FLACB200_MD5_STEP(FLACB200_G, D, A, B, C, 10, 0x02441453u, 9)
        FLACB200_MD5_STEP(FLACB200_G, C, D, A, B, 15, 0xd8a1e681u, 14) FLACB200_MD5_STEP(FLACB200_G, B, C, D, A, 4, 0xe7d3fbc8u, 20)
        FLACB200_MD5_STEP(FLACB200_G, A, B, C, D, 9, 0x21e1cde6u, 5)  FLACB200_MD5_STEP(FLACB200_G, D, A, B, C, 14, 0xc33707d6u, 9)
        FLACB200_MD5_STEP(FLACB200_G, C, D, A, B, 3, 0xf4d50d87u, 14) FLACB200_MD5_STEP(FLACB200_G, B, C, D, A, 8, 0x455a14edu, 20)
        FLACB200_MD5_STEP(FLACB200_G, A, B, C, D, 13, 0xa9e3e905u, 5) FLACB200_MD5_STEP(FLACB200_G, D, A, B, C, 2, 0xfcefa3f8u, 9)
        FLACB200_MD5_STEP(FLACB200_G, C, D, A, B, 7, 0x676f02d9u, 14) FLACB200_MD5_STEP(FLACB200_G, B, C, D, A, 12, 0x8d2a4c8au, 20)
        FLACB200_MD5_STEP(FLACB200_MD5_H, A, B, C, D, 5, 0xfffa3942u, 4)  FLACB200_MD5_STEP(FLACB200_MD5_H, D, A, B, C, 8, 0x8771f681u, 11)
        FLACB200_MD5_STEP(FLACB200_MD5_H, C, D, A, B, 11, 0x6d9d6122u, 16) FLACB200_MD5_STEP(FLACB200_MD5_H, B, C, D, A, 14, 0xfde5380cu, 23)
        FLACB200_MD5_STEP(FLACB200_MD5_H, A, B, C, D, 1, 0xa4beea44u, 4)  FLACB200_MD5_STEP(FLACB200_MD5_H, D, A, B, C, 4, 0x4bdecfa9u, 11)
        FLACB200_MD5_STEP(FLACB200_MD5_H, C, D, A, B, 7, 0xf6bb4b60u, 16) FLACB200_MD5_STEP(FLACB200_MD5_H, B, C, D, A, 10, 0xbebfbc70u, 23)
        FLACB200_MD5_STEP(FLACB200_MD5_H, A, B, C, D, 13, 0x289b7ec6u, 4) FLACB200_MD5_STEP(FLACB200_MD5_H, D, A, B, C, 0, 0xeaa127fau, 11)
        FLACB200_MD5_STEP(FLACB200_MD5_H, C, D, A, B, 3, 0xd4ef3085u, 16) FLACB200_MD5_STEP(FLACB200_MD5_H, B, C, D, A, 6, 0x04881d05u, 23)
        FLACB200_MD5_STEP(FLACB200_MD5_H, A, B, C, D, 9, 0xd9d4d039u, 4)  FLACB200_MD5_STEP(FLACB200_MD5_H, D, A, B, C, 12, 0xe6db99e5u, 11)
        FLACB200_MD5_STEP(FLACB200_MD5_H, C, D, A, B, 15, 0x1fa27cf8u, 16) FLACB200_MD5_STEP(FLACB200_MD5_H, B, C, D, A, 2, 0xc4ac5665u, 23)
        FLACB200_MD5_STEP(FLACB200_I, A, B, C, D, 0, 0xf4292244u, 6)  FLACB200_MD5_STEP(FLACB200_I, D, A, B, C, 7, 0x432aff97u, 10)
        FLACB200_MD5_STEP(FLACB200_I, C, D, A, B, 14, 0xab9423a7u, 15) FLACB200_MD5_STEP(FLACB200_I, B, C, D, A, 5, 0xfc93a039u, 21)
        FLACB200_MD5_STEP(FLACB200_I, A, B, C, D, 12, 0x655b59c3u, 6) FLACB200_MD5_STEP(FLACB200_I, D, A, B, C, 3, 0x8f0ccc92u, 10)
        FLACB200_MD5_STEP(FLACB200_I, C, D, A, B, 10, 0xffeff47du, 15) FLACB200_MD5_STEP(FLACB200_I, B, C, D, A, 1, 0x85845dd1u, 21)
        FLACB200_MD5_STEP(FLACB200_I, A, B, C, D, 8, 0x6fa87e4fu, 6)  FLACB200_MD5_STEP(FLACB200_I, D, A, B, C, 15, 0xfe2ce6e0u, 10)
        FLACB200_MD5_STEP(FLACB200_I, C, D, A, B, 6, 0xa3014314u, 15) FLACB200_MD5_STEP(FLACB200_I, B, C, D, A, 13, 0x4e0811a1u, 21)
        FLACB200_MD5_STEP(FLACB200_I, A, B, C, D, 4, 0xf7537e82u, 6)  FLACB200_MD5_STEP(FLACB200_I, D, A, B, C, 11, 0xbd3af235u, 10)
        FLACB200_MD5_STEP(FLACB200_I, C, D, A, B, 2, 0x2ad7d2bbu, 15) FLACB200_MD5_STEP(FLACB200_I, B, C, D, A, 9, 0xeb86d391u, 21)
#undef FLACB200_MD5_STEP
#undef FLACB200_F
#undef FLACB200_G
#undef FLACB200_MD5_H
#undef FLACB200_I
        a += A; b += B; c += C; d += D;
    }

    void update(const uint8_t* p, size_t n)
    {
        total += n;
        if (ntail) {
            const size_t take = std::min<size_t>(64 - ntail, n);
            memcpy(tail + ntail, p, take);
            ntail += (uint32_t)take;
            p += take;
            n -= take;
            if (ntail < 64) return;
            block(tail);
            ntail = 0;
        }
        while (n >= 64) {
            block(p);
            p += 64;
            n -= 64;
        }
        if (n) {
            memcpy(tail, p, n);
            ntail = (uint32_t)n;
        }
    }

    void final(uint8_t out[16]) const   // does not disturb the running state (md5.clone().finalize(), :2100)
    {
        Md5 t = *this;
        const uint64_t bits = t.total * 8;
        uint8_t pad[72] = {0x80};
        const size_t padlen = (t.ntail < 56 ? 56 : 120) - t.ntail;
        t.update(pad, padlen);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (8 * i));
        t.update(len, 8);
        const uint32_t v[4] = {t.a, t.b, t.c, t.d};
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(v[i] >> (8 * k));
    }
};

void put_be(uint8_t* p, uint64_t v, int bytes)
{
    for (int i = 0; i < bytes; i++) p[i] = (uint8_t)(v >> (8 * (bytes - 1 - i)));
}

uint64_t get_be(const uint8_t* p, int bytes)
{
    uint64_t v = 0;
    for (int i = 0; i < bytes; i++) v = (v << 8) | p[i];
    return v;
}

struct SeekPt {
    uint64_t sample_offset, byte_offset;
    uint32_t frame_samples;
};

// growable byte buffer, pinned when a device is present (cudaMemcpyAsync from pageable memory is staged and slow)
// Pinning host memory costs milliseconds (cudaHostAlloc of the 15 MB a one-minute stream decodes to: 3-5 ms, more than the
// decode itself), and readers/writers are opened and closed per stream: released pinned buffers are parked here and handed
// to the next handle that asks for about that size.
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<uint8_t*, size_t>> parked;
    size_t parked_bytes = 0;
    static constexpr size_t LIMIT = (size_t)512 << 20;
    uint8_t* take(size_t want, size_t* cap)
    {
        std::lock_guard<std::mutex> g(mu);
        size_t best = parked.size();
        for (size_t i = 0; i < parked.size(); i++)
            if (parked[i].second >= want && parked[i].second <= want * 2 + (1 << 20) && (best == parked.size() || parked[i].second < parked[best].second)) best = i;
        if (best == parked.size()) return nullptr;
        uint8_t* p = parked[best].first;
        *cap = parked[best].second;
        parked_bytes -= parked[best].second;
        parked.erase(parked.begin() + (long)best);
        return p;
    }
    void give(uint8_t* p, size_t cap)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            if (parked_bytes + cap <= LIMIT && parked.size() < 64) {
                parked.emplace_back(p, cap);
                parked_bytes += cap;
                return;
            }
        }
        flacb200_host_free(p);
    }
};
static PinnedPool g_pinned_pool;

struct HostBuf {
    uint8_t* p = nullptr;
    size_t cap = 0, len = 0;
    bool pinned = false;
    ~HostBuf() { release(); }
    void release()
    {
        if (!p) return;
        if (pinned) g_pinned_pool.give(p, cap);
        else free(p);
        p = nullptr;
        cap = len = 0;
    }
    bool reserve(size_t want, bool try_pinned)
    {
        if (want <= cap) return true;
        size_t ncap = std::max<size_t>(want, cap + cap / 2 + 4096);
        uint8_t* np = nullptr;
        bool npinned = false;
        if (try_pinned) {
            np = g_pinned_pool.take(ncap, &ncap);
            if (!np) np = (uint8_t*)flacb200_host_alloc(ncap);
            npinned = np != nullptr;
        }
        if (!np) np = (uint8_t*)malloc(ncap);
        if (!np) return false;
        if (len) memcpy(np, p, len);
        const size_t keep = len;
        release();
        p = np;
        cap = ncap;
        len = keep;
        pinned = npinned;
        return true;
    }
};

}   // namespace

// =================================================================================================
// writer
// =================================================================================================
struct flacb200_writer {
    flacb200_engine* engine = nullptr;
    flacb200_writer_options opt{};
    uint32_t rate = 0, bps = 0, channels = 0, bytes_per_sample = 0, block_size = 0, launch_frames = 0;
    uint64_t total = 0;             // expected inter-channel samples, 0 = unknown
    size_t pcm_frame_bytes = 0, block_bytes = 0;
    HostBuf pending;                // little-endian packed PCM not yet encoded (whole samples)
    uint8_t partial[8] = {0};            // bytes of an incomplete sample handed to write_bytes
    uint32_t npartial = 0;
    HostBuf stage;                  // output of one flacb200_encode call
    std::vector<uint8_t> ready;     // frames not yet drained
    std::vector<uint8_t> drained;   // storage behind the pointer returned by drain
    std::vector<uint32_t> sizes;
    std::vector<SeekPt> points;     // one per frame (Encoder::encode pushes one per frame, :1999)
    std::vector<uint8_t> header;
    size_t n_placeholders = 0;
    bool have_seektable = false, finalized = false, failed = false;
    uint64_t pcm_frames_encoded = 0, frames = 0, frame_bytes = 0, next_frame_number = 0;
    uint32_t min_frame = 0, max_frame = 0, launches = 0;
    Md5 md5;
    uint8_t md5_final[16] = {0};
    bool md5_known = false;
};

namespace {

// SeekTableInterval::filter (src/encode.rs:1338-1358)
template <class GetPoint>
void seek_filter(const flacb200_writer_options& o, uint32_t rate, size_t n, GetPoint get, std::vector<size_t>& keep)
{
    keep.clear();
    if (o.seektable_kind == 1) {
        const uint64_t nth = (uint64_t)((uint32_t)(uint8_t)o.seektable_n * rate);
        uint64_t offset = 0;
        for (size_t i = 0; i < n; i++) {
            const SeekPt p = get(i);
            if (offset >= p.sample_offset && offset < p.sample_offset + p.frame_samples) {
                offset += nth;
                keep.push_back(i);
            }
        }
    } else if (o.seektable_kind == 2) {
        const size_t step = o.seektable_n ? o.seektable_n : 1;
        for (size_t i = 0; i < n; i += step) keep.push_back(i);
    }
}

void put_seekpoint(uint8_t* p, const SeekPt* s)
{
    if (s) {
        put_be(p, s->sample_offset, 8);
        put_be(p + 8, s->byte_offset, 8);
        put_be(p + 16, s->frame_samples, 2);
    } else {   // SeekPoint::Placeholder
        put_be(p, ~0ull, 8);
        put_be(p + 8, 0, 8);
        put_be(p + 16, 0, 2);
    }
}

// write_blocks (src/metadata/mod.rs:904-976): "fLaC", STREAMINFO, [SEEKTABLE], [PADDING] in the order Encoder::new
// sorts them (:1944-1951); at finalize the defined seek points replace the placeholders (:2041-2051) or, when no
// placeholder table exists, a table is carved out of the padding and appended after it (:2053-2072).
void build_header(flacb200_writer& w, bool final_pass)
{
    const flacb200_writer_options& o = w.opt;
    const bool padding = o.padding > 0;
    size_t pad_body = padding ? (size_t)o.padding : 0;
    std::vector<size_t> keep;
    bool table_after_padding = false;
    if (final_pass && o.seektable_kind) {
        seek_filter(o, w.rate, w.points.size(), [&](size_t i) { return w.points[i]; }, keep);
        if (!w.have_seektable && padding) {
            if (keep.size() > MAX_SEEK_POINTS) keep.resize(MAX_SEEK_POINTS);
            const size_t table_size = 4 + 18 * keep.size();   // MetadataBlock::total_size(): header + body
            if (pad_body >= table_size) {
                table_after_padding = true;
                pad_body -= table_size;
            }
        }
    }
    std::vector<uint8_t>& h = w.header;
    h.clear();
    h.insert(h.end(), {'f', 'L', 'a', 'C'});
    auto block_header = [&](uint8_t type, bool last, size_t body) {
        uint8_t b[4];
        b[0] = (uint8_t)((last ? 0x80 : 0) | type);
        put_be(b + 1, body, 3);
        h.insert(h.end(), b, b + 4);
    };
    // STREAMINFO (src/metadata/mod.rs:1742-1760)
    block_header(0, !(w.have_seektable || padding), 34);
    {
        uint8_t b[34];
        put_be(b, w.block_size, 2);
        put_be(b + 2, w.block_size, 2);
        put_be(b + 4, w.min_frame, 3);
        put_be(b + 7, w.max_frame, 3);
        const uint64_t total = final_pass ? w.pcm_frames_encoded : w.total;
        const uint64_t v = ((uint64_t)w.rate << 44) | ((uint64_t)(w.channels - 1) << 41) | ((uint64_t)(w.bps - 1) << 36) | (total & 0xFFFFFFFFFull);
        put_be(b + 10, v, 8);
        if (w.md5_known) memcpy(b + 18, w.md5_final, 16);
        else memset(b + 18, 0, 16);
        h.insert(h.end(), b, b + 34);
    }
    if (w.have_seektable) {
        block_header(3, !padding, 18 * w.n_placeholders);
        const size_t at = h.size();
        h.resize(at + 18 * w.n_placeholders);
        for (size_t i = 0; i < w.n_placeholders; i++)
            put_seekpoint(h.data() + at + 18 * i, (final_pass && i < keep.size()) ? &w.points[keep[i]] : nullptr);
    }
    if (padding) {
        block_header(1, !table_after_padding, pad_body);
        h.resize(h.size() + pad_body, 0);
        if (table_after_padding) {
            block_header(3, true, 18 * keep.size());
            const size_t at = h.size();
            h.resize(at + 18 * keep.size());
            for (size_t i = 0; i < keep.size(); i++) put_seekpoint(h.data() + at + 18 * i, &w.points[keep[i]]);
        }
    }
}

int writer_reserve_pending(flacb200_writer& w, size_t extra)
{
    if (!w.pending.reserve(w.pending.len + extra, w.engine != nullptr)) return FLACB200_E_OUT_OF_MEMORY;
    return 0;
}

// Encodes the first n_pcm inter-channel samples of `pending` (whole blocks, or the final short block).
int writer_encode(flacb200_writer& w, uint64_t n_pcm)
{
    if (n_pcm == 0) return 0;
    if (!w.engine) return FLACB200_E_NO_DEVICE;
    const size_t nbytes = (size_t)n_pcm * w.pcm_frame_bytes;
    const uint64_t nblocks = (n_pcm + w.block_size - 1) / w.block_size;
    flacb200_stream_params prm{};
    prm.sample_rate = w.rate;
    prm.bits_per_sample = w.bps;
    prm.channels = w.channels;
    flacb200_segment seg{0, n_pcm, w.next_frame_number};
    const size_t bound = flacb200_encode_bound(&w.opt.frame, &prm, &seg, 1);
    if (!w.stage.reserve(bound, true)) return FLACB200_E_OUT_OF_MEMORY;
    const size_t first = w.sizes.size();
    w.sizes.resize(first + nblocks);
    // MD5 of exactly these bytes, concurrently with the GPU (FlacByteWriter::write :369, update_md5 :1292)
    std::future<void> md5_job = std::async(std::launch::async, [&w, nbytes] { w.md5.update(w.pending.p, nbytes); });
    uint64_t nf = 0, total = 0;
    flacb200_engine_set_keep_info(w.engine, 0);
    const int rc = flacb200_encode(w.engine, &w.opt.frame, &prm, w.pending.p, nbytes, FLACB200_PCM_BYTES_LE, FLACB200_HOST, 0, &seg, 1,
                                   w.stage.p, w.stage.cap, FLACB200_HOST, w.sizes.data() + first, nblocks, &nf, &total);
    md5_job.get();
    w.launches++;
    if (rc) {
        w.sizes.resize(first);
        return rc;
    }
    w.ready.insert(w.ready.end(), w.stage.p, w.stage.p + total);
    uint64_t done = 0;
    for (uint64_t f = 0; f < nf; f++) {
        const uint32_t s = w.sizes[first + f];
        const uint32_t n = (uint32_t)std::min<uint64_t>(w.block_size, n_pcm - done);
        w.points.push_back(SeekPt{w.pcm_frames_encoded + done, w.frame_bytes, n});   // :1999-2003 (offsets count from the first frame)
        w.frame_bytes += s;
        if (s < MAX_FRAME_SIZE && s != 0) {   // encode_frame tail (:2413-2436)
            w.min_frame = w.min_frame == 0 ? s : std::min(w.min_frame, s);
            w.max_frame = w.max_frame == 0 ? s : std::max(w.max_frame, s);
        }
        done += n;
    }
    w.pcm_frames_encoded += n_pcm;
    w.frames += nf;
    w.next_frame_number += nf;
    // keep what was not encoded
    const size_t rest = w.pending.len - nbytes;
    if (rest) memmove(w.pending.p, w.pending.p + nbytes, rest);
    w.pending.len = rest;
    return 0;
}

// Encoder::encode's running-total check (:2006-2011), applied to the blocks that become complete with this write
int writer_after_append(flacb200_writer& w, bool force)
{
    const uint64_t whole = w.pending.len / w.block_bytes;
    if (w.total && w.pcm_frames_encoded + whole * w.block_size > w.total) {
        w.failed = true;
        return E_EXCESSIVE_TOTAL_SAMPLES;
    }
    if (whole && (force || whole >= w.launch_frames)) {
        const int rc = writer_encode(w, whole * w.block_size);
        if (rc) w.failed = true;
        return rc;
    }
    return 0;
}

}   // namespace

extern "C" {

void flacb200_writer_options_default(flacb200_writer_options* o)
{
    memset(o, 0, sizeof(*o));
    flacb200_options_default(&o->frame);
    o->padding = 4096;       // Options::default(): Padding { size: 4096 } (:1392)
    o->seektable_kind = 1;   // SeekTableInterval::default(): every 10 seconds (:1329)
    o->seektable_n = 10;
}

void flacb200_writer_options_fast(flacb200_writer_options* o)
{
    flacb200_writer_options_default(o);
    flacb200_options_fast(&o->frame);
}

void flacb200_writer_options_best(flacb200_writer_options* o)
{
    flacb200_writer_options_default(o);
    flacb200_options_best(&o->frame);
}

int flacb200_total_from_bytes(uint64_t total_bytes, uint32_t bits_per_sample, uint32_t channels, uint64_t* pcm_frames)
{
    if (bits_per_sample < 1 || bits_per_sample > 32) return E_INVALID_BPS;
    const uint64_t bytes_per_sample = (bits_per_sample + 7) / 8;
    if (channels == 0 || total_bytes % channels || (total_bytes / channels) % bytes_per_sample) return E_SAMPLES_NOT_DIVISIBLE;   // :170-175
    const uint64_t n = total_bytes / channels / bytes_per_sample;
    if (n == 0) return E_INVALID_TOTAL_BYTES;
    if (pcm_frames) *pcm_frames = n;
    return 0;
}

int flacb200_total_from_samples(uint64_t total_samples, uint32_t channels, uint64_t* pcm_frames)
{
    if (channels == 0 || total_samples % channels) return E_SAMPLES_NOT_DIVISIBLE;   // :516-519
    const uint64_t n = total_samples / channels;
    if (n == 0) return E_INVALID_TOTAL_SAMPLES;
    if (pcm_frames) *pcm_frames = n;
    return 0;
}

int flacb200_writer_open(flacb200_engine* engine, const flacb200_writer_options* opt, uint32_t sample_rate, uint32_t bits_per_sample,
                         uint32_t channels, uint64_t total_pcm_frames, flacb200_writer** out)
{
    if (!opt || !out) return FLACB200_E_BAD_ARGUMENT;
    if (bits_per_sample < 1 || bits_per_sample > 32) return E_INVALID_BPS;   // SignedBitCount<32>::try_from (:151)
    if (opt->frame.block_size < 16) return FLACB200_E_BAD_ARGUMENT;          // OptionsError::InvalidBlockSize (:1418)
    if (sample_rate >= (1u << 20)) return E_INVALID_SAMPLE_RATE;             // :1899-1902
    if (channels < 1 || channels > 8) return E_EXCESSIVE_CHANNELS;           // :1904-1908
    if (total_pcm_frames >= MAX_SAMPLES) return E_EXCESSIVE_TOTAL_SAMPLES;   // :1909-1915
    flacb200_writer* w = new flacb200_writer();
    w->engine = engine;
    w->opt = *opt;
    w->rate = sample_rate;
    w->bps = bits_per_sample;
    w->channels = channels;
    w->total = total_pcm_frames;
    w->bytes_per_sample = (bits_per_sample + 7) / 8;
    w->block_size = opt->frame.block_size;
    w->launch_frames = opt->launch_frames ? opt->launch_frames : 4096;
    w->pcm_frame_bytes = (size_t)w->bytes_per_sample * channels;
    w->block_bytes = w->pcm_frame_bytes * w->block_size;
    // placeholder SEEKTABLE sized from the expected total (:1920-1939)
    if (total_pcm_frames && opt->seektable_kind) {
        const uint64_t bs = w->block_size, nblocks = (total_pcm_frames + bs - 1) / bs;
        std::vector<size_t> keep;
        seek_filter(*opt, sample_rate, (size_t)nblocks, [&](size_t i) {   // EncoderSeekPoint::placeholders (:2131)
            const uint64_t off = (uint64_t)i * bs;
            return SeekPt{off, 0, (uint32_t)std::min<uint64_t>(bs, total_pcm_frames - off)};
        }, keep);
        w->n_placeholders = std::min(keep.size(), MAX_SEEK_POINTS);
        w->have_seektable = true;
    }
    build_header(*w, false);
    *out = w;
    return 0;
}

void flacb200_writer_close(flacb200_writer* w) { delete w; }

int flacb200_writer_header(flacb200_writer* w, const uint8_t** bytes, size_t* len)
{
    if (!w || !bytes || !len) return FLACB200_E_BAD_ARGUMENT;
    *bytes = w->header.data();
    *len = w->header.size();
    return 0;
}

int flacb200_writer_write_bytes(flacb200_writer* w, const uint8_t* pcm, size_t n, int big_endian)
{
    if (!w || (!pcm && n)) return FLACB200_E_BAD_ARGUMENT;
    if (w->finalized || w->failed) return FLACB200_E_BAD_ARGUMENT;
    const uint32_t B = w->bytes_per_sample;
    int rc = writer_reserve_pending(*w, n + 4);
    if (rc) return rc;
    uint8_t* dst = w->pending.p + w->pending.len;
    size_t added = 0;
    auto emit = [&](const uint8_t* s) {   // Endianness::bytes_to_le (src/byteorder.rs:181)
        if (big_endian) for (uint32_t k = 0; k < B; k++) dst[added + k] = s[B - 1 - k];
        else memcpy(dst + added, s, B);
        added += B;
    };
    if (w->npartial) {
        while (n && w->npartial < B) {
            w->partial[(w->npartial++) & 7] = *pcm++;
            n--;
        }
        if (w->npartial == B) {
            emit(w->partial);
            w->npartial = 0;
        }
    }
    const size_t whole = n / B;
    if (!big_endian || B == 1) {
        memcpy(dst + added, pcm, whole * B);
        added += whole * B;
    } else {
        for (size_t i = 0; i < whole; i++) emit(pcm + i * B);
    }
    for (size_t i = whole * B; i < n; i++) w->partial[(w->npartial++) & 7] = pcm[i];
    w->pending.len += added;
    return writer_after_append(*w, false);
}

int flacb200_writer_write_samples(flacb200_writer* w, const int32_t* s, size_t n)
{
    if (!w || (!s && n)) return FLACB200_E_BAD_ARGUMENT;
    if (w->finalized || w->failed) return FLACB200_E_BAD_ARGUMENT;
    const uint32_t B = w->bytes_per_sample;
    int rc = writer_reserve_pending(*w, n * B);
    if (rc) return rc;
    uint8_t* dst = w->pending.p + w->pending.len;
    for (size_t i = 0; i < n; i++) {   // update_md5's byte form (:1292-1318): the low B bytes, little endian
        const uint32_t v = (uint32_t)s[i];
        for (uint32_t k = 0; k < B; k++) dst[i * B + k] = (uint8_t)(v >> (8 * k));
    }
    w->pending.len += n * B;
    return writer_after_append(*w, false);
}

int flacb200_writer_write_channels(flacb200_writer* w, const int32_t* const* ch, uint32_t nch, size_t n)
{
    if (!w || (!ch && nch)) return FLACB200_E_BAD_ARGUMENT;
    if (w->finalized || w->failed) return FLACB200_E_BAD_ARGUMENT;
    if (nch != w->channels) return E_CHANNEL_COUNT_MISMATCH;   // FlacChannelWriter::write (:845-849)
    const uint32_t B = w->bytes_per_sample;
    int rc = writer_reserve_pending(*w, n * nch * B);
    if (rc) return rc;
    uint8_t* dst = w->pending.p + w->pending.len;
    for (size_t i = 0; i < n; i++)
        for (uint32_t c = 0; c < nch; c++) {
            const uint32_t v = (uint32_t)ch[c][i];
            for (uint32_t k = 0; k < B; k++) *dst++ = (uint8_t)(v >> (8 * k));
        }
    w->pending.len += n * nch * B;
    return writer_after_append(*w, false);
}

int flacb200_writer_drain(flacb200_writer* w, const uint8_t** frames, size_t* len)
{
    if (!w || !frames || !len) return FLACB200_E_BAD_ARGUMENT;
    w->drained.swap(w->ready);
    w->ready.clear();
    *frames = w->drained.data();
    *len = w->drained.size();
    return 0;
}

int flacb200_writer_flush(flacb200_writer* w)
{
    if (!w) return FLACB200_E_BAD_ARGUMENT;
    if (w->finalized || w->failed) return 0;
    return writer_after_append(*w, true);
}

int flacb200_writer_finalize(flacb200_writer* w)
{
    if (!w) return FLACB200_E_BAD_ARGUMENT;
    if (w->finalized) return 0;   // Finalized::Finalized: second call is a no-op (:236)
    if (w->failed) return FLACB200_E_BAD_ARGUMENT;
    w->finalized = true;
    // whole blocks first, then the final short block truncated to whole PCM frames (:240-258, :591-609)
    const uint64_t pcm = w->pending.len / w->pcm_frame_bytes;
    if (w->total && w->pcm_frames_encoded + pcm > w->total) return E_EXCESSIVE_TOTAL_SAMPLES;
    int rc = writer_encode(*w, pcm);
    if (rc) return rc;
    w->pending.len = 0;
    // Encoder::finalize_inner (:2024-2110)
    if (w->total) {
        if (w->total != w->pcm_frames_encoded) return E_SAMPLE_COUNT_MISMATCH;
    } else {
        if (w->pcm_frames_encoded >= MAX_SAMPLES) return E_EXCESSIVE_TOTAL_SAMPLES;
        if (w->pcm_frames_encoded == 0) return E_NO_SAMPLES;
    }
    w->md5.final(w->md5_final);
    w->md5_known = true;
    build_header(*w, true);
    return 0;
}

int flacb200_writer_get_stats(flacb200_writer* w, flacb200_writer_stats* s)
{
    if (!w || !s) return FLACB200_E_BAD_ARGUMENT;
    memset(s, 0, sizeof(*s));
    s->pcm_frames_written = w->pcm_frames_encoded;
    s->frames_written = w->frames;
    s->frame_bytes_written = w->frame_bytes;
    s->min_frame_size = w->min_frame;
    s->max_frame_size = w->max_frame;
    s->launches = w->launches;
    if (w->md5_known) memcpy(s->md5, w->md5_final, 16);
    return 0;
}

void flacb200_md5(const uint8_t* data, size_t len, uint8_t out[16])
{
    Md5 m;
    m.update(data, len);
    m.final(out);
}

}   // extern "C"

// =================================================================================================
// reader: Decoder (src/decode.rs:1311-1491) behind the FlacByteReader / FlacSampleReader / FlacChannelReader facades
// =================================================================================================
// The reference's Decoder pulls one frame at a time out of `R: Read`.  Here the unit of GPU work is a WINDOW: a run of
// bytes starting at a frame boundary (at most window_bytes of them, decoding to at most window_pcm samples) goes through
// one flacb200_decode call; flacb200_decode_last_frames says where its frames lie, the reader hands them out one frame at
// a time (fill_buf) and starts the next window where the last good frame ended.  A frame cut by the window's end -- or
// one that does not fit the PCM budget -- fails in this window and is simply decoded again as the first frame of the next;
// an error that is still there at the head of a window is genuine and is returned at that point: after every good frame
// in front of it has been delivered, as with the serial reader.  Memory is bounded by the window, not by the stream.
// The bytes come from a caller-owned file image, or are fed chunk by chunk (flacb200_reader_feed: an `R: Read`).
struct flacb200_reader {
    flacb200_engine* engine = nullptr;
    const uint8_t* image = nullptr;   // image mode: the whole file, caller-owned
    size_t image_len = 0;
    bool fed_mode = false, fed_eof = false, meta_done = false;
    std::vector<uint8_t> fed;         // feed mode: fed[0] is the byte at absolute offset fed_base
    uint64_t fed_base = 0;
    flacb200_streaminfo si{};
    std::vector<flacb200_seekpoint> seektable;
    size_t window_bytes = (size_t)32 << 20;
    uint64_t window_pcm = (uint64_t)8 << 20;   // inter-channel samples per window
    uint64_t next_byte = 0;           // absolute offset of the next frame to decode
    uint64_t decoded_samples = 0;     // inter-channel samples in front of the current window's end
    HostBuf win;                      // the window's PCM: int32 interleaved, or packed little-endian bytes --
    int win_kind = FLACB200_PCM_I32_INTERLEAVED;   // whichever the access that triggered the decode asked for (a byte reader
                                      // gets its bytes with one memcpy, a sample reader its i32 without a conversion pass)
    std::vector<int32_t> conv;        // current frame as i32 when the window holds bytes
    std::vector<flacb200_frame_entry> frames;
    size_t cur = 0;                   // frame being handed out
    uint32_t cur_off = 0;             // inter-channel samples of it already consumed
    // a read that ended inside a PCM frame (io::Read of an odd byte count, FlacSampleReader::read of one sample of a stereo
    // stream: the reference's buffers are byte- and sample-granular): `part` units of the PCM frame at cur_off are gone
    uint32_t part = 0;                // bytes (byte kinds) or samples (i32) consumed of that PCM frame
    int part_kind = 0;                // the layout those units were counted in
    bool at_end = false;
    int sticky_error = 0;
    bool seekable_source = false;     // feed mode over an `R: Read + Seek`: a seek asks the caller to reposition its source
    bool seeking = false;             // flacb200_reader_seek in progress (NEED_SEEK / NEED_DATA in between)
    uint64_t seek_sample = 0, seek_pos = 0, want_offset = 0;
    bool verifying = false;           // flacb200_reader_verify in progress (a fed reader returns FLACB200_NEED_DATA in between)
    Md5 verify_md5;
    std::vector<int32_t> planar;      // FlacChannelReader view of the current frame
    std::vector<const int32_t*> planar_ptrs;
};

namespace {

// BlockIterator (src/metadata/mod.rs:482-646) reduced to what a decoder needs: STREAMINFO first, at most one
// SEEKTABLE, block types validated, every other block skipped.
int parse_metadata(const uint8_t* f, size_t len, flacb200_streaminfo* si, std::vector<flacb200_seekpoint>* table)
{
    if (len < 4) return E_IO;
    if (memcmp(f, "fLaC", 4) != 0) return E_MISSING_FLAC_TAG;
    size_t p = 4;
    bool first = true, seektable_seen = false;
    memset(si, 0, sizeof(*si));
    if (table) table->clear();
    for (;;) {
        if (p + 4 > len) return first ? E_MISSING_STREAMINFO : E_IO;
        const bool last = (f[p] >> 7) != 0;
        const uint32_t type = f[p] & 0x7F;
        const size_t blen = (size_t)get_be(f + p + 1, 3);
        p += 4;
        if (first) {
            if (type != 0 || blen != 34) return E_MISSING_STREAMINFO;
            if (p + blen > len) return E_IO;
            const uint8_t* b = f + p;
            si->min_block_size = (uint16_t)get_be(b, 2);
            si->max_block_size = (uint16_t)get_be(b + 2, 2);
            si->min_frame_size = (uint32_t)get_be(b + 4, 3);
            si->max_frame_size = (uint32_t)get_be(b + 7, 3);
            const uint64_t v = get_be(b + 10, 8);
            si->sample_rate = (uint32_t)(v >> 44);
            si->channels = (uint32_t)((v >> 41) & 7) + 1;
            si->bits_per_sample = (uint32_t)((v >> 36) & 31) + 1;
            si->total_samples = v & 0xFFFFFFFFFull;
            memcpy(si->md5, b + 18, 16);
            first = false;
        } else {
            if (type >= 7 && type <= 126) return E_RESERVED_METADATA_BLOCK;   // src/metadata/mod.rs:313
            if (type == 127) return E_INVALID_METADATA_BLOCK;
            if (p + blen > len) return E_IO;
            if (type == 0) return E_MULTIPLE_STREAMINFO;
            if (type == 3) {
                if (seektable_seen) return E_MULTIPLE_SEEKTABLE;
                seektable_seen = true;
                if (blen % 18) return E_INVALID_SEEKTABLE_SIZE;   // :2005
                uint64_t last_off = 0;
                bool have_last = false;
                for (size_t i = 0; i < blen / 18; i++) {
                    const uint8_t* s = f + p + 18 * i;
                    flacb200_seekpoint sp{get_be(s, 8), get_be(s + 8, 8), (uint32_t)get_be(s + 16, 2), 0};
                    sp.placeholder = sp.sample_offset == ~0ull;
                    if (!sp.placeholder) {   // defined points must increase (Contiguous, :2001-2003)
                        if (have_last && sp.sample_offset <= last_off) return E_INVALID_SEEKTABLE_POINT;
                        last_off = sp.sample_offset;
                        have_last = true;
                    }
                    if (table) table->push_back(sp);
                }
                si->n_seekpoints = (uint32_t)(blen / 18);
            }
        }
        p += blen;
        if (last) break;
    }
    si->frames_start = p;
    return 0;
}

// bytes available from absolute offset `off`
inline const uint8_t* reader_bytes(const flacb200_reader& r, uint64_t off, size_t* avail)
{
    if (!r.fed_mode) {
        *avail = off < r.image_len ? r.image_len - (size_t)off : 0;
        return r.image + off;
    }
    const uint64_t end = r.fed_base + r.fed.size();
    *avail = off < end ? (size_t)(end - off) : 0;
    return r.fed.data() + (off - r.fed_base);
}

// feed mode: the metadata blocks must have arrived before anything else can happen.  Returns FLACB200_NEED_DATA while they
// have not.
int reader_ensure_meta(flacb200_reader& r)
{
    if (r.meta_done) return 0;
    const int rc = parse_metadata(r.fed.data(), r.fed.size(), &r.si, &r.seektable);
    if ((rc == E_IO || (rc == E_MISSING_STREAMINFO && r.fed.size() < 42)) && !r.fed_eof) return FLACB200_NEED_DATA;
    if (rc) return rc;
    r.meta_done = true;
    r.next_byte = r.si.frames_start;
    return 0;
}

// Decodes the next window.  0: frames[] refilled (or at_end set); FLACB200_NEED_DATA: feed mode, nothing decodable is
// buffered yet; else the error the serial reader would return at this point.
int reader_next_window(flacb200_reader& r, int want_kind)
{
    r.frames.clear();
    r.cur = 0;
    r.cur_off = 0;
    r.part = 0;
    if (r.sticky_error) return r.sticky_error;
    if (r.at_end) return 0;
    if (!r.engine) return FLACB200_E_NO_DEVICE;
    const flacb200_streaminfo& si = r.si;
    const uint64_t total = si.total_samples;
    if (total && r.decoded_samples == total) {   // Some(0) => Ok(None)  (src/decode.rs:1402)
        r.at_end = true;
        return 0;
    }
    // (a last frame that overshot the announced total: `total - current_sample` has wrapped in the reference (:1400), its reader
    // goes on until the bytes end and fails there with Io; here the rest is decoded as a stream of unknown length, then Io)
    const bool overshot = total && r.decoded_samples > total;
    flacb200_stream_params prm{};
    prm.sample_rate = si.sample_rate;
    prm.bits_per_sample = si.bits_per_sample;
    prm.channels = si.channels;
    prm.max_block_size = si.max_block_size;
    size_t want = r.window_bytes;
    uint64_t cap_pcm = std::max<uint64_t>(r.window_pcm, (uint64_t)si.max_block_size);
    for (int attempt = 0; attempt < 24; attempt++) {
        size_t avail = 0;
        const uint8_t* src = reader_bytes(r, r.next_byte, &avail);
        const bool source_done = !r.fed_mode || r.fed_eof;
        if (avail == 0) {
            if (!source_done) return FLACB200_NEED_DATA;
            if (total) return r.sticky_error = E_IO;   // FrameHeader::read hits EOF with samples outstanding (or after an overshoot)
            r.at_end = true;                           // unsized stream: EOF at a frame boundary ends it (:1416)
            return 0;
        }
        const size_t take = std::min(avail, want);
        const bool last_window = take == avail && source_done;
        const uint64_t remaining = (total && !overshot) ? total - r.decoded_samples : 0;
        // the PCM budget of the window; with a known total the final window must be able to hold everything that remains
        uint64_t cap = cap_pcm;
        if (total && !overshot) cap = std::min<uint64_t>(cap, remaining + si.max_block_size);
        r.win_kind = want_kind == FLACB200_PCM_I32_INTERLEAVED ? FLACB200_PCM_I32_INTERLEAVED : FLACB200_PCM_BYTES_LE;
        const size_t fb = (size_t)si.channels * (r.win_kind == FLACB200_PCM_I32_INTERLEAVED ? 4 : (si.bits_per_sample + 7) / 8);
        if (!r.win.reserve((size_t)cap * fb + 64, true)) return FLACB200_E_OUT_OF_MEMORY;
        flacb200_decode_segment seg{0, take, 0, remaining};
        uint64_t nf = 0, ns = 0, bad = 0;
        const int rc = flacb200_decode(r.engine, &prm, src, take, FLACB200_HOST, &seg, 1, r.win.p, (size_t)cap * fb, r.win_kind, FLACB200_HOST, 0,
                                       &nf, &ns, &bad);
        if (rc < 0 && rc != FLACB200_E_OUTPUT_TOO_SMALL) return rc;   // CUDA / argument errors are not stream errors
        uint64_t ntab = 0;
        r.frames.resize((size_t)nf);
        if (nf) {
            const int rt = flacb200_decode_last_frames(r.engine, r.frames.data(), r.frames.size(), &ntab);
            if (rt) return rt;
            if (ntab != nf) return FLACB200_E_BAD_ARGUMENT;   // (cannot happen: both count the frames of the walk)
        }
        if (nf) {   // deliver these; whatever stopped the walk shows up again at the head of the next window
            const flacb200_frame_entry& l = r.frames.back();
            r.next_byte += l.byte_offset + l.byte_length;
            r.decoded_samples += ns;
            return 0;
        }
        // nothing decoded
        if (rc == 0) {   // unsized stream whose last bytes are no frame (fewer than 16: EOF inside a header, :1416), or an empty window
            if (last_window) {
                if (overshot) return r.sticky_error = E_IO;
                r.at_end = true;
                return 0;
            }
            if (r.fed_mode && take == avail) return FLACB200_NEED_DATA;
            want *= 2;
            continue;
        }
        const bool cut = rc == E_IO || rc == FLACB200_E_OUTPUT_TOO_SMALL;   // the first frame did not fit the window
        if (cut && !last_window && rc == E_IO) {
            if (r.fed_mode && take == avail) return FLACB200_NEED_DATA;
            want *= 2;
            continue;
        }
        if (rc == FLACB200_E_OUTPUT_TOO_SMALL) {
            cap_pcm *= 2;
            continue;
        }
        return r.sticky_error = rc;   // genuine: the frame at next_byte fails
    }
    return FLACB200_E_OUTPUT_TOO_SMALL;
}

// makes frames[cur] a frame with unconsumed samples; 0 with at_end set when the stream is over
int reader_current_frame(flacb200_reader& r, int want_kind)
{
    for (;;) {
        if (r.cur < r.frames.size()) {
            if (r.cur_off < r.frames[r.cur].block_size) return 0;
            r.cur++;
            r.cur_off = 0;
            continue;
        }
        if (r.at_end) return 0;
        const int rc = reader_next_window(r, want_kind);
        if (rc) return rc;
        if (r.frames.empty() && r.at_end) return 0;
    }
}

// the unconsumed part of the current frame as int32 (converted into r.conv when the window holds packed bytes)
inline const int32_t* reader_frame_samples(flacb200_reader& r)
{
    const flacb200_frame_entry& f = r.frames[r.cur];
    const size_t ch = r.si.channels, first = (size_t)(f.pcm_offset + r.cur_off) * ch;
    if (r.win_kind == FLACB200_PCM_I32_INTERLEAVED) return reinterpret_cast<const int32_t*>(r.win.p) + first;
    const size_t B = (r.si.bits_per_sample + 7) / 8, n = (size_t)(f.block_size - r.cur_off) * ch;
    const uint32_t sh = 32 - 8 * (uint32_t)B;
    r.conv.resize(n);
    const uint8_t* src = r.win.p + first * B;
    for (size_t i = 0; i < n; i++) {
        uint32_t v = 0;
        for (size_t k = 0; k < B; k++) v |= (uint32_t)src[i * B + k] << (8 * k);
        r.conv[i] = (int32_t)(v << sh) >> sh;
    }
    return r.conv.data();
}

// the same span as packed little-endian bytes, when the window holds them
inline const uint8_t* reader_frame_bytes(const flacb200_reader& r)
{
    const flacb200_frame_entry& f = r.frames[r.cur];
    return r.win.p + (size_t)(f.pcm_offset + r.cur_off) * r.si.channels * ((r.si.bits_per_sample + 7) / 8);
}

// Decoder::seek (src/decode.rs:1452-1491): the last defined seek point at or before `sample`, else the first frame
uint64_t reader_seek_point(flacb200_reader& r, uint64_t sample)
{
    uint64_t at_sample = 0, at_byte = 0;
    for (size_t i = r.seektable.size(); i-- > 0;) {
        const flacb200_seekpoint& p = r.seektable[i];
        if (!p.placeholder && p.sample_offset <= sample) {
            at_sample = p.sample_offset;
            at_byte = p.byte_offset;
            break;
        }
    }
    r.next_byte = r.si.frames_start + at_byte;
    r.decoded_samples = at_sample;
    r.frames.clear();
    r.cur = 0;
    r.cur_off = 0;
    r.part = 0;
    r.at_end = false;
    r.sticky_error = 0;
    return at_sample;
}

}   // namespace

extern "C" {

int flacb200_read_streaminfo(const uint8_t* flac, size_t len, flacb200_streaminfo* si)
{
    if (!flac || !si) return FLACB200_E_BAD_ARGUMENT;
    return parse_metadata(flac, len, si, nullptr);
}

int flacb200_reader_open(flacb200_engine* engine, const uint8_t* flac, size_t len, flacb200_reader** out)
{
    if (!flac || !out) return FLACB200_E_BAD_ARGUMENT;
    flacb200_reader* r = new flacb200_reader();
    r->engine = engine;
    r->image = flac;
    r->image_len = len;
    const int rc = parse_metadata(flac, len, &r->si, &r->seektable);
    if (rc) {
        delete r;
        return rc;
    }
    r->meta_done = true;
    r->next_byte = r->si.frames_start;
    *out = r;
    return 0;
}

int flacb200_reader_open_stream(flacb200_engine* engine, flacb200_reader** out)
{
    if (!out) return FLACB200_E_BAD_ARGUMENT;
    flacb200_reader* r = new flacb200_reader();
    r->engine = engine;
    r->fed_mode = true;
    *out = r;
    return 0;
}

int flacb200_reader_feed(flacb200_reader* r, const uint8_t* bytes, size_t len, int eof)
{
    if (!r || !r->fed_mode || (!bytes && len)) return FLACB200_E_BAD_ARGUMENT;
    if (r->fed_eof && len) return FLACB200_E_BAD_ARGUMENT;
    // bytes in front of the next frame have been decoded: drop them (a window's PCM lives in its own buffer)
    if (r->meta_done && r->next_byte > r->fed_base) {
        const size_t drop = (size_t)std::min<uint64_t>(r->next_byte - r->fed_base, r->fed.size());
        r->fed.erase(r->fed.begin(), r->fed.begin() + (long)drop);
        r->fed_base += drop;
    }
    r->fed.insert(r->fed.end(), bytes, bytes + len);
    if (eof) r->fed_eof = true;
    return 0;
}

void flacb200_reader_close(flacb200_reader* r) { delete r; }

int flacb200_reader_set_window(flacb200_reader* r, size_t window_bytes, uint64_t window_pcm_frames)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    if (window_bytes) r->window_bytes = std::max<size_t>(window_bytes, 64);
    if (window_pcm_frames) r->window_pcm = window_pcm_frames;
    return 0;
}

int flacb200_reader_info(flacb200_reader* r, flacb200_streaminfo* si)
{
    if (!r || !si) return FLACB200_E_BAD_ARGUMENT;
    if (r->fed_mode) {
        const int rc = reader_ensure_meta(*r);
        if (rc) return rc;
    }
    *si = r->si;
    return 0;
}

int flacb200_reader_seektable(flacb200_reader* r, flacb200_seekpoint* points, size_t capacity, size_t* n_points)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    if (r->fed_mode) {
        const int rc = reader_ensure_meta(*r);
        if (rc) return rc;
    }
    if (n_points) *n_points = r->seektable.size();
    if (points)
        for (size_t i = 0; i < r->seektable.size() && i < capacity; i++) points[i] = r->seektable[i];
    return 0;
}

// FlacSampleReader::fill_buf (src/decode.rs:466-486): the unconsumed samples of the current frame, interleaved
int flacb200_reader_fill_buf(flacb200_reader* r, const int32_t** samples, size_t* n_samples)
{
    if (!r || !samples || !n_samples) return FLACB200_E_BAD_ARGUMENT;
    *samples = nullptr;
    *n_samples = 0;
    if (r->fed_mode) {
        const int rm = reader_ensure_meta(*r);
        if (rm) return rm;
    }
    const int rc = reader_current_frame(*r, FLACB200_PCM_I32_INTERLEAVED);
    if (rc) return rc;
    if (r->cur >= r->frames.size()) return 0;   // end of stream: an empty buffer
    if (r->part && r->part_kind != FLACB200_PCM_I32_INTERLEAVED) return FLACB200_E_BAD_ARGUMENT;   // (a byte read stopped inside a PCM frame)
    *samples = reader_frame_samples(*r) + r->part;
    *n_samples = (size_t)(r->frames[r->cur].block_size - r->cur_off) * r->si.channels - r->part;
    return 0;
}

// FlacSampleReader::consume (:487): n_samples counts all channels; any amount up to what fill_buf returned (VecDeque::drain)
int flacb200_reader_consume(flacb200_reader* r, size_t n_samples)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    if (r->cur >= r->frames.size()) return n_samples ? FLACB200_E_BAD_ARGUMENT : 0;
    if (r->part && r->part_kind != FLACB200_PCM_I32_INTERLEAVED) return FLACB200_E_BAD_ARGUMENT;
    const size_t ch = r->si.channels;
    const size_t left = (size_t)(r->frames[r->cur].block_size - r->cur_off) * ch - r->part;
    if (n_samples > left) return FLACB200_E_BAD_ARGUMENT;
    const size_t adv = r->part + n_samples;
    r->cur_off += (uint32_t)(adv / ch);
    r->part = (uint32_t)(adv % ch);
    r->part_kind = FLACB200_PCM_I32_INTERLEAVED;
    return 0;
}

// FlacChannelReader::fill_buf (:917-944): one slice per channel over the unconsumed part of the current frame
int flacb200_reader_fill_channels(flacb200_reader* r, const int32_t* const** channels, size_t* n_per_channel)
{
    if (!r || !channels || !n_per_channel) return FLACB200_E_BAD_ARGUMENT;
    *n_per_channel = 0;
    if (r->fed_mode) {
        const int rm = reader_ensure_meta(*r);
        if (rm) return rm;
    }
    const int rc = reader_current_frame(*r, FLACB200_PCM_I32_INTERLEAVED);
    if (rc) return rc;
    if (r->part) return FLACB200_E_BAD_ARGUMENT;   // (an interleaved read stopped inside a PCM frame: not a channel reader's state)
    const size_t ch = r->si.channels;
    r->planar_ptrs.assign(ch, nullptr);
    *channels = r->planar_ptrs.data();
    if (r->cur >= r->frames.size()) return 0;   // vec![&[]; channels]
    const size_t n = r->frames[r->cur].block_size - r->cur_off;
    r->planar.resize(n * ch);
    const int32_t* s = reader_frame_samples(*r);
    for (size_t c = 0; c < ch; c++) {
        int32_t* d = r->planar.data() + c * n;
        for (size_t i = 0; i < n; i++) d[i] = s[i * ch + c];
        r->planar_ptrs[c] = d;
    }
    *n_per_channel = n;
    return 0;
}

int flacb200_reader_consume_channels(flacb200_reader* r, size_t n_per_channel)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    return flacb200_reader_consume(r, n_per_channel * r->si.channels);
}

int flacb200_reader_read(flacb200_reader* r, void* out, size_t capacity, int pcm_kind, size_t* n_out)
{
    if (!r || (!out && capacity) || !n_out) return FLACB200_E_BAD_ARGUMENT;
    if (pcm_kind != FLACB200_PCM_BYTES_LE && pcm_kind != FLACB200_PCM_BYTES_BE && pcm_kind != FLACB200_PCM_I32_INTERLEAVED)
        return FLACB200_E_BAD_ARGUMENT;
    *n_out = 0;
    if (r->fed_mode) {
        const int rm = reader_ensure_meta(*r);
        if (rm) return rm;
    }
    const size_t B = (r->si.bits_per_sample + 7) / 8, ch = r->si.channels;
    const bool as_i32 = pcm_kind == FLACB200_PCM_I32_INTERLEAVED, be = pcm_kind == FLACB200_PCM_BYTES_BE;
    const size_t unit = as_i32 ? 4 : 1;         // bytes of `out` per unit (a sample, or a byte of the packed layouts)
    const size_t U = as_i32 ? ch : ch * B;      // units of one PCM frame
    size_t units = 0;                           // units delivered
    // one PCM frame (the one at cur_off) in the caller's layout
    auto pcm_frame = [&](uint8_t* tmp) {
        if (as_i32) memcpy(tmp, reader_frame_samples(*r), ch * 4);
        else if (r->win_kind == FLACB200_PCM_BYTES_LE) {
            const uint8_t* src = reader_frame_bytes(*r);
            for (size_t i = 0; i < ch; i++)
                for (size_t k = 0; k < B; k++) tmp[i * B + k] = src[i * B + (be ? B - 1 - k : k)];
        } else {
            const int32_t* sm = reader_frame_samples(*r);
            for (size_t i = 0; i < ch; i++)
                for (size_t k = 0; k < B; k++) tmp[i * B + k] = (uint8_t)((uint32_t)sm[i] >> (8 * (be ? B - 1 - k : k)));
        }
    };
    uint8_t tmp[8 * 4];
    if (r->part && capacity) {   // the rest of the PCM frame the previous read stopped in
        if (r->part_kind != pcm_kind) return FLACB200_E_BAD_ARGUMENT;
        const int rc = reader_current_frame(*r, pcm_kind);
        if (rc) return rc;
        if (r->cur < r->frames.size()) {
            pcm_frame(tmp);
            const size_t n = std::min<size_t>(capacity, U - r->part);
            memcpy(out, tmp + r->part * unit, n * unit);
            units = n;
            r->part += (uint32_t)n;
            if (r->part == U) {
                r->part = 0;
                r->cur_off++;
            }
        }
    }
    size_t room = r->part ? 0 : ((capacity - units) / U) * ch;   // single-channel samples of the whole PCM frames that still fit
    const size_t head_bytes = units * unit;
    size_t done = 0;
    bool stopped = false;   // end of stream, or an error that surfaces at the next call
    while (room) {
        const int rc = reader_current_frame(*r, pcm_kind);
        if (rc) {
            if (done || units) { stopped = true; break; }   // deliver what precedes the error; it surfaces at the next call
            return rc;
        }
        if (r->cur >= r->frames.size()) { stopped = true; break; }   // end of stream
        const size_t n = std::min<size_t>(room, (size_t)(r->frames[r->cur].block_size - r->cur_off) * ch);
        if (!as_i32 && r->win_kind == FLACB200_PCM_BYTES_LE) {   // Frame::to_buf (src/audio.rs:110-134) was done on the device
            const uint8_t* src = reader_frame_bytes(*r);
            uint8_t* o = (uint8_t*)out + head_bytes + done * B;
            if (!be || B == 1) memcpy(o, src, n * B);
            else
                for (size_t i = 0; i < n; i++)
                    for (size_t k = 0; k < B; k++) o[i * B + k] = src[i * B + B - 1 - k];
        } else {
            const int32_t* s = reader_frame_samples(*r);
            if (as_i32) memcpy((uint8_t*)out + head_bytes + done * 4, s, n * 4);
            else {
                uint8_t* o = (uint8_t*)out + head_bytes + done * B;
                for (size_t i = 0; i < n; i++) {
                    const uint32_t v = (uint32_t)s[i];
                    for (size_t k = 0; k < B; k++) o[i * B + k] = (uint8_t)(v >> (8 * (be ? B - 1 - k : k)));
                }
            }
        }
        r->cur_off += (uint32_t)(n / ch);
        done += n;
        room -= n;
    }
    units += as_i32 ? done : done * B;
    // what is left of the caller's buffer holds less than a PCM frame: hand out the head of the next one
    const size_t rest = capacity - units;
    if (!stopped && !r->part && rest && rest < U) {
        const int rc = reader_current_frame(*r, pcm_kind);
        if (rc && !units) return rc;
        if (!rc && r->cur < r->frames.size()) {
            pcm_frame(tmp);
            memcpy((uint8_t*)out + units * unit, tmp, rest * unit);
            units += rest;
            r->part = (uint32_t)rest;
            r->part_kind = pcm_kind;
        }
    }
    *n_out = units;
    return 0;
}

// FlacSampleReader::seek / FlacChannelReader::seek (:823-860, :1021-1057): Decoder::seek to the last seek point at or before
// the sample (the stream start without a table), then frames are decoded and skipped up to the sample itself; running out of
// stream first is InvalidSeek.  (FlacByteReader's io::Seek (:715-820) is the same walk in bytes.)
int flacb200_reader_seek(flacb200_reader* r, uint64_t pcm_frame)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    if (r->fed_mode && !r->seekable_source) return E_IO;   // a plain `R: Read` is not seekable (frames_start: None -> NotSeekable)
    if (r->fed_mode) {
        const int rm = reader_ensure_meta(*r);
        if (rm) return rm;
    }
    if (!(r->seeking && r->seek_sample == pcm_frame)) {
        r->seek_pos = reader_seek_point(*r, pcm_frame);
        r->seek_sample = pcm_frame;
        r->seeking = true;
        if (r->fed_mode) {   // the buffered bytes belong to another place in the file: the caller repositions its source
            r->fed.clear();
            r->fed_base = r->next_byte;
            r->fed_eof = false;
            r->want_offset = r->next_byte;
            return FLACB200_NEED_SEEK;
        }
    }
    while (pcm_frame > r->seek_pos) {
        const int rc = reader_current_frame(*r, r->win_kind);
        if (rc) {
            if (rc != FLACB200_NEED_DATA) r->seeking = false;
            return rc;
        }
        if (r->cur >= r->frames.size()) {
            r->seeking = false;
            return E_INVALID_SEEK;
        }
        const uint64_t take = std::min<uint64_t>(r->frames[r->cur].block_size - r->cur_off, pcm_frame - r->seek_pos);
        r->cur_off += (uint32_t)take;
        r->seek_pos += take;
    }
    r->seeking = false;
    return 0;
}

int flacb200_reader_set_seekable(flacb200_reader* r, int seekable)
{
    if (!r) return FLACB200_E_BAD_ARGUMENT;
    r->seekable_source = seekable != 0;
    return 0;
}

int flacb200_reader_wanted_offset(flacb200_reader* r, uint64_t* offset)
{
    if (!r || !offset) return FLACB200_E_BAD_ARGUMENT;
    *offset = r->want_offset;
    return 0;
}

// verify_reader (src/decode.rs:1291-1309): decode everything from the current position, MD5 of the little-endian PCM
int flacb200_reader_verify(flacb200_reader* r, int* result, uint8_t md5_out[16])
{
    if (!r || !result) return FLACB200_E_BAD_ARGUMENT;
    if (r->fed_mode) {
        const int rm = reader_ensure_meta(*r);
        if (rm) return rm;
    }
    if (!r->verifying) {
        if (!r->fed_mode) reader_seek_point(*r, 0);   // the whole stream, wherever the reads have got to
        else if (r->decoded_samples || r->next_byte != r->si.frames_start || r->cur < r->frames.size())
            return FLACB200_E_BAD_ARGUMENT;   // a fed stream is verified from its start (seek to 0 first when the source can)
        r->verify_md5 = Md5();
        r->verifying = true;
    }
    const size_t B = (r->si.bits_per_sample + 7) / 8, ch = r->si.channels;
    Md5& m = r->verify_md5;
    std::vector<uint8_t> buf;
    for (;;) {
        const int rc = reader_current_frame(*r, FLACB200_PCM_BYTES_LE);
        if (rc) {
            if (rc != FLACB200_NEED_DATA) r->verifying = false;
            return rc;
        }
        if (r->cur >= r->frames.size()) break;
        const size_t n = (size_t)(r->frames[r->cur].block_size - r->cur_off) * ch;
        if (r->win_kind == FLACB200_PCM_BYTES_LE) {
            // frames of a window are contiguous: hash to the end of the window in one go
            const flacb200_frame_entry& l = r->frames.back();
            const uint8_t* a = reader_frame_bytes(*r);
            const uint8_t* z = r->win.p + (size_t)(l.pcm_offset + l.block_size) * ch * B;
            m.update(a, (size_t)(z - a));
            r->cur = r->frames.size() - 1;
        } else {
            const int32_t* s = reader_frame_samples(*r);
            buf.resize(n * B);
            for (size_t i = 0; i < n; i++)
                for (size_t k = 0; k < B; k++) buf[i * B + k] = (uint8_t)((uint32_t)s[i] >> (8 * k));
            m.update(buf.data(), buf.size());
        }
        r->cur_off = r->frames[r->cur].block_size;
    }
    uint8_t sum[16];
    m.final(sum);
    r->verifying = false;
    if (md5_out) memcpy(md5_out, sum, 16);
    static const uint8_t zero[16] = {0};
    if (memcmp(r->si.md5, zero, 16) == 0) *result = 2;        // Verified::NoMD5
    else *result = memcmp(r->si.md5, sum, 16) == 0 ? 0 : 1;   // MD5Match / MD5Mismatch
    return 0;
}

}   // extern "C"

// =================================================================================================
// FlacStreamReader (src/decode.rs:1149-1240): subset frames without metadata, parameters read from every frame header
// =================================================================================================
// The reference scans for a sync code, parses ONE frame and returns it with the parameters of its header, which may change
// from frame to frame.  Here the host does the scan and parses the first header (a few bytes of bit twiddling), then the GPU
// decodes the RUN of frames that follows with the same channel count and sample width in one flacb200_decode call
// (params.subset = 1); the frames of the run are handed out one by one.  A frame with other parameters, or bytes that are no
// frame, end the run -- the next read starts a new scan there, as the serial reader's next read would.
struct flacb200_stream_reader {
    flacb200_engine* engine = nullptr;
    std::vector<uint8_t> fed;
    bool eof = false;
    size_t pos = 0;                      // scan position in fed
    HostBuf win;                         // decoded run: int32 interleaved
    std::vector<flacb200_frame_entry> frames;
    std::vector<uint32_t> rates;         // sample rate of every frame of the run (the one parameter a run may mix)
    size_t cur = 0;
    size_t run_base = 0;                 // offset in fed of the run's first byte
    uint32_t channels = 0, bps = 0;
    size_t window_bytes = (size_t)16 << 20;
};

namespace {

struct SubsetHeader {
    uint32_t block_size, sample_rate, channels, bps, hdr_len;
};

// FrameHeader::read_subset (src/stream.rs:166-181, :214-240) over p[0..avail).  Returns 0, E_IO (ran out of bytes) or the
// parse error; *consumed = bytes the reference's reader has taken from the stream at that point (its BitReader pulls whole
// bytes as it goes: a failed parse leaves the stream behind the last byte it touched).
int parse_subset_header(const uint8_t* p, size_t avail, SubsetHeader* h, size_t* consumed)
{
    auto need = [&](size_t n) { return avail >= n; };
    *consumed = std::min<size_t>(avail, 2);
    if (!need(2)) return E_IO;
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return 23;   // InvalidSyncCode
    *consumed = std::min<size_t>(avail, 3);
    if (!need(3)) return E_IO;
    const uint32_t bsc = p[2] >> 4, src = p[2] & 15;
    if (bsc == 0) return 24;                                 // InvalidBlockSize
    uint32_t rate = 0, rate_kind = 0;
    static const uint32_t rates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    if (src == 0) return 27;                                 // NonSubsetSampleRate
    if (src <= 11) rate = rates[src];
    else if (src <= 14) rate_kind = src - 11;
    else return 26;                                          // InvalidSampleRate
    *consumed = std::min<size_t>(avail, 4);
    if (!need(4)) return E_IO;
    const uint32_t ca = p[3] >> 4, bpc = (p[3] >> 1) & 7;
    if (ca > 10) return 31;                                  // InvalidChannels
    uint32_t bps;
    switch (bpc) {
    case 0: return 28;                                       // NonSubsetBitsPerSample
    case 1: bps = 8; break;
    case 2: bps = 12; break;
    case 3: return 33;                                       // InvalidBitsPerSample
    case 4: bps = 16; break;
    case 5: bps = 20; break;
    case 6: bps = 24; break;
    default: bps = 32; break;
    }
    size_t n = 4;
    *consumed = std::min<size_t>(avail, 5);
    if (!need(5)) return E_IO;
    const uint32_t f0 = p[4];
    uint32_t ones = 0;
    while (ones < 8 && (f0 & (0x80u >> ones))) ones++;
    if (ones == 0) n += 1;
    else {
        if (ones == 1 || ones > 7) return 36;                // InvalidFrameNumber
        for (uint32_t i = 1; i < ones; i++) {
            *consumed = std::min<size_t>(avail, 4 + i + 1);
            if (!need(4 + i + 1)) return E_IO;
            if ((p[4 + i] & 0xC0) != 0x80) return 36;
        }
        n += ones;
    }
    uint32_t bs;
    if (bsc == 6) {
        *consumed = std::min<size_t>(avail, n + 1);
        if (!need(n + 1)) return E_IO;
        bs = (uint32_t)p[n] + 1;
        n += 1;
    } else if (bsc == 7) {
        *consumed = std::min<size_t>(avail, n + 2);
        if (!need(n + 2)) return E_IO;
        const uint32_t v = ((uint32_t)p[n] << 8) | p[n + 1];
        if (v == 0xFFFF) return 24;
        bs = v + 1;
        n += 2;
    } else {
        bs = bsc == 1 ? 192u : bsc <= 5 ? (576u << (bsc - 2)) : (256u << (bsc - 8));
    }
    if (rate_kind == 1) {
        *consumed = std::min<size_t>(avail, n + 1);
        if (!need(n + 1)) return E_IO;
        rate = (uint32_t)p[n] * 1000;
        n += 1;
    } else if (rate_kind) {
        *consumed = std::min<size_t>(avail, n + 2);
        if (!need(n + 2)) return E_IO;
        rate = ((uint32_t)p[n] << 8) | p[n + 1];
        if (rate_kind == 3) rate *= 10;
        n += 2;
    }
    *consumed = std::min<size_t>(avail, n + 1);
    if (!need(n + 1)) return E_IO;
    n += 1;
    uint8_t crc = 0;
    for (size_t i = 0; i < n; i++) {   // Crc8 (src/crc.rs:104), poly 0x07
        crc ^= p[i];
        for (int b = 0; b < 8; b++) crc = (uint8_t)((crc & 0x80) ? ((crc << 1) ^ 0x07) : (crc << 1));
    }
    if (crc != 0) return 39;                                 // Crc8Mismatch
    h->block_size = bs;
    h->sample_rate = rate;
    h->channels = ca <= 7 ? ca + 1 : 2;
    h->bps = bps;
    h->hdr_len = (uint32_t)n;
    return 0;
}

// Finds the next frame the reference's scan loop (src/decode.rs:1186-1222) would accept, starting at r.pos, and decodes the
// run of frames behind it.  0: frames[] refilled; FLACB200_NEED_DATA; E_IO at the end of the stream; or the frame's error.
int stream_reader_next_run(flacb200_stream_reader& r)
{
    r.frames.clear();
    r.rates.clear();
    r.cur = 0;
    if (!r.engine) return FLACB200_E_NO_DEVICE;
    const uint8_t* f = r.fed.data();
    const size_t len = r.fed.size();
    SubsetHeader h{};
    for (;;) {
        // skip_until(0xFF): consumes through the first 0xFF
        const uint8_t* q = r.pos < len ? (const uint8_t*)memchr(f + r.pos, 0xFF, len - r.pos) : nullptr;
        if (!q) {
            r.pos = len;
            return r.eof ? E_IO : FLACB200_NEED_DATA;   // "eof looking for frame sync" (:1198)
        }
        const size_t at = (size_t)(q - f);
        if (at + 1 >= len) {   // the byte behind the 0xFF is not here yet
            r.pos = at;
            return r.eof ? E_IO : FLACB200_NEED_DATA;
        }
        if ((f[at + 1] >> 1) != 0x7C) {   // Ok(_) => continue: nothing but the 0xFF was consumed
            r.pos = at + 1;
            continue;
        }
        size_t consumed = 0;
        const int rc = parse_subset_header(f + at, len - at, &h, &consumed);
        if (rc == E_IO && !r.eof) {   // the header is not complete yet
            r.pos = at;
            return FLACB200_NEED_DATA;
        }
        if (rc != 0) {   // `if let Ok(header)` fails: scan on behind what the parser took
            r.pos = at + std::max<size_t>(consumed, 1);
            continue;
        }
        r.pos = at;
        break;
    }
    // ---- decode the run ----
    flacb200_stream_params prm{};
    prm.sample_rate = h.sample_rate;
    prm.bits_per_sample = h.bps;
    prm.channels = h.channels;
    prm.subset = 1;
    r.channels = h.channels;
    r.bps = h.bps;
    size_t want = r.window_bytes;
    uint64_t cap_pcm = (uint64_t)4 << 20;
    for (int attempt = 0; attempt < 24; attempt++) {
        const size_t avail = len - r.pos, take = std::min(avail, want);
        const bool last_window = take == avail && r.eof;
        cap_pcm = std::max<uint64_t>(cap_pcm, h.block_size);
        if (!r.win.reserve((size_t)cap_pcm * h.channels * 4 + 64, true)) return FLACB200_E_OUT_OF_MEMORY;
        flacb200_decode_segment seg{0, take, 0, 0};
        uint64_t nf = 0, ns = 0, bad = 0;
        const int rc = flacb200_decode(r.engine, &prm, f + r.pos, take, FLACB200_HOST, &seg, 1, r.win.p, (size_t)cap_pcm * h.channels * 4,
                                       FLACB200_PCM_I32_INTERLEAVED, FLACB200_HOST, 0, &nf, &ns, &bad);
        if (rc < 0 && rc != FLACB200_E_OUTPUT_TOO_SMALL) return rc;
        if (nf) {
            uint64_t ntab = 0;
            r.frames.resize((size_t)nf);
            const int rt = flacb200_decode_last_frames(r.engine, r.frames.data(), r.frames.size(), &ntab);
            if (rt) return rt;
            if (ntab != nf) return FLACB200_E_BAD_ARGUMENT;
            r.run_base = r.pos;
            for (const flacb200_frame_entry& e : r.frames) {   // the rate is each frame's own
                SubsetHeader fh{};
                size_t c = 0;
                parse_subset_header(f + r.pos + e.byte_offset, len - r.pos - (size_t)e.byte_offset, &fh, &c);
                r.rates.push_back(fh.sample_rate);
            }
            const flacb200_frame_entry& l = r.frames.back();
            r.pos += (size_t)(l.byte_offset + l.byte_length);
            return 0;
        }
        if (rc == FLACB200_E_OUTPUT_TOO_SMALL) {
            cap_pcm *= 2;
            continue;
        }
        if ((rc == E_IO || rc == 0) && !last_window) {   // the first frame is cut by the window / by what has been fed
            if (take == avail) return FLACB200_NEED_DATA;
            want *= 2;
            continue;
        }
        // the frame at the head fails: the serial reader returns that error and has consumed the frame's bytes; the next
        // read scans on behind its sync code
        r.pos += 2;
        return rc ? rc : E_IO;
    }
    return FLACB200_E_OUTPUT_TOO_SMALL;
}

}   // namespace

extern "C" {

int flacb200_stream_reader_open(flacb200_engine* engine, flacb200_stream_reader** out)
{
    if (!out) return FLACB200_E_BAD_ARGUMENT;
    flacb200_stream_reader* r = new flacb200_stream_reader();
    r->engine = engine;
    *out = r;
    return 0;
}

void flacb200_stream_reader_close(flacb200_stream_reader* r) { delete r; }

int flacb200_stream_reader_feed(flacb200_stream_reader* r, const uint8_t* bytes, size_t len, int eof)
{
    if (!r || (!bytes && len) || (r->eof && len)) return FLACB200_E_BAD_ARGUMENT;
    if (r->cur >= r->frames.size() && r->pos) {   // nothing borrowed from the buffer: drop what has been scanned
        r->fed.erase(r->fed.begin(), r->fed.begin() + (long)r->pos);
        r->pos = 0;
    }
    r->fed.insert(r->fed.end(), bytes, bytes + len);
    if (eof) r->eof = true;
    return 0;
}

int flacb200_stream_reader_read(flacb200_stream_reader* r, flacb200_framebuf* out)
{
    if (!r || !out) return FLACB200_E_BAD_ARGUMENT;
    memset(out, 0, sizeof(*out));
    if (r->cur >= r->frames.size()) {
        const int rc = stream_reader_next_run(*r);
        if (rc) return rc;
    }
    const flacb200_frame_entry& e = r->frames[r->cur];
    out->samples = reinterpret_cast<const int32_t*>(r->win.p) + (size_t)e.pcm_offset * r->channels;
    out->n_samples = (size_t)e.block_size * r->channels;
    out->sample_rate = r->rates[r->cur];
    out->channels = r->channels;
    out->bits_per_sample = r->bps;
    out->block_size = e.block_size;
    r->cur++;
    return 0;
}

// FlacStreamWriter::write (src/encode.rs:1094-1274): one subset frame from one call's interleaved samples
int flacb200_stream_write(flacb200_engine* e, const flacb200_options* opt, uint32_t sample_rate, uint32_t channels, uint32_t bits_per_sample,
                          const int32_t* samples, size_t n_samples, uint64_t frame_number, uint8_t* out, size_t out_capacity, size_t* out_len)
{
    if (!e || !opt || !out_len || (!samples && n_samples)) return FLACB200_E_BAD_ARGUMENT;
    *out_len = 0;
    if (channels == 0 || n_samples % channels) return E_SAMPLES_NOT_DIVISIBLE;   // :1103
    const size_t n = n_samples / channels;
    if (n == 0) return 0;
    if (n > 65535) return 24;   // InvalidBlockSize: a frame holds at most 65535 samples per channel (:1118)
    flacb200_options o = *opt;
    o.block_size = (uint16_t)n;
    flacb200_stream_params prm{};
    prm.sample_rate = sample_rate;
    prm.bits_per_sample = bits_per_sample;
    prm.channels = channels;
    prm.subset = 1;
    flacb200_segment seg{0, n, frame_number};
    uint64_t nf = 0, total = 0;
    flacb200_engine_set_keep_info(e, 0);
    const int rc = flacb200_encode(e, &o, &prm, samples, n_samples * 4, FLACB200_PCM_I32_INTERLEAVED, FLACB200_HOST, 0, &seg, 1, out, out_capacity,
                                   FLACB200_HOST, nullptr, 0, &nf, &total);
    if (rc) return rc;
    *out_len = (size_t)total;
    return 0;
}

}   // extern "C"

// The metadata blocks of a finished stream from its frame sizes: Encoder::finalize_inner's bookkeeping (src/encode.rs:2024-2110)
// without a writer handle -- for callers that encoded the frames of many streams in one batch (batch.cpp).
extern "C" int flacb200_build_stream_header(const flacb200_writer_options* opt, uint32_t sample_rate, uint32_t bits_per_sample, uint32_t channels,
                                            uint64_t total_pcm_frames, int total_known_at_open, const uint32_t* frame_sizes, uint64_t n_frames,
                                            const uint8_t md5[16], uint8_t* out, size_t capacity, size_t* len)
{
    if (!opt || !len || (!frame_sizes && n_frames)) return FLACB200_E_BAD_ARGUMENT;
    flacb200_writer* w = nullptr;
    int rc = flacb200_writer_open(nullptr, opt, sample_rate, bits_per_sample, channels, total_known_at_open ? total_pcm_frames : 0, &w);
    if (rc) return rc;
    if (frame_sizes) {
        uint64_t done = 0, bytes = 0;
        for (uint64_t f = 0; f < n_frames; f++) {
            const uint32_t n = (uint32_t)std::min<uint64_t>(w->block_size, total_pcm_frames - done), s = frame_sizes[f];
            w->points.push_back(SeekPt{done, bytes, n});
            bytes += s;
            if (s < MAX_FRAME_SIZE && s != 0) {
                w->min_frame = w->min_frame == 0 ? s : std::min(w->min_frame, s);
                w->max_frame = w->max_frame == 0 ? s : std::max(w->max_frame, s);
            }
            done += n;
        }
        w->pcm_frames_encoded = total_pcm_frames;
        if (md5) {
            memcpy(w->md5_final, md5, 16);
            w->md5_known = true;
        }
        build_header(*w, true);
    }
    *len = w->header.size();
    rc = 0;
    if (out) {
        if (capacity < w->header.size()) rc = FLACB200_E_OUTPUT_TOO_SMALL;
        else memcpy(out, w->header.data(), w->header.size());
    }
    flacb200_writer_close(w);
    return rc;
}
