// flacb200.hpp -- header-only C++ mirror of flac-codec's writer/reader facades over the C ABI of libflacb200.so.
//
// Same type names, constructor arguments and error behaviour as the reference
//   FlacByteWriter / FlacSampleWriter / FlacChannelWriter   src/encode.rs:103-893
//   FlacSampleReader / FlacByteReader / verify              src/decode.rs:103-620, :1282-1309
// with std::ostream (seekable) / std::istream standing in for `W: Write + Seek` / `R: Read`.
// Errors are thrown as flacb200::Error carrying the ordinal of the matching flac_codec::Error variant.
#pragma once
#include <cstdint>
#include <istream>
#include <iterator>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/flacb200_stream.h"

namespace flacb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* what) : std::runtime_error(std::string(what) + ": " + flacb200_strerror(c)), code(c) {}
};

inline void check(int rc, const char* what)
{
    if (rc != 0) throw Error(rc, what);
}

// flac_codec::encode::Options (src/encode.rs:1363-1672)
struct Options {
    flacb200_writer_options o;
    Options() { flacb200_writer_options_default(&o); }
    static Options fast() { Options r; flacb200_writer_options_fast(&r.o); return r; }
    static Options best() { Options r; flacb200_writer_options_best(&r.o); return r; }
    Options& block_size(uint16_t n) { if (n < 16) throw std::invalid_argument("block size must be >= 16"); o.frame.block_size = n; return *this; }
    Options& max_lpc_order(std::optional<uint8_t> n)
    {
        if (n && (*n == 0 || *n > 32)) throw std::invalid_argument("maximum LPC order must be <= 32");
        o.frame.max_lpc_order = n.value_or(0);
        return *this;
    }
    Options& max_partition_order(uint32_t n) { if (n > 15) throw std::invalid_argument("max partition order must be <= 15"); o.frame.max_partition_order = (uint8_t)n; return *this; }
    Options& mid_side(bool on) { o.frame.mid_side = on; return *this; }
    Options& fast_channel_correlation(bool fast) { o.frame.exhaustive_channel_correlation = !fast; return *this; }
    Options& padding(uint32_t size) { o.padding = size ? (int32_t)size : -1; return *this; }
    Options& no_padding() { o.padding = -1; return *this; }
    Options& seektable_seconds(uint8_t s) { o.seektable_kind = s ? 1 : 0; o.seektable_n = s; return *this; }
    Options& seektable_frames(size_t n) { o.seektable_kind = n ? 2 : 0; o.seektable_n = (uint32_t)n; return *this; }
    Options& no_seektable() { o.seektable_kind = 0; return *this; }
    Options& launch_frames(uint32_t n) { o.launch_frames = n; return *this; }   // engine extension: blocks per GPU launch
};

class Engine {
public:
    explicit Engine(int device = 0) { check(flacb200_engine_create(device, &e_), "flacb200_engine_create"); }
    ~Engine() { flacb200_engine_destroy(e_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    flacb200_engine* get() const { return e_; }

private:
    flacb200_engine* e_ = nullptr;
};

// Encoder<W> (src/encode.rs:1860-2110)
class Encoder {
public:
    Encoder(std::ostream& w, Engine& eng, const Options& opt, uint32_t rate, uint32_t bps, uint8_t channels, uint64_t total_pcm_frames)
        : w_(w)
    {
        check(flacb200_writer_open(eng.get(), &opt.o, rate, bps, channels, total_pcm_frames, &h_), "Encoder::new");
        start_ = w_.tellp();
        put_header();
    }
    ~Encoder()
    {
        try { finalize(); } catch (...) {}   // Drop finalises and swallows errors (:2113)
        flacb200_writer_close(h_);
    }
    Encoder(const Encoder&) = delete;
    Encoder& operator=(const Encoder&) = delete;

    void finalize()   // finalize_inner (:2024)
    {
        if (finalized_) return;
        finalized_ = true;
        check(flacb200_writer_finalize(h_), "finalize");
        drain();
        const std::ostream::pos_type end = w_.tellp();
        w_.seekp(start_);
        put_header();
        w_.seekp(end);
    }
    void after(int rc, const char* what)
    {
        check(rc, what);
        drain();
    }
    flacb200_writer* handle() const { return h_; }
    std::ostream& sink() { return w_; }

private:
    void put_header()
    {
        const uint8_t* p = nullptr;
        size_t n = 0;
        check(flacb200_writer_header(h_, &p, &n), "writer_header");
        w_.write(reinterpret_cast<const char*>(p), (std::streamsize)n);
    }
    void drain()
    {
        const uint8_t* p = nullptr;
        size_t n = 0;
        check(flacb200_writer_drain(h_, &p, &n), "writer_drain");
        if (n) w_.write(reinterpret_cast<const char*>(p), (std::streamsize)n);
    }
    std::ostream& w_;
    flacb200_writer* h_ = nullptr;
    std::ostream::pos_type start_;
    bool finalized_ = false;
};

inline uint64_t total_from_samples(std::optional<uint64_t> t, uint32_t channels)
{
    uint64_t n = 0;
    if (t) check(flacb200_total_from_samples(*t, channels, &n), "FlacSampleWriter::new");
    return n;
}
inline uint64_t total_from_bytes(std::optional<uint64_t> t, uint32_t bps, uint32_t channels)
{
    uint64_t n = 0;
    if (t) check(flacb200_total_from_bytes(*t, bps, channels, &n), "FlacByteWriter::new");
    return n;
}

// FlacSampleWriter<W> (src/encode.rs:431-628)
class FlacSampleWriter {
public:
    FlacSampleWriter(std::ostream& w, Engine& eng, const Options& opt, uint32_t rate, uint32_t bps, uint8_t channels,
                     std::optional<uint64_t> total_samples = std::nullopt)
        : enc_(w, eng, opt, rate, bps, channels, total_from_samples(total_samples, channels)) {}
    void write(const int32_t* samples, size_t n) { enc_.after(flacb200_writer_write_samples(enc_.handle(), samples, n), "FlacSampleWriter::write"); }
    void write(const std::vector<int32_t>& s) { write(s.data(), s.size()); }
    void finalize() { enc_.finalize(); }

private:
    Encoder enc_;
};

// FlacByteWriter<W, E> (src/encode.rs:103-405)
class FlacByteWriter {
public:
    FlacByteWriter(std::ostream& w, Engine& eng, const Options& opt, uint32_t rate, uint32_t bps, uint8_t channels,
                   std::optional<uint64_t> total_bytes = std::nullopt, bool big_endian = false)
        : enc_(w, eng, opt, rate, bps, channels, total_from_bytes(total_bytes, bps, channels)), big_(big_endian) {}
    size_t write(const uint8_t* buf, size_t n)   // always consumes the whole slice (:387)
    {
        enc_.after(flacb200_writer_write_bytes(enc_.handle(), buf, n, big_), "FlacByteWriter::write");
        return n;
    }
    void flush()
    {
        enc_.after(flacb200_writer_flush(enc_.handle()), "flush");
        enc_.sink().flush();
    }
    void finalize() { enc_.finalize(); }

private:
    Encoder enc_;
    bool big_;
};

// FlacChannelWriter<W> (src/encode.rs:713-893)
class FlacChannelWriter {
public:
    FlacChannelWriter(std::ostream& w, Engine& eng, const Options& opt, uint32_t rate, uint32_t bps, uint8_t channels,
                      std::optional<uint64_t> total_samples = std::nullopt)
        : enc_(w, eng, opt, rate, bps, channels, check_total(total_samples)) {}
    void write(const std::vector<std::vector<int32_t>>& channels)
    {
        const size_t n = channels.empty() ? 0 : channels[0].size();
        std::vector<const int32_t*> ptrs;
        for (const auto& c : channels) {
            if (c.size() != n) throw Error(65, "FlacChannelWriter::write");   // ChannelLengthMismatch (:851)
            ptrs.push_back(c.data());
        }
        enc_.after(flacb200_writer_write_channels(enc_.handle(), ptrs.data(), (uint32_t)ptrs.size(), n), "FlacChannelWriter::write");
    }
    void finalize() { enc_.finalize(); }

private:
    static uint64_t check_total(std::optional<uint64_t> t)
    {
        if (t && *t == 0) throw Error(63, "FlacChannelWriter::new");   // InvalidTotalSamples
        return t.value_or(0);
    }
    Encoder enc_;
};

enum class Verified { MD5Match, MD5Mismatch, NoMD5 };

// FlacSampleReader<R> / FlacByteReader<R, E> / FlacChannelReader<R> (src/decode.rs:103-1065) over a std::istream: the handle
// is FED from the stream chunk by chunk as its decode windows ask for bytes (FLACB200_NEED_DATA) and repositions a seekable
// stream when a seek asks for it (FLACB200_NEED_SEEK) -- the file is never held in memory.
class FlacReader {
public:
    FlacReader(std::istream& r, Engine& eng, bool seekable = true) : in_(r), chunk_(1 << 20)
    {
        check(flacb200_reader_open_stream(eng.get(), &h_), "FlacReader::new");
        if (seekable) check(flacb200_reader_set_seekable(h_, 1), "set_seekable");
        call([&] { return flacb200_reader_info(h_, &si_); }, "BlockList::read");
    }
    ~FlacReader() { flacb200_reader_close(h_); }
    FlacReader(const FlacReader&) = delete;
    FlacReader& operator=(const FlacReader&) = delete;
    // Metadata trait (src/metadata/mod.rs:48-105)
    uint8_t channel_count() const { return (uint8_t)si_.channels; }
    uint32_t sample_rate() const { return si_.sample_rate; }
    uint32_t bits_per_sample() const { return si_.bits_per_sample; }
    std::optional<uint64_t> total_samples() const { return si_.total_samples ? std::optional<uint64_t>(si_.total_samples) : std::nullopt; }
    const flacb200_streaminfo& streaminfo() const { return si_; }
    size_t read(int32_t* samples, size_t n)   // FlacSampleReader::read
    {
        size_t got = 0;
        call([&] { return flacb200_reader_read(h_, samples, n, FLACB200_PCM_I32_INTERLEAVED, &got); }, "FlacSampleReader::read");
        return got;
    }
    size_t read_bytes(uint8_t* buf, size_t n, bool big_endian = false)   // FlacByteReader::read
    {
        size_t got = 0;
        call([&] { return flacb200_reader_read(h_, buf, n, big_endian ? FLACB200_PCM_BYTES_BE : FLACB200_PCM_BYTES_LE, &got); }, "FlacByteReader::read");
        return got;
    }
    // FlacSampleReader::fill_buf / consume (:466-492): the unconsumed interleaved samples of the current frame
    std::pair<const int32_t*, size_t> fill_buf()
    {
        const int32_t* p = nullptr;
        size_t n = 0;
        call([&] { return flacb200_reader_fill_buf(h_, &p, &n); }, "fill_buf");
        return {p, n};
    }
    void consume(size_t n_samples) { check(flacb200_reader_consume(h_, n_samples), "consume"); }
    // FlacChannelReader::fill_buf / consume (:917-949): one pointer per channel, n samples each
    std::pair<const int32_t* const*, size_t> fill_channels()
    {
        const int32_t* const* pp = nullptr;
        size_t n = 0;
        call([&] { return flacb200_reader_fill_channels(h_, &pp, &n); }, "FlacChannelReader::fill_buf");
        return {pp, n};
    }
    void consume_channels(size_t n_per_channel) { check(flacb200_reader_consume_channels(h_, n_per_channel), "consume"); }
    void seek(uint64_t sample) { call([&] { return flacb200_reader_seek(h_, sample); }, "seek"); }
    Verified verify()
    {
        int res = 0;
        call([&] { return flacb200_reader_verify(h_, &res, nullptr); }, "verify");
        return res == 0 ? Verified::MD5Match : res == 1 ? Verified::MD5Mismatch : Verified::NoMD5;
    }

private:
    template <class F>
    void call(F f, const char* what)
    {
        for (;;) {
            const int rc = f();
            if (rc == FLACB200_NEED_DATA) {
                if (eof_) throw Error(1, what);   // Io: the stream ended inside a frame
                in_.read(chunk_.data(), (std::streamsize)chunk_.size());
                const size_t n = (size_t)in_.gcount();
                eof_ = n == 0;
                check(flacb200_reader_feed(h_, reinterpret_cast<const uint8_t*>(chunk_.data()), n, eof_), "feed");
            } else if (rc == FLACB200_NEED_SEEK) {
                uint64_t off = 0;
                check(flacb200_reader_wanted_offset(h_, &off), "wanted_offset");
                in_.clear();
                in_.seekg((std::streamoff)off);
                eof_ = false;
            } else {
                check(rc, what);
                return;
            }
        }
    }
    std::istream& in_;
    std::vector<char> chunk_;
    bool eof_ = false;
    flacb200_reader* h_ = nullptr;
    flacb200_streaminfo si_{};
};

// FlacStreamWriter<W> (src/encode.rs:1063-1290): one subset frame per call, parameters in every header
class FlacStreamWriter {
public:
    FlacStreamWriter(std::ostream& w, Engine& eng, const Options& o) : w_(w), eng_(eng), o_(o) {}
    void write(uint32_t sample_rate, uint8_t channels, uint32_t bits_per_sample, const int32_t* samples, size_t n_samples)
    {
        buf_.resize(n_samples * 5 + 1024);
        size_t n = 0;
        check(flacb200_stream_write(eng_.get(), &o_.o.frame, sample_rate, channels, bits_per_sample, samples, n_samples, frame_number_, buf_.data(),
                                    buf_.size(), &n), "FlacStreamWriter::write");
        if (n) {
            frame_number_++;
            w_.write(reinterpret_cast<const char*>(buf_.data()), (std::streamsize)n);
        }
    }

private:
    std::ostream& w_;
    Engine& eng_;
    Options o_;
    uint64_t frame_number_ = 0;
    std::vector<uint8_t> buf_;
};

// FlacStreamReader<R> (src/decode.rs:1149-1268): read() returns the next frame with the parameters of its own header
class FlacStreamReader {
public:
    FlacStreamReader(std::istream& r, Engine& eng) : in_(r), chunk_(1 << 20) { check(flacb200_stream_reader_open(eng.get(), &h_), "FlacStreamReader::new"); }
    ~FlacStreamReader() { flacb200_stream_reader_close(h_); }
    FlacStreamReader(const FlacStreamReader&) = delete;
    FlacStreamReader& operator=(const FlacStreamReader&) = delete;
    flacb200_framebuf read()
    {
        flacb200_framebuf fb{};
        for (;;) {
            const int rc = flacb200_stream_reader_read(h_, &fb);
            if (rc != FLACB200_NEED_DATA) {
                check(rc, "FlacStreamReader::read");
                return fb;
            }
            if (eof_) throw Error(1, "eof looking for frame sync");
            in_.read(chunk_.data(), (std::streamsize)chunk_.size());
            const size_t n = (size_t)in_.gcount();
            eof_ = n == 0;
            check(flacb200_stream_reader_feed(h_, reinterpret_cast<const uint8_t*>(chunk_.data()), n, eof_), "feed");
        }
    }

private:
    std::istream& in_;
    std::vector<char> chunk_;
    bool eof_ = false;
    flacb200_stream_reader* h_ = nullptr;
};

}   // namespace flacb200
