/*
 * flacb200_stream.h -- stream-level C ABI of libflacb200.so: the host side of the reference's writer and
 * reader facades, on top of the batch frame engine of flacb200.h.
 *
 * What each handle replaces in tuffy/flac-codec 1.3.2 (paths relative to the reference tree):
 *   flacb200_writer   FlacByteWriter / FlacSampleWriter / FlacChannelWriter -> Encoder
 *                     src/encode.rs:103-405, :431-628, :713-893, Encoder::new/encode/finalize_inner :1882-2110
 *   flacb200_reader   FlacByteReader / FlacSampleReader -> Decoder
 *                     src/decode.rs:103-371, :384-620, Decoder::read_frame :1388, verify :1282-1309
 *   (FlacStreamWriter / FlacStreamReader, src/encode.rs:1063-1290 and src/decode.rs:1158-1268, have no stream state
 *    besides the frame number: they map to flacb200_encode / flacb200_decode with params.subset = 1.)
 *
 * The handles keep the reference's observable behaviour -- same metadata blocks ("fLaC", STREAMINFO, SEEKTABLE
 * placeholder, PADDING; src/encode.rs:1920-1951, src/metadata/mod.rs:904-976), same buffering of partial blocks, same
 * MD5 and seek points, same errors (return value = 1-based ordinal of flac_codec::Error, as in flacb200.h) -- but
 * encode `launch_frames` blocks per GPU launch instead of one frame per call.  Byte sinks/sources stay with the
 * caller (the Rust shim owns `W: Write + Seek` / `R: Read`): a writer hands out byte ranges to append and, at
 * finalize, the metadata to rewrite at the stream start; a reader is given the whole file image.
 *
 * Handles are not thread-safe (the reference's `&mut self`).  Without a GPU every data call fails with
 * FLACB200_E_NO_DEVICE; only the metadata helpers (writer header, reader open/info, md5) run on the host alone.
 */
#ifndef FLACB200_STREAM_H
#define FLACB200_STREAM_H

#include "flacb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* flac_codec::encode::Options (src/encode.rs:1363-1408): frame options + the container options */
typedef struct flacb200_writer_options {
    flacb200_options frame;
    int32_t padding;          /* PADDING block size in bytes; < 0 = no padding block (Options::no_padding :1510) */
    uint32_t seektable_kind;  /* 0 none, 1 every seektable_n seconds, 2 every seektable_n frames (SeekTableInterval :1321) */
    uint32_t seektable_n;
    uint32_t launch_frames;   /* blocks buffered per GPU launch; 0 = default (4096) */
} flacb200_writer_options;

void flacb200_writer_options_default(flacb200_writer_options* o); /* Options::default(): padding 4096, seektable 10 s */
void flacb200_writer_options_fast(flacb200_writer_options* o);
void flacb200_writer_options_best(flacb200_writer_options* o);

typedef struct flacb200_writer flacb200_writer;

/* Encoder::new (src/encode.rs:1882).  total_pcm_frames: inter-channel samples expected, 0 = unknown
 * (the facades' total_bytes / total_samples divided down by the caller; flacb200_total_from_bytes/_samples
 * reproduce their checks).  engine may be NULL for header-only use. */
int flacb200_writer_open(flacb200_engine* engine, const flacb200_writer_options* opt, uint32_t sample_rate,
                         uint32_t bits_per_sample, uint32_t channels, uint64_t total_pcm_frames, flacb200_writer** out);
void flacb200_writer_close(flacb200_writer* w); /* does NOT finalize (a Rust Drop impl calls finalize first) */

/* FlacByteWriter::new / FlacSampleWriter::new argument checks (src/encode.rs:170-178, :516-521):
 * 0 and *pcm_frames on success, else SamplesNotDivisibleByChannels / InvalidTotalBytes / InvalidTotalSamples */
int flacb200_total_from_bytes(uint64_t total_bytes, uint32_t bits_per_sample, uint32_t channels, uint64_t* pcm_frames);
int flacb200_total_from_samples(uint64_t total_samples, uint32_t channels, uint64_t* pcm_frames);

/* The metadata blocks as they stand now ("fLaC" + STREAMINFO [+ SEEKTABLE] [+ PADDING]): what Encoder::new writes
 * before the first frame, and after flacb200_writer_finalize what finalize_inner rewrites at the stream start
 * (same length unless a SEEKTABLE was carved out of the PADDING, in which case it is still the same length).
 * The pointer stays valid until the next call on the writer. */
int flacb200_writer_header(flacb200_writer* w, const uint8_t** bytes, size_t* len);

/* FlacByteWriter::write (:347), FlacSampleWriter::write (:560), FlacChannelWriter::write (:832).
 * Input is copied; whole blocks are encoded whenever launch_frames of them are buffered. */
int flacb200_writer_write_bytes(flacb200_writer* w, const uint8_t* pcm, size_t n_bytes, int big_endian);
int flacb200_writer_write_samples(flacb200_writer* w, const int32_t* interleaved, size_t n_samples);
int flacb200_writer_write_channels(flacb200_writer* w, const int32_t* const* channels, uint32_t n_channels,
                                   size_t n_per_channel);

/* Frames completed since the last drain, to be appended to the byte sink.  Valid until the next call on w. */
int flacb200_writer_drain(flacb200_writer* w, const uint8_t** frames, size_t* len);
/* Encode everything buffered that forms whole blocks now (std::io::Write::flush never emits a partial block). */
int flacb200_writer_flush(flacb200_writer* w);

/* finalize_inner (:236, :587, :2024): encodes the final short block, checks/updates the sample count
 * (SampleCountMismatch, NoSamples, ExcessiveTotalSamples), stores the MD5, fills the seek table.  Afterwards
 * flacb200_writer_drain returns the last frames and flacb200_writer_header the final metadata. */
int flacb200_writer_finalize(flacb200_writer* w);

typedef struct flacb200_writer_stats {
    uint64_t pcm_frames_written, frames_written, frame_bytes_written;
    uint32_t min_frame_size, max_frame_size;
    uint32_t launches;       /* flacb200_encode calls issued */
    uint8_t md5[16];
} flacb200_writer_stats;
int flacb200_writer_get_stats(flacb200_writer* w, flacb200_writer_stats* s);

/* ---------------------------------------------------------------------------------------------------------- */

/* Streaminfo (src/metadata/mod.rs:1640-1760) + where the frames start */
typedef struct flacb200_streaminfo {
    uint16_t min_block_size, max_block_size;
    uint32_t min_frame_size, max_frame_size;
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;
    uint64_t total_samples;   /* inter-channel samples, 0 = unknown */
    uint8_t md5[16];
    uint64_t frames_start;    /* byte offset of the first frame */
    uint32_t n_seekpoints;
    uint32_t reserved;
} flacb200_streaminfo;

/* metadata walk of FlacByteReader::new -> BlockList::read (src/metadata/mod.rs:482-646); host only */
int flacb200_read_streaminfo(const uint8_t* flac, size_t len, flacb200_streaminfo* si);

typedef struct flacb200_seekpoint {
    uint64_t sample_offset, byte_offset;
    uint32_t frame_samples, placeholder;
} flacb200_seekpoint;

typedef struct flacb200_reader flacb200_reader;

/* Returned by the reader calls in feed mode when the bytes buffered so far hold no complete frame (or not yet all
 * metadata blocks): feed more and call again.  Not an error of the stream. */
#define FLACB200_NEED_DATA (-10)
/* Returned by flacb200_reader_seek on a fed reader whose source can seek (flacb200_reader_set_seekable): reposition the
 * source to the absolute file offset flacb200_reader_wanted_offset reports, then call flacb200_reader_seek again with the
 * same sample and keep feeding from there. */
#define FLACB200_NEED_SEEK (-11)

/* The reader is the reference's Decoder (src/decode.rs:1311-1491) behind FlacByteReader / FlacSampleReader /
 * FlacChannelReader.  It decodes WINDOW by window: a run of bytes that starts at a frame boundary goes through one
 * flacb200_decode call, its frames are handed out one at a time, and the next window starts where the last good frame
 * ended -- memory is bounded by the window (default 32 MiB of frames / 8 Mi inter-channel samples), not by the stream.
 * An error of the stream is returned when the reader gets to the failing frame, after every frame in front of it has been
 * delivered (Decoder::read_frame's behaviour), and is sticky from then on.
 *
 * Two sources:
 *   flacb200_reader_open         `R: Read + Seek` (new_seekable / open): a file image that stays valid while the reader lives
 *   flacb200_reader_open_stream  `R: Read` (new): bytes arrive through flacb200_reader_feed; calls return FLACB200_NEED_DATA
 *                                until enough has been fed; decoded bytes are dropped; not seekable
 * engine may be NULL for metadata-only use. */
int flacb200_reader_open(flacb200_engine* engine, const uint8_t* flac, size_t len, flacb200_reader** out);
int flacb200_reader_open_stream(flacb200_engine* engine, flacb200_reader** out);
int flacb200_reader_feed(flacb200_reader* r, const uint8_t* bytes, size_t len, int eof);
void flacb200_reader_close(flacb200_reader* r);
/* window limits: bytes of frames / inter-channel samples per GPU call; 0 keeps the current value */
int flacb200_reader_set_window(flacb200_reader* r, size_t window_bytes, uint64_t window_pcm_frames);
int flacb200_reader_info(flacb200_reader* r, flacb200_streaminfo* si);
int flacb200_reader_seektable(flacb200_reader* r, flacb200_seekpoint* points, size_t capacity, size_t* n_points);

/* FlacSampleReader::read / FlacByteReader::read (src/decode.rs:274-303, :420-450): up to `capacity` single-channel samples
 * (or bytes) from the current position, whole PCM frames only; *n_out = how many were delivered (0 at the end of the stream).
 * pcm_kind: FLACB200_PCM_BYTES_LE / _BE (capacity and *n_out in bytes) or FLACB200_PCM_I32_INTERLEAVED (in samples). */
int flacb200_reader_read(flacb200_reader* r, void* out, size_t capacity, int pcm_kind, size_t* n_out);
/* FlacSampleReader::fill_buf / consume (:466-492): the unconsumed interleaved samples of the current frame (a borrow that
 * lives until the next call on r; empty at the end of the stream); consume counts samples of all channels. */
int flacb200_reader_fill_buf(flacb200_reader* r, const int32_t** samples, size_t* n_samples);
int flacb200_reader_consume(flacb200_reader* r, size_t n_samples);
/* FlacChannelReader::fill_buf / consume (:917-949): one pointer per channel over the unconsumed part of the current frame
 * (*channels: array of streaminfo.channels pointers, n_per_channel samples each), consume counts samples per channel. */
int flacb200_reader_fill_channels(flacb200_reader* r, const int32_t* const** channels, size_t* n_per_channel);
int flacb200_reader_consume_channels(flacb200_reader* r, size_t n_per_channel);
/* FlacSampleReader::seek / FlacChannelReader::seek / io::Seek of FlacByteReader (:715-860, :1021-1057): Decoder::seek
 * (:1452-1491) repositions to the last SEEKTABLE point at or before the sample (the first frame without a table) and the
 * frames up to the sample are decoded and skipped.  pcm_frame: inter-channel samples; beyond the end -> InvalidSeek. */
int flacb200_reader_seek(flacb200_reader* r, uint64_t pcm_frame);
/* feed mode over an `R: Read + Seek` (new_seekable without holding the file in memory): see FLACB200_NEED_SEEK */
int flacb200_reader_set_seekable(flacb200_reader* r, int seekable);
int flacb200_reader_wanted_offset(flacb200_reader* r, uint64_t* offset);
/* verify_reader (src/decode.rs:1291-1309): decode from the current position to the end, MD5 of the little-endian PCM against
 * STREAMINFO.  *result: 0 MD5Match, 1 MD5Mismatch, 2 NoMD5 (all-zero sum stored) */
int flacb200_reader_verify(flacb200_reader* r, int* result, uint8_t md5_out[16]);

/* ---------------------------------------------------------------------------------------------------------- */
/* FlacStreamWriter::write (src/encode.rs:1094-1274): one subset frame (no metadata, parameters in every header) from the
 * n_samples interleaved samples of one call; frame_number is the writer's running count.  The frame goes to out. */
int flacb200_stream_write(flacb200_engine* e, const flacb200_options* opt, uint32_t sample_rate, uint32_t channels,
                          uint32_t bits_per_sample, const int32_t* samples, size_t n_samples, uint64_t frame_number,
                          uint8_t* out, size_t out_capacity, size_t* out_len);

/* FlacStreamReader (src/decode.rs:1149-1268): subset frames without metadata; every FrameBuf carries the parameters of its
 * own frame header (FrameBuf, :1253-1268) -- the caller passes none.  Bytes are fed like an `R: BufRead`; read returns the
 * next frame, FLACB200_NEED_DATA when more bytes are needed, the frame's error, or Io at the end of the stream ("eof looking
 * for frame sync", :1198).  Runs of frames with the same channel count and sample width are decoded in one GPU call. */
typedef struct flacb200_stream_reader flacb200_stream_reader;
typedef struct flacb200_framebuf {
    const int32_t* samples;   /* interleaved; valid until the next call on the reader */
    size_t n_samples;         /* block_size * channels */
    uint32_t sample_rate, channels, bits_per_sample, block_size;
} flacb200_framebuf;
int flacb200_stream_reader_open(flacb200_engine* engine, flacb200_stream_reader** out);
void flacb200_stream_reader_close(flacb200_stream_reader* r);
int flacb200_stream_reader_feed(flacb200_stream_reader* r, const uint8_t* bytes, size_t len, int eof);
int flacb200_stream_reader_read(flacb200_stream_reader* r, flacb200_framebuf* out);

/* MD5 as the encoder/verify use it (host); exported for the shim and the tests */
void flacb200_md5(const uint8_t* data, size_t len, uint8_t out[16]);
/* The same sum for many buffers at once: eight streams advance in the lanes of one AVX2 register per host thread
 * (csrc/md5_mb.h); threads = 0: all hardware threads.  digests: 16 bytes per buffer. */
void flacb200_md5_many(const uint8_t* const* data, const size_t* len, size_t n, uint8_t* digests, unsigned threads);

/* ---------------------------------------------------------------------------------------------------------- */
/* Whole-file batches over one or several GPUs: the file-level fan-out the reference does with rayon
 * (examples/flac2wav.rs:31-38, examples/flac-split.rs:84-87) -- many independent streams, each one a complete
 * FlacByteWriter / FlacSampleWriter run (Encoder::new .. finalize_inner, src/encode.rs:1882-2110) or FlacByteReader run.
 *
 * Tracks are dealt to the devices by size; every device has its own host thread and engine and works through its share
 * in sub-batches (upload of sub-batch k + 1, kernels of k and download of k - 1 overlap).  The only cross-device step is
 * on the host: per-track frame sizes are summed (an exclusive scan per track) to place each track's frames behind its
 * metadata blocks in its own file image -- no collective, no device-to-device traffic.  The MD5 of
 * every track is computed meanwhile on the remaining host threads (flacb200_md5_many); STREAMINFO (min/max frame size,
 * total samples, MD5), SEEKTABLE and PADDING are then written as Encoder::finalize_inner would.
 */
/* The metadata blocks of a finished stream from its frame sizes -- Encoder::finalize_inner's bookkeeping (src/encode.rs:2024-2110:
 * min/max frame size, total samples, MD5, seek points filtered by the table interval, placeholder table or a table carved from
 * the padding) without a writer handle, for callers that encoded the frames of many streams in one batch.  frame_sizes NULL:
 * the blocks as Encoder::new writes them (same length).  out NULL: only *len. */
int flacb200_build_stream_header(const flacb200_writer_options* opt, uint32_t sample_rate, uint32_t bits_per_sample,
                                 uint32_t channels, uint64_t total_pcm_frames, int total_known_at_open,
                                 const uint32_t* frame_sizes, uint64_t n_frames, const uint8_t md5[16], uint8_t* out,
                                 size_t capacity, size_t* len);

typedef struct flacb200_track {
    const void* pcm;            /* host memory (pinned memory uploads at full link speed); interleaved samples */
    uint64_t n_pcm_frames;      /* inter-channel samples */
    uint32_t sample_rate, bits_per_sample, channels;
    int32_t pcm_kind;           /* FLACB200_PCM_BYTES_LE, _BYTES_BE or _I32_INTERLEAVED */
} flacb200_track;

typedef struct flacb200_file {
    uint8_t* data;              /* in: NULL (the library allocates; release with flacb200_files_free) or a caller buffer */
    size_t capacity;            /* in: size of the caller buffer (flacb200_encode_batch_bound says how much can be needed) */
    size_t len;                 /* out: bytes of the complete .flac file */
    int32_t status;             /* out: 0 or the error of this track (same codes as the writer calls) */
    uint32_t frames;            /* out */
    uint8_t md5[16];            /* out: the STREAMINFO signature */
} flacb200_file;

size_t flacb200_encode_batch_bound(const flacb200_track* track, const flacb200_writer_options* opt);
/* devices: CUDA device ordinals (NULL / n_devices 0: device 0).  Returns 0 when every track was encoded (files[i].status
 * all 0), else the first failing track's status. */
int flacb200_encode_batch(const flacb200_track* tracks, size_t n_tracks, const flacb200_writer_options* opt,
                          const int* devices, int n_devices, flacb200_file* files);
void flacb200_files_free(flacb200_file* files, size_t n_files);

typedef struct flacb200_pcm {
    void* data;                 /* in: NULL (library allocates) or caller buffer; out: interleaved PCM */
    size_t capacity;
    size_t len;                 /* out: bytes */
    int32_t status;             /* out: 0 or this stream's error (flac_codec::Error ordinal) */
    int32_t verified;           /* out (when verify != 0): 0 MD5Match, 1 MD5Mismatch, 2 NoMD5 */
    flacb200_streaminfo info;   /* out */
} flacb200_pcm;

/* flac2wav-style fan-out: decodes n complete .flac images.  pcm_kind: FLACB200_PCM_BYTES_LE / _BE / _I32_INTERLEAVED.
 * verify != 0 also checks every stream's MD5 (verify_reader, src/decode.rs:1291). */
int flacb200_decode_batch(const uint8_t* const* flac, const size_t* flac_len, size_t n_files, int pcm_kind, int verify,
                          const int* devices, int n_devices, flacb200_pcm* out);
void flacb200_pcm_free(flacb200_pcm* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
