/*
 * flacb200.h -- C ABI of the B200-native FLAC frame engine (libflacb200.so).
 *
 * This is the drop-in boundary for the hot path of tuffy/flac-codec 1.3.2: the batch form of
 *   Encoder::encode(&Frame) -> encode_frame(..)          src/encode.rs:1997, :2259
 *   FlacStreamWriter::write(rate, ch, bps, &[i32])       src/encode.rs:1094
 *   Decoder::read_frame() -> read_subframes(..)          src/decode.rs:1388, :1494
 *   Frame::fill_from_buf / Frame::to_buf                 src/audio.rs:149, :110
 * The reference has no FFI seam of its own (#![forbid(unsafe_code)]); a Rust shim that keeps the
 * crate's public writer/reader types and calls these entry points is in rust/ and INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types
 *   - return value: 0 = OK; > 0 = 1-based ordinal of the matching flac_codec::Error variant
 *     (src/lib.rs:57-193, e.g. 40 = Crc16Mismatch); < 0 = -(1000 + cudaError_t) or FLACB200_E_*
 *   - an engine is bound to one CUDA device and one stream and is NOT thread-safe (same rule as the
 *     reference's `&mut self`); use one engine per host thread / per GPU
 *   - there is no CPU fallback: every entry point fails with FLACB200_E_NO_DEVICE without a GPU
 */
#ifndef FLACB200_H
#define FLACB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLACB200_E_NO_DEVICE (-1)
#define FLACB200_E_BAD_ARGUMENT (-2)
#define FLACB200_E_OUT_OF_MEMORY (-3)
#define FLACB200_E_OUTPUT_TOO_SMALL (-4)
#define FLACB200_E_CUDA_BASE (-1000)

/* where a buffer lives */
#define FLACB200_HOST 0
#define FLACB200_DEVICE 1

/* PCM layouts (Frame::fill_from_buf / fill_from_samples / fill_from_channels, src/audio.rs:149-225) */
#define FLACB200_PCM_BYTES_LE 0      /* interleaved, ceil(bps/8) bytes per sample, little endian */
#define FLACB200_PCM_BYTES_BE 1      /* interleaved, big endian */
#define FLACB200_PCM_I32_INTERLEAVED 2
#define FLACB200_PCM_I32_PLANAR 3    /* channel c starts at element c * planar_stride */

/* Mirrors flac_codec::encode::Options / EncoderOptions (src/encode.rs:1363-1374, :1701-1709):
 * the fields the frame engine reads.  Container options stay with the host-side writers. */
typedef struct flacb200_options {
    uint16_t block_size;                    /* Options::block_size, >= 16 (ignored for subset frames) */
    uint8_t max_lpc_order;                  /* 0 = None, else 1..=32 */
    uint8_t max_partition_order;            /* 0..=6 (the reference panics above 6) */
    uint8_t mid_side;                       /* Options::mid_side */
    uint8_t exhaustive_channel_correlation; /* Options::exhaustive_channel_correlation */
    uint8_t window_kind;                    /* 0 Rectangle, 1 Hann, 2 Tukey(tukey_p)  (Window, :1713) */
    uint8_t reserved0;
    float tukey_p;
} flacb200_options;

void flacb200_options_default(flacb200_options* o); /* Options::default()  :1376 */
void flacb200_options_fast(flacb200_options* o);    /* Options::fast()     :1635 */
void flacb200_options_best(flacb200_options* o);    /* Options::best()     :1649 */

typedef struct flacb200_stream_params {
    uint32_t sample_rate;     /* Hz, < 2^20 */
    uint32_t bits_per_sample; /* 1..=32 */
    uint32_t channels;        /* 1..=8 */
    uint32_t subset;          /* 1 = FrameHeader::write_subset / read_subset semantics (FlacStreamWriter/Reader) */
    /* decode only (STREAMINFO cross-checks, src/stream.rs:279-313); 0 = unknown */
    uint32_t max_block_size;
    uint32_t reserved;
} flacb200_stream_params;

/* A run of PCM that is cut into blocks of options.block_size (the last block may be short), numbered
 * first_frame_number, first_frame_number + 1, ...  One segment = (part of) one stream/track. */
typedef struct flacb200_segment {
    uint64_t pcm_offset;         /* index of the first inter-channel sample (PCM frame) in the pcm buffer */
    uint64_t n_pcm_frames;       /* inter-channel samples in this segment */
    uint64_t first_frame_number; /* FrameNumber of the first block */
} flacb200_segment;

typedef struct flacb200_engine flacb200_engine;

int flacb200_engine_create(int device, flacb200_engine** out);
void flacb200_engine_destroy(flacb200_engine* e);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the engine's own. */
int flacb200_engine_set_stream(flacb200_engine* e, void* cuda_stream);
/* Frames processed per internal launch group (bounds scratch memory); 0 = default */
int flacb200_engine_set_chunk_frames(flacb200_engine* e, uint32_t frames);
/* Runtime knobs for tests and experiments.  Their defaults come from the environment
 * (FLACB200_LEGACY, FLACB200_BATCH_BYTES, FLACB200_NO_BATCH, FLACB200_DEBUG), which is read once, at
 * flacb200_engine_create; keys: "legacy", "batch_bytes", "no_batch", "debug", "fused" (FLACB200_FUSED: the single-kernel
 * stereo frame encoder k_frame4 instead of k_analyze3 + k_decide + k_scan + k_pack3), "lpc_overlap". */
int flacb200_engine_set_option(flacb200_engine* e, const char* key, uint64_t value);
/* Keep per-subframe decisions of each encode call for flacb200_encode_last_info (default on, calls of
 * at most 65536 frames); switch off on throughput paths -- it costs a device-to-host copy per group. */
int flacb200_engine_set_keep_info(flacb200_engine* e, int enable);

/*
 * Encode: batch form of Encoder::encode / encode_frame (src/encode.rs:1997, :2259) and, with
 * params->subset, of FlacStreamWriter::write (:1094).
 *   pcm / pcm_location / pcm_kind : input samples (host or device memory)
 *   out / out_location            : receives the frames back to back (host or device memory)
 *   frame_bytes (host, optional)  : byte size of every frame, in segment order
 * Frames are byte-identical to what the reference encoder emits for the same blocks and options
 * (see DESIGN.md "Parity" for the one libm caveat).
 */
int flacb200_encode(flacb200_engine* e, const flacb200_options* opt, const flacb200_stream_params* params,
                    const void* pcm, size_t pcm_bytes, int pcm_kind, int pcm_location, uint64_t planar_stride,
                    const flacb200_segment* segments, size_t n_segments, void* out, size_t out_capacity,
                    int out_location, uint32_t* frame_bytes, size_t frame_bytes_capacity, uint64_t* n_frames,
                    uint64_t* total_bytes);

/* Worst-case output size for flacb200_encode with these arguments (every subframe VERBATIM). */
size_t flacb200_encode_bound(const flacb200_options* opt, const flacb200_stream_params* params,
                             const flacb200_segment* segments, size_t n_segments);

/* What the encoder chose per subframe (debug / parity tooling; mirrors the oracle's fo_subframe_info). */
typedef struct flacb200_subframe_info {
    int32_t type;  /* 0 CONSTANT, 1 VERBATIM, 2 FIXED, 3 LPC */
    int32_t order;
    int32_t wasted;
    int32_t bps;
    int32_t precision;
    int32_t shift;
    int32_t coefs[32];
    int32_t coding_method;
    int32_t partition_order;
    uint8_t rice[64];
    uint8_t kind[64];
    uint64_t bits;
} flacb200_subframe_info;

typedef struct flacb200_frame_info {
    int32_t channel_assignment;
    int32_t channels;
    uint32_t frame_bytes;
    flacb200_subframe_info sub[8];
} flacb200_frame_info;

/* Copies the decisions of the most recent flacb200_encode call (at most `capacity` frames). */
int flacb200_encode_last_info(flacb200_engine* e, flacb200_frame_info* infos, size_t capacity, uint64_t* n_frames);

/* One run of consecutive frames of one stream inside the `frames` buffer */
typedef struct flacb200_decode_segment {
    uint64_t byte_offset;    /* first byte of the first frame */
    uint64_t byte_length;    /* bytes of frame data */
    uint64_t pcm_offset;     /* index of the first inter-channel sample in the output buffer */
    uint64_t n_pcm_frames;   /* samples expected (STREAMINFO total_samples); 0 = until the bytes end */
} flacb200_decode_segment;

/*
 * Decode: batch form of Decoder::read_frame / read_subframes (src/decode.rs:1388, :1494) followed by
 * Frame::to_buf (src/audio.rs:110).  CRC-8, CRC-16 and every structural check of the reference are
 * enforced; the first failing frame's error is returned (and its index in *bad_frame).
 *   frame_offsets (host, optional): receives the byte offset of every frame found, per segment order
 */
int flacb200_decode(flacb200_engine* e, const flacb200_stream_params* params, const void* frames,
                    size_t frames_bytes, int frames_location, const flacb200_decode_segment* segments,
                    size_t n_segments, void* pcm_out, size_t pcm_out_bytes, int pcm_kind, int pcm_location,
                    uint64_t planar_stride, uint64_t* n_frames, uint64_t* n_pcm_frames, uint64_t* bad_frame);

/* Where the frames of the most recent flacb200_decode call were found, in stream order (the frames the serial reader
 * would have visited, up to the first error): what Decoder::read_frame learns one frame at a time -- byte position,
 * block size (FrameHeader::block_size), running sample position (Decoder::current_sample, src/decode.rs:1432).  The
 * readers of flacb200_stream.h use it to hand out frame-sized buffers (fill_buf) and to continue a stream window by
 * window.  Valid for a call that was not split into batches (fewer than four segments, or device buffers). */
typedef struct flacb200_frame_entry {
    uint64_t byte_offset;  /* of the frame's sync code in the `frames` buffer */
    uint64_t pcm_offset;   /* inter-channel sample index in the PCM output where the frame's samples start */
    uint32_t byte_length;  /* header .. CRC-16 */
    uint32_t block_size;   /* inter-channel samples */
} flacb200_frame_entry;
int flacb200_decode_last_frames(flacb200_engine* e, flacb200_frame_entry* table, size_t capacity, uint64_t* n_entries);

/*
 * MD5 of many streams at once: the STREAMINFO signature the reference computes while encoding (update_md5,
 * src/encode.rs:1292-1318) and checks in verify (src/decode.rs:1291-1309) -- MD5 over the samples as little-endian
 * interleaved bytes, ceil(bits_per_sample / 8) bytes each.  One digest (16 bytes, host memory) per segment;
 * segment.first_frame_number is ignored.  MD5 is serial per stream: the kernel runs one thread per segment, so this
 * pays for batches (hundreds of tracks), not for a single stream.
 */
int flacb200_md5_batch(flacb200_engine* e, const void* pcm, size_t pcm_bytes, int pcm_kind, int pcm_location,
                 uint64_t planar_stride, uint32_t channels, uint32_t bits_per_sample,
                 const flacb200_segment* segments, size_t n_segments, uint8_t* digests);

/* Device-side timing of the most recent encode/decode call, measured with CUDA events on the
 * engine's stream (milliseconds; kernels only, no copies). */
typedef struct flacb200_timings {
    float total_ms;        /* first kernel launch to last kernel end */
    float h2d_ms, d2h_ms;  /* copies, when the call had host buffers */
    float kernel_ms[8];    /* encode: planes, lpc, residual, decide+scan+zero, pack+crc16; decode: see DESIGN.md */
    uint32_t kernel_launches[8];
    uint32_t launches;     /* kernels launched by the call */
} flacb200_timings;
int flacb200_set_profiling(flacb200_engine* e, int enable); /* per-kernel events cost a few us per launch */
int flacb200_last_timings(flacb200_engine* e, flacb200_timings* t);

/* Deterministic synthetic PCM (mixed sinusoids, chirp, noise, silence gap, square burst),
 * generated on the device: bit-identical to tests/flacb200_testutil.py::synth_pcm.
 * Writes n_pcm_frames * channels samples of track `track` as packed little-endian bytes. */
int flacb200_synth_pcm(flacb200_engine* e, void* pcm_device, uint64_t first_track, uint64_t n_tracks,
                       uint64_t n_pcm_frames, uint32_t channels, uint32_t sample_rate, uint32_t bits_per_sample,
                       uint64_t seed);

/* Parity tooling: evaluates the engine's device-side log (fn = 0) / log2 (fn = 1) -- the restatement of glibc's
 * functions that the LPC order estimate and coefficient shift use (src/encode.rs:3674, :3360) -- over n doubles
 * (host pointers), so that a test can compare them bit for bit with the C library. */
int flacb200_debug_libm(flacb200_engine* e, int fn, const double* in, double* out, size_t n);

/* pinned host memory for the end-to-end path */
void* flacb200_host_alloc(size_t bytes);
void flacb200_host_free(void* p);
void* flacb200_device_alloc(flacb200_engine* e, size_t bytes);
void flacb200_device_free(flacb200_engine* e, void* p);
int flacb200_memcpy(flacb200_engine* e, void* dst, const void* src, size_t bytes, int kind /* 1 H2D, 2 D2H, 3 D2D */);
int flacb200_synchronize(flacb200_engine* e);

const char* flacb200_strerror(int code);
const char* flacb200_version(void);

#ifdef __cplusplus
}
#endif
#endif
