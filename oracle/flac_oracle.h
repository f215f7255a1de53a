/*
 * flac_oracle.h -- CPU restatement of tuffy/flac-codec's frame encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement of
 * the reference crate's algorithm (src/encode.rs, src/decode.rs, src/stream.rs,
 * src/crc.rs, src/audio.rs of flac-codec 1.3.2).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call it.  The product
 * (flac_codec_b200/, libflacb200.so) never links or loads it.
 *
 * Parity pin: the reference itself cannot be built here (no Rust toolchain), so this
 * restatement is pinned by the reference's own known-answer tests and fixtures:
 *   encoder maths   src/encode.rs:3216-3272, 3404-3476, 3503-3527, 3591-3653, 3704-3745
 *   decoder maths   src/decode.rs:1754-1798
 *   bit layout      doc-test byte strings in src/stream.rs (:107-128, :1645-1677, ...)
 *   whole streams   tests/data/sine.flac (MD5 831671b8...), all-frames.flac, cuesheet.flac
 * (see tests/test_oracle_kat.py).  Encoder OUTPUT BYTES are not pinned by any reference
 * test ("parity unpinned" for compressed bytes): fidelity there rests on the KATs above,
 * on line-by-line restatement, and on lossless round trips through the pinned decoder.
 */
#ifndef FLAC_ORACLE_H
#define FLAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: ordinal of the matching variant of flac_codec::Error (src/lib.rs:57-193), 1-based */
enum {
    FO_OK = 0,
    FO_ERR_IO = 1,
    FO_ERR_MISSING_FLAC_TAG = 3,
    FO_ERR_MISSING_STREAMINFO = 4,
    FO_ERR_INVALID_METADATA_BLOCK = 15,
    FO_ERR_SHORT_BLOCK = 21,
    FO_ERR_INVALID_SYNC_CODE = 23,
    FO_ERR_INVALID_BLOCK_SIZE = 24,
    FO_ERR_BLOCK_SIZE_MISMATCH = 25,
    FO_ERR_INVALID_SAMPLE_RATE = 26,
    FO_ERR_NON_SUBSET_SAMPLE_RATE = 27,
    FO_ERR_NON_SUBSET_BPS = 28,
    FO_ERR_SAMPLE_RATE_MISMATCH = 29,
    FO_ERR_EXCESSIVE_CHANNELS = 30,
    FO_ERR_INVALID_CHANNELS = 31,
    FO_ERR_CHANNELS_MISMATCH = 32,
    FO_ERR_INVALID_BPS = 33,
    FO_ERR_BPS_MISMATCH = 35,
    FO_ERR_INVALID_FRAME_NUMBER = 36,
    FO_ERR_EXCESSIVE_FRAME_NUMBER = 38,
    FO_ERR_CRC8_MISMATCH = 39,
    FO_ERR_CRC16_MISMATCH = 40,
    FO_ERR_INVALID_SUBFRAME_HEADER = 41,
    FO_ERR_INVALID_SUBFRAME_HEADER_TYPE = 42,
    FO_ERR_EXCESSIVE_WASTED_BITS = 43,
    FO_ERR_INVALID_CODING_METHOD = 45,
    FO_ERR_INVALID_PARTITION_ORDER = 46,
    FO_ERR_INVALID_FIXED_ORDER = 47,
    FO_ERR_INVALID_LPC_ORDER = 48,
    FO_ERR_INVALID_QLP_PRECISION = 49,
    FO_ERR_NEGATIVE_LPC_SHIFT = 50,
    FO_ERR_NO_BEST_LPC_ORDER = 51,
    FO_ERR_INSUFFICIENT_LPC_SAMPLES = 52,
    FO_ERR_ZERO_LP_COEFFICIENTS = 53,
    FO_ERR_LP_NEGATIVE_SHIFT = 54,
    FO_ERR_EXCESSIVE_TOTAL_SAMPLES = 57,
    FO_ERR_NO_SAMPLES = 58,
    FO_ERR_SAMPLE_COUNT_MISMATCH = 59,
    FO_ERR_RESIDUAL_OVERFLOW = 60,
    FO_ERR_SAMPLES_NOT_DIVISIBLE = 61,
};

/* Mirrors flac_codec::encode::Options (src/encode.rs:1363-1408) minus the metadata list. */
typedef struct fo_options {
    uint16_t block_size;          /* >= 16 */
    uint8_t max_lpc_order;        /* 0 = None, else 1..=32 */
    uint8_t max_partition_order;  /* 0..=15 (values > 6 overflow the reference's 64-entry ArrayVec) */
    uint8_t mid_side;
    uint8_t exhaustive_channel_correlation;
    uint8_t window_kind;          /* 0 = Rectangle, 1 = Hann, 2 = Tukey(tukey_p) */
    float tukey_p;
    uint8_t seektable_kind;       /* 0 = none, 1 = every n seconds, 2 = every n frames */
    uint32_t seektable_n;
    int32_t padding;              /* PADDING block body size, < 0 = no PADDING block */
} fo_options;

void fo_options_default(fo_options* o); /* Options::default() src/encode.rs:1376 */
void fo_options_fast(fo_options* o);    /* Options::fast()    src/encode.rs:1635 */
void fo_options_best(fo_options* o);    /* Options::best()    src/encode.rs:1649 */

/* What the encoder chose for one subframe; used to debug parity against the GPU path. */
typedef struct fo_subframe_info {
    int32_t type;            /* 0 CONSTANT, 1 VERBATIM, 2 FIXED, 3 LPC */
    int32_t order;
    int32_t wasted;
    int32_t bps;             /* effective bps of the subframe after wasted bits */
    int32_t precision;
    int32_t shift;
    int32_t coefs[32];
    int32_t coding_method;   /* 0 = 4-bit rice params, 1 = 5-bit */
    int32_t partition_order;
    uint8_t rice[64];        /* per partition: rice parameter, or escape width for escaped */
    uint8_t kind[64];        /* per partition: 0 standard, 1 escaped, 2 all-zero ("Constant") */
    uint64_t bits;           /* exact size of the subframe in bits */
} fo_subframe_info;

typedef struct fo_frame_info {
    int32_t channel_assignment; /* header code: 0..7 independent, 8 left/side, 9 side/right, 10 mid/side */
    int32_t channels;
    uint32_t frame_bytes;
    fo_subframe_info sub[8];
} fo_frame_info;

typedef struct fo_encoder fo_encoder; /* per-thread scratch ("EncodingCaches", src/encode.rs) */
fo_encoder* fo_encoder_new(void);
void fo_encoder_free(fo_encoder* e);

/*
 * encode_frame (src/encode.rs:2259) for one block of planar samples.
 *  subset != 0 selects FrameHeader::write_subset semantics (FlacStreamWriter, :1094) where
 *  the sample rate / bps must be expressible in the frame header.
 * Returns bytes written (>0) or a negative error code (-FO_ERR_*).
 */
int64_t fo_encode_frame(fo_encoder* e, const fo_options* opt, uint32_t sample_rate, uint32_t bps,
                        uint32_t channels, uint64_t frame_number, const int32_t* const* planar,
                        uint32_t nsamples, int subset, uint8_t* out, size_t out_cap,
                        fo_frame_info* info /* may be NULL */);

/*
 * Whole-stream encode as FlacSampleWriter / FlacByteWriter would produce it
 * (src/encode.rs:103-628, 1882-2110): "fLaC", STREAMINFO, [SEEKTABLE], [PADDING], frames.
 * interleaved: n_pcm_frames * channels samples.  total_known mirrors passing Some(total).
 * nthreads > 1 encodes frames concurrently (CPU-baseline use; output identical).
 * frame_sizes (may be NULL) receives each frame's byte size; *n_frames their count.
 */
int64_t fo_encode_stream(const fo_options* opt, uint32_t sample_rate, uint32_t bps, uint32_t channels,
                         const int32_t* interleaved, uint64_t n_pcm_frames, int total_known,
                         int nthreads, uint8_t* out, size_t out_cap, uint32_t* frame_sizes,
                         size_t frame_sizes_cap, uint64_t* n_frames);
/* frames only (no container), used by batch baselines and the frame-level parity tests */
int64_t fo_encode_frames_only(const fo_options* opt, uint32_t sample_rate, uint32_t bps, uint32_t channels,
                              const int32_t* interleaved, uint64_t n_pcm_frames, uint64_t first_frame_number,
                              int nthreads, uint8_t* out, size_t out_cap, uint32_t* frame_sizes,
                              size_t frame_sizes_cap, uint64_t* n_frames, fo_frame_info* infos /* may be NULL */);

/* STREAMINFO as the decoder sees it */
typedef struct fo_streaminfo {
    uint16_t min_block_size, max_block_size;
    uint32_t min_frame_size, max_frame_size;
    uint32_t sample_rate;
    uint8_t channels;
    uint8_t bps;
    uint64_t total_samples; /* 0 = unknown */
    uint8_t md5[16];
    uint64_t frames_start;  /* byte offset of the first frame in the file */
} fo_streaminfo;

typedef struct fo_frame_header {
    uint32_t block_size;
    uint32_t sample_rate;
    uint32_t bps;
    uint32_t channels;
    uint32_t channel_assignment;
    uint32_t blocking_strategy;
    uint64_t frame_number;
    uint32_t header_bytes;
} fo_frame_header;

/* metadata walk: "fLaC" + blocks (src/metadata/mod.rs:482-646); only STREAMINFO is interpreted */
int fo_read_streaminfo(const uint8_t* flac, size_t len, fo_streaminfo* si);

/*
 * Decoder::read_frame (src/decode.rs:1388) for one frame starting at data[0].
 * si == NULL selects read_subset semantics (FlacStreamReader).
 * planar_out receives channels x block_size samples, channel-major with stride block_size.
 * Returns bytes consumed (>0) or negative error code.
 */
int64_t fo_decode_frame(const uint8_t* data, size_t len, const fo_streaminfo* si, uint64_t remaining_or_0,
                        int32_t* planar_out, size_t planar_cap, fo_frame_header* hdr);

/*
 * Whole-stream decode (FlacSampleReader::read_to_end equivalent): interleaved i32 out.
 * md5_out (may be NULL) receives the MD5 of the decoded little-endian PCM bytes (verify(), :1282).
 * Returns number of interleaved samples written or negative error.
 */
int64_t fo_decode_stream(const uint8_t* flac, size_t len, int32_t* interleaved_out, size_t out_cap,
                         fo_streaminfo* si_out, uint8_t md5_out[16]);
/* multi-threaded frame decode given known frame offsets (CPU-baseline use) */
int64_t fo_decode_frames_mt(const uint8_t* frames, const uint64_t* offsets, uint64_t n_frames,
                            const fo_streaminfo* si, int nthreads, int32_t* interleaved_out, size_t out_cap);

/* ---- pieces exported so the reference's unit KATs can be replayed verbatim ---- */
int fo_autocorrelate(const double* windowed, uint32_t n, uint32_t max_lpc_order, double* out /* max+1 */);
/* returns number of orders; coeffs[(o-1)*32 + j], errors[o-1] */
int fo_lp_coefficients(const double* autoc, uint32_t n_autoc, double* coeffs, double* errors);
int fo_subframe_bits_by_order(uint32_t bps, uint32_t precision, uint32_t sample_count, const double* errors,
                              uint32_t n_orders, double* bits_out);
int fo_quantize(uint32_t order, const double* coeffs, uint32_t precision, int32_t* qcoefs, uint32_t* shift);
int fo_lpc_residuals(uint32_t order, uint32_t shift, const int32_t* qcoefs, const int32_t* samples, uint32_t n,
                     int32_t* residuals);
void fo_predict(const int64_t* coefficients, uint32_t order, uint32_t shift, int32_t* channel, uint32_t n);
void fo_window(const fo_options* opt, uint32_t n, double* out);
uint32_t fo_rice_parameter_f64(uint64_t sum, uint32_t samples); /* ceil(log2(sum/samples)) as the reference computes it */
int64_t fo_decode_stream_ex(const uint8_t* flac, size_t len, int32_t* out, size_t out_cap, fo_streaminfo* si_out, uint8_t md5_out[16],
                            uint64_t* frames_done, uint64_t* samples_done);
/* the C library's log (fn 0) / log2 (fn 1) over an array: what f64::ln / f64::log2 of the reference resolve to on Linux */
void fo_libm(int fn, const double* in, double* out, size_t n);
uint8_t fo_crc8(const uint8_t* p, size_t n);
uint16_t fo_crc16(const uint8_t* p, size_t n);
void fo_md5(const uint8_t* p, size_t n, uint8_t out[16]);
/* frame-number codec (src/stream.rs:1246-1326); returns byte count or negative error */
int fo_write_frame_number(uint64_t v, uint8_t out[7]);
int fo_read_frame_number(const uint8_t* p, size_t n, uint64_t* v);
/* PCM bytes <-> i32 (src/audio.rs:110-187, src/byteorder.rs) */
void fo_bytes_to_samples(const uint8_t* bytes, size_t n_samples, uint32_t bytes_per_sample, int big_endian, int32_t* out);
void fo_samples_to_bytes(const int32_t* samples, size_t n_samples, uint32_t bytes_per_sample, int big_endian, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
