// decode.cuh -- structures shared by the decode kernels (decode_kernels.cu) and the host engine (engine.cu).
#pragma once
#include "common.cuh"

namespace flacb200 {

struct DecCfg {
    uint32_t channels, bps, sample_rate, subset, max_block_size;
    uint32_t nslots;    // scratch planes per frame: channels (+1 "high word" plane for a 33-bit side channel)
    uint32_t bstride;   // samples per scratch plane (multiple of 4)
    uint32_t pcm_kind, bytes_per_sample;
    uint32_t nseg;
    unsigned long long planar_stride;
    unsigned long long nbytes;       // size of the frames buffer
    unsigned long long out_samples;  // capacity of the PCM output in inter-channel samples
};

// Scratch planes (decoded subframes before stereo restoration).  The 32 candidates of a bundle (candidate index / 32)
// are interleaved in units of four samples: a warp whose lanes walk 32 frames (k_parse) or 32 subframes (k_restore) in
// lockstep then reads and writes 512 contiguous bytes per step instead of 16 bytes in each of 32 planes.
// Sample s of plane (c, ch) lives at int32 index plane_base(cfg, c, ch) + plane_off(s); groups of four samples
// (s % 4 == 0) are 16-byte aligned.
__host__ __device__ inline size_t plane_base(const DecCfg& cfg, uint32_t c, uint32_t ch)
{
    return (((size_t)(c >> 5) * cfg.nslots + ch) * (cfg.bstride >> 2) * 32 + (c & 31)) * 4;
}
__host__ __device__ inline uint32_t plane_off(uint32_t s) { return (s >> 2) * 128 + (s & 3); }

struct DecSeg {
    unsigned long long byte_off, byte_end, pcm_off, n_pcm;
};

struct FrameCand {
    unsigned long long off;
    uint32_t block_size;
    uint32_t seg;
    uint8_t hdr_len, assignment, pad0, pad1;
    uint32_t pad2;
};

struct DecRec {
    unsigned long long end;   // one past the frame's CRC-16
    uint32_t err;             // 0 or the flac_codec::Error ordinal
    uint32_t wide;            // 1: the side channel is 33 bits wide, its high words are in the extra plane
};

// per subframe, written by k_parse and read by k_restore (decode_parse.cu)
struct SubRec {
    uint8_t kind;      // 0 constant, 1 verbatim, 2 fixed, 3 lpc; 0xFF: frame left to k_decode (33-bit side channel)
    uint8_t order, shift, wasted;
    int16_t coef[32];
};

struct ChainState {
    unsigned long long expect_off;    // a chain continues into the next group at this byte offset ...
    unsigned long long seg_samples;   // ... with this many samples of its segment already decoded
    unsigned long long frames_total, samples_total;
    unsigned long long err_frame;
    uint32_t expect_seg, active, err, pad;
};


}   // namespace flacb200
