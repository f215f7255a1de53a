// tiles.cuh -- PCM intake shared by the register-tiled encode kernels (k_lpc2, k_analyze, k_pack2, k_pack3):
// Frame::fill_from_buf / fill_from_samples / fill_from_channels (src/audio.rs:149-225) straight from the caller's
// buffer into 16-sample register tiles, with 128-bit loads where the layout allows.
#pragma once
#include "common.cuh"

namespace flacb200 {

// ------------------------------------------------------------------------------------------------
// PCM load: Frame::fill_from_buf / fill_from_samples / fill_from_channels  (src/audio.rs:149-225)
// ------------------------------------------------------------------------------------------------
__device__ inline int32_t load_pcm_sample(const uint8_t* __restrict__ pcm, const EncCfg& c, unsigned long long idx, uint32_t ch)
{
    switch (c.pcm_kind) {
    case 2: return reinterpret_cast<const int32_t*>(pcm)[idx * c.channels + ch];
    case 3: return reinterpret_cast<const int32_t*>(pcm)[(unsigned long long)ch * c.planar_stride + idx];
    default: break;
    }
    const uint8_t* p = pcm + (idx * c.channels + ch) * c.bytes_per_sample;
    uint32_t v;
    switch (c.bytes_per_sample) {
    case 1: return (int32_t)(int8_t)p[0];
    case 2: {
        uint32_t raw = *reinterpret_cast<const uint16_t*>(p);
        if (c.pcm_kind == 1) raw = ((raw & 0xff) << 8) | (raw >> 8);
        return (int32_t)(int16_t)raw;
    }
    case 3:
        v = c.pcm_kind == 1 ? ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2]
                            : ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
        return (int32_t)(v << 8) >> 8;
    default:
        v = c.pcm_kind == 1 ? ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]
                            : ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
        return (int32_t)v;
    }
}

constexpr int AN_THREADS = 256;
constexpr int AN_SPT = 16;                      // samples per thread
constexpr int AN_TILE = AN_THREADS * AN_SPT;    // 4096: largest block of the fast path
constexpr int AN_PAD = 32;                      // zero samples in front of every plane (history of the first thread)
constexpr int AN_STRIDE = AN_TILE + AN_PAD;

// ---- packed PCM -> 16 consecutive inter-channel samples of C channels (Frame::fill_from_buf, src/audio.rs:149-187) ----
template <int C, int B>
__device__ inline void load16(const uint8_t* __restrict__ p, bool big_endian, int32_t* __restrict__ v)
{
    constexpr int NB = 16 * C * B, NW = NB / 4;
    uint32_t w[NW + 1];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < NW / 4; k++) {
            const uint4 t = reinterpret_cast<const uint4*>(p)[k];
            w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
        }
    } else if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = reinterpret_cast<const uint32_t*>(p)[k];
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++)
            w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
    }
    w[NW] = 0;
#pragma unroll
    for (int k = 0; k < 16 * C; k++) {
        const int bo = k * B;
        const uint32_t raw = __funnelshift_r(w[bo >> 2], w[(bo >> 2) + 1], (bo & 3) * 8);   // sample bytes in memory order
        if (big_endian) v[k] = (int32_t)__byte_perm(raw, 0, 0x0123) >> (32 - 8 * B);
        else v[k] = (int32_t)(raw << (32 - 8 * B)) >> (32 - 8 * B);
    }
}

template <int C>
__device__ inline void load16_any(const uint8_t* __restrict__ p, uint32_t bytes_per_sample, bool big_endian, int32_t* __restrict__ v)
{
    switch (bytes_per_sample) {
    case 1: load16<C, 1>(p, big_endian, v); break;
    case 2: load16<C, 2>(p, big_endian, v); break;
    case 3: load16<C, 3>(p, big_endian, v); break;
    default: load16<C, 4>(p, big_endian, v); break;
    }
}

// Fills the thread's 16 samples of up to two channels (ch0, ch0 + 1 when C == 2) of the block; samples past n are 0.
template <int C>
__device__ inline void load_thread_samples(const EncCfg& cfg, const FrameDesc& d, const uint8_t* __restrict__ pcm, uint32_t i0, uint32_t ch0,
                                           int32_t* __restrict__ a, int32_t* __restrict__ b)
{
    const bool full = i0 + AN_SPT <= d.n;
    if (full && cfg.pcm_kind != 3 && cfg.channels == (uint32_t)C) {
        int32_t v[16 * C];
        load16_any<C>(pcm + (d.pcm_off + i0) * (unsigned long long)(C * cfg.bytes_per_sample), cfg.bytes_per_sample, cfg.pcm_kind == 1, v);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            a[e] = v[e * C];
            if (C == 2) b[e] = v[e * C + 1];
        }
        return;
    }
    if (full && cfg.pcm_kind == 3) {
        load16<1, 4>(pcm + ((unsigned long long)ch0 * cfg.planar_stride + d.pcm_off + i0) * 4ull, false, a);
        if (C == 2) load16<1, 4>(pcm + ((unsigned long long)(ch0 + 1) * cfg.planar_stride + d.pcm_off + i0) * 4ull, false, b);
        return;
    }
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const bool ok = i0 + e < d.n;
        a[e] = ok ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0) : 0;
        if (C == 2) b[e] = ok ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, ch0 + 1) : 0;
    }
}

__device__ inline void store16(int32_t* __restrict__ dst, const int32_t* __restrict__ v)
{
#pragma unroll
    for (int k = 0; k < 4; k++) reinterpret_cast<int4*>(dst)[k] = make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// the candidate's 16 samples starting at i0 (zeros past the block end), before the wasted-bit shift
template <bool STEREO>
__device__ inline void aw_load_tile(const EncCfg& cfg, const FrameDesc& d, const uint8_t* __restrict__ pcm, uint32_t slot, uint32_t i0, int32_t* x)
{
    if (STEREO) {
        int32_t a[16], b[16];
        load_thread_samples<2>(cfg, d, pcm, i0, 0, a, b);
        switch (slot) {   // uniform across the warp
        case 0:
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = a[e];
            break;
        case 1:
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = b[e];
            break;
        case 2:
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = (a[e] + b[e]) >> 1;   // :2721
            break;
        default:
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = a[e] - b[e];          // :2734
            break;
        }
    } else if (cfg.channels == 1) {
        int32_t b[1];
        load_thread_samples<1>(cfg, d, pcm, i0, 0, x, b);
    } else {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = i0 + e < d.n ? load_pcm_sample(pcm, cfg, d.pcm_off + i0 + e, slot) : 0;
    }
}

}   // namespace flacb200
