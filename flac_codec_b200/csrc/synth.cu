// synth.cu -- deterministic integer-only synthetic PCM generated on the device:
// mixed sinusoids + chirp + noise, a digital-silence gap and a full-scale square burst per track.
// Bit-identical to tests/flacb200_testutil.py::synth_pcm (the numpy statement the tests compare with).
#include "common.cuh"

namespace flacb200 {

__device__ inline unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthParams {
    unsigned long long delta[8][3];
    unsigned long long d0, dd;
    unsigned long long n_pcm_frames, first_track, n_tracks, seed;
    uint32_t channels, sample_rate, bps, bytes_per_sample;
};

__global__ void __launch_bounds__(256) k_synth(SynthParams p, const int32_t* __restrict__ lut, uint8_t* __restrict__ pcm)
{
    const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = p.n_tracks * p.n_pcm_frames;
    if (gid >= total) return;
    const unsigned long long t = gid / p.n_pcm_frames, idx = gid % p.n_pcm_frames;
    const unsigned long long track = p.first_track + t;
    const long long full = 1ll << (p.bps - 1);
    const int amp[3] = {307, 204, 51};   // int(0.30 * 1024), int(0.20 * 1024), int(0.05 * 1024)
    long long base = 0;
    uint8_t* dst = pcm + gid * p.channels * p.bytes_per_sample;
    for (uint32_t c = 0; c < p.channels; c++) {
        long long acc = 0;
        for (int k = 0; k < 3; k++) {
            const unsigned long long phase0 = ((track * 977ull + c * 131ull) << 20) & 0xFFFFFFFFull;
            const unsigned long long ph = (phase0 + idx * p.delta[c][k]) & 0xFFFFFFFFull;
            acc += ((long long)lut[ph >> 20] * amp[k]) >> 10;
        }
        const unsigned long long ph = (idx * p.d0 + ((idx * idx) >> 1) * p.dd) & 0xFFFFFFFFull;
        acc += ((long long)lut[ph >> 20] * 154) >> 10;
        long long sig = p.bps <= 31 ? (acc >> (31 - p.bps)) : (acc << (p.bps - 31));
        if (c == 1) sig = (base * 819) >> 10;
        if (c == 0) base = sig;
        const unsigned long long cid = track * 8 + c + 1;
        const unsigned long long r = splitmix64(p.seed * 0x9E3779B97F4A7C15ull * cid + idx);
        sig += (long long)(r >> 56) - 128;
        if (idx >= p.sample_rate / 2 && idx < p.sample_rate) sig = 0;
        const unsigned long long b0 = (unsigned long long)p.sample_rate * 3 / 2;
        if (idx >= b0 && idx < b0 + 4096) sig = (((idx >> 5) & 1) == 0) ? full - 1 : -full;
        sig = sig < -full ? -full : (sig > full - 1 ? full - 1 : sig);
        const uint32_t v = (uint32_t)(int32_t)sig;
        for (uint32_t b = 0; b < p.bytes_per_sample; b++) dst[c * p.bytes_per_sample + b] = (uint8_t)(v >> (8 * b));
    }
}

cudaError_t launch_synth(uint8_t* pcm, unsigned long long first_track, unsigned long long n_tracks, unsigned long long n_pcm_frames,
                         uint32_t channels, uint32_t sample_rate, uint32_t bps, unsigned long long seed, const int32_t* lut, cudaStream_t st)
{
    SynthParams p{};
    for (uint32_t c = 0; c < channels; c++) {
        const double f[3] = {220.0, 440.0 * (1.0 + (double)c / 16.0), 3520.0};
        for (int k = 0; k < 3; k++) p.delta[c][k] = (unsigned long long)(f[k] / (double)sample_rate * 4294967296.0);
    }
    p.d0 = (unsigned long long)(100.0 / (double)sample_rate * 4294967296.0);
    p.dd = (unsigned long long)((8000.0 - 100.0) / (double)sample_rate * 4294967296.0 / 4194304.0);
    p.n_pcm_frames = n_pcm_frames;
    p.first_track = first_track;
    p.n_tracks = n_tracks;
    p.seed = seed;
    p.channels = channels;
    p.sample_rate = sample_rate;
    p.bps = bps;
    p.bytes_per_sample = (bps + 7) / 8;
    const unsigned long long total = n_tracks * n_pcm_frames;
    if (total == 0) return cudaSuccess;
    count_launch(), k_synth<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p, lut, pcm);
    return cudaGetLastError();
}

}   // namespace flacb200
