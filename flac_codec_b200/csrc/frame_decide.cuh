// frame_decide.cuh -- what encode_frame decides once every subframe candidate has its exact size: the channel assignment
// (src/encode.rs:2747-2819), the frame header with its CRC-8 (src/stream.rs:242-276, :1266-1326), the bit offset of every
// subframe and the frame size.  Shared by k_decide (thread per frame) and k_frame4 (inside the fused frame kernel).
#pragma once
#include "common.cuh"

namespace flacb200 {

__device__ inline int frame_number_bytes(unsigned long long v, uint8_t* out)   // src/stream.rs:1266-1326
{
    if (v <= 0x7F) { out[0] = (uint8_t)v; return 1; }
    int bytes = v <= 0x7FF ? 2 : v <= 0xFFFF ? 3 : v <= 0x1FFFFF ? 4 : v <= 0x3FFFFFF ? 5 : v <= 0x7FFFFFFFull ? 6 : 7;
    const uint32_t lead = 7 - bytes;
    out[0] = (uint8_t)((0xFFu << (8 - bytes)) | (lead ? (uint32_t)((v >> (6 * (bytes - 1))) & ((1u << lead) - 1u)) : 0u));
    for (int i = 1; i < bytes; i++) out[i] = (uint8_t)(0x80 | ((v >> (6 * (bytes - 1 - i))) & 0x3F));
    return bytes;
}

// c: the frame's candidate records (cfg.nslots of them); abssum4: the four abs sums of the fast channel correlation
__device__ inline void decide_frame(const EncCfg& cfg, const FrameDesc& d, const CandRec* c, const unsigned long long* abssum4, FrameRec& fr)
{
    fr.err = 0;
    uint32_t assignment;
    if (cfg.mode == MODE_INDEPENDENT) {
        assignment = cfg.channels - 1;
        for (uint32_t k = 0; k < cfg.channels; k++) fr.slot[k] = (uint8_t)k;
        fr.nsub = (uint8_t)cfg.channels;
    } else {
        if (cfg.mode == MODE_EXH_MID_SIDE) {   // [Independent, LeftSide, SideRight, MidSide], first minimum (:2747-2768)
            const unsigned long long t[4] = {(unsigned long long)c[0].bits + c[1].bits, (unsigned long long)c[0].bits + c[3].bits,
                                             (unsigned long long)c[3].bits + c[1].bits, (unsigned long long)c[2].bits + c[3].bits};
            int b = 0;
            for (int k = 1; k < 4; k++) if (t[k] < t[b]) b = k;
            assignment = b == 0 ? 1u : (b == 1 ? 8u : (b == 2 ? 9u : 10u));
        } else if (cfg.mode == MODE_EXH_SIDE) {   // [Independent, LeftSide, SideRight] (:2803-2819)
            const unsigned long long t[3] = {(unsigned long long)c[0].bits + c[1].bits, (unsigned long long)c[0].bits + c[3].bits,
                                             (unsigned long long)c[3].bits + c[1].bits};
            int b = 0;
            for (int k = 1; k < 3; k++) if (t[k] < t[b]) b = k;
            assignment = b == 0 ? 1u : (b == 1 ? 8u : 9u);
        } else {
            assignment = fast_assignment(abssum4, cfg.mode == MODE_FAST_MID_SIDE);
        }
        assignment_slots(assignment, &fr.slot[0], &fr.slot[1]);
        fr.nsub = 2;
    }
    fr.assignment = (uint8_t)assignment;
    // ---- frame header (src/stream.rs:242-276) ----
    uint8_t* h = fr.hdr;
    int hl = 0;
    const uint32_t n = d.n;
    uint32_t bsc, bs_extra = 0;
    switch (n) {   // src/stream.rs:537-560
    case 192: bsc = 1; break;   case 576: bsc = 2; break;    case 1152: bsc = 3; break;  case 2304: bsc = 4; break;
    case 4608: bsc = 5; break;  case 256: bsc = 8; break;    case 512: bsc = 9; break;   case 1024: bsc = 10; break;
    case 2048: bsc = 11; break; case 4096: bsc = 12; break;  case 8192: bsc = 13; break; case 16384: bsc = 14; break;
    case 32768: bsc = 15; break;
    default:
        if (n <= 256) { bsc = 6; bs_extra = 8; } else { bsc = 7; bs_extra = 16; }
    }
    uint32_t src, rate_kind = 0;
    const uint32_t rate = cfg.sample_rate;
    switch (rate) {   // src/stream.rs:779-802
    case 88200: src = 1; break;  case 176400: src = 2; break; case 192000: src = 3; break; case 8000: src = 4; break;
    case 16000: src = 5; break;  case 22050: src = 6; break;  case 24000: src = 7; break;  case 32000: src = 8; break;
    case 44100: src = 9; break;  case 48000: src = 10; break; case 96000: src = 11; break;
    default:
        if (rate % 1000 == 0 && rate / 1000 < 255) { src = 12; rate_kind = 1; }
        else if (rate % 10 == 0 && rate / 10 < 65535) { src = 14; rate_kind = 3; }
        else if (rate < 65535) { src = 13; rate_kind = 2; }
        else src = 0;
    }
    uint32_t bpc;
    switch (cfg.bps) {   // src/stream.rs:1136-1149
    case 8: bpc = 1; break; case 12: bpc = 2; break; case 16: bpc = 4; break; case 20: bpc = 5; break;
    case 24: bpc = 6; break; case 32: bpc = 7; break; default: bpc = 0;
    }
    h[hl++] = 0xFF;
    h[hl++] = 0xF8;
    h[hl++] = (uint8_t)((bsc << 4) | src);
    h[hl++] = (uint8_t)((assignment << 4) | (bpc << 1));
    hl += frame_number_bytes(d.fnum, h + hl);
    if (bs_extra == 8) h[hl++] = (uint8_t)(n - 1);
    else if (bs_extra == 16) { h[hl++] = (uint8_t)((n - 1) >> 8); h[hl++] = (uint8_t)(n - 1); }
    if (rate_kind == 1) h[hl++] = (uint8_t)(rate / 1000);
    else if (rate_kind == 2) { h[hl++] = (uint8_t)(rate >> 8); h[hl++] = (uint8_t)rate; }
    else if (rate_kind == 3) { h[hl++] = (uint8_t)((rate / 10) >> 8); h[hl++] = (uint8_t)(rate / 10); }
    uint8_t crc = 0;
    for (int i = 0; i < hl; i++) crc = crc8_update(crc, h[i]);
    h[hl++] = crc;
    fr.hdr_len = (uint8_t)hl;
    unsigned long long bits = (unsigned long long)hl * 8;
    for (uint32_t k = 0; k < fr.nsub; k++) {
        fr.sub_bit[k] = (uint32_t)bits;
        const CandRec& cr = c[fr.slot[k]];
        if (cr.type == 0xFF) fr.err = 1;
        bits += cr.bits;
    }
    for (uint32_t k = fr.nsub; k < MAX_CH; k++) { fr.sub_bit[k] = 0; fr.slot[k] = 0; }
    fr.frame_bytes = (uint32_t)((bits + 7) / 8) + 2;   // byte align + CRC-16 (:2408-2409)
    fr.out_off = 0;
    fr.pad = 0;
}

}   // namespace flacb200
