// wav2flac / flac2wav -- the reference's example front ends (examples/wav2flac.rs:37-130, examples/flac2wav.rs:40-97)
// on top of flacb200.hpp: RIFF/WAVE PCM <-> .flac through the GPU engine.
//   flacb200_wav2flac encode [--best|--fast] in.wav out.flac
//   flacb200_wav2flac decode in.flac out.wav
#include <cstring>
#include <fstream>
#include <iostream>

#include "flacb200.hpp"

namespace {

uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
void put32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
void put16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

int encode(const flacb200::Options& opt, const char* in, const char* out)
{
    std::ifstream f(in, std::ios::binary);
    if (!f) { std::cerr << "cannot open " << in << "\n"; return 2; }
    uint8_t hdr[12];
    f.read((char*)hdr, 12);
    if (!f || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) { std::cerr << "not a RIFF/WAVE file\n"; return 2; }
    uint32_t rate = 0, bps = 0, channels = 0, data_len = 0;
    for (;;) {   // chunk walk: "fmt " then "data"
        uint8_t ch[8];
        f.read((char*)ch, 8);
        if (!f) { std::cerr << "no data chunk\n"; return 2; }
        const uint32_t len = le32(ch + 4);
        if (!memcmp(ch, "fmt ", 4)) {
            if (len < 16 || len > 4096) { std::cerr << "bad fmt chunk\n"; return 2; }   // WAVEFORMAT is 16 bytes, EXTENSIBLE 40
            std::vector<uint8_t> fmt(len);
            f.read((char*)fmt.data(), len);
            if ((uint32_t)f.gcount() != len) { std::cerr << "short fmt chunk\n"; return 2; }
            const uint16_t tag = le16(fmt.data());
            if (tag != 1 && tag != 0xFFFE) { std::cerr << "not PCM\n"; return 2; }
            channels = le16(fmt.data() + 2);
            rate = le32(fmt.data() + 4);
            bps = le16(fmt.data() + 14);
            if (len & 1) f.ignore(1);
        } else if (!memcmp(ch, "data", 4)) {
            data_len = len;
            break;
        } else {
            f.ignore(len + (len & 1));
        }
    }
    if (channels < 1 || channels > 8 || bps < 1 || bps > 32) { std::cerr << "unsupported channel count / sample width\n"; return 2; }
    flacb200::Engine eng(0);
    std::ofstream o(out, std::ios::binary | std::ios::trunc);
    flacb200::FlacByteWriter w(o, eng, opt, rate, bps, (uint8_t)channels, data_len);
    std::vector<uint8_t> buf(1 << 22);
    uint64_t left = data_len;
    const bool unsigned8 = bps <= 8;   // WAVE stores 8-bit samples unsigned
    while (left) {
        const size_t n = (size_t)std::min<uint64_t>(left, buf.size());
        f.read((char*)buf.data(), (std::streamsize)n);
        if ((size_t)f.gcount() != n) { std::cerr << "short read\n"; return 2; }
        if (unsigned8) for (size_t i = 0; i < n; i++) buf[i] ^= 0x80;
        w.write(buf.data(), n);
        left -= n;
    }
    w.finalize();
    return 0;
}

int decode(const char* in, const char* out)
{
    std::ifstream f(in, std::ios::binary);
    if (!f) { std::cerr << "cannot open " << in << "\n"; return 2; }
    flacb200::Engine eng(0);
    flacb200::FlacReader r(f, eng);
    const uint32_t B = (r.bits_per_sample() + 7) / 8;
    std::ofstream o(out, std::ios::binary | std::ios::trunc);
    uint8_t hdr[44] = {0};
    o.write((char*)hdr, 44);
    std::vector<uint8_t> buf(1 << 22);
    uint64_t total = 0;
    for (;;) {
        const size_t n = r.read_bytes(buf.data(), buf.size() - buf.size() % B);
        if (!n) break;
        if (B == 1) for (size_t i = 0; i < n; i++) buf[i] ^= 0x80;
        o.write((char*)buf.data(), (std::streamsize)n);
        total += n;
    }
    memcpy(hdr, "RIFF", 4);
    put32(hdr + 4, (uint32_t)(36 + total));
    memcpy(hdr + 8, "WAVEfmt ", 8);
    put32(hdr + 16, 16);
    put16(hdr + 20, 1);
    put16(hdr + 22, r.channel_count());
    put32(hdr + 24, r.sample_rate());
    put32(hdr + 28, r.sample_rate() * r.channel_count() * B);
    put16(hdr + 32, (uint16_t)(r.channel_count() * B));
    put16(hdr + 34, (uint16_t)r.bits_per_sample());
    memcpy(hdr + 36, "data", 4);
    put32(hdr + 40, (uint32_t)total);
    o.seekp(0);
    o.write((char*)hdr, 44);
    return 0;
}

}   // namespace

int main(int argc, char** argv)
{
    try {
        if (argc >= 4 && !strcmp(argv[1], "encode")) {
            flacb200::Options opt;
            int a = 2;
            if (!strcmp(argv[a], "--best")) { opt = flacb200::Options::best(); a++; }
            else if (!strcmp(argv[a], "--fast")) { opt = flacb200::Options::fast(); a++; }
            if (argc - a != 2) { std::cerr << "usage: encode [--best|--fast] in.wav out.flac\n"; return 1; }
            return encode(opt, argv[a], argv[a + 1]);
        }
        if (argc == 4 && !strcmp(argv[1], "decode")) return decode(argv[2], argv[3]);
        std::cerr << "usage: flacb200_wav2flac encode [--best|--fast] in.wav out.flac | decode in.flac out.wav\n";
        return 1;
    } catch (const flacb200::Error& e) {
        std::cerr << "error " << e.code << ": " << e.what() << "\n";
        return 3;
    }
}
