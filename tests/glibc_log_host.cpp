// Host build of flac_codec_b200/csrc/glibc_log.cuh for tests/test_glibc_log_port.py (test infrastructure):
// the same operation sequence the device runs, with the C library's exact fma() and no contraction (-ffp-contract=off).
#include <cstddef>
#include "../flac_codec_b200/csrc/glibc_log.cuh"

extern "C" void port_libm(int fn, const double* in, double* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = fn == 0 ? flacb200::glibc_log(in[i]) : flacb200::glibc_log2(in[i]);
}
