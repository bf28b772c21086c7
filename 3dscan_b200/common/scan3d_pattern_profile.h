// scan3d_pattern_profile.h -- one row (or column) of a projector pattern, exactly as the
// reference's generator writes it (1/pattern_generator.cpp:291-397 fringe, :56-197 Gray,
// :490-507 inverse Gray).  A pattern is constant along its stripes, so this 1-D profile IS the
// pattern; the host library (scan3d_synth_pattern_row) and the device generator
// (scan3d_generate_patterns) both expand it.  Pinned against the reference's Generated_patterns
// images by tests/test_abi.py (tests/golden/pattern_kat.npz).
//
// Host code on purpose: the fringe value goes through libm's cosf and a float->uchar truncation,
// which a device cosf would not reproduce bit for bit; a profile is a few KB per pattern.
#ifndef SCAN3D_PATTERN_PROFILE_H
#define SCAN3D_PATTERN_PROFILE_H
#include <math.h>
#include <stdint.h>

namespace s3d_profile {

constexpr double kPiTrue = 3.14159265358979323846;

// B = binary digits of code_number, index 0 = MSB of an M-bit word; G0 = B0, Gi = B(i-1)^B(i)
inline int gray_bit(int code_number, int i, int M)
{
    const int b_i = (code_number >> (M - 1 - i)) & 1;
    const int b_prev = i == 0 ? 0 : (code_number >> (M - i)) & 1;
    return b_i ^ b_prev;
}

// kind 0: fringe pattern k of an N-step set (N = n_or_m); kind 1 / 2: Gray / inverse Gray bit plane
// k of an M-bit code (M = n_or_m).  fw = fringe width in projector pixels.
inline void pattern_profile(int kind, int n_or_m, int fw, int k, int length, uint8_t* out)
{
    if (kind == 0) {
        const int N = n_or_m;
        for (int c = 0; c < length; c++) {
            const float q = (float)c / (float)fw;
            double arg;
            // the reference's expressions, token for token (Pi is the textual macro 22.0/7.0)
            if (N == 3) arg = q * 2.0 * 22.0 / 7.0 - 22.0 / 7.0 - ((22.0 / 7.0) / 2.0) + (22.0 / 7.0 / 2.0) * (float)k;
            else if (N == 4) arg = q * (2.0 * 22.0 / 7.0) - 22.0 / 7.0 + (22.0 / 7.0 / 2.0) * (float)k;
            else if (N == 5) arg = q * (2.0 * 22.0 / 7.0) - 22.0 / 7.0 - 2.0 * ((22.0 / 7.0) / 2) + ((22.0 / 7.0) / 2) * (float)k;
            else arg = q * (2.0 * 22.0 / 7.0) - 22.0 / 7.0 + 2.0 * kPiTrue * k / N;
            const float t = 127.0f + 128.0f * cosf((float)arg);
            out[c] = (unsigned char)t;
        }
        return;
    }
    const int M = n_or_m;
    for (int c = 0; c < length; c++) {
        const int g = gray_bit(c / fw, k, M) * 255;
        out[c] = (uint8_t)(kind == 1 ? g : 255 - g);
    }
}

}  // namespace s3d_profile
#endif
