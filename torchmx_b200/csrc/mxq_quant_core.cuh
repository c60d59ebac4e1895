// Per-thread core of the MX quantizer (K1), shared with the decode GEMM that quantizes its activation on the fly (K3c):
// block amax -> E8M0 shared exponent -> packed hardware conversions, including the NaN-block rules.  Keeping ONE copy of
// this arithmetic is what makes the fused path bit-identical to torchmx::quantize_mx by construction.
#pragma once
#include "mxq_common.cuh"

namespace mxq {

// max over the two u16 halves of |bits| for all words -> exponent field of the largest magnitude
__device__ __forceinline__ uint32_t umax16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }

template <int ELEM, int NW>
__device__ __forceinline__ void convert_words(const uint32_t (&w)[NW], int s, uint32_t (&out)[(ELEM == MXQ_ELEM_E2M1) ? NW / 4 : NW / 2]) {
    // w[i] holds elements 2i (low half) and 2i+1 (high half) as bf16 bit patterns
    const float inv = inv_scale_f32(s);
    if constexpr (ELEM == MXQ_ELEM_INT8) {
        constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: fma rounds x*inv to an integer (RNE) in the low mantissa bits
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
            uint32_t b[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float lo = fmaf(__uint_as_float(w[2 * i + j] << 16), inv, kMagic);
                float hi = fmaf(__uint_as_float(w[2 * i + j] & 0xFFFF0000u), inv, kMagic);
                lo = fminf(fmaxf(lo, kMagic - 127.0f), kMagic + 127.0f);
                hi = fminf(fmaxf(hi, kMagic - 127.0f), kMagic + 127.0f);
                b[2 * j] = __float_as_uint(lo);
                b[2 * j + 1] = __float_as_uint(hi);
            }
            out[i] = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
        }
    } else if constexpr (ELEM == MXQ_ELEM_E2M1) {
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ww = w[4 * i + j];
                const uint32_t byte = cvt_e2m1_byte(__uint_as_float(ww << 16) * inv, __uint_as_float(ww & 0xFFFF0000u) * inv);
                acc |= byte << (8 * j);
            }
            out[i] = acc;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
            const uint32_t w0 = w[2 * i], w1 = w[2 * i + 1];
            const uint32_t p0 = cvt_pair<ELEM>(__uint_as_float(w0 << 16) * inv, __uint_as_float(w0 & 0xFFFF0000u) * inv);
            const uint32_t p1 = cvt_pair<ELEM>(__uint_as_float(w1 << 16) * inv, __uint_as_float(w1 & 0xFFFF0000u) * inv);
            out[i] = p0 | (p1 << 16);
        }
    }
}

// NaN-scale block: all codes +0 (simulated, mx_quantization_utils.py:473), or the hw_exact quirk
template <int ELEM, int NW>
__device__ __forceinline__ void nanblock_words(const uint32_t (&w)[NW], bool hw_exact, uint32_t (&out)[(ELEM == MXQ_ELEM_E2M1) ? NW / 4 : NW / 2]) {
    constexpr int NO = (ELEM == MXQ_ELEM_E2M1) ? NW / 4 : NW / 2;
#pragma unroll
    for (int i = 0; i < NO; ++i) out[i] = 0;
    if constexpr (ELEM != MXQ_ELEM_INT8 && ELEM != MXQ_ELEM_E5M2) {
        if (hw_exact) {
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const uint32_t c0 = hw_exact_nanblock_code<ELEM>(w[i] & 0xFFFF);
                const uint32_t c1 = hw_exact_nanblock_code<ELEM>(w[i] >> 16);
                if constexpr (ELEM == MXQ_ELEM_E2M1) out[i / 4] |= ((c0 << 4) | c1) << (8 * (i % 4));
                else out[i / 2] |= (c0 | (c1 << 8)) << (16 * (i % 2));
            }
        }
    }
}


// one thread, one whole 32-element block held as 16 words of bf16 pairs: scale byte + 32 code bytes (16 for fp4)
template <int ELEM>
__device__ __forceinline__ int quantize_block32(const uint32_t (&w)[16], bool hw_exact, uint32_t (&out)[(ELEM == MXQ_ELEM_E2M1) ? 4 : 8]) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) m = umax16x2(m, w[i] & 0x7FFF7FFFu);
    m = max(m & 0xFFFFu, m >> 16);
    const int s = shared_exp_from_maxE<ELEM>((int)(m >> 7));
    if (s != 255) convert_words<ELEM, 16>(w, s, out);
    else nanblock_words<ELEM, 16>(w, hw_exact, out);
    return s;
}

}  // namespace mxq
