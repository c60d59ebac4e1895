// K1b: SwiGLU gating fused with the MX quantization of its result.
//
// In the reference an MLP block computes  down_proj(act_fn(gate_proj(x)) * up_proj(x))  (torchmx/layers/mx_llama_attention.py:
// 19-59, with transformers' LlamaMLP / Qwen2MLP forward), and MXInferenceLinear.forward quantizes the product on entry
// (torchmx/layers/mx_linear.py:63-66): three launches -- silu, mul, K1 -- that write and re-read a [tokens, intermediate] bf16
// tensor twice.  This kernel reads gate and up once and writes the codes + scales down_proj consumes:
//     h = bf16( g / (1 + exp(-g)) )        (aten silu: fp32 arithmetic on the bf16 input, one rounding; evaluated by the five-
//                                           instruction sequence of mxq_silu.cuh, checked against the plain formula for all
//                                           65536 bf16 inputs on the hardware: tools/silu_check.cu)
//     y = bf16( h * u )                    (aten mul)
//     codes, scales = quantize_mx(y)       (K1's arithmetic, mxq_quant_core.cuh)
// bit-identical to the three-launch chain.  Algorithmic traffic 2 + 2 + 1 + 1/32 B per element.  gate / up may be column
// slices of one stacked projection output (row strides given), so a single gate+up GEMM can feed it without a copy.
#include "mxq_quant_core.cuh"
#include "mxq_silu.cuh"

#include <cmath>

namespace mxq {

template <int ELEM>
__global__ void __launch_bounds__(256) silu_mul_quantize_kernel(const uint16_t* __restrict__ gate, const uint16_t* __restrict__ up, int64_t n_blocks,
                                                                 int blocks_per_row, int64_t ld_gate, int64_t ld_up, uint8_t* __restrict__ codes,
                                                                 uint8_t* __restrict__ scales, uint32_t flags) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int NO = (ELEM == MXQ_ELEM_E2M1) ? 4 : 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_blocks; c += stride) {
        const int64_t row = c / blocks_per_row;
        const int64_t col = (c - row * blocks_per_row) * 32;
        const uint8_t* pg = reinterpret_cast<const uint8_t*>(gate + row * ld_gate + col);
        const uint8_t* pu = reinterpret_cast<const uint8_t*>(up + row * ld_up + col);
        uint32_t g[16], u[16], w[16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const u32x8 a = ldg256_stream(pg + 32 * j), b = ldg256_stream(pu + 32 * j);
#pragma unroll
            for (int k = 0; k < 8; ++k) { g[8 * j + k] = a.v[k]; u[8 * j + k] = b.v[k]; }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float g0 = __uint_as_float(g[i] << 16), g1 = __uint_as_float(g[i] & 0xFFFF0000u);
            const uint32_t h = pack_bf16x2(silu_bf16_input(g0), silu_bf16_input(g1));
            w[i] = pack_bf16x2(__uint_as_float(h << 16) * __uint_as_float(u[i] << 16), __uint_as_float(h & 0xFFFF0000u) * __uint_as_float(u[i] & 0xFFFF0000u));
        }
        uint32_t out[NO];
        const int sc = quantize_block32<ELEM>(w, (flags & MXQ_FLAG_HW_EXACT) != 0, out);
        uint8_t* dst = codes + c * (NO * 4);
        if constexpr (NO == 4) stg128_stream(dst, make_uint4(out[0], out[1], out[2], out[3]));
        else {
            u32x8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = out[k];
            stg256_stream(dst, o);
        }
        scales[c] = (uint8_t)sc;
    }
}

cudaError_t launch_silu_mul_quantize(const void* gate, const void* up, int64_t rows, int64_t cols, int64_t ld_gate, int64_t ld_up, int elem,
                                     unsigned flags, void* codes, uint8_t* scales, int sm_count, cudaStream_t stream) {
    const int bpr = (int)(cols / 32);
    const int64_t n_blocks = rows * bpr;
    if (n_blocks == 0) return cudaSuccess;
    const int64_t want = (n_blocks + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 64;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const uint16_t* g = (const uint16_t*)gate;
    const uint16_t* u = (const uint16_t*)up;
    uint8_t* c8 = (uint8_t*)codes;
    switch (elem) {
    case MXQ_ELEM_E4M3: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_E4M3>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    case MXQ_ELEM_E3M2: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_E3M2>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    case MXQ_ELEM_E2M3: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_E2M3>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    case MXQ_ELEM_E2M1: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_E2M1>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    case MXQ_ELEM_INT8: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_INT8>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    case MXQ_ELEM_E5M2: launch_pdl(silu_mul_quantize_kernel<MXQ_ELEM_E5M2>, dim3(grid), dim3(256), 0, stream, g, u, n_blocks, bpr, ld_gate, ld_up, c8, scales, (uint32_t)flags); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mxq
