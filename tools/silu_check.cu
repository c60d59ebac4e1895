// Which cheap evaluation of silu equals aten's  bf16( g / (1 + expf(-g)) )  for EVERY bf16 input g on the real hardware?
// K1b (csrc/mxq_act_quant.cu) sees only bf16 gate values, so the question has 65536 cases and is answered exhaustively here:
// each variant is evaluated for all 65536 bit patterns and its bf16 rounding compared with the rounding of the plain formula
// (what the three-launch chain computes).  Build and run on a B200:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I torchmx_b200/csrc tools/silu_check.cu -o /tmp/silu_check && /tmp/silu_check
// Recorded result: profiles/r2_k1b_silu_check.json
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "mxq_silu.cuh"

__device__ __forceinline__ uint16_t bf16_rn(float f) {
    uint16_t r;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ float ex2a(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int NV = 7;
__device__ float variant(int v, float g) {
    switch (v) {
    case 0: {  // one multiply into ex2, approximate reciprocal
        const float d = 1.0f + ex2a(g * -1.4426950408889634f);
        return g * rcpa(d);
    }
    case 1: {  // + one Newton step on the reciprocal (guarded for d = inf)
        const float d = 1.0f + ex2a(g * -1.4426950408889634f);
        const float r0 = rcpa(d);
        const float r = __fmaf_rn(r0, __fmaf_rn(-d, r0, 1.0f), r0);
        return g * (d > 3.0e38f ? r0 : r);
    }
    case 2: {  // + residual correction of the quotient
        const float d = 1.0f + ex2a(g * -1.4426950408889634f);
        const float r0 = rcpa(d);
        const float r = __fmaf_rn(r0, __fmaf_rn(-d, r0, 1.0f), r0);
        const float q = g * r;
        const float qq = __fmaf_rn(r, __fmaf_rn(-d, q, g), q);
        return d > 1.0e37f ? g * r0 : qq;
    }
    case 3: {  // expf of the library, approximate reciprocal
        const float d = 1.0f + expf(-g);
        return g * rcpa(d);
    }
    case 4: {  // expf of the library, refined + corrected quotient
        const float d = 1.0f + expf(-g);
        const float r0 = rcpa(d);
        const float r = __fmaf_rn(r0, __fmaf_rn(-d, r0, 1.0f), r0);
        const float q = g * r;
        const float qq = __fmaf_rn(r, __fmaf_rn(-d, q, g), q);
        return d > 1.0e37f ? g / d : qq;
    }
    case 5: {  // two-term argument reduction in front of ex2 (the rounding of g * log2e is what hurts for g << 0), approximate reciprocal
        const float t = g * -1.4426950408889634f;
        const float lo = __fmaf_rn(g, -1.4426950408889634f, -t) + g * -1.925963033500011e-8f;  // what the product and the constant rounded away
        const float e0 = ex2a(t);
        const float e = __fmaf_rn(e0, lo * 0.6931471805599453f, e0);
        const float d = 1.0f + e;
        return g * rcpa(d);
    }
    default:  // the variant K1b ships
        return mxq::silu_bf16_input(g);
    }
}

__global__ void k(unsigned* mism, unsigned* first) {
    const unsigned bits = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 65535
    const float g = __uint_as_float(bits << 16);
    const uint16_t want = bf16_rn(g / (1.0f + expf(-g)));
    for (int v = 0; v < NV; ++v) {
        const uint16_t got = bf16_rn(variant(v, g));
        if (got != want) {
            const unsigned slot = atomicAdd(&mism[v], 1u);
            if (slot < 8) first[v * 8 + slot] = bits | ((unsigned)got << 16);
        }
    }
}

int main() {
    unsigned *d_m, *d_f, h_m[NV], h_f[NV * 8];
    cudaMalloc(&d_m, sizeof(h_m)); cudaMalloc(&d_f, sizeof(h_f));
    cudaMemset(d_m, 0, sizeof(h_m)); cudaMemset(d_f, 0, sizeof(h_f));
    k<<<256, 256>>>(d_m, d_f);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_m, d_m, sizeof(h_m), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_f, d_f, sizeof(h_f), cudaMemcpyDeviceToHost);
    const char* names[NV] = {"ex2_rcp", "ex2_rcp_newton", "ex2_rcp_newton_residual", "expf_rcp", "expf_rcp_newton_residual", "ex2_two_term_rcp", "shipped"};
    printf("{\"cuda\": \"%s\", \"inputs\": 65536, \"mismatches\": {", cudaGetErrorString(e));
    for (int v = 0; v < NV; ++v) {
        printf("%s\"%s\": {\"count\": %u, \"first_inputs_hex\": [", v ? ", " : "", names[v], h_m[v]);
        for (unsigned i = 0; i < (h_m[v] < 8 ? h_m[v] : 8); ++i) printf("%s\"%04x->%04x\"", i ? ", " : "", h_f[v * 8 + i] & 0xFFFF, h_f[v * 8 + i] >> 16);
        printf("]}");
    }
    printf("}}\n");
    return e == cudaSuccess && h_m[NV - 1] == 0 ? 0 : 1;
}
