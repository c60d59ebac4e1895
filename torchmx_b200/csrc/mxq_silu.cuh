// silu of a bf16 value, as aten computes it on a bf16 tensor: fp32 arithmetic  g / (1 + expf(-g)),  ONE rounding to bf16 by the caller.
// The input has only 65536 possible values, so the cheap evaluation below is not argued to be within some error bound of the
// plain formula: tools/silu_check.cu evaluates both for every bf16 bit pattern on the hardware and compares the bf16 roundings
// (recorded: profiles/r2_k1b_silu_check.json, zero differences; tests/test_gpu_silu_check.py reruns it).
#pragma once
#include <cuda_runtime.h>

namespace mxq {

// Everything is carried at a quarter of its size -- d/4 = 1/4 + 2^(-g log2(e) - 2) -- so that the reciprocal stays a normal number
// for the three inputs (-87.5, -88, -88.5) whose denominator exceeds 2^126 (rcp.approx.ftz would flush 1/d to zero); g/4 is exact
// (bf16 values have 8 significant bits, the build does not flush fp32 subnormals).  Past g = -88.72, where expf overflows and the
// plain formula gives -0, 4/d is subnormal and IS flushed: -0 as well.
__device__ __forceinline__ float silu_bf16_input(float g) {
    float e4, r4;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e4) : "f"(__fmaf_rn(g, -1.4426950408889634f, -2.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r4) : "f"(0.25f + e4));
    return (g * 0.25f) * r4;
}

}  // namespace mxq
