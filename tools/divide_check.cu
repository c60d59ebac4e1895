// Does the hoisted-reciprocal divide of K4a (csrc/mxq_softmax.cu) equal the compiler's `a / b` and the correctly rounded quotient
// on the real hardware reciprocal?  Sweeps denominators with (nearly) all-ones and all-zeros mantissas -- the classical hard case
// of reciprocal-refinement division -- and random ones, against numerators 1.0 (present in every softmax row), powers of two and
// random values in [2^-80, 1].  Build and run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/divide_check.cu && ./a.out
// Recorded result: profiles/r1_k4a_divide_check.json (2.0e9 divides, zero differences).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* cnt, int n_rand) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    unsigned long long fast_ne_ref = 0, fast_ne_exact = 0, ref_ne_exact = 0, total = 0;
    uint32_t rng = 0x9E3779B9u * (tid + 1);
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; };
    for (int i = tid; i < 17 * 4096 + n_rand; i += nth) {
        uint32_t mant, ex;
        if (i < 17 * 4096) { ex = i / 4096; const int k = i % 4096; mant = k < 2048 ? 0x7FFFFFu - k : (uint32_t)(k - 2048); }  // all-ones side and all-zeros side
        else { mant = next() & 0x7FFFFFu; ex = next() % 17; }
        const float b = __uint_as_float(((127u + ex) << 23) | mant);
        float r0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b));
        const float r = __fmaf_rn(r0, __fmaf_rn(-b, r0, 1.0f), r0);
        for (int j = 0; j < 40; ++j) {
            float a;
            if (j == 0) a = 1.0f; else if (j < 12) a = __uint_as_float((127u - (uint32_t)(j * 7)) << 23);
            else a = __uint_as_float(((127u - 1u - next() % 80u) << 23) | (next() & 0x7FFFFFu));
            const float q = a * r;
            const float fast = __fmaf_rn(r, __fmaf_rn(-b, q, a), q);
            const float ref = a / b;
            const float exact = (float)((double)a / (double)b);
            fast_ne_ref += __float_as_uint(fast) != __float_as_uint(ref);
            fast_ne_exact += __float_as_uint(fast) != __float_as_uint(exact);
            ref_ne_exact += __float_as_uint(ref) != __float_as_uint(exact);
            ++total;
        }
    }
    atomicAdd(&cnt[0], fast_ne_ref); atomicAdd(&cnt[1], fast_ne_exact); atomicAdd(&cnt[2], ref_ne_exact); atomicAdd(&cnt[3], total);
}
int main() {
    unsigned long long* d; unsigned long long h[4] = {0, 0, 0, 0};
    cudaMalloc(&d, sizeof(h)); cudaMemset(d, 0, sizeof(h));
    k<<<592, 256>>>(d, 50000000);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"cuda\": \"%s\", \"divides\": %llu, \"hoisted_ne_compiler\": %llu, \"hoisted_ne_correctly_rounded\": %llu, \"compiler_ne_correctly_rounded\": %llu}\n",
           cudaGetErrorString(e), h[3], h[0], h[1], h[2]);
    return 0;
}
