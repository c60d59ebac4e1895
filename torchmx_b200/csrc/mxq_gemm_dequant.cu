// K3d: MX matmul for operands the block-scaled tensor-core instruction cannot take -- int8 elements, block sizes other than
// 32, blocks that do not run along the contraction (README matmul: B blocked along N), padded tensors, K % 128 != 0, arbitrary
// strides.  The reference handles every MX matmul this way (torchmx/ops.py:29-41, 60-68, 99-119): dequantize both operands to
// bf16, then a bf16 GEMM with fp32 accumulation.  Here the two steps are one kernel:
//
//   D[b][m][n] = bf16( sum_k bf16(dec(A[b][m][k]) * 2^(sa-127)) * bf16(dec(B[b][n][k]) * 2^(sb-127)) (+ bias[n]) )
//
// * 8 producer warps read element codes and scales straight from the caller's (strided) tensors, dequantize with K2's
//   arithmetic (exact decode, exact fp32 product, ONE rounding to bf16: bit-identical to mxq_dequantize) and write bf16
//   K-major tiles into shared memory in the 128B-swizzled layout the tensor core reads: the bf16 operands never exist in HBM;
// * one thread issues tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 accumulator in TMEM, M = 128, N = 128, K = 16 per
//   instruction), a 4-stage mbarrier ring decouples it from the producers;
// * the producer warps then drain the accumulator (tcgen05.ld), add the bias, round once to bf16 and store.
// Only the accumulation order differs from the reference's recipe.
#include <cstring>

#include "mxq_tc.cuh"

namespace mxq {
namespace gemm {
namespace dq {

constexpr int TILE = 128;       // output tile: 128 x 128
constexpr int BK = 64;          // bf16 elements per stage and row = one 128-byte swizzle row
constexpr int STAGES = 3;       // bf16 operand stages (what the tensor core reads)
constexpr int RAW_STAGES = 6;   // raw code stages (what TMA writes): the deep ring that covers the DRAM / L2 latency
constexpr int kProducerThreads = 512;             // 8 warps per operand: the dequantization is latency-bound per warp (one chunk chain at
                                                  // a time), so twice the warps of the first version is nearly twice the rate
constexpr int kPerOperand = kProducerThreads / 2;
constexpr int CHUNKS = TILE * 4 / kPerOperand;    // 16-element chunks per thread and stage (fast path)
constexpr int kMmaWarp = kProducerThreads / 32, kTmaWarp = kMmaWarp + 1;
constexpr int kThreads = kProducerThreads + 64;  // + MMA warp + TMA warp
constexpr int STAGE_BYTES = TILE * BK * 2;  // 16 KB per operand
constexpr int RAW_BYTES = TILE * BK;        // 8 KB per operand (fp4 uses half of it)

struct Smem {
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * STAGE_BYTES;
    static constexpr int OFF_RAW_A = OFF_B + STAGES * STAGE_BYTES;
    static constexpr int OFF_RAW_B = OFF_RAW_A + RAW_STAGES * RAW_BYTES;
    static constexpr int OFF_LUT = OFF_RAW_B + RAW_STAGES * RAW_BYTES;  // 2 x 256 fp32: decoded value of every code byte, per operand
    static constexpr int OFF_BAR = OFF_LUT + 2 * 256 * 4;
    static constexpr int NUM_BARS = 2 * STAGES + 2 * RAW_STAGES + 1;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;
    static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(RAW_STAGES * 2 * STAGE_BYTES <= OFF_LUT, "direct mode lays its operand ring over the operand + raw area");
};

struct Operand {
    const uint8_t* codes; const uint8_t* scales;
    int64_t row_stride, k_stride, batch_stride, srow_stride, sk_stride, sbatch_stride;
    int rows;         // M (A) or N (B)
    int elem, block_size, along_k;
    int bs_shift;     // log2(block_size) when it is a power of two, else -1
    int fast;         // codes K-contiguous, 16-byte aligned rows, blocks of >= 16 along K: raw tiles arrive by TMA
    int tma_batched;  // the raw tensor map has a batch dimension (batch stride != 0)
};

struct Params {
    Operand a, b;
    const uint16_t* bias; uint16_t* d;
    int64_t ldd, d_batch;
    int M, N, K, m_blocks, n_blocks;
    int direct;  // the operands ARE bf16 K-major matrices (dequantized once by K2): their tiles arrive by TMA, nobody dequantizes
};

__device__ __forceinline__ float lut_entry(int elem, uint32_t code) {
    switch (elem) {
    case MXQ_ELEM_E4M3: return decode_one<MXQ_ELEM_E4M3>(code);
    case MXQ_ELEM_E3M2: return decode_one<MXQ_ELEM_E3M2>(code);
    case MXQ_ELEM_E2M3: return decode_one<MXQ_ELEM_E2M3>(code);
    case MXQ_ELEM_E2M1: return decode_one<MXQ_ELEM_E2M1>(code);
    case MXQ_ELEM_E5M2: return decode_one<MXQ_ELEM_E5M2>(code);
    default: return decode_one<MXQ_ELEM_INT8>(code);
    }
}

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- generic path: eight consecutive k of one operand row -> eight bf16 (one 16-byte shared-memory chunk).  `r` is the row
// inside the operand, k0 a multiple of 8.  Rows / k past the edge give zeros.  Any strides, any block size / orientation.
__device__ __forceinline__ uint4 dequant_chunk(const Operand& op, const uint8_t* codes, const uint8_t* scales, const float* lut, int r, int k0, int K) {
    float f[8];
    if (r >= op.rows || k0 >= K) return make_uint4(0u, 0u, 0u, 0u);
    const bool fp4 = op.elem == MXQ_ELEM_E2M1;
    uint32_t c[8];
    // ---- codes
    if (op.k_stride == 1 && (op.along_k || !fp4) && k0 + 8 <= K) {  // K-contiguous: eight codes are 8 (fp4: 4) adjacent bytes
        if (fp4) {
            const uint8_t* p = codes + (int64_t)r * op.row_stride + (k0 >> 1);
            uint32_t w;
            if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) w = *reinterpret_cast<const uint32_t*>(p);
            else w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // the earlier element sits in the HIGH nibble (torchmx/utils.py:145)
                c[2 * j] = (w >> (8 * j + 4)) & 0xF;
                c[2 * j + 1] = (w >> (8 * j)) & 0xF;
            }
        } else {
            const uint8_t* p = codes + (int64_t)r * op.row_stride + k0;
            uint32_t lo, hi;
            if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
                const uint2 v = *reinterpret_cast<const uint2*>(p);
                lo = v.x; hi = v.y;
            } else {
                lo = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                hi = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j] = (lo >> (8 * j)) & 0xFF; c[4 + j] = (hi >> (8 * j)) & 0xFF; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            if (k >= K) { c[j] = 0x100; continue; }  // marker: contributes zero
            if (!fp4) {
                c[j] = codes[(int64_t)r * op.row_stride + (int64_t)k * op.k_stride];
            } else if (op.along_k) {
                const uint32_t byte = codes[(int64_t)r * op.row_stride + (int64_t)(k >> 1) * op.k_stride];
                c[j] = (k & 1) ? (byte & 0xF) : (byte >> 4);
            } else {  // packed along the row index
                const uint32_t byte = codes[(int64_t)(r >> 1) * op.row_stride + (int64_t)k * op.k_stride];
                c[j] = (r & 1) ? (byte & 0xF) : (byte >> 4);
            }
        }
    }
    // ---- scales + product (exact in fp32; s == 255 -> NaN for the whole block, also for zero codes)
    if (op.along_k && op.bs_shift >= 3) {  // the chunk lies inside one block
        const float sc = scale_f32(scales[(int64_t)r * op.srow_stride + (int64_t)(k0 >> op.bs_shift) * op.sk_stride]);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c[j] & 0x100) ? 0.0f : lut[c[j]] * sc;
    } else {
        const int rb = op.along_k ? r : (op.bs_shift >= 0 ? (r >> op.bs_shift) : r / op.block_size);
        int last = -1;
        float sc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            if (c[j] & 0x100) { f[j] = 0.0f; continue; }
            const int kb = op.along_k ? (op.bs_shift >= 0 ? (k >> op.bs_shift) : k / op.block_size) : k;
            if (kb != last) {
                sc = scale_f32(scales[(int64_t)rb * op.srow_stride + (int64_t)kb * op.sk_stride]);
                last = kb;
            }
            f[j] = lut[c[j]] * sc;
        }
    }
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---- fast path: sixteen consecutive codes of one block (one 16-byte raw chunk; fp4: 8 bytes) times one scale -> 16 bf16 ----
// Same arithmetic as K2 (exact decode, exact fp32 product, one rounding to bf16).
template <int ELEM>
__device__ __forceinline__ void dequant16(const uint32_t (&raw)[4], int s_byte, uint32_t (&out)[8]) {
    const float sc = scale_f32(s_byte);
    if constexpr (ELEM == MXQ_ELEM_INT8) {
        if (s_byte <= 230) {
            // v + 128 placed in the low mantissa byte of 2^23: f = 2^23 + 128 + v exactly; (f - (2^23 + 128)) * sc as one FMA whose
            // exact result v * sc is representable, so the single rounding of the FMA does not round at all
            const float neg_c = -8388736.0f * sc;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t x = raw[w] ^ 0x80808080u;
                const float f0 = fmaf(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540)), sc, neg_c);
                const float f1 = fmaf(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7541)), sc, neg_c);
                const float f2 = fmaf(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7542)), sc, neg_c);
                const float f3 = fmaf(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7543)), sc, neg_c);
                out[2 * w] = pack_bf16x2(f0, f1);
                out[2 * w + 1] = pack_bf16x2(f2, f3);
            }
        } else {  // 2^23 * sc would overflow: plain conversion (also the NaN scale)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = (float)(int)(int8_t)(raw[w] >> (8 * j)) * sc;
                out[2 * w] = pack_bf16x2(f[0], f[1]);
                out[2 * w + 1] = pack_bf16x2(f[2], f[3]);
            }
        }
    } else if constexpr (ELEM == MXQ_ELEM_E2M1) {  // raw[0], raw[1]: eight bytes = sixteen codes, earlier element in the HIGH nibble
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t h = decode_e2m1_byte_f16x2((raw[w] >> (8 * j)) & 0xFF);
                out[4 * w + j] = pack_bf16x2(f16hi_to_f32(h) * sc, f16lo_to_f32(h) * sc);
            }
    } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t h0 = decode_pair_f16x2<ELEM>(raw[w] & 0xFFFF), h1 = decode_pair_f16x2<ELEM>(raw[w] >> 16);
            out[2 * w] = pack_bf16x2(f16lo_to_f32(h0) * sc, f16hi_to_f32(h0) * sc);
            out[2 * w + 1] = pack_bf16x2(f16lo_to_f32(h1) * sc, f16hi_to_f32(h1) * sc);
        }
    }
}

__device__ __forceinline__ void dequant16_rt(int elem, const uint32_t (&raw)[4], int s_byte, uint32_t (&out)[8]) {
    switch (elem) {  // uniform over the CTA
    case MXQ_ELEM_E4M3: dequant16<MXQ_ELEM_E4M3>(raw, s_byte, out); break;
    case MXQ_ELEM_E3M2: dequant16<MXQ_ELEM_E3M2>(raw, s_byte, out); break;
    case MXQ_ELEM_E2M3: dequant16<MXQ_ELEM_E2M3>(raw, s_byte, out); break;
    case MXQ_ELEM_E2M1: dequant16<MXQ_ELEM_E2M1>(raw, s_byte, out); break;
    case MXQ_ELEM_E5M2: dequant16<MXQ_ELEM_E5M2>(raw, s_byte, out); break;
    default: dequant16<MXQ_ELEM_INT8>(raw, s_byte, out); break;
    }
}

__global__ void __launch_bounds__(kThreads, 1) mx_gemm_dequant_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                      const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::OFF_BAR);
    uint64_t* full = bars;                           // both bf16 tiles of the stage written (count 256: every producer thread)
    uint64_t* empty = bars + STAGES;                 // MMAs of the stage retired (count 1, tcgen05.commit)
    uint64_t* raw_full = bars + 2 * STAGES;          // raw code tiles landed (count 1 + tx)
    uint64_t* raw_empty = raw_full + RAW_STAGES;     // raw stage read into registers (count 4 per TMA-fed operand: one per warp)
    uint64_t* tmem_full = raw_empty + RAW_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::OFF_TMEM_PTR);
    float* lut = reinterpret_cast<float*>(smem + Smem::OFF_LUT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_batch = p.m_blocks * p.n_blocks;
    const int b = blockIdx.x / tiles_per_batch;
    const int t = blockIdx.x - b * tiles_per_batch;
    const int nb = t / p.m_blocks, mb = t - nb * p.m_blocks;  // m fastest: neighbouring CTAs share a B panel in L2
    const int k_steps = (p.K + BK - 1) / BK;
    const int n_fast = p.a.fast + p.b.fast;

    if (threadIdx.x < 256) {
        lut[threadIdx.x] = lut_entry(p.a.elem, threadIdx.x);
        lut[256 + threadIdx.x] = lut_entry(p.b.elem, threadIdx.x);
    }
    static_assert(kPerOperand == 256 || kPerOperand == 128, "thread <-> (row, chunk) maps assume 128 or 256 producer threads per operand");
    if (warp == kMmaWarp) {
        if (elect_one()) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&full[i], kProducerThreads);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < RAW_STAGES; ++i) {
                mbar_init(&raw_full[i], 1);
                // (direct mode: the raw ring's barriers and shared memory serve as a 6-deep ring of bf16 operand stages)
                mbar_init(&raw_empty[i], p.direct ? 1 : (kPerOperand / 32) * (n_fast > 0 ? n_fast : 1));
            }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<128>(tmem_ptr);
    }
    if (warp == kTmaWarp && elect_one()) {
        if (p.a.fast || p.direct) tma_prefetch_desc(&map_a);
        if (p.b.fast || p.direct) tma_prefetch_desc(&map_b);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < kMmaWarp) {
        // ================= producers: warps 0..7 build the A tile, warps 8..15 the B tile =================
        const bool is_b = threadIdx.x >= kPerOperand;
        const Operand& op = is_b ? p.b : p.a;
        const int tidx = threadIdx.x & (kPerOperand - 1);
        const int row0 = (is_b ? nb : mb) * TILE;
        const uint8_t* codes = op.codes + (int64_t)b * op.batch_stride;
        const uint8_t* scales = op.scales + (int64_t)b * op.sbatch_stride;
        const float* my_lut = lut + (is_b ? 256 : 0);
        uint8_t* tile_base = smem + (is_b ? Smem::OFF_B : Smem::OFF_A);
        const uint8_t* raw_base = smem + (is_b ? Smem::OFF_RAW_B : Smem::OFF_RAW_A);
        uint32_t stage = 0, phase = 0, rs = 0, rphase = 0;
        if (op.fast) {
            // TMA-fed: the raw tile is [128 rows][64 B] (fp4: 32 B), no swizzle.  Work item = (row, 16-element chunk q); a warp
            // reads 512 (256) contiguous raw bytes per instruction.  item = i * 128 + tidx -> row = item >> 2, q = item & 3.
            const bool fp4 = op.elem == MXQ_ELEM_E2M1;
            const int q = tidx & 3;
            int rows_i[CHUNKS];
            const uint8_t* sc_ptr[CHUNKS];
#pragma unroll
            for (int i = 0; i < CHUNKS; ++i) {
                rows_i[i] = (i * kPerOperand + tidx) >> 2;
                const int r = min(row0 + rows_i[i], op.rows - 1);  // rows past the edge carry zero codes; any in-range scale will do
                sc_ptr[i] = scales + (int64_t)r * op.srow_stride;
            }
            auto load_scales = [&](int ks, int (&sb)[CHUNKS]) {
                const int kq = ks * BK + 16 * q;
                const bool live = ks < k_steps && kq < p.K;
                const int64_t off = (int64_t)(kq >> op.bs_shift) * op.sk_stride;
#pragma unroll
                for (int i = 0; i < CHUNKS; ++i) sb[i] = live ? (int)__ldg(sc_ptr[i] + off) : 127;
            };
            int sb_cur[CHUNKS], sb_nxt[CHUNKS];
            load_scales(0, sb_cur);
            for (int ks = 0; ks < k_steps; ++ks) {
                load_scales(ks + 1, sb_nxt);
                mbar_wait(&raw_full[rs], rphase);
                const uint8_t* raw = raw_base + rs * RAW_BYTES;
                uint32_t rw[CHUNKS][4];
#pragma unroll
                for (int i = 0; i < CHUNKS; ++i) {
                    if (fp4) {
                        const uint2 v = *reinterpret_cast<const uint2*>(raw + rows_i[i] * 32 + q * 8);
                        rw[i][0] = v.x; rw[i][1] = v.y; rw[i][2] = 0; rw[i][3] = 0;
                    } else {
                        const uint4 v = *reinterpret_cast<const uint4*>(raw + rows_i[i] * 64 + q * 16);
                        rw[i][0] = v.x; rw[i][1] = v.y; rw[i][2] = v.z; rw[i][3] = v.w;
                    }
                }
                fence_proxy_async_smem();  // our reads of the raw stage are ordered before the TMA refill
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[rs]);
                if (++rs == RAW_STAGES) { rs = 0; rphase ^= 1; }
                const bool dead = ks * BK + 16 * q >= p.K;  // the whole chunk lies past K: zeros whatever the scale says
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* dst = tile_base + stage * STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < CHUNKS; ++i) {
                    uint32_t o[8];
                    if (dead || row0 + rows_i[i] >= op.rows) {  // past K / past the last row (decode: most of the 128-token tile): zeros
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = 0u;
                    } else {
                        dequant16_rt(op.elem, rw[i], sb_cur[i], o);
                    }
                    const int row = rows_i[i];
                    uint8_t* drow = dst + row * 128;
                    *reinterpret_cast<uint4*>(drow + (((2 * q) ^ (row & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(drow + (((2 * q + 1) ^ (row & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
                for (int i = 0; i < CHUNKS; ++i) sb_cur[i] = sb_nxt[i];
            }
        } else {
            // generic: two threads per row of the tile (four 8-element chunks each), element-wise addressing through the operand's strides
            constexpr int TPR = kPerOperand / TILE;
            const int row = tidx / TPR, c0 = (tidx % TPR) * (8 / TPR);
            const int r = row0 + row;
            for (int ks = 0; ks < (p.direct ? 0 : k_steps); ++ks) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* dst = tile_base + stage * STAGE_BYTES + row * 128;
#pragma unroll 4
                for (int c = c0; c < c0 + 8 / TPR; ++c)  // 16-byte chunk c of the row lives at chunk c ^ (row & 7) (SWIZZLE_128B)
                    *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = dequant_chunk(op, codes, scales, my_lut, r, ks * BK + c * 8, p.K);
                fence_proxy_async_smem();
                mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        // ================= epilogue: warp (quadrant, column group) drains 32 rows x 32 columns =================
        const int quad = warp & 3, cgrp = warp >> 2;
        if (k_steps > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        const int m = mb * TILE + quad * 32 + lane;
        uint16_t* drow = p.d + (int64_t)b * p.d_batch + (int64_t)m * p.ldd;
#pragma unroll 1
        for (int c = 0; c < TILE / 32 / (kMmaWarp / 4); ++c) {
            uint32_t v[32];
            if (k_steps > 0) {
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (cgrp * (TILE / 32 / (kMmaWarp / 4)) + c) * 32, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            const int col0 = nb * TILE + (cgrp * (TILE / 32 / (kMmaWarp / 4)) + c) * 32;
            if (m < p.M && col0 < p.N) {
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < p.N) f[i] += __uint_as_float((uint32_t)p.bias[col0 + i] << 16);
                }
                if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(drow + col0) & 15) == 0)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
                        o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
                        o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
                        o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
                        *reinterpret_cast<uint4*>(drow + col0 + 8 * i) = o;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < p.N) drow[col0 + i] = (uint16_t)pack_bf16x2(f[i], 0.0f);
                }
            }
        }
        tc_fence_before();
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
        // kind::f16 descriptor: fp32 accumulator (bit 4), bf16 A and B (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
        constexpr uint64_t HI_OPERAND = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutSw128 << 61);
        // direct mode: RAW_STAGES stages of [A tile | B tile] laid over the whole operand + raw area (nobody dequantizes, so the TMA
        // ring is all there is between DRAM latency and the tensor core: 6 x 32 KB in flight instead of 3)
        const uint32_t a_lo0 = smem_u32(smem + Smem::OFF_A) >> 4, b_lo0 = (p.direct ? smem_u32(smem + STAGE_BYTES) : smem_u32(smem + Smem::OFF_B)) >> 4;
        const uint32_t stage_step = (p.direct ? 2 * STAGE_BYTES : STAGE_BYTES) >> 4;
        const uint32_t n_stages = p.direct ? RAW_STAGES : STAGES;
        uint64_t* const full_b = p.direct ? raw_full : full;
        uint64_t* const empty_b = p.direct ? raw_empty : empty;
        uint32_t stage = 0, phase = 0;
        for (int ks = 0; ks < k_steps; ++ks) {
            mbar_wait(&full_b[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = a_lo0 + stage * stage_step, b_lo = b_lo0 + stage * stage_step;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)  // 16 bf16 = 32 bytes further along the swizzle row
                    tc_mma_f16(tmem_base, HI_OPERAND | (a_lo + k * 2), HI_OPERAND | (b_lo + k * 2), idesc, (ks | k) != 0);
                tc_commit(&empty_b[stage]);
                if (ks == k_steps - 1) tc_commit(tmem_full);
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        tc_fence_before();
    } else {
        // ================= TMA producer: bf16 operand tiles (direct mode), or the raw code tiles of operands in the fast form =================
        if (p.direct) {
            if (elect_one()) {
                uint32_t stage = 0, phase = 0;
                for (int ks = 0; ks < k_steps; ++ks) {
                    mbar_wait(&raw_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&raw_full[stage], 2 * STAGE_BYTES);
                    tma_load_3d(&map_a, &raw_full[stage], smem + stage * 2 * STAGE_BYTES, ks * BK, mb * TILE, b);
                    tma_load_3d(&map_b, &raw_full[stage], smem + stage * 2 * STAGE_BYTES + STAGE_BYTES, ks * BK, nb * TILE, b);
                    if (++stage == RAW_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (n_fast > 0 && elect_one()) {
            const uint32_t a_bytes = p.a.fast ? (p.a.elem == MXQ_ELEM_E2M1 ? TILE * 32 : TILE * 64) : 0;
            const uint32_t b_bytes = p.b.fast ? (p.b.elem == MXQ_ELEM_E2M1 ? TILE * 32 : TILE * 64) : 0;
            uint32_t rs = 0, rphase = 0;
            for (int ks = 0; ks < k_steps; ++ks) {
                mbar_wait(&raw_empty[rs], rphase ^ 1);
                mbar_arrive_expect_tx(&raw_full[rs], a_bytes + b_bytes);
                if (p.a.fast)
                    tma_load_3d(&map_a, &raw_full[rs], smem + Smem::OFF_RAW_A + rs * RAW_BYTES, ks * (p.a.elem == MXQ_ELEM_E2M1 ? 32 : 64), mb * TILE,
                                p.a.tma_batched ? b : 0);
                if (p.b.fast)
                    tma_load_3d(&map_b, &raw_full[rs], smem + Smem::OFF_RAW_B + rs * RAW_BYTES, ks * (p.b.elem == MXQ_ELEM_E2M1 ? 32 : 64), nb * TILE,
                                p.b.tma_batched ? b : 0);
                if (++rs == RAW_STAGES) { rs = 0; rphase ^= 1; }
            }
        }
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc<128>(tmem_base);
    }
}

static void fill_operand(Operand& o, const mxq_operand_t& s, int64_t rows, int64_t K, int64_t batch) {
    o.codes = (const uint8_t*)s.codes; o.scales = s.scales;
    o.row_stride = s.row_stride; o.k_stride = s.k_stride; o.batch_stride = s.batch_stride;
    o.srow_stride = s.srow_stride; o.sk_stride = s.sk_stride; o.sbatch_stride = s.sbatch_stride;
    o.rows = (int)rows; o.elem = s.elem; o.block_size = s.block_size; o.along_k = s.blocked_along_k ? 1 : 0;
    o.bs_shift = -1;
    for (int sft = 0; sft < 31; ++sft)
        if ((1 << sft) == s.block_size) o.bs_shift = sft;
    // TMA-fed form: K-contiguous codes, blocks of a power of two >= 16 along K (a 16-element chunk never straddles a block), rows
    // and batches on 16-byte boundaries
    o.fast = (K > 0 && s.k_stride == 1 && o.along_k && o.bs_shift >= 4 && ((uintptr_t)s.codes % 16) == 0 && s.row_stride % 16 == 0 && s.row_stride > 0 &&
              (batch <= 1 || s.batch_stride % 16 == 0) && s.batch_stride >= 0) ? 1 : 0;
    o.tma_batched = (batch > 1 && s.batch_stride > 0) ? 1 : 0;
}

}  // namespace dq

// D[b] = A[b] B[b]^T (+ bias) on bf16 operands that were dequantized ONCE (K2): the same MMA loop and epilogue with the operand
// tiles arriving by TMA.  For large M x N this beats dequantizing inside the GEMM, which repeats the work for every output tile.
int launch_gemm_bf16(const void* a, int64_t lda, int64_t a_bs, const void* b, int64_t ldb, int64_t b_bs, const void* bias, void* d, int64_t ldd, int64_t d_bs,
                     int64_t batch, int64_t M, int64_t N, int64_t K, int device, cudaStream_t stream, char* msg, size_t msg_len) {
    using namespace dq;
    if (M > 0x7FFFFFFF || N > 0x7FFFFFFF || K > 0x7FFFFFFF || K % 8 || lda % 8 || ldb % 8 || a_bs % 8 || b_bs % 8 || ((uintptr_t)a % 16) || ((uintptr_t)b % 16)) {
        snprintf(msg, msg_len, "bf16 operands need K %% 8 == 0 and 16-byte aligned rows");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    dq::Params p;
    memset(&p, 0, sizeof(p));
    p.a.rows = (int)M; p.b.rows = (int)N;
    p.bias = (const uint16_t*)bias; p.d = (uint16_t*)d;
    p.ldd = ldd; p.d_batch = d_bs;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.m_blocks = (int)((M + TILE - 1) / TILE);
    p.n_blocks = (int)((N + TILE - 1) / TILE);
    p.direct = 1;
    const int64_t ctas = (int64_t)p.m_blocks * p.n_blocks * batch;
    if (ctas > 0x7FFFFFFF) { snprintf(msg, msg_len, "too many output tiles"); return MXQ_ERR_UNSUPPORTED_SHAPE; }
    CUtensorMap maps[2];
    if (!cached_bf16_operand_map(&maps[0], a, K, M, batch, lda, a_bs, TILE, device) || !cached_bf16_operand_map(&maps[1], b, K, N, batch, ldb, b_bs, TILE, device)) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled failed");
        return MXQ_ERR_CUDA;
    }
    cudaError_t e = ensure_smem_attr((const void*)mx_gemm_dequant_kernel, Smem::DYN_BYTES, device);
    if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    mx_gemm_dequant_kernel<<<(unsigned)ctas, dq::kThreads, Smem::DYN_BYTES, stream>>>(maps[0], maps[1], p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch (bf16 gemm): %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    return MXQ_OK;
}

int launch_gemm_dequant(const mxq_gemm_dequant_args_t* a, int device, cudaStream_t stream, char* msg, size_t msg_len) {
    using namespace dq;
    if (a->M > 0x7FFFFFFF || a->N > 0x7FFFFFFF || a->K > 0x7FFFFFFF) { snprintf(msg, msg_len, "extent too large"); return MXQ_ERR_UNSUPPORTED_SHAPE; }
    dq::Params p;
    fill_operand(p.a, a->a, a->M, a->K, a->batch);
    fill_operand(p.b, a->b, a->N, a->K, a->batch);
    p.bias = (const uint16_t*)a->bias; p.d = (uint16_t*)a->d;
    p.ldd = a->ldd; p.d_batch = a->d_batch_stride;
    p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K;
    p.m_blocks = (int)((a->M + TILE - 1) / TILE);
    p.n_blocks = (int)((a->N + TILE - 1) / TILE);
    p.direct = 0;
    const int64_t ctas = (int64_t)p.m_blocks * p.n_blocks * a->batch;
    if (ctas > 0x7FFFFFFF) { snprintf(msg, msg_len, "too many output tiles"); return MXQ_ERR_UNSUPPORTED_SHAPE; }
    CUtensorMap maps[2];
    memset(maps, 0, sizeof(maps));
    for (int side = 0; side < 2; ++side) {
        Operand& o = side ? p.b : p.a;
        if (!o.fast) continue;
        const bool fp4 = o.elem == MXQ_ELEM_E2M1;
        const int64_t k_bytes = fp4 ? (a->K + 1) / 2 : a->K;
        if (!cached_raw_map(&maps[side], o.codes, k_bytes, o.rows, o.tma_batched ? a->batch : 1, o.row_stride, o.batch_stride, fp4 ? 32 : 64, TILE, device))
            o.fast = 0;  // the driver refused the layout: element-wise path
    }
    cudaError_t e = ensure_smem_attr((const void*)mx_gemm_dequant_kernel, Smem::DYN_BYTES, device);
    if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    mx_gemm_dequant_kernel<<<(unsigned)ctas, dq::kThreads, Smem::DYN_BYTES, stream>>>(maps[0], maps[1], p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch (dequant gemm): %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    return MXQ_OK;
}

}  // namespace gemm
}  // namespace mxq
