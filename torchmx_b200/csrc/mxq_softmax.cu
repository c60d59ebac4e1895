// K4a: attention probabilities in one pass -- scale, additive mask, softmax and MX quantization of P.
//
// Replaces the reference op chain torchmx/layers/mx_llama_attention.py:214-239
//     attn_weights = matmul(q_mx, k_mx^T) / sqrt(head_dim)      (bf16; CUDA evaluates x * (1/sqrt) in fp32, rounds to bf16)
//     attn_weights = attn_weights + causal_mask                 (bf16)
//     attn_weights = softmax(attn_weights, dim=-1, dtype=float32).to(bf16)
//     attn_weights = MXTensor.to_mx(attn_weights, elem_dtype, 32)
// which moves ~3.3 GB per Llama-8B layer at 2048 tokens (scores 268 MB, read and written by every step, fp32 in the middle);
// this kernel reads the bf16 scores once and writes codes + scales: 2 + 1 + 1/32 B per element.
//
// Every intermediate rounding of the chain is reproduced (two bf16 roundings before the softmax, exp(x - max) / sum in fp32
// with expf and an IEEE divide, one bf16 rounding after); what differs from the unfused chain is the ORDER of the fp32 row
// sum, i.e. the last ulp of the denominator, so a probability can land on the other side of a bf16 rounding boundary once in
// ~1e4 elements.  The quantization itself is K1's arithmetic (mxq_quant_core.cuh), NaN rows included.
//
// Layout: one thread owns one 32-element MX block of a row (32 fp32 values in registers).  Rows of 8 .. 256 blocks (the
// prefill sizes) are laid out 4 rows x 8 blocks per warp, so that causally hidden blocks fill whole warps; shorter rows share
// a warp lane-segment-wise, longer ones span whole warps.  Row reductions: shuffles + one shared-memory exchange.  L <= 32768.
#include <cstdio>

#include "mxq_softmax_core.cuh"

namespace mxq {

struct SoftmaxParams {
    const uint16_t* scores;
    const uint16_t* mask;
    int64_t mask_sb, mask_sh, mask_sq;  // element strides of the mask over (batch, head, query row); kv stride is 1
    int64_t rows;
    int L, tpr;                         // row length, threads (= MX blocks) per row
    int layout;                         // thread <-> (row, block) mapping: 0 = A, 1 = C, 2 = B (see the kernel)
    int heads, q_len;
    int causal, causal_offset;          // kv index j of query row q is visible iff j <= q + causal_offset
    int mask_vec;                       // mask rows are 16-byte aligned -> 128-bit loads
    float scaling;
    uint8_t* codes;
    uint8_t* scales;
    uint32_t flags;
};

using sm::max_nan;

template <int ELEM, int MAXT>
__global__ void __launch_bounds__(MAXT) softmax_quantize_kernel(const SoftmaxParams p) {
    __shared__ float red[32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t row;
    int t;
    bool live;
    int seg_base = 0;       // layout A: first lane of this row's segment
    int warps_per_row = 1;  // layouts B, C: warps that share a row
    const int layout = p.layout;
    if (layout == 0) {
        // A (rows of at most 32 blocks): a row is tpr consecutive lanes, 32 / tpr rows per warp
        const int rpw = 32 / p.tpr;
        const int r = lane / p.tpr;
        t = lane - r * p.tpr;
        seg_base = r * p.tpr;
        row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * rpw + r;
        live = r < rpw && row < p.rows;
    } else if (layout == 1) {
        // C (masked rows of 8 .. 256 blocks): a warp holds 8 consecutive blocks of 4 consecutive rows, a row spans ceil(tpr / 8) warps.
        // Blocks hidden by a causal mask are then (nearly) uniform across the warp, which skips their arithmetic as a whole.
        warps_per_row = (p.tpr + 7) >> 3;
        const int g = warp / warps_per_row;  // 4-row group within the CTA
        t = (warp - g * warps_per_row) * 8 + (lane & 7);
        row = ((int64_t)blockIdx.x * ((blockDim.x >> 5) / warps_per_row) + g) * 4 + (lane >> 3);
        live = t < p.tpr && row < p.rows;
    } else {
        // B (everything else): a row is ceil(tpr / 32) whole warps
        warps_per_row = (p.tpr + 31) >> 5;
        const int r = warp / warps_per_row;
        t = (warp - r * warps_per_row) * 32 + lane;
        row = (int64_t)blockIdx.x * ((blockDim.x >> 5) / warps_per_row) + r;
        live = t < p.tpr && row < p.rows;
    }

    // ---- load 32 scores, apply scale and mask with the reference's bf16 roundings --------------------------------------
    // causal: `vis` = how many of this block's 32 positions the query row may see.  A block with vis == 0 is never loaded:
    // its probabilities are exp(-inf - max) / sum, i.e. +0 (or NaN when the row is NaN), which is written directly below.
    float x[32];
    int vis = live ? 32 : 0;
    if (live && p.causal) {
        const int q = (int)(row % p.q_len);
        vis = min(32, max(0, q + p.causal_offset + 1 - t * 32));
    }
    if (vis > 0) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.scores) + (row * p.L + (int64_t)t * 32) * 2;
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const u32x8 v = ldg256_stream(src + 32 * j);
#pragma unroll
            for (int k = 0; k < 8; ++k) w[8 * j + k] = v.v[k];
        }
        sm::scale_round(w, p.scaling, x);
        if (p.mask) {
            const int64_t q = row % p.q_len, bh = row / p.q_len;
            const int64_t h = bh % p.heads, b = bh / p.heads;
            uint32_t mw[16];
            sm::load_mask(p.mask + b * p.mask_sb + h * p.mask_sh + q * p.mask_sq + (int64_t)t * 32, p.mask_vec != 0, mw);
            sm::add_mask(x, mw);
        }
        if (vis < 32) sm::hide_from(x, vis);
    }

    // ---- row max ---------------------------------------------------------------------------------------------------
    // NaN-propagating max: a NaN score makes row_max NaN, every exp(x - NaN) NaN and the whole row NaN -- the same outcome
    // as the unfused softmax (whose max drops NaNs but whose sum picks them up).
    float m = -INFINITY;  // a block the query row cannot see at all (vis == 0) never touches x[]
    float lo = INFINITY;  // smallest entry: tells below whether any exp() can be denormal-ish without a per-element test
    if (vis > 0) sm::block_max(x, m, lo);
    auto row_reduce = [&](float v, bool is_max) -> float {
        if (layout == 0) {
            float acc = is_max ? -INFINITY : 0.0f;
            for (int j = 0; j < p.tpr; ++j) {
                const float o = __shfl_sync(0xFFFFFFFFu, v, (seg_base + j) & 31);
                acc = is_max ? max_nan(acc, o) : acc + o;
            }
            return acc;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            if (layout == 1 && d > 4) continue;  // C: the row's 8 lanes only
            const float o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
            v = is_max ? max_nan(v, o) : v + o;
        }
        __syncthreads();  // red[] may still be read from the previous reduction
        const int slot = layout == 1 ? (lane >> 3) : 0;  // C keeps one partial per row of the warp
        if (layout == 1 ? (lane & 7) == 0 : lane == 0) red[warp * 4 + slot] = v;
        __syncthreads();
        const int w0 = (warp / warps_per_row) * warps_per_row;
        float acc = red[w0 * 4 + slot];
        for (int j = 1; j < warps_per_row; ++j) acc = is_max ? max_nan(acc, red[(w0 + j) * 4 + slot]) : acc + red[(w0 + j) * 4 + slot];
        return acc;
    };
    const float row_max = row_reduce(m, true);

    // ---- exp, row sum, normalise, round to bf16 ----------------------------------------------------------------------
    // A block whose largest entry sits more than 110 below the row max has exp() == +0 for every element (expf underflows to
    // zero below about -104): hidden by the causal rule, by a -inf / finfo.min additive mask, or simply negligible.  Its
    // exponentials, divides and conversions are skipped; NaNs fail the comparison and take the full path.
    float s = 0.0f;
    const bool dead = sm::block_dead(vis, m, row_max);
    if (!dead) s = sm::exp_sum(x, row_max);
    const float row_sum = row_reduce(s, false);
    if (!live) return;
    const int64_t blk = row * p.tpr + t;
    constexpr int NO = (ELEM == MXQ_ELEM_E2M1) ? 4 : 8;
    if (dead) {
        // hidden block: +0 everywhere -> codes 0, scale = shared exponent of an all-zero block; NaN row -> scale 255, codes 0
        uint8_t* dst = p.codes + blk * (NO * 4);
        if constexpr (NO == 4) stg128_stream(dst, make_uint4(0, 0, 0, 0));
        else {
            u32x8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = 0;
            stg256_stream(dst, o);
        }
        p.scales[blk] = (uint8_t)sm::dead_block_scale<ELEM>(row_max, row_sum);
        return;
    }
    uint32_t w[16];
    sm::normalize(x, lo, row_max, row_sum, w);

    // ---- K1's block quantizer ------------------------------------------------------------------------------------------
    uint32_t out[NO];
    const int sc = quantize_block32<ELEM>(w, (p.flags & MXQ_FLAG_HW_EXACT) != 0, out);
    uint8_t* dst = p.codes + blk * (NO * 4);
    if constexpr (NO == 4) stg128_stream(dst, make_uint4(out[0], out[1], out[2], out[3]));
    else {
        u32x8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = out[k];
        stg256_stream(dst, o);
    }
    p.scales[blk] = (uint8_t)sc;
}

int launch_softmax_quantize(const mxq_softmax_args_t* a, cudaStream_t stream, char* msg, size_t msg_len) {
    if (a->kv_len % 32 || a->kv_len > 32768) {
        snprintf(msg, msg_len, "kv_len %lld must be a multiple of 32 and at most 32768", (long long)a->kv_len);
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    if (((uintptr_t)a->scores % 32) || ((uintptr_t)a->codes % 32)) {
        snprintf(msg, msg_len, "scores and codes must be 32-byte aligned");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    SoftmaxParams p;
    p.scores = (const uint16_t*)a->scores;
    p.mask = (const uint16_t*)a->mask;
    p.mask_sb = a->mask_stride_b; p.mask_sh = a->mask_stride_h; p.mask_sq = a->mask_stride_q;
    p.rows = a->batch * a->heads * a->q_len;
    p.L = (int)a->kv_len;
    p.tpr = p.L / 32;
    p.heads = (int)a->heads; p.q_len = (int)a->q_len;
    p.causal = a->causal ? 1 : 0;
    p.causal_offset = (int)(a->kv_len - a->q_len);
    p.mask_vec = a->mask && ((uintptr_t)a->mask % 16 == 0) && (a->mask_stride_b % 8 == 0) && (a->mask_stride_h % 8 == 0) && (a->mask_stride_q % 8 == 0);
    p.scaling = a->scaling;
    p.codes = (uint8_t*)a->codes; p.scales = a->scales; p.flags = a->flags;
    int threads;
    int64_t rows_per_cta;
    const bool masked = p.causal || p.mask;
    p.layout = sm::sum_layout(p.tpr, masked);  // (K4b adds the block sums of a row in the order this choice implies)
    if (p.layout == 0) {
        threads = 256;
        rows_per_cta = (int64_t)(32 / p.tpr) * (threads / 32);
    } else if (p.layout == 1) {  // measured on [32, 2048, 2048]: 6 % faster than layout 2 with a causal mask, 9 % slower without any mask
        const int wpr = (p.tpr + 7) / 8;
        const int groups = wpr >= 8 ? 1 : 8 / wpr;
        threads = wpr * groups * 32;
        rows_per_cta = 4 * groups;
    } else {
        const int wpr = (p.tpr + 31) / 32;
        const int rpc = wpr >= 8 ? 1 : 8 / wpr;
        threads = wpr * rpc * 32;
        rows_per_cta = rpc;
    }
    const int64_t grid = (p.rows + rows_per_cta - 1) / rows_per_cta;
    if (grid > 0x7FFFFFFF) {
        snprintf(msg, msg_len, "too many rows (%lld)", (long long)p.rows);
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
#define MXQ_SOFTMAX_CASE(E)                                                                              \
    case E:                                                                                              \
        if (threads <= 256) softmax_quantize_kernel<E, 256><<<(unsigned)grid, threads, 0, stream>>>(p);  \
        else softmax_quantize_kernel<E, 1024><<<(unsigned)grid, threads, 0, stream>>>(p);                \
        break;
    switch (a->elem) {
        MXQ_SOFTMAX_CASE(MXQ_ELEM_E4M3)
        MXQ_SOFTMAX_CASE(MXQ_ELEM_E3M2)
        MXQ_SOFTMAX_CASE(MXQ_ELEM_E2M3)
        MXQ_SOFTMAX_CASE(MXQ_ELEM_E2M1)
        MXQ_SOFTMAX_CASE(MXQ_ELEM_INT8)
        MXQ_SOFTMAX_CASE(MXQ_ELEM_E5M2)
    default: snprintf(msg, msg_len, "unknown element type %d", a->elem); return MXQ_ERR_INVALID;
    }
#undef MXQ_SOFTMAX_CASE
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(msg, msg_len, "launch: %s", cudaGetErrorString(e));
        return MXQ_ERR_CUDA;
    }
    return MXQ_OK;
}

}  // namespace mxq
