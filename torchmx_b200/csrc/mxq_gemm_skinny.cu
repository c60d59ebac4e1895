// K3c: skinny MX GEMM for decode-sized activations (M <= 128 rows), HBM-bound weight streaming.
//
//   D[t][n] = sum_k ( X[t][k] * 2^(sfx[t][k/32]-127) ) * ( W[n][k] * 2^(sfw[n][k/32]-127) ) (+ bias[n])
//
// Same contraction as K3 (torchmx/ops.py:29-41 with a [tokens, K] activation and an [N, K] weight), but the roles of
// the MMA operands are swapped: the WEIGHT rows are the MMA M dimension (128 per CTA) and the tokens are the MMA N
// dimension (32 / 64 / 128 columns), so a CTA streams a 128-row slab of W exactly once and the tensor core never
// idles on padding rows.  What bounds the kernel is reading W once from HBM: the grid is n_tiles x S CTAs where the
// S CTAs of a cluster split K, so that ~one CTA per SM is streaming even when N / 128 is far below the SM count
// (q/k/v/o projections).  The S partial accumulators are reduced through distributed shared memory in a fixed
// order (deterministic), each CTA finishing 1/S of the tokens.
//
// Per CTA: warp 0 TMA producer (W 128 x 128 B + X N_TOK x 128 B per K block, STAGES-deep ring), warps 1 and 8 MMA issuers
// (tcgen05.mma cta_group::1 kind::mxf8f6f4.block_scale, M=128, N=N_TOK), warps 2/3 scale-factor loaders (W rows /
// token rows, four K blocks per ring stage), warps 4..7 epilogue.
//
// Two issuers, two accumulators: at these widths the tensor core spends ~20 cycles on a K block, but the instruction chain
// that issues it (barrier polls, two tcgen05.cp, four tcgen05.mma, one to three tcgen05.commit) takes one thread ~540 cycles
// (measured: profiles/r2_skinny_issue_loop.txt) -- more than streaming the K block's 8 / 12 KB of packed 4 / 6-bit weights
// takes, so the packed formats ran no faster than one-byte codes.  For N_TOK <= 64 the even K blocks therefore go to warp 1
// and accumulator 0, the odd ones to warp 8 and accumulator 1 (one-CTA-per-SM variants) (own scale-factor columns in TMEM each); the epilogue adds
// the two accumulators in that order, so the result is deterministic.
#include "mxq_quant_core.cuh"
#include "mxq_tc.cuh"

namespace mxq {
namespace gemm {
namespace skinny {

constexpr int TILE_W = 128;  // weight rows per CTA
constexpr int kThreads = 288;
constexpr int SF_KB_BYTES = 512;  // one K block of scale factors for (up to) 128 rows

struct Params {
    const uint8_t* sfx; const uint8_t* sfw; const uint16_t* bias; uint16_t* d;
    uint16_t* d_mc;  // multicast alias of the output on every rank: add (multimem.red) instead of store
    const uint16_t* x_hp; int64_t ldx; int x_flags;  // fused mode: bf16 activation, quantized to e4m3 / block 32 in the kernel
    int64_t ld_sfx, ld_sfw, ldd;
    int M, N, K, splits;
    int tile_rows;  // weight rows a CTA streams (<= TILE_W, multiple of 8): chosen so that the tiles spread evenly over the SMs
    uint32_t idesc_fmt, tx_w, tx_x;  // element formats of the descriptor; bytes a W / X box posts on the mbarrier
    int pdl;      // launched with programmatic stream serialization: X / its scales may only be read after griddepcontrol.wait
    int w_static; // ... and W / its scales too, unless the caller vouches that no earlier kernel writes them (MXQ_GEMM_B_STATIC)
    int pf_dist;  // L2 prefetch distance of the W stream, in K blocks (developer builds; 0 = off)
    int exp;           // developer builds: timing experiments (wrong results)
    long long* trace;  // developer builds: 8 globaltimer + 8 clock64 stamps per CTA (tools/skinny_trace.py)
    int sf_tma;   // scales are 16-byte aligned with a 16-byte multiple row pitch: fetch them with TMA (deep prefetch)
};

constexpr int RAW_X = 2;  // raw scale ring of the (L2-resident) token scales

template <int N_TOK, int STAGES, int RAW_W>
struct Smem {
    static constexpr int W_STAGE = TILE_W * BLOCK_K;  // 16 KB
    static constexpr int X_STAGE = N_TOK * BLOCK_K;   // 4 / 8 / 16 KB
    static constexpr int SF_STAGE = SF_KB * SF_KB_BYTES;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_X = OFF_W + STAGES * W_STAGE;
    static constexpr int OFF_SFW = OFF_X + STAGES * X_STAGE;
    static constexpr int OFF_SFX = OFF_SFW + SF_STAGES * SF_STAGE;
    static constexpr int OFF_RAW_W = OFF_SFX + SF_STAGES * SF_STAGE;  // TMA landing buffers of the scales: [128 rows][16 B] each
    static constexpr int OFF_RAW_X = OFF_RAW_W + RAW_W * 2048;
    // split-K, deep variants: receive buffer the S CTAs of a cluster PUSH their partial sums into -- [source split][token of
    // mine][128 rows] fp32, never aliased with the rings (a peer may push while this CTA's MMAs still read its W stages)
    static constexpr bool PUSH = N_TOK <= 64 && STAGES >= 6;
    static constexpr int RECV_BYTES = PUSH ? (N_TOK + 8) * TILE_W * 4 : 0;  // S * ceil(N_TOK / S) <= N_TOK + S - 1 token slots
    static constexpr int OFF_RECV = OFF_RAW_X + RAW_X * 2048;
    static constexpr int OFF_BAR = OFF_RECV + RECV_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 2 * SF_STAGES + 1 + RAW_W + RAW_X;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;
    // split-K partials (fp32 [N_TOK][128], token-major so that a warp writes / reads 128 contiguous bytes) reuse the
    // W ring once every MMA has retired
    static_assert(STAGES * W_STAGE >= N_TOK * TILE_W * 4, "partial-sum buffer does not fit the W ring");
    static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget");
};

// (volatile keeps it between the two cluster barriers, which are volatile asm statements themselves; no "memory" clobber, so
// that the loads of an item are in flight together)
__device__ __forceinline__ void st_dsmem_v4(uint32_t cluster_addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t cluster_addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
    return v;
}

#ifdef MXQ_DEV
#define SK_EXP(bit) ((p.exp & (bit)) != 0)  // timing experiments of developer builds (wrong results); compiled out of the shipped library
__device__ __forceinline__ long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; }
#define SK_TRACE(i) if (p.trace != nullptr) { p.trace[blockIdx.x * 16 + (i)] = gtimer(); p.trace[blockIdx.x * 16 + 8 + (i)] = clock64(); }
#else
#define SK_EXP(bit) false
#define SK_TRACE(i)
#endif

template <int N_TOK, int STAGES, int RAW_W>
__global__ void __launch_bounds__(kThreads, STAGES <= 4 ? 2 : 1) mx_gemm_skinny_kernel(  // (the 4-stage variants exist to run two CTAs per SM)
    const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ CUtensorMap map_sfw,
                                                                  const __grid_constant__ CUtensorMap map_sfx, const Params p) {
    using L = Smem<N_TOK, STAGES, RAW_W>;
    constexpr bool BATCH_COMMITS = STAGES >= 6;  // release the ring stages of a scale-factor group together (see the header)
    // (two CTAs per SM -- the 4-stage variants -- already have two issuers per SM; measured no gain there, and the fused
    // activation quantizer lost 10 %.  A second 128-column accumulator would not fit twice either.)
    constexpr int ISSUERS = (N_TOK <= 64 && STAGES >= 6) ? 2 : 1;
    constexpr int TMEM_NEED = ISSUERS * (N_TOK + 16);
    constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : 256);
    constexpr uint32_t TM_SF = ISSUERS * N_TOK, SF_BUF_COLS = 8;  // per issuer: two buffers of (4 W + 4 X) scale-factor columns

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                       // TMA bytes landed              (count 1 + tx)
    uint64_t* empty = bars + STAGES;             // MMAs of the stage retired     (count 1, tcgen05.commit)
    uint64_t* sf_full = bars + 2 * STAGES;       // scale factors in smem         (count 2: both loader warps)
    uint64_t* sf_empty = sf_full + SF_STAGES;    // MMAs using the SF stage retired (count 1, tcgen05.commit)
    uint64_t* tmem_full = sf_empty + SF_STAGES;  // accumulator complete          (count 1, tcgen05.commit)
    uint64_t* raw_w = tmem_full + 1;             // raw W-scale boxes landed      (count 1 + tx)
    uint64_t* raw_x = raw_w + RAW_W;             // raw token-scale boxes landed  (count 1 + tx)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { SK_TRACE(0) }
    const int S = p.splits;
    const int tile = blockIdx.x / S;
    const int split = (int)cluster_ctarank();  // == blockIdx.x % S: the cluster spans the K splits of one tile
    const int k_blocks_total = p.K / BLOCK_K;
    const int kb0 = (int)((int64_t)k_blocks_total * split / S), kb1 = (int)((int64_t)k_blocks_total * (split + 1) / S);
    const int k_blocks = kb1 - kb0;
    const int n0 = tile * p.tile_rows;  // (rows tile_rows .. 127 of the MMA are stale shared memory: their accumulator lanes are never stored)

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
    }
    if (warp == 1 && elect_one()) {
        // fused activation quantization: the four epilogue warps produce the X tile and its scales themselves and arrive
        // on the stage barriers in place of the X TMA bytes / the token-scale loader warp
        const uint32_t xq = p.x_hp != nullptr ? 4u : 0u;
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], xq ? 2 : 1);  // producer (+ tx bytes) and, in fused mode, the quantizer warp that owns the K block
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < SF_STAGES; ++i) {
            mbar_init(&sf_full[i], xq ? 1 + xq : 2);
            mbar_init(&sf_empty[i], ISSUERS);  // every issuer commits once per scale-factor stage
        }
        mbar_init(tmem_full, ISSUERS);
        for (int i = 0; i < RAW_W; ++i) mbar_init(&raw_w[i], 1);
        for (int i = 0; i < RAW_X; ++i) mbar_init(&raw_x[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 3) tmem_alloc<TMEM_COLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 0) { SK_TRACE(1) }
    pdl_launch_dependents();  // the next launch of the stream (if it opted in) may set itself up and start streaming ITS weights

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            // Programmatic dependent launch: this grid may be running before the preceding kernels of the stream have
            // finished.  A layer's pre-quantized weight does not depend on them (the caller says so: w_static), so the first
            // ring of W tiles is requested right away; everything that reads X (and, transitively, every write of D) comes
            // after pdl_wait().  Without that promise -- F.linear(to_mx(x), to_mx(w)): the quantize kernel that is still
            // writing W lets its dependents start early -- nothing is read before the wait.
            const int n_pre = (p.pdl && p.w_static) ? min(k_blocks, STAGES) : 0;
            const bool xq = p.x_hp != nullptr;
            const uint32_t tx = p.tx_w + (xq ? 0u : p.tx_x);
            for (int i = 0; i < n_pre; ++i) {
                mbar_arrive_expect_tx(&full[i], tx);
                tma_load_3d(&map_w, &full[i], smem + L::OFF_W + i * L::W_STAGE, (kb0 + i) * BLOCK_K, n0, 0);
            }
            if (p.pdl) pdl_wait();
            SK_TRACE(2)
            for (int i = 0; i < n_pre; ++i) {
                if (!xq) tma_load_3d(&map_x, &full[i], smem + L::OFF_X + i * L::X_STAGE, (kb0 + i) * BLOCK_K, 0, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            // The TMA unit keeps only a few tens of KB of loads in flight per SM, which at DRAM latency is ~1/3 of the HBM
            // rate; L2 prefetches are fire-and-forget, so W is pulled DRAM -> L2 `pf_dist` K blocks ahead of the ring and
            // the ring's own loads become L2 hits.
            const int pf_dist = p.pf_dist;
            for (int kb = kb0; kb < kb1 && kb < kb0 + pf_dist; ++kb) tma_prefetch_l2_3d(&map_w, kb * BLOCK_K, n0, 0);
            for (int kb = kb0 + n_pre; kb < kb1; ++kb) {
                if (kb + pf_dist < kb1) tma_prefetch_l2_3d(&map_w, (kb + pf_dist) * BLOCK_K, n0, 0);
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], tx);
                tma_load_3d(&map_w, &full[stage], smem + L::OFF_W + stage * L::W_STAGE, kb * BLOCK_K, n0, 0);
                if (!xq) tma_load_3d(&map_x, &full[stage], smem + L::OFF_X + stage * L::X_STAGE, kb * BLOCK_K, 0, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 8) {
        // ================= MMA issuers (the whole warp runs the loop, one elected lane issues) =================
        // issuer `me` takes K blocks me, me + ISSUERS, ... (slots me, me + ISSUERS of every 4-K-block scale-factor stage) into
        // accumulator `me`; with one issuer (N_TOK = 128) warp 8 only reports on the barriers
        const int me = warp == 1 ? 0 : 1;
        const bool active = me < ISSUERS;
        const uint32_t idesc = make_idesc(TILE_W, N_TOK) | p.idesc_fmt;
        constexpr uint64_t HI_OPERAND = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutSw128 << 61);
        constexpr uint64_t HI_SF = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutNone << 61);
        const uint32_t w_lo0 = smem_u32(smem + L::OFF_W) >> 4, x_lo0 = smem_u32(smem + L::OFF_X) >> 4;
        const uint32_t sfw_lo0 = smem_u32(smem + L::OFF_SFW) >> 4, sfx_lo0 = smem_u32(smem + L::OFF_SFX) >> 4;
        const uint32_t tm_acc = tmem_base + (uint32_t)(me * N_TOK);
        const uint32_t tm_sf0 = tmem_base + TM_SF + (uint32_t)(me * 2) * SF_BUF_COLS;
        const int n_groups = (k_blocks + SF_KB - 1) / SF_KB;
        uint32_t sfs = 0, sf_phase = 0, sf_sel = 0;
        bool first = true;
        if (active) {
            for (int g = 0; g < n_groups; ++g) {
                mbar_wait(&sf_full[sfs], sf_phase);
#pragma unroll
                for (int j = me; j < SF_KB; j += ISSUERS) {
                    const int kb = g * SF_KB + j;
                    if (kb < k_blocks) {
                        const uint32_t stage = (uint32_t)(kb % STAGES), phase = (uint32_t)((kb / STAGES) & 1);
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t w_lo = w_lo0 + stage * (L::W_STAGE >> 4), x_lo = x_lo0 + stage * (L::X_STAGE >> 4);
                            const uint32_t sf_off = sfs * (L::SF_STAGE >> 4) + j * (SF_KB_BYTES >> 4);
                            const uint32_t tm_sfw = tm_sf0 + sf_sel * SF_BUF_COLS, tm_sfx = tm_sfw + 4;
                            if (!SK_EXP(8)) {
                            tc_copy_sf(tm_sfw, HI_SF | (sfw_lo0 + sf_off));
                            tc_copy_sf(tm_sfx, HI_SF | (sfx_lo0 + sf_off));
                            }
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                if (!SK_EXP(8))
                                tc_mma_mx(tm_acc, HI_OPERAND | (w_lo + k * (UMMA_K >> 4)), HI_OPERAND | (x_lo + k * (UMMA_K >> 4)), idesc_with_sf(idesc, k, k),
                                          !(first && k == 0), tm_sfw, tm_sfx);
                            if (!BATCH_COMMITS) tc_commit(&empty[stage]);
                        }
                        __syncwarp();
                        if (first && me == 0 && lane == 0) { SK_TRACE(3) }
                        first = false;
                        sf_sel ^= 1;
                    }
                }
                if (elect_one()) {
                    if (BATCH_COMMITS) {
#pragma unroll
                        for (int j = me; j < SF_KB; j += ISSUERS) {
                            const int kb = g * SF_KB + j;
                            if (kb < k_blocks) { if (SK_EXP(16)) mbar_arrive(&empty[kb % STAGES]); else tc_commit(&empty[kb % STAGES]); }
                        }
                    }
                    tc_commit(&sf_empty[sfs]);
                }
                __syncwarp();
                if (++sfs == SF_STAGES) { sfs = 0; sf_phase ^= 1; }
            }
            if (me == 0 && lane == 0) { SK_TRACE(4) }
            if (elect_one()) tc_commit(tmem_full);  // (an issuer without a K block of its own -- k_blocks == 1 -- arrives at once)
            __syncwarp();
        }
    } else if (warp == 2 || warp == 3) {
        // ================= scale-factor loaders (warp 2: the 128 W rows, warp 3: the token rows) =================
        uint32_t sfs = 0, sf_phase = 0;
        auto arrive = [&](uint32_t st) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sf_full[st]);
        };
        if (warp == 3 && p.x_hp != nullptr) {
            // fused mode: the token scales are produced by the quantizer (epilogue) warps
        } else if (p.pdl && (warp == 3 || !p.w_static)) {
            pdl_wait();  // the token scales (and, unless static, the weight scales) are written by a preceding kernel
        }
        if (warp == 3 && p.x_hp != nullptr) {
        } else if (p.sf_tma) {
            if (warp == 2)
                sf_tma_tile4<RAW_W>(&map_sfw, kb0 * 4, n0, k_blocks, smem + L::OFF_RAW_W, raw_w, smem + L::OFF_SFW, sf_empty, sfs, sf_phase, lane, arrive);
            else
                sf_tma_tile4<RAW_X>(&map_sfx, kb0 * 4, 0, k_blocks, smem + L::OFF_RAW_X, raw_x, smem + L::OFF_SFX, sf_empty, sfs, sf_phase, lane, arrive);
        } else if (warp == 2) {
            sf_load_tile4<1>(p.sfw + (int64_t)kb0 * 4, p.ld_sfw, n0, p.N, k_blocks, smem + L::OFF_SFW, SF_KB_BYTES, sf_empty, sfs, sf_phase, lane, arrive);
        } else {
            sf_load_tile4<1>(p.sfx + (int64_t)kb0 * 4, p.ld_sfx, 0, p.M, k_blocks, smem + L::OFF_SFX, SF_KB_BYTES, sf_empty, sfs, sf_phase, lane, arrive);
        }
    } else if (warp < 8) {
        if constexpr (N_TOK <= 64) if (p.x_hp != nullptr) {
            // ================= fused activation quantization (K1 arithmetic, one thread per MX block) =================
            // Quantizer warp w owns K blocks w, w+4, ... (= slot w of every 4-K-block scale-factor stage), so four K blocks
            // are in flight at once and each warp can afford to wait for its own L2 round trip.  Within a K block lane l
            // takes sub-block l & 3 of token rows (l >> 2) + 8 i: four lanes read one row's 256 contiguous bytes.  Codes go
            // straight into the 128B-swizzled K-major X tile the MMA reads (16-byte chunk c of row t at chunk c ^ (t & 7)),
            // the scale byte into the tcgen05.cp chunk layout.
            const int qw = warp - 4;
            const bool hw_exact = (p.x_flags & MXQ_FLAG_HW_EXACT) != 0;
            const int j = lane & 3;
            if (p.pdl) pdl_wait();  // the activation is written by the preceding kernel of the stream
            const int n_groups = (k_blocks + SF_KB - 1) / SF_KB;
            uint32_t sfs = 0, sf_phase = 0;
            for (int g = 0; g < n_groups; ++g) {
                const int kb = g * SF_KB + qw;
                const bool live = kb < k_blocks;
                const uint32_t stage = (uint32_t)(kb % STAGES), par = (uint32_t)((kb / STAGES) & 1);
                uint8_t* xt = smem + L::OFF_X + stage * L::X_STAGE;
                uint8_t* sfx = smem + L::OFF_SFX + sfs * L::SF_STAGE + qw * SF_KB_BYTES;
                mbar_wait(&sf_empty[sfs], sf_phase ^ 1);
#pragma unroll 1
                for (int half = 0; half < N_TOK / 32; ++half) {
                    uint32_t w[4][16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int t = half * 32 + (lane >> 2) + 8 * i;
                        if (live && t < p.M) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.x_hp + (int64_t)t * p.ldx + (int64_t)(kb0 + kb) * BLOCK_K + j * 32);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 v = src[q];
                                w[i][4 * q] = v.x; w[i][4 * q + 1] = v.y; w[i][4 * q + 2] = v.z; w[i][4 * q + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 16; ++q) w[i][q] = 0;
                        }
                    }
                    if (live && half == 0) mbar_wait(&empty[stage], par ^ 1);  // the loads above are already in flight
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int t = half * 32 + (lane >> 2) + 8 * i;
                            uint32_t out[8];
                            const int sc = quantize_block32<MXQ_ELEM_E4M3>(w[i], hw_exact, out);
                            *reinterpret_cast<uint4*>(xt + t * 128 + (((2 * j) ^ (t & 7)) << 4)) = make_uint4(out[0], out[1], out[2], out[3]);
                            *reinterpret_cast<uint4*>(xt + t * 128 + (((2 * j + 1) ^ (t & 7)) << 4)) = make_uint4(out[4], out[5], out[6], out[7]);
                            sfx[(t & 31) * 16 + (t >> 5) * 4 + j] = (uint8_t)sc;
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (live) mbar_arrive(&full[stage]);
                    mbar_arrive(&sf_full[sfs]);  // every quantizer warp reports for every group, with or without a K block in it
                }
                if (++sfs == SF_STAGES) { sfs = 0; sf_phase ^= 1; }
            }
        }
        // ================= epilogue, part 1: accumulator -> global (S == 1) or -> partial-sum buffer (S > 1) =================
        const int quad = warp & 3;
        const int r = quad * 32 + lane;  // accumulator lane == weight row within the tile
        const int n = n0 + r;
        const bool row_ok = r < p.tile_rows && n < p.N;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (threadIdx.x == 128) { SK_TRACE(5) }
        const bool two_acc = ISSUERS == 2 && k_blocks > 1;  // accumulator 1 holds the odd K blocks (never written when there is only one)
        float* part = reinterpret_cast<float*>(smem + L::OFF_W);
        const float bias = (S == 1 && p.bias != nullptr && row_ok) ? __uint_as_float((uint32_t)p.bias[n] << 16) : 0.0f;
#pragma unroll 1
        for (int c = 0; c < N_TOK / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + c * 32, v);
            if (two_acc) {
                uint32_t v1[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + N_TOK + c * 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v1[i]));
            } else {
                tmem_ld_wait();
            }
            if (S == 1) {
                if (p.d_mc != nullptr) {
                    // fused tensor-parallel all-reduce: adjacent lanes (weight rows n, n+1) pair up into one bf16x2 add
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int t = c * 32 + i;
                        const float mine = __uint_as_float(v[i]) + bias;
                        const float next = __shfl_down_sync(0xFFFFFFFFu, mine, 1);
                        if (!(lane & 1) && t < p.M && row_ok) multimem_red_add_bf16x2(p.d_mc + (int64_t)t * p.ldd + n, pack_bf16x2(mine, next));
                    }
                } else if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int t = c * 32 + i;
                        if (t < p.M) p.d[(int64_t)t * p.ldd + n] = (uint16_t)pack_bf16x2(__uint_as_float(v[i]) + bias, 0.0f);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) part[(c * 32 + i) * TILE_W + r] = __uint_as_float(v[i]);
            }
        }
        tc_fence_before();
        if constexpr (L::PUSH) {
            if (S > 1) {
                // push: the staged partials (token-major, this CTA's own W ring) go out as 16-byte remote stores -- posted, no
                // round trip -- to the CTA that finishes the token (token t belongs to split t % S)
                asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps: staging complete
                const int e = quad * 32 + lane;
                const int slots = (N_TOK + S - 1) / S;
                const uint32_t recv_addr = smem_u32(smem + L::OFF_RECV);
                for (int item = e; item < p.M * (TILE_W / 4); item += 128) {
                    const int t = item / (TILE_W / 4), r0 = (item % (TILE_W / 4)) * 4;
                    const float4 v = *reinterpret_cast<const float4*>(part + t * TILE_W + r0);
                    st_dsmem_v4(mapa_shared(recv_addr + (uint32_t)(((split * slots + t / S) * TILE_W + r0) * 4), (uint32_t)(t % S)), v);
                }
            }
        }
        if (threadIdx.x == 128) { SK_TRACE(6) }
    }

    if (S > 1) {
        // ================= epilogue, part 2: fixed-order reduction of the S partials through DSMEM =================
        __syncwarp();
        cluster_sync_all();  // every CTA's partials are written (and every MMA has retired, so the W ring was free to reuse)
        MXQ_DEV_ONLY(if (threadIdx.x == 128 && SK_EXP(32)) { SK_TRACE(1) })
        if (warp >= 4 && warp < 8) {
            // This CTA finishes tokens split, split + S, ...  One work item = one token x four consecutive weight rows: the S
            // partials -- already here when the variant pushes; otherwise S 16-byte DSMEM loads, all in flight together (a remote
            // load costs ~40 cycles of issue bandwidth whatever its width, plus one round trip) -- summed in split order, one
            // 8-byte store.
            const int e = (warp & 3) * 32 + lane;
            const uint32_t part_addr = smem_u32(smem + L::OFF_W);
            const int n_mine = split < p.M ? (p.M - split + S - 1) / S : 0;
            const bool vec_store = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.d) & 7) == 0);
            const int n_lim = min(p.N, n0 + p.tile_rows);  // (tile_rows is a multiple of 8: a group of four rows is inside or outside)
            for (int item = e; item < n_mine * (TILE_W / 4); item += 128) {
                const int t = split + (item / (TILE_W / 4)) * S, r0 = (item % (TILE_W / 4)) * 4, n = n0 + r0;
                float4 v[8];
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (s < S) {
                        if constexpr (L::PUSH) v[s] = *reinterpret_cast<const float4*>(smem + L::OFF_RECV + ((s * ((N_TOK + S - 1) / S) + item / (TILE_W / 4)) * TILE_W + r0) * 4);
                        else v[s] = ld_dsmem_v4(mapa_shared(part_addr + (uint32_t)(t * TILE_W + r0) * 4u, (uint32_t)s));
                    }
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (s < S) { acc[0] += v[s].x; acc[1] += v[s].y; acc[2] += v[s].z; acc[3] += v[s].w; }
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n + i < n_lim) acc[i] += __uint_as_float((uint32_t)p.bias[n + i] << 16);
                }
                const uint32_t lo = pack_bf16x2(acc[0], acc[1]), hi = pack_bf16x2(acc[2], acc[3]);
                if (p.d_mc != nullptr) {
                    if (n + 1 < n_lim) multimem_red_add_bf16x2(p.d_mc + (int64_t)t * p.ldd + n, lo);
                    if (n + 3 < n_lim) multimem_red_add_bf16x2(p.d_mc + (int64_t)t * p.ldd + n + 2, hi);
                } else if (vec_store && n + 3 < n_lim) {
                    *reinterpret_cast<uint2*>(p.d + (int64_t)t * p.ldd + n) = make_uint2(lo, hi);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n + i < n_lim) p.d[(int64_t)t * p.ldd + n + i] = (uint16_t)((i < 2 ? lo : hi) >> ((i & 1) * 16));
                }
            }
        }
        MXQ_DEV_ONLY(if (threadIdx.x == 128 && SK_EXP(32)) { SK_TRACE(2) })
        __syncwarp();
        if constexpr (L::PUSH) __syncthreads();  // (after the barrier above nobody touches a peer's shared memory any more)
        else cluster_sync_all();                // nobody exits while a peer may still read its partials
        MXQ_DEV_ONLY(if (threadIdx.x == 128 && SK_EXP(32)) { SK_TRACE(3) })
    } else {
        __syncwarp();
        __syncthreads();
    }
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
        if (lane == 0) { SK_TRACE(7) }
    }
}

template <int N_TOK, int STAGES, int RAW_W>
static int launch(const mxq_gemm_args_t* a, int splits, int tile_rows, int device, cudaStream_t stream, char* msg, size_t msg_len) {
    using L = Smem<N_TOK, STAGES, RAW_W>;
    CUtensorMap mw, mx;
    const bool w_ok = cached_operand_map(&mw, a->b_codes, a->K, a->N, 1, a->ldb, 0, tile_rows, a->b_format, device);
    const bool xq = a->x_bf16 != nullptr;
    if (xq) mx = mw;  // unused placeholders in fused-quantization mode
    if (!w_ok || (!xq && !cached_operand_map(&mx, a->a_codes, a->K, a->M, 1, a->lda, 0, N_TOK, a->a_format, device))) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled failed (driver entry point missing or invalid strides)");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    // scales by TMA when their layout allows it (every split then starts on a 16-byte boundary: splits are 4-K-block aligned)
    CUtensorMap msw = mw, msx = mx;
    const int k_blocks_total = (int)(a->K / BLOCK_K);
    int sf_tma = (xq || (((uintptr_t)a->sfa % 16 == 0) && (a->ld_sfa % 16 == 0))) && ((uintptr_t)a->sfb % 16 == 0) && (a->ld_sfb % 16 == 0) &&
                 (k_blocks_total % (splits * SF_KB) == 0);
    MXQ_DEV_ONLY(if (dev_env("MXQ_SKINNY_NO_SFTMA")) sf_tma = 0;)
    if (sf_tma && (!cached_scale_map(&msw, a->sfb, a->K / 32, a->N, a->ld_sfb, device) || (!xq && !cached_scale_map(&msx, a->sfa, a->K / 32, a->M, a->ld_sfa, device)))) sf_tma = 0;
    auto kernel = mx_gemm_skinny_kernel<N_TOK, STAGES, RAW_W>;
    cudaError_t e = ensure_smem_attr((const void*)kernel, L::DYN_BYTES, device);
    if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    Params p;
    p.sfx = a->sfa; p.sfw = a->sfb; p.bias = (const uint16_t*)a->bias; p.d = (uint16_t*)a->d;
    p.d_mc = (uint16_t*)a->d_multicast;
    p.x_hp = (const uint16_t*)a->x_bf16; p.ldx = a->ldx; p.x_flags = a->x_quant_flags;
    p.ld_sfx = a->ld_sfa; p.ld_sfw = a->ld_sfb; p.ldd = a->ldd;
    p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K; p.splits = splits;
    // the weights are the MMA A operand here, the tokens the B operand
    p.idesc_fmt = idesc_formats(a->b_format, xq ? MXQ_OPERAND_E4M3_BYTES : a->a_format);
    p.tx_w = tile_rows * BLOCK_K * operand_bits(a->b_format) / 8;
    p.tile_rows = tile_rows;
    p.tx_x = N_TOK * BLOCK_K * operand_bits(a->a_format) / 8;
    p.sf_tma = sf_tma;
    p.pdl = (a->flags & MXQ_GEMM_NO_PDL) ? 0 : 1;
    p.w_static = (a->flags & MXQ_GEMM_B_STATIC) ? 1 : 0;
    p.pf_dist = 0;  // measured: no gain on B200
    MXQ_DEV_ONLY(p.pf_dist = dev_env("MXQ_SKINNY_PF");)
    p.trace = nullptr;
    p.exp = 0;
    MXQ_DEV_ONLY(p.exp = dev_env("MXQ_SKINNY_EXP");)
    MXQ_DEV_ONLY({ static int launch_no = 0; const char* tp = getenv("MXQ_SKINNY_TRACE"); if (tp) p.trace = reinterpret_cast<long long*>(strtoull(tp, nullptr, 0)) + (size_t)(launch_no++ % 16) * 16384; })
    const int n_tiles = (int)((a->N + tile_rows - 1) / tile_rows);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_tiles * splits), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = L::DYN_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)splits;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, kernel, mw, mx, msw, msx, p);
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch (skinny): %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    return MXQ_OK;
}

}  // namespace skinny

// M <= 128, batch == 1.  Returns MXQ_ERR_UNSUPPORTED_SHAPE when the caller should use the general kernels.
int launch_gemm_skinny(const mxq_gemm_args_t* a, int sm_count, int device, cudaStream_t stream, char* msg, size_t msg_len) {
    using namespace skinny;
    if (a->batch != 1 || a->M > 128 || a->K % BLOCK_K) return MXQ_ERR_UNSUPPORTED_SHAPE;
    if (a->x_bf16 != nullptr && (a->M > 64 || ((uintptr_t)a->x_bf16 % 16) || (a->ldx % 8))) {
        snprintf(msg, msg_len, "fused activation quantization needs M <= 64 and 16-byte aligned activation rows");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    if (a->d_multicast != nullptr && (((uintptr_t)a->d_multicast % 4) || (a->ldd % 2) || (a->N % 2))) {
        snprintf(msg, msg_len, "d_multicast needs an even N / ldd and a 4-byte aligned buffer");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    const int k_blocks = (int)(a->K / BLOCK_K);
    if (a->split_k > 8) {
        snprintf(msg, msg_len, "split_k=%d exceeds the portable cluster size 8", a->split_k);
        return MXQ_ERR_INVALID;
    }
    // K splits: K is split across a cluster until about one CTA per SM is streaming (>= 4 K blocks per CTA, 4-K-block aligned
    // split points: the scale-factor loaders read 16 bytes = 4 K blocks per row).
    // Tile height stays 128 rows.  Shorter tiles that spread evenly over the SMs (14336 rows = 112 tiles of 128 on 148 SMs, or
    // 138 tiles of 104) were measured SLOWER: what a CTA pays per K block (~0.27 us per SM, see profiles/r2_skinny_issue_loop.txt)
    // does not shrink with the rows in it, so more tiles just means more K-block iterations (28672 x 4096: 26 -> 31 us), and
    // clusters of 4 fit at most ~36 times on the chip, so a 37th cluster is a second wave (4096 x 4096: 6.9 -> 11.2 us).
    // The kernel keeps the tile height a parameter (developer builds: MXQ_SKINNY_ROWS).
    int best_rows = TILE_W, best_splits = 1;
    {
        const int n_tiles = (int)((a->N + TILE_W - 1) / TILE_W);
        int splits = 1;
        if (a->split_k > 0) splits = a->split_k < k_blocks ? a->split_k : k_blocks;
        else
            while (splits < 8 && n_tiles * splits * 2 <= sm_count && k_blocks % (splits * 2 * 4) == 0 && k_blocks / (splits * 2) >= 4) splits *= 2;
        best_splits = splits;
    }
    MXQ_DEV_ONLY(if (dev_env("MXQ_SKINNY_ROWS")) { best_rows = dev_env("MXQ_SKINNY_ROWS"); })
    const int splits = best_splits, rows = best_rows;
    const int n_tiles = (int)((a->N + rows - 1) / rows);
    const bool deep = (int64_t)n_tiles * splits <= sm_count;  // one CTA per SM: spend the shared memory on a deeper ring
    if (a->M <= 32) return deep ? launch<32, 8, 8>(a, splits, rows, device, stream, msg, msg_len) : launch<32, 4, 4>(a, splits, rows, device, stream, msg, msg_len);
    if (a->M <= 64) return deep ? launch<64, 6, 8>(a, splits, rows, device, stream, msg, msg_len) : launch<64, 4, 4>(a, splits, rows, device, stream, msg, msg_len);
    return deep ? launch<128, 6, 4>(a, splits, rows, device, stream, msg, msg_len) : launch<128, 4, 4>(a, splits, rows, device, stream, msg, msg_len);
}

}  // namespace gemm
}  // namespace mxq
