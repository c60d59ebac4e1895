// K3: block-scaled MX GEMM on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   D[b][m][n] = sum_k ( A[b][m][k] * 2^(sfa[b][m][k/32]-127) ) * ( B[b][n][k] * 2^(sfb[b][n][k/32]-127) ) (+ bias[n])
//
// replaces the reference's dequantize -> bf16 aten op recipe (torchmx/ops.py:29-41, 60-68, 99-119):
// the element codes stay 1 byte (E4M3 container), the E8M0 scales are applied by the MMA itself
// (tcgen05.mma.kind::mxf8f6f4.block_scale, scale factors in TMEM), accumulation is fp32 in TMEM.
//
// Kernel anatomy (one CTA per SM, persistent over 128 x BLOCK_N output tiles, 8 warps):
//   warp 0      TMA producer: A [128 x 128 B] and B [BLOCK_N x 128 B] K-major tiles, 128B swizzle,
//               STAGES-deep mbarrier ring
//   warp 1      MMA issuer (one elected lane): per 128-wide K block, tcgen05.cp the scale factors
//               smem -> TMEM, then four K=32 MMAs; tcgen05.commit releases the stage / publishes
//               the accumulator
//   warp 2      scale-factor loader: reads the reference-layout scales ([rows, K/32] uint8, row-major)
//               straight from global and writes them into the 32x16B layout tcgen05.cp expects
//               (row r of a 128-row group -> 16-byte chunk r%32, word r/32), two K blocks ahead
//   warp 3      TMEM allocator
//   warps 4..7  epilogue: tcgen05.ld the fp32 accumulator (lane = row), + bias, -> bf16, 64-byte row
//               segments to global
#include "mxq_tc.cuh"

namespace mxq {
namespace gemm {

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int A_STAGE = BLOCK_M * BLOCK_K;           // 16 KB
    static constexpr int B_STAGE = BLOCK_N * BLOCK_K;           // 16 / 32 KB
    static constexpr int SFA_STAGE = 512;                       // 128 rows x 4 bytes
    static constexpr int SFB_STAGE = (BLOCK_N / 128) * 512;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
    static constexpr int OFF_SFA = OFF_B + STAGES * B_STAGE;
    static constexpr int OFF_SFB = OFF_SFA + STAGES * SFA_STAGE;
    static constexpr int OFF_BAR = OFF_SFB + STAGES * SFB_STAGE;
    static constexpr int NUM_BARS = 3 * STAGES + 2;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for the manual 1024-byte alignment
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) mx_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                              const Params p) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    constexpr int SF_COLS_A = 4, SF_COLS_B = BLOCK_N / 32;
    // two scale-factor buffers in TMEM: the tcgen05.cp of K block k+1 must not wait for the MMAs of K block k to finish
    // reading theirs (measured on 1024 x 1024 x 8192, one tile per CTA: 296 -> 241 ns per K block)
    constexpr int SF_BUF = SF_COLS_A + SF_COLS_B;
    constexpr int TMEM_COLS = (BLOCK_N + 2 * SF_BUF) <= 256 ? 256 : 512;
    constexpr uint32_t TM_SFA = BLOCK_N, TM_SFB = BLOCK_N + SF_COLS_A;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                    // TMA bytes landed          (count 1 + tx)
    uint64_t* sf_full = bars + STAGES;        // scale factors in smem     (count 64: both loader warps)
    uint64_t* empty = bars + 2 * STAGES;      // MMAs of the stage retired (count 1, tcgen05.commit)
    uint64_t* tmem_full = bars + 3 * STAGES;  // accumulator complete      (count 1, tcgen05.commit)
    uint64_t* tmem_empty = tmem_full + 1;     // accumulator drained       (count 128)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_blocks = p.K / BLOCK_K;
    const int tiles_per_batch = p.m_blocks * p.n_blocks;
    const int num_tiles = tiles_per_batch * p.batch;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&sf_full[i], 64);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kEpilogueThreads);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<TMEM_COLS>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto tile_coords = [&](int tile, int& b, int& mb, int& nb) {
        b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        nb = t / p.m_blocks;  // m fastest: concurrently running CTAs share one B panel
        mb = t - nb * p.m_blocks;
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int b, mb, nb;
                tile_coords(tile, b, mb, nb);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], p.tx_a + (BLOCK_N / 128) * p.tx_b);
                    tma_load_3d(&map_a, &full[stage], smem + L::OFF_A + stage * L::A_STAGE, kb * BLOCK_K, mb * BLOCK_M, b);
                    tma_load_3d(&map_b, &full[stage], smem + L::OFF_B + stage * L::B_STAGE, kb * BLOCK_K, nb * BLOCK_N, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N) | p.idesc_fmt;
        uint32_t stage = 0, phase = 0, acc_phase = 0, sf_sel = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tmem_empty, acc_phase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < k_blocks; ++kb, sf_sel ^= 1) {
                mbar_wait(&full[stage], phase);
                mbar_wait(&sf_full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t tm_sfa = tmem_base + TM_SFA + sf_sel * SF_BUF, tm_sfb = tmem_base + TM_SFB + sf_sel * SF_BUF;
                    const uint32_t a_addr = smem_u32(smem + L::OFF_A + stage * L::A_STAGE);
                    const uint32_t b_addr = smem_u32(smem + L::OFF_B + stage * L::B_STAGE);
                    const uint32_t sfa_addr = smem_u32(smem + L::OFF_SFA + stage * L::SFA_STAGE);
                    const uint32_t sfb_addr = smem_u32(smem + L::OFF_SFB + stage * L::SFB_STAGE);
                    tc_copy_sf(tm_sfa, smem_desc(sfa_addr, 128, kLayoutNone));
#pragma unroll
                    for (int i = 0; i < BLOCK_N / 128; ++i) tc_copy_sf(tm_sfb + 4 * i, smem_desc(sfb_addr + 512 * i, 128, kLayoutNone));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major SW128 tile: 8-row groups are 1024 B apart; advancing K inside the swizzle row = +32 B
                        const uint64_t da = smem_desc(a_addr + k * UMMA_K, 1024, kLayoutSw128);
                        const uint64_t db = smem_desc(b_addr + k * UMMA_K, 1024, kLayoutSw128);
                        tc_mma_mx(tmem_base, da, db, idesc_with_sf(idesc, k, k), (kb | k) != 0, tm_sfa, tm_sfb);
                    }
                    tc_commit(&empty[stage]);
                    if (kb == k_blocks - 1) tc_commit(tmem_full);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc_phase ^= 1;
        }
    } else if (warp == 2 || warp == 3) {
        // ================= scale-factor loaders (warp 2: A rows, warp 3: B rows) =================
        const bool is_a = warp == 2;
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int b, mb, nb;
            tile_coords(tile, b, mb, nb);
            auto arrive = [&](uint32_t st) {
                fence_proxy_async_smem();
                mbar_arrive(&sf_full[st]);
            };
            if (is_a)
                sf_load_tile<1, STAGES>(p.sfa + (int64_t)b * p.sfa_batch, p.ld_sfa, mb * BLOCK_M, p.M, k_blocks, smem + L::OFF_SFA, L::SFA_STAGE, empty,
                                        stage, phase, lane, arrive);
            else
                sf_load_tile<BLOCK_N / 128, STAGES>(p.sfb + (int64_t)b * p.sfb_batch, p.ld_sfb, nb * BLOCK_N, p.N, k_blocks, smem + L::OFF_SFB,
                                                    L::SFB_STAGE, empty, stage, phase, lane, arrive);
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int b, mb, nb;
            tile_coords(tile, b, mb, nb);
            mbar_wait(tmem_full, acc_phase);
            tc_fence_after();
            const int row = mb * BLOCK_M + quad * 32 + lane;
            uint16_t* drow = p.d + (int64_t)b * p.d_batch + (int64_t)row * p.ldd;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + c * 32, v);
                tmem_ld_wait();
                const int col0 = nb * BLOCK_N + c * 32;
                if (row < p.M && col0 < p.N) {
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
            mbar_arrive(tmem_empty);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ---- K3b: CTA-pair variant (cta_group::2) -------------------------------------------------------------------
// Two CTAs of a cluster (same TPC) compute one 256 x 256 output tile: each CTA stages its own 128 rows of A
// and 128 of the 256 B rows, the leader's single thread issues M=256 MMAs that read both CTAs' shared memory,
// and each CTA ends up with its 128 x 256 half of the accumulator in its own TMEM.  Per CTA and MMA this
// halves the B bytes written by TMA and read by the tensor core (24 KB -> 16 KB of shared-memory traffic per
// 128-cycle MMA), which is what bounds the single-CTA kernel with 1-byte operands.
//
// TMEM (512 columns): two accumulator slots at columns 0 and 192 that overlap in [192,256).  The epilogue drains
// the overlapping 64 columns first and then hands the slot back, so the MMAs of the next tile (other slot) run
// under the rest of the epilogue.  Scale factors: two 12-column buffers (SFA 4 + SFB 8) from column 448.
//
// Two rings: operand stages (one 128-wide K block each, TMA) and scale-factor stages (FOUR K blocks each -- the
// loader warps read 16 B = four K blocks per row anyway), so the issuing thread and the loaders synchronise on
// scale factors once per 2048 MMA cycles instead of once per 512.
//
// Barriers: full / sf_full / tmem_empty live in the leader CTA (TMA complete_tx and the loader / epilogue warps of
// both CTAs arrive there); empty / sf_empty / tmem_full exist in both CTAs and are signalled by multicast
// tcgen05.commit.
namespace pair {
constexpr int TILE_M = 256, TILE_N = 256;
constexpr uint32_t ACC_SLOT1 = 192;
constexpr uint32_t TM_SF = 448, SF_BUF_COLS = 12;

// EW = epilogue warps: 4 (one per TMEM lane quadrant; large K, the mainloop hides the epilogue) or 8 (quadrant x column half;
// small K, where the single-warp-per-SMSP latency chain TMEM -> bf16 -> staging of a 64 KB output tile bounds the kernel)
// NBUF = staging buffers per epilogue warp: a TMA store takes ~1.4 k cycles from issue until it has read its 4 KB of shared
// memory; with 2 buffers a warp waits ~600 cycles per 64-column group for its second-to-last store (measured by clock64 stamps)
template <int STAGES, int EW, int NBUF>
struct Smem {
    static constexpr int A_STAGE = 128 * BLOCK_K;  // 16 KB
    static constexpr int B_STAGE = 128 * BLOCK_K;  // 16 KB: this CTA's half of the 256 B rows
    static constexpr int SFA_KB = 512, SFB_KB = 1024;  // per K block: own 128 A rows / all 256 B rows
    static constexpr int SFA_STAGE = SF_KB * SFA_KB, SFB_STAGE = SF_KB * SFB_KB;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_STAGE;
    static constexpr int OFF_SFA = OFF_B + STAGES * B_STAGE;
    static constexpr int OFF_SFB = OFF_SFA + SF_STAGES * SFA_STAGE;
    static constexpr int EPI_WARP = NBUF * 4096;  // per epilogue warp: NBUF buffers of 32 rows x 64 bf16 columns (128B-swizzled rows)
    static constexpr int OFF_EPI = OFF_SFB + SF_STAGES * SFB_STAGE;
    static constexpr int OFF_BAR = OFF_EPI + EW * EPI_WARP;
    static constexpr int NUM_BARS = 2 * STAGES + 2 * SF_STAGES + 2;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM_PTR + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int STAGES, int EW, int NBUF, bool MXF4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
    mx_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_d,
                        const Params p, const int group_m, const int tma_store) {
    using L = Smem<STAGES, EW, NBUF>;
    extern __shared__ uint8_t smem_raw[];
    // the dynamic shared window starts at the same offset in both CTAs, so the aligned carve-up matches too
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                              // leader: both CTAs' TMA bytes landed  (count 1 + tx)
    uint64_t* empty = bars + STAGES;                    // both:   MMAs of the stage retired    (count 1, multicast commit)
    uint64_t* sf_full = bars + 2 * STAGES;              // leader: scale factors in both smems  (count 4: two loader warps x two CTAs)
    uint64_t* sf_empty = sf_full + SF_STAGES;           // both:   MMAs using the SF stage retired (count 1, multicast commit)
    uint64_t* tmem_full = sf_empty + SF_STAGES;         // both:   accumulator complete         (count 1, multicast commit)
    uint64_t* tmem_empty = tmem_full + 1;               // leader: accumulator slot reusable    (count 8: epilogue warps of both CTAs)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int k_blocks = p.K / BLOCK_K;  // 128-element K blocks = one 32-bit scale word per row
    // operand stages: one 128-byte swizzle row per matrix row = 128 one-byte (or expanded 6 / 4-bit) elements, or, for the
    // dense 4-bit streams of kind::mxf4, 256 elements = two K blocks
    constexpr int KBS = MXF4 ? 2 : 1;
    const int k_stages = k_blocks / KBS;
    const int tiles_per_batch = p.m_blocks * p.n_blocks;
    const int num_tiles = tiles_per_batch * p.batch;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (tma_store) tma_prefetch_desc(&map_d);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < SF_STAGES; ++i) {
            mbar_init(&sf_full[i], 4);
            mbar_init(&sf_empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * EW);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 3) tmem_alloc_pair<512>(tmem_ptr);
    tc_fence_before();
    cluster_sync_all();  // the peer's barriers must be initialised before anything is posted on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    MXQ_DEV_ONLY(if (p.trace != nullptr && leader && threadIdx.x == 128) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[640 + pair_id] = (long long)t;
    })

    // grouped rasterisation: group_m row-blocks x all column-blocks at a time, row-block fastest, so one wave of
    // pairs touches ~group_m A panels and ~num_pairs/group_m B panels (both stay in L2) instead of every A panel
    auto tile_coords = [&](int tile, int& b, int& mb, int& nb) {
        b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        const int per_group = group_m * p.n_blocks;
        const int g = t / per_group;
        const int first_m = g * group_m;
        const int gm = min(group_m, p.m_blocks - first_m);
        const int r = t - g * per_group;
        nb = r / gm;
        mb = first_m + (r - nb * gm);
    };

    if (warp == 0) {
        // ================= TMA producer (both CTAs; completion bytes are posted on the leader's barrier) =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                int b, mb, nb;
                tile_coords(tile, b, mb, nb);
                for (int kb = 0; kb < k_stages; ++kb) {  // (coordinates: elements, or bytes of the dense 4-bit stream)
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (p.tx_a + p.tx_b));
                    const uint32_t full_leader = mapa_shared(smem_u32(&full[stage]), 0);
                    tma_load_3d_pair(&map_a, full_leader, smem + L::OFF_A + stage * L::A_STAGE, kb * BLOCK_K, mb * TILE_M + (int)rank * 128, b);
                    tma_load_3d_pair(&map_b, full_leader, smem + L::OFF_B + stage * L::B_STAGE, kb * BLOCK_K, nb * TILE_N + (int)rank * 128, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA; the whole warp runs the loop so that every operand stays in
        // uniform registers -- a single-lane loop makes ptxas wrap each tcgen05 instruction in a lane-serialising
        // R2UR loop -- and one elected lane issues) =================
        if (leader) {
            const uint32_t idesc = make_idesc(TILE_M, TILE_N) | p.idesc_fmt;
            // descriptor words: the high halves are constants, the low halves are (shared address >> 4) + offsets
            constexpr uint64_t HI_OPERAND = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutSw128 << 61);
            constexpr uint64_t HI_SF = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutNone << 61);
            const uint32_t a_lo0 = smem_u32(smem + L::OFF_A) >> 4, b_lo0 = smem_u32(smem + L::OFF_B) >> 4;
            const uint32_t sfa_lo0 = smem_u32(smem + L::OFF_SFA) >> 4, sfb_lo0 = smem_u32(smem + L::OFF_SFB) >> 4;
            uint32_t stage = 0, phase = 0, sfs = 0, sf_phase = 0, sf_j = 0, acc_phase = 0, slot = 0, sf_sel = 0;
            MXQ_DEV_ONLY(const bool tracing = p.trace != nullptr && pair_id == 0 && lane == 0; int tile_iter = 0;)
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 0] = clock64();)
                mbar_wait(tmem_empty, acc_phase ^ 1);
                tc_fence_after();
                MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 1] = clock64();)
                const uint32_t tmem_d = tmem_base + (slot ? ACC_SLOT1 : 0u);
                for (int kb = 0; kb < k_stages; ++kb) {
                    if (sf_j == 0) mbar_wait(&sf_full[sfs], sf_phase);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    MXQ_DEV_ONLY(if (tracing && kb == 0) p.trace[tile_iter * 8 + 2] = clock64();)
                    const bool last = kb == k_stages - 1;
                    const bool sf_done = sf_j + KBS >= SF_KB || last;
                    if (elect_one()) {
                        const uint32_t a_lo = a_lo0 + stage * (L::A_STAGE >> 4), b_lo = b_lo0 + stage * (L::B_STAGE >> 4);
                        const uint32_t tm_sf = tmem_base + TM_SF + sf_sel * (KBS * SF_BUF_COLS);
#pragma unroll
                        for (int j = 0; j < KBS; ++j) {  // scale words of the stage's K block(s): SFA 4 columns, SFB 8 columns each
                            const uint32_t sfa_lo = sfa_lo0 + sfs * (L::SFA_STAGE >> 4) + (sf_j + j) * (L::SFA_KB >> 4);
                            const uint32_t sfb_lo = sfb_lo0 + sfs * (L::SFB_STAGE >> 4) + (sf_j + j) * (L::SFB_KB >> 4);
                            const uint32_t tm_sfa = tm_sf + j * SF_BUF_COLS, tm_sfb = tm_sfa + 4;
                            tc_copy_sf_pair(tm_sfa, HI_SF | sfa_lo);
                            tc_copy_sf_pair(tm_sfb, HI_SF | sfb_lo);
                            tc_copy_sf_pair(tm_sfb + 4, HI_SF | (sfb_lo + (512 >> 4)));
                        }
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {  // advancing K inside the 128B swizzle row = +32 B
                            if constexpr (MXF4) {
                                // 64 elements per instruction: K block k / 2 of the stage, scale bytes {0,1} or {2,3} of its word
                                const uint32_t tm_sfa = tm_sf + (k >> 1) * SF_BUF_COLS;
                                tc_mma_mxf4_pair(tmem_d, HI_OPERAND | (a_lo + k * (UMMA_K >> 4)), HI_OPERAND | (b_lo + k * (UMMA_K >> 4)),
                                                 idesc_with_sf(idesc, (k & 1) * 2, (k & 1) * 2), (kb | k) != 0, tm_sfa, tm_sfa + 4);
                            } else {
                                tc_mma_mx_pair(tmem_d, HI_OPERAND | (a_lo + k * (UMMA_K >> 4)), HI_OPERAND | (b_lo + k * (UMMA_K >> 4)),
                                               idesc_with_sf(idesc, k, k), (kb | k) != 0, tm_sf, tm_sf + 4);
                            }
                        }
                        tc_commit_pair(&empty[stage]);
                        if (sf_done) tc_commit_pair(&sf_empty[sfs]);
                        if (last) tc_commit_pair(tmem_full);
                    }
                    __syncwarp();
                    if (sf_done) {
                        sf_j = 0;
                        if (++sfs == SF_STAGES) { sfs = 0; sf_phase ^= 1; }
                    } else {
                        sf_j += KBS;
                    }
                    sf_sel ^= 1;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 3] = clock64(); ++tile_iter;)
                acc_phase ^= 1;
                slot ^= 1;
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ================= scale-factor loaders (both CTAs; warp 2: own 128 A rows, warp 3: all 256 B rows) =================
        const bool is_a = warp == 2;
        uint32_t sfs = 0, sf_phase = 0;
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
            int b, mb, nb;
            tile_coords(tile, b, mb, nb);
            auto arrive = [&](uint32_t st) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&sf_full[st]), 0));
            };
            if (is_a)
                sf_load_tile4<1>(p.sfa + (int64_t)b * p.sfa_batch, p.ld_sfa, mb * TILE_M + (int)rank * 128, p.M, k_blocks, smem + L::OFF_SFA, L::SFA_KB,
                                 sf_empty, sfs, sf_phase, lane, arrive);
            else
                sf_load_tile4<2>(p.sfb + (int64_t)b * p.sfb_batch, p.ld_sfb, nb * TILE_N, p.N, k_blocks, smem + L::OFF_SFB, L::SFB_KB, sf_empty, sfs,
                                 sf_phase, lane, arrive);
        }
    } else {
        // ================= epilogue (both CTAs, own 128 x 256 accumulator half) =================
        const int quad = warp & 3;
        const uint32_t tmem_empty_leader = mapa_shared(smem_u32(tmem_empty), 0);
        uint32_t acc_phase = 0, slot = 0;
        MXQ_DEV_ONLY(const bool tracing = p.trace != nullptr && pair_id == 0 && leader && warp == 4 && lane == 0; int tile_iter = -1;)
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
            int b, mb, nb;
            tile_coords(tile, b, mb, nb);
            mbar_wait(tmem_full, acc_phase);
            tc_fence_after();
            MXQ_DEV_ONLY(++tile_iter; if (tracing) p.trace[tile_iter * 8 + 4] = clock64();)
            const int row = mb * TILE_M + (int)rank * 128 + quad * 32 + lane;
            uint16_t* drow = p.d + (int64_t)b * p.d_batch + (int64_t)row * p.ldd;
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (slot ? ACC_SLOT1 : 0u);
            if (tma_store) {
                // Coalesced path: 64-column groups go registers -> 128B-swizzled shared rows -> one TMA store per warp and
                // group (32 rows x 128 B; rows / columns past the matrix edge are clipped by the tensor map).  Direct
                // 16-byte st.global from one row per lane costs 32 partial-sector L2 writes per instruction and was
                // measured to stall the TMA loads of the next tile.
                uint8_t* ebuf = smem + L::OFF_EPI + (warp - 4) * L::EPI_WARP;
                // The two accumulator slots share TMEM columns [192,256): 64-column group 3 of slot 0, group 0 of slot 1.  With
                // four warps each drains all four groups, the shared one first; with eight, warp (quadrant, half) drains the two
                // groups of its column half -- the half that holds the shared group takes it first, the other half never touches
                // shared columns and hands the slot over at once.
                constexpr int GROUPS_PER_WARP = 16 / EW;
                const int half = (warp - 4) >> 2;
                const bool owns = EW == 4 || (slot ? half == 0 : half == 1);
                const int first_g = EW == 4 ? (slot ? 0 : 3) : (owns ? (slot ? 0 : 3) : 2 * half);
                const int step_g = (EW == 8 && owns && !slot) ? -1 : 1;  // (3, 2) for the owning half of slot 0
                if (!owns) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
                }
#pragma unroll 1
                for (int h = 0; h < GROUPS_PER_WARP; ++h) {
                    const int g = (first_g + step_g * h) & 3;
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(tmem_acc + g * 64, v0);
                    tmem_ld_32x32b_x32(tmem_acc + g * 64 + 32, v1);
                    tmem_ld_wait();
                    if (h == 0 && owns) {  // the columns shared with the other slot are in registers: the next tile's MMAs may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
                        MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 5] = clock64();)
                    }
                    const int col0 = nb * TILE_N + g * 64;
                    uint8_t* buf = ebuf + (NBUF == 2 ? (h & 1) : h) * 4096;  // NBUF == 4 goes with four groups per warp and tile
                    if (p.d_mc == nullptr && lane == 0) tma_store_wait_read<NBUF - 1>();  // the store that last read this buffer (two groups ago) is done
                    __syncwarp();
                    uint32_t pk[32];
                    // two separate loops: with the bias test inside one loop ptxas predicates the whole bias path (80 ISETP +
                    // 320 predicated instructions per group) and the no-bias case pays ~600 issue cycles for nothing (measured)
                    if (p.bias == nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            pk[i] = pack_bf16x2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
                            pk[16 + i] = pack_bf16x2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float f0 = __uint_as_float(v0[2 * i]), f1 = __uint_as_float(v0[2 * i + 1]);
                            float f2 = __uint_as_float(v1[2 * i]), f3 = __uint_as_float(v1[2 * i + 1]);
                            const int c = col0 + 2 * i;
                            if (c < p.N) f0 += __uint_as_float((uint32_t)p.bias[c] << 16);
                            if (c + 1 < p.N) f1 += __uint_as_float((uint32_t)p.bias[c + 1] << 16);
                            if (c + 32 < p.N) f2 += __uint_as_float((uint32_t)p.bias[c + 32] << 16);
                            if (c + 33 < p.N) f3 += __uint_as_float((uint32_t)p.bias[c + 33] << 16);
                            pk[i] = pack_bf16x2(f0, f1);
                            pk[16 + i] = pack_bf16x2(f2, f3);
                        }
                    }
                    // lane = row of the 32-row box; 16-byte chunk c of the row lives at chunk (c ^ (row & 7)) (SWIZZLE_128B)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    if (p.d_mc != nullptr) {
                        // Fused tensor-parallel all-reduce: this rank's partial tile is ADDED into the output buffer of every
                        // rank through the NVLink multicast address (the sum is formed in the switch).  Read the staging rows
                        // back with eight lanes per row so each instruction carries four complete 128-byte row segments.
                        __syncwarp();
                        const int row_base = mb * TILE_M + (int)rank * 128 + quad * 32;
                        const int c = lane & 7;
                        if (col0 + 8 * c < p.N) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int rr = 4 * j + (lane >> 3);
                                const uint4 val = *reinterpret_cast<const uint4*>(buf + rr * 128 + ((c ^ (rr & 7)) << 4));
                                if (row_base + rr < p.M) multimem_red_add_bf16x8(p.d_mc + (int64_t)(row_base + rr) * p.ldd + col0 + 8 * c, val);
                            }
                        }
                        continue;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&map_d, buf, col0, mb * TILE_M + (int)rank * 128 + quad * 32, b);
                        tma_store_commit();
                    }
                }
                MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 6] = clock64();)
                acc_phase ^= 1;
                slot ^= 1;
                continue;
            }
            const int first = slot ? 0 : 6;  // the two 32-column chunks inside [192,256) of TMEM come first
            auto store_chunk = [&](const uint32_t (&v)[32], int c) {
                const int col0 = nb * TILE_N + c * 32;
                if (row < p.M && col0 < p.N MXQ_DEV_ONLY(&& !(p.dbg & 1))) {
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
            };
            {
                uint32_t v0[32], v1[32];
                tmem_ld_32x32b_x32(tmem_acc + first * 32, v0);
                tmem_ld_32x32b_x32(tmem_acc + (first + 1) * 32, v1);
                tmem_ld_wait();
                // the columns shared with the other slot are in registers: the next tile's MMAs may start
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader);
                MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 5] = clock64();)
                store_chunk(v0, first);
                store_chunk(v1, first + 1);
            }
#pragma unroll 1
            for (int ci = 2; ci < 8; ++ci) {
                const int c = (first + ci) & 7;
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + c * 32, v);
                tmem_ld_wait();
                store_chunk(v, c);
            }
            MXQ_DEV_ONLY(if (tracing) p.trace[tile_iter * 8 + 6] = clock64();)
            acc_phase ^= 1;
            slot ^= 1;
        }
    }

    MXQ_DEV_ONLY(if (p.trace != nullptr && leader && threadIdx.x == 128) {  // epilogue warp 4 of every leader: when did this pair finish (ns)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[512 + pair_id] = (long long)t;
    })
    if (tma_store && warp >= 4 && lane == 0) tma_store_wait<0>();  // this thread's bulk stores have left shared memory and are performed
    __syncwarp();  // single-lane roles (producer, MMA issuer) rejoin their warp before the aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still post on its barriers / read its smem
    if (warp == 3) tmem_dealloc_pair<512>(tmem_base);
}
}  // namespace pair

// ---- host side ------------------------------------------------------------------------------------------
static void fill_params(Params& p, const mxq_gemm_args_t* a, int tile_m, int tile_n) {
    p.sfa = a->sfa; p.sfb = a->sfb; p.bias = (const uint16_t*)a->bias; p.d = (uint16_t*)a->d;
    p.d_mc = nullptr;
    p.ld_sfa = a->ld_sfa; p.ld_sfb = a->ld_sfb; p.sfa_batch = a->sfa_batch_stride; p.sfb_batch = a->sfb_batch_stride;
    p.ldd = a->ldd; p.d_batch = a->d_batch_stride;
    p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K; p.batch = (int)a->batch;
    p.idesc_fmt = idesc_formats(a->a_format, a->b_format);
    p.tx_a = 128 * BLOCK_K * operand_bits(a->a_format) / 8;
    p.tx_b = 128 * BLOCK_K * operand_bits(a->b_format) / 8;
    p.m_blocks = (int)((a->M + tile_m - 1) / tile_m);
    p.n_blocks = (int)((a->N + tile_n - 1) / tile_n);
    MXQ_DEV_ONLY(p.trace = nullptr; p.dbg = 0;)
}

template <int BLOCK_N, int STAGES>
static int launch_cfg(const mxq_gemm_args_t* a, const CUtensorMap& ma, const CUtensorMap& mb, int sm_count, int device, cudaStream_t stream, char* msg,
                      size_t msg_len) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    {
        const cudaError_t e = ensure_smem_attr((const void*)mx_gemm_kernel<BLOCK_N, STAGES>, L::DYN_BYTES, device);
        if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    }
    Params p;
    fill_params(p, a, BLOCK_M, BLOCK_N);
    const int64_t tiles = (int64_t)p.m_blocks * p.n_blocks * p.batch;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    mx_gemm_kernel<BLOCK_N, STAGES><<<grid, kThreads, L::DYN_BYTES, stream>>>(ma, mb, p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    return MXQ_OK;
}


template <int STAGES, int EW, int NBUF, bool MXF4>
static int launch_pair(const mxq_gemm_args_t* a, const CUtensorMap& ma, const CUtensorMap& mb, int sm_count, int device, cudaStream_t stream, char* msg,
                       size_t msg_len) {
    using L = pair::Smem<STAGES, EW, NBUF>;
    {
        const cudaError_t e = ensure_smem_attr((const void*)pair::mx_gemm_pair_kernel<STAGES, EW, NBUF, MXF4>, L::DYN_BYTES, device);
        if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    }
    Params p;
    fill_params(p, a, pair::TILE_M, pair::TILE_N);
    p.d_mc = (uint16_t*)a->d_multicast;
    if (MXF4) {  // dense 4-bit streams: a stage is 128 rows x 128 bytes (256 elements) of each operand
        p.idesc_fmt = kIdescMxf4Formats;
        p.tx_a = p.tx_b = 128 * 128;
    }
    MXQ_DEV_ONLY(p.dbg = dev_env("MXQ_GEMM_DBG");
                 p.trace = getenv("MXQ_GEMM_TRACE") ? reinterpret_cast<long long*>(strtoull(getenv("MXQ_GEMM_TRACE"), nullptr, 0)) : nullptr;)
    const int64_t tiles = (int64_t)p.m_blocks * p.n_blocks * p.batch;
    const int max_pairs = sm_count / 2;
    const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
    int group_m = 8;
    MXQ_DEV_ONLY(if (dev_env("MXQ_GEMM_GM") > 0) group_m = dev_env("MXQ_GEMM_GM");)
    // coalesced TMA-store epilogue needs a 16-byte aligned D with a 16-byte multiple row pitch; otherwise direct stores
    CUtensorMap md;
    int tma_store = ((uintptr_t)a->d % 16 == 0) && (a->ldd % 8 == 0) && (a->d_batch_stride % 8 == 0);
    MXQ_DEV_ONLY(if (p.dbg & 2) tma_store = 0;)
    if (a->d_multicast != nullptr) {  // fused all-reduce epilogue: staging path without the tensor map
        if (((uintptr_t)a->d_multicast % 16) || (a->ldd % 8) || (a->N % 8) || a->batch != 1) {
            snprintf(msg, msg_len, "d_multicast needs a 16-byte aligned buffer, N %% 8 == 0, ldd %% 8 == 0 and batch == 1");
            return MXQ_ERR_UNSUPPORTED_SHAPE;
        }
        tma_store = 1;
        md = ma;
    } else if (tma_store && !cached_d_map(&md, a->d, a->N, a->M, a->batch, a->ldd, a->d_batch_stride, device)) tma_store = 0;
    if (!tma_store) md = ma;  // unused placeholder
    if (EW == 8 && !tma_store) return MXQ_ERR_UNSUPPORTED_SHAPE;  // the eight-warp epilogue exists for the staged path only (caller retries with EW = 4)
    pair::mx_gemm_pair_kernel<STAGES, EW, NBUF, MXF4><<<2 * pairs, 128 + 32 * EW, L::DYN_BYTES, stream>>>(ma, mb, md, p, group_m, tma_store);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch (pair): %s", cudaGetErrorString(e)); return MXQ_ERR_CUDA; }
    return MXQ_OK;
}

int launch_gemm_skinny(const mxq_gemm_args_t* a, int sm_count, int device, cudaStream_t stream, char* msg, size_t msg_len);  // mxq_gemm_skinny.cu

}  // namespace gemm

int launch_gemm(const mxq_gemm_args_t* a, int sm_count, int device, cudaStream_t stream, char* msg, size_t msg_len) {
    using namespace gemm;
    if (a->K <= 0 || a->K % BLOCK_K) { snprintf(msg, msg_len, "K=%lld is not a positive multiple of %d", (long long)a->K, BLOCK_K); return MXQ_ERR_UNSUPPORTED_SHAPE; }
    if ((a->lda % 16) || (a->ldb % 16) || ((uintptr_t)a->a_codes % 16) || ((uintptr_t)a->b_codes % 16) || (a->a_batch_stride % 16) || (a->b_batch_stride % 16)) {
        snprintf(msg, msg_len, "operand pointers / strides must be 16-byte aligned for TMA");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    for (int side = 0; side < 2; ++side) {
        const int fmt = side ? a->b_format : a->a_format;
        const int64_t ld = side ? a->ldb : a->lda, bs = side ? a->b_batch_stride : a->a_batch_stride;
        const uintptr_t base = (uintptr_t)(side ? a->b_codes : a->a_codes);
        if (fmt != MXQ_OPERAND_E4M3_BYTES && fmt != MXQ_OPERAND_E5M2_BYTES && ((ld % 32) || (bs % 32) || (base % 32))) {
            snprintf(msg, msg_len, "packed operands need 32-byte aligned pointers / strides");
            return MXQ_ERR_UNSUPPORTED_SHAPE;
        }
    }
    if ((a->ld_sfa % 4) || (a->ld_sfb % 4) || ((uintptr_t)a->sfa % 4) || ((uintptr_t)a->sfb % 4) || (a->sfa_batch_stride % 4) || (a->sfb_batch_stride % 4)) {
        snprintf(msg, msg_len, "scale pointers / strides must be 4-byte aligned");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    if (a->M > 0x7FFFFFFF || a->N > 0x7FFFFFFF || a->K > 0x7FFFFFFF || a->batch > 0x7FFFFFFF) { snprintf(msg, msg_len, "extent too large"); return MXQ_ERR_UNSUPPORTED_SHAPE; }
    bool skinny_ok = a->M <= 128 && a->batch == 1;
    MXQ_DEV_ONLY(if (dev_env("MXQ_NO_SKINNY")) skinny_ok = false;)
    if (skinny_ok) {  // decode-sized activations: weight-streaming kernel (K3c)
        const int rc = launch_gemm_skinny(a, sm_count, device, stream, msg, msg_len);
        if (rc != MXQ_ERR_UNSUPPORTED_SHAPE) return rc;
    }
    if (a->x_bf16 != nullptr) {
        snprintf(msg, msg_len, "fused activation quantization is implemented by the skinny (decode) kernel only");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    if (a->d_multicast != nullptr && !(a->M > 128 && a->N > 128)) {
        snprintf(msg, msg_len, "d_multicast (fused all-reduce) is implemented by the CTA-pair and skinny kernels only");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    int cfg = 0;  // developer builds: <BLOCK_N><STAGES>, 2<STAGES> = CTA pair
    MXQ_DEV_ONLY(cfg = dev_env("MXQ_GEMM_CFG");)
    // Under-filled grids: when the 256x256 pair tiles would occupy at most a quarter of the SM pairs, 128x128 tiles spread the
    // same work over 4x as many SMs (measured, K = 4096, before / after: 2048x1024 16.8 -> 13.9 us, 1024x1024 16.6 -> 12.8, 512x4096 17.1 -> 13.7).
    // MXQ_GEMM_WIDE_TILES switches the rule off; a fused all-reduce epilogue exists in the pair kernel only.
    const int64_t pair_tiles = ((a->M + 255) / 256) * ((a->N + 255) / 256) * a->batch;
    const bool underfilled = !(a->flags & MXQ_GEMM_WIDE_TILES) && a->M > 128 && a->N > 128 && pair_tiles * 4 <= sm_count && a->d_multicast == nullptr;
    const bool wide = a->N > 128 && !underfilled;
    CUtensorMap ma, mb;
    const bool use_pair = wide && a->M > 128 && sm_count >= 2 && (cfg == 0 || cfg / 10 == 2 || a->d_multicast != nullptr);
    if (use_pair) {
        // fp4 x fp4: kind::mxf4 reads the dense 4-bit streams (no 16-byte slot expansion) -> plain byte maps over K / 2 bytes per row
        const bool mxf4 = a->a_format == MXQ_OPERAND_E2M1_PACKED && a->b_format == MXQ_OPERAND_E2M1_PACKED && a->K % 256 == 0 && !(a->flags & MXQ_GEMM_NO_MXF4);
        const int64_t k_a = mxf4 ? a->K / 2 : a->K, k_b = mxf4 ? a->K / 2 : a->K;
        const int fmt_a = mxf4 ? MXQ_OPERAND_E4M3_BYTES : a->a_format, fmt_b = mxf4 ? MXQ_OPERAND_E4M3_BYTES : a->b_format;
        if (!cached_operand_map(&ma, a->a_codes, k_a, a->M, a->batch, a->lda, a->a_batch_stride, 128, fmt_a, device) ||
            !cached_operand_map(&mb, a->b_codes, k_b, a->N, a->batch, a->ldb, a->b_batch_stride, 128, fmt_b, device)) {
            snprintf(msg, msg_len, "cuTensorMapEncodeTiled failed (driver entry point missing or invalid strides)");
            return MXQ_ERR_UNSUPPORTED_SHAPE;
        }
        if (mxf4) return launch_pair<5, 4, 2, true>(a, ma, mb, sm_count, device, stream, msg, msg_len);
        MXQ_DEV_ONLY(if (cfg == 24) return launch_pair<4, 4, 2, false>(a, ma, mb, sm_count, device, stream, msg, msg_len);)
        // short K loops are bound by the epilogue (a 64 KB output tile per CTA for a few hundred MMA cycles): a shallower operand
        // ring; long K loops hide the epilogue and want the deeper ring.  (Eight epilogue warps / four staging buffers per warp
        // were measured and dropped once the bias code was un-predicated.)
        if (a->K / BLOCK_K <= 2 && cfg != 25) {
            const int rc = launch_pair<3, 4, 2, false>(a, ma, mb, sm_count, device, stream, msg, msg_len);
            if (rc != MXQ_ERR_UNSUPPORTED_SHAPE) return rc;
        }
        return launch_pair<5, 4, 2, false>(a, ma, mb, sm_count, device, stream, msg, msg_len);
    }
    if (!cached_operand_map(&ma, a->a_codes, a->K, a->M, a->batch, a->lda, a->a_batch_stride, BLOCK_M, a->a_format, device) ||
        !cached_operand_map(&mb, a->b_codes, a->K, a->N, a->batch, a->ldb, a->b_batch_stride, wide ? 256 : 128, a->b_format, device)) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled failed (driver entry point missing or invalid strides)");
        return MXQ_ERR_UNSUPPORTED_SHAPE;
    }
    if (wide) return launch_cfg<256, 4>(a, ma, mb, sm_count, device, stream, msg, msg_len);
    return launch_cfg<128, 6>(a, ma, mb, sm_count, device, stream, msg, msg_len);
}

}  // namespace mxq
