// dcn_pair.cuh — the DCNv2Pack site of EDVR (conv_offset + sigmoid/split + modulated deformable 3x3 conv, stride 1, pad 1)
// as ONE kernel issued by CTA PAIRS, with the conv_offset GEMM of the NEXT half tile running on the tensor cores WHILE the
// gather warps sample the current one.
//
// Replaces (paths under /root/reference/basicsr/models/):
//   DCNv2Pack.forward ........................ archs/arch_util.py:243-257   (conv_offset + chunk/cat + sigmoid + dcn)
//   modulated_deform_conv_cuda_forward ....... ops/dcn/src/deform_conv_cuda.cpp:490-569
//   modulated_deformable_im2col_gpu_kernel ... ops/dcn/src/deform_conv_cuda_kernel.cu:570-633 (+ bilinear :467-497)
//
// What profiles/r02_ncu_dcn_site_fused_* said about the single-CTA kernel (dcn_site.cuh): per 128-pixel tile the two phases
// ran back to back - ~18 K cycles of conv_offset MMAs during which the 16 gather warps idle, then ~15 K cycles of gather
// during which the tensor pipe is 25 % busy - and every SM re-streamed all 811 KB of weights (conv_offset 516 KB + DCN
// 295 KB) per tile from L2.  This kernel changes three things:
//   1. CTA pairs (tcgen05 cta_group::2): the two CTAs of a cluster each own one 16x8-pixel tile (M = 2 x 128 rows per MMA)
//      and HALF of the rows of both weight matrices (B split along N), so each SM streams 406 KB per tile, not 811.
//   2. The conv_offset accumulator is split into two column halves of 112 (deformable groups [0, dg/2) and [dg/2, dg)).
//      The gather consumes half 0 during the first half of the tile's channels and half 1 during the second, so the
//      tensor cores refill half h for tile T+1 as soon as the gather of tile T has read it: phase A of the next tile
//      overlaps phase B of this one with no extra tensor memory (2 x 128 accumulator + 2 x 112 offset columns = 480).
//      Two issuer warps (one per phase) feed the tensor pipe independently.
//   3. The sampling window is staged per 32-channel chunk ({32 ch, 24 x, 31 y} box, 64-byte TMA swizzle): half the bytes
//      per buffer of the 64-channel window, which pays for the second staging ring AND for a wider margin (-6 .. +7 px in x,
//      +-6 px in y served from shared memory instead of +-3).  A (chunk, tap) stage is gathered by 8 warps; the two
//      groups of 8 gather warps take alternate stages.
// Thread mapping of the gather: lane = output pixel = TMEM lane; warp (q, kp) covers TMEM lane quarter q and the K-atom
// pair kp (16 channels) of the 32-channel chunk; the (dh, dw, mask logit) triple of (pixel, group, tap) comes straight
// out of tensor memory (tcgen05.ld); offsets and masks never exist in HBM.  Sampling arithmetic is fp32 (coordinates,
// bilinear weights); the blend is packed fp16 (weights rounded to 11 bits) and the gathered column is the fp16 operand.
// Measured (one L1 launch over 28 frames of 180x320, offsets ~N(0, sigma^2) px): 1686 / 2220 / 3177 us at sigma 0.02 / 3 / 10
// against 1937 / 2724 / 3695 for dcn_site.cuh; parity <= 6.0e-4 (profiles/r02_ncu_dcn_pair_*,
// r02_dcn_sweep_pair_final.json, r02_dcn_pair_role_timing.txt).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dcn_fused.cuh"
#include "dcn_site.cuh"
#include "epilogue.cuh"

namespace eb {

// Role timing (compile with -DDP_PROF): every role's lane 0 accumulates the cycles it spends inside each kind of wait and
// adds them to PP.prof[] when it finishes; read back with eb_dcn_pair_prof_read (tools/dp_prof.py).
#ifdef DP_PROF
#define DP_T0() const long long dp_t0_ = clock64()
#define DP_ACC(VAR_) VAR_ += clock64() - dp_t0_
#define DP_TIMED(VAR_, STMT_) do { const long long dp_t_ = clock64(); STMT_; VAR_ += clock64() - dp_t_; } while (0)
#define DP_FLUSH(IDX_, VAR_) do { if (PP.prof != nullptr && lane == 0) atomicAdd(PP.prof + (IDX_), static_cast<unsigned long long>(VAR_)); } while (0)
#else
#define DP_TIMED(VAR_, STMT_) do { STMT_; } while (0)
#define DP_FLUSH(IDX_, VAR_) do { } while (0)
#endif

#ifndef DP_CFG_RY          // development overrides (nvcc -DDP_CFG_...): A/B builds of the ring depths and the window margin
#define DP_CFG_RY 6
#endif
#ifndef DP_CFG_A_STAGES
#define DP_CFG_A_STAGES 4
#endif
#ifndef DP_CFG_WO_STAGES
#define DP_CFG_WO_STAGES 4
#endif
constexpr int DP_RY = DP_CFG_RY;                                   // rows of offset served from shared memory above / below the 3x3 reach
constexpr int DP_XL = 6;                                   // columns to the left (to the right: DP_WW - 11 - DP_XL = 7)
constexpr int DP_WW = 24;                                  // window columns; a multiple of 8 keeps the swizzle phase of row y+1
constexpr int DP_WH = DC_TILE_H + 3 + 2 * DP_RY;           // 31 rows
constexpr int DP_WIN_TX = DP_WH * DP_WW * 64;              // bytes one window box delivers (32 channels per pixel)
constexpr int DP_WIN_BYTES = ((DP_WIN_TX + 1023) / 1024) * 1024;
constexpr int DP_A_STAGES = DP_CFG_A_STAGES;                             // gathered operand: (32-channel chunk, tap) = 4 K-atom planes x 128 px; the
                                                           // DCN weights of the stage travel in the same ring slot (one barrier pair)
constexpr int DP_A_LBO = 128 * 16;
constexpr int DP_A_STAGE = 4 * DP_A_LBO;                   // 8192
constexpr int DP_W_STAGES = DP_A_STAGES;                   // this CTA's half (BN/2 rows) of the DCN weights of one stage
constexpr int DP_W_STAGE = 4 * 64 * 16;                    // 4096 at BN = 128
constexpr int DP_F_STAGES = 3;                             // 18x10 halo of 32 offset-feature channels (phase A operand A)
constexpr int DP_F_STAGE = DS_F_STAGE;                     // 11520
constexpr int DP_OFF_HALF = 112;                           // conv_offset columns per half: (dg/2) * 27 <= 112
constexpr int DP_WO_ROWS = DP_OFF_HALF / 2;                // rows of a half held by each CTA
constexpr int DP_WO_TAP = 4 * DP_WO_ROWS * 16;             // 3584: (32-channel chunk, tap) of one half, one CTA
constexpr int DP_WO_STAGE = 3 * DP_WO_TAP;                 // 10752: three taps per stage - one barrier round trip (~90 cycles for an
                                                           // already-complete try_wait) per 6 MMAs; per-tap stages left the A issuer
                                                           // wait-bound at ~150 cycles per 112 cycles of MMA work (r02_ncu_dcn_pair_v2)
constexpr int DP_WO_STAGES = DP_CFG_WO_STAGES;
constexpr int DP_THREADS = 32 * 25;   // warps: 0 B-side producer, 1 B issuer, 2-5 epilogue, 6-21 gather, 22 forwarder, 23 A-side producer, 24 A issuer
constexpr int DP_MISC_BYTES = 128 * 4 + 256 * 4 + 1024;    // DCN bias, conv_offset bias, barriers
constexpr int DP_SMEM_BYTES = 2 * DP_WIN_BYTES + DP_A_STAGES * DP_A_STAGE + DP_W_STAGES * DP_W_STAGE + DP_F_STAGES * DP_F_STAGE +
                              DP_WO_STAGES * DP_WO_STAGE + DP_MISC_BYTES;
static_assert(DP_SMEM_BYTES <= 232448, "dcn_pair: shared memory budget");
static_assert(DP_WW % 8 == 0 && DP_WW >= DC_TILE_W + 3 + DP_XL, "dcn_pair: window width");
static_assert(DP_A_STAGES < 9 && DP_A_STAGES % 2 == 0, "dcn_pair: an even ring depth keeps each gather warp group on its own slots");

struct DpParams {
    DcnParams d;               // x view, shapes, dg / cpg, epilogue; d.wpack = DCN weights in the CTA-pair layout (eb_pack_weight_pair)
    CUtensorMap tmap_x;        // x as {pix_stride, W, H, N} fp16, box {32, DP_WW, DP_WH, 1}, 64-byte swizzle, zero fill
    CUtensorMap tmap_f;        // offset features as {8, W, H, pix_stride/8, N}, box {8, 10, 18, 4, 1} (K-atom planes)
    CUtensorMap tmap_w;        // packed DCN weights as rows of 256 fp16, box = one (chunk, tap) stage of one CTA (BN / 16 rows)
    CUtensorMap tmap_wo;       // packed conv_offset weights as rows of 256 fp16, box = 7 rows = one stage of one CTA
    int f_ch_off;
    const __half* wo_pack;     // conv_offset weights [half 2][cta 2][chunk32][tap][k16 2][plane 2][56 rows][8] fp16 (behind tmap_wo)
    const float* bo;           // conv_offset bias in column order [half][112], 224 entries
    float* absmean;            // optional: += sum |offset|
    unsigned long long* prof;  // role timing counters (builds with -DDP_PROF only)
    int hint_crit, hint_idle, hint_gather;  // suspend-time hints (ns) of the pipeline-critical / the long waits (mbar_wait_hint)
    int dbg;                   // profiling ablations (EDVR_B200_DP_DBG): 1 gather skips sampling, 2 no phase-A MMAs, 4 no conv_offset
                               // weight copies, 8 no DCN weight copies, 16 no window copies, 32 no phase-B MMAs, 64 no halo copies
};

template <int EK, bool TWO>
__global__ void __launch_bounds__(DP_THREADS, 1) dcn_pair_kernel(const __grid_constant__ DpParams PP) {
    const DcnParams& P = PP.d;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();     // the swizzled window boxes need this alignment
    uint8_t* win_smem = smem;                                             // [2][DP_WIN_BYTES]
    uint8_t* a_smem = win_smem + 2 * DP_WIN_BYTES;                        // [DP_A_STAGES][DP_A_STAGE]
    uint8_t* w_smem = a_smem + DP_A_STAGES * DP_A_STAGE;                  // [DP_W_STAGES][DP_W_STAGE]
    uint8_t* f_smem = w_smem + DP_W_STAGES * DP_W_STAGE;                  // [DP_F_STAGES][DP_F_STAGE]
    uint8_t* wo_smem = f_smem + DP_F_STAGES * DP_F_STAGE;                 // [DP_WO_STAGES][DP_WO_STAGE]
    float* bias_s = reinterpret_cast<float*>(wo_smem + DP_WO_STAGES * DP_WO_STAGE);    // [128]
    float* bo_s = bias_s + 128;                                           // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bo_s + 256);
    // LEADER = used in the even CTA only, signalled by both CTAs (TMA bytes of the odd CTA's boxes are counted on the even
    // CTA's barrier by the .cta_group::2 copies; plain arrivals cross through mbar_arrive_remote_cta)
    uint64_t* full = bars;                          // [A]  LEADER: per CTA one forwarder arrival (operand gathered) + the weight bytes
    uint64_t* empty = full + DP_A_STAGES;           // [A]  both: B issuer commit (stage consumed: gather warps and weight producer)
    uint64_t* gathered = empty + DP_A_STAGES;       // [A]  local: 8 gather warps
    uint64_t* win_full = gathered + DP_A_STAGES;    // [2]  local: window bytes
    uint64_t* win_empty = win_full + 2;             // [2]  local: 16 gather warps
    uint64_t* acc_full = win_empty + 2;             // [2]  both: B issuer commit
    uint64_t* acc_empty = acc_full + 2;             // [2]  LEADER: 4 + 4 epilogue warps
    uint64_t* f_full = acc_empty + 2;               // [F]  LEADER: halo bytes of both CTAs
    uint64_t* f_empty = f_full + DP_F_STAGES;       // [F]  both: A issuer commit
    uint64_t* wo_full = f_empty + DP_F_STAGES;      // [WO] LEADER: conv_offset weight bytes of both CTAs
    uint64_t* wo_empty = wo_full + DP_WO_STAGES;    // [WO] both: A issuer commit
    uint64_t* off_full = wo_empty + DP_WO_STAGES;   // [2]  both: offsets of column half h complete (A issuer commit)
    uint64_t* off_empty = off_full + 2;             // [2]  LEADER: 16 + 16 gather warps have read half h of this tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(off_empty + 2);
    static_assert((3 * DP_A_STAGES + 8 + 2 * DP_F_STAGES + 2 * DP_WO_STAGES + 4) * 8 + 8 <= 1024, "dcn_pair: barrier area");
    static_assert(DP_WO_STAGE % 512 == 0, "dcn_pair: a conv_offset weight stage is a whole number of 512-byte tensor-map rows");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_x = (P.Wo + DC_TILE_W - 1) / DC_TILE_W;
    const int tiles_y = (P.Ho + DC_TILE_H - 1) / DC_TILE_H;
    const int total_tiles = P.N * tiles_y * tiles_x;
    const int npairs = (total_tiles + 1) >> 1;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int n_iter = cluster_id < npairs ? (npairs - cluster_id + nclusters - 1) / nclusters : 0;
    const int nc = P.C / 32;                        // 32-channel chunks, even
    const int nch = nc >> 1;                        // chunks per offset half
    const int nstages = 9 * nc;                     // (chunk, tap) stages per tile, even
    const int halfN = P.BN >> 1;                    // DCN weight rows held by this CTA
    const uint32_t w_stage_bytes = static_cast<uint32_t>(halfN) * 64u;
    constexpr uint32_t TM_OFF = 256;                // TMEM columns: [0,128) [128,256) DCN accumulators, [256,368) [368,480) offset halves

#define DP_TILE_OF(IT_) (2 * (cluster_id + (IT_) * nclusters) + static_cast<int>(rank))
#define DP_LIVE(IT_) (DP_TILE_OF(IT_) < total_tiles)

    const bool has_bias = P.epi.bias != nullptr;
    if (has_bias && threadIdx.x < P.BN) bias_s[threadIdx.x] = P.epi.bias[threadIdx.x];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bo_s[i] = i < 2 * DP_OFF_HALF ? PP.bo[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DP_A_STAGES; ++i) { mbar_init(&full[i], 4); mbar_init(&empty[i], 1); mbar_init(&gathered[i], 8); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&win_full[i], 1); mbar_init(&win_empty[i], 16);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
            mbar_init(&off_full[i], 1); mbar_init(&off_empty[i], 32);
        }
        for (int i = 0; i < DP_F_STAGES; ++i) { mbar_init(&f_full[i], 2); mbar_init(&f_empty[i], 1); }
        for (int i = 0; i < DP_WO_STAGES; ++i) { mbar_init(&wo_full[i], 2); mbar_init(&wo_empty[i], 1); }
        fence_barrier_init();
        tma_prefetch_desc(&PP.tmap_x);
        tma_prefetch_desc(&PP.tmap_f);
        tma_prefetch_desc(&PP.tmap_w);
        tma_prefetch_desc(&PP.tmap_wo);
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                    // both CTAs' barriers and tensor memory exist before any remote traffic
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch (same contract as conv_pair.cuh): the next kernel of the stream may start its prologue on
    // SMs this grid has left; everything above touched only constants, everything below waits for the predecessor.
    pdl_trigger();
    pdl_wait();

    // this CTA's share of a two-CTA stage: one arrival + its byte count on the EVEN CTA's barrier
    auto expect_on_leader = [&](uint64_t* bar, uint32_t bytes) {
        if (leader) mbar_arrive_expect_tx(bar, bytes); else mbar_arrive_expect_tx_remote(bar, bytes, 0);
    };
    const int dbg = PP.dbg;
    const uint32_t hc = static_cast<uint32_t>(PP.hint_crit), hi = static_cast<uint32_t>(PP.hint_idle), hg = static_cast<uint32_t>(PP.hint_gather);
    auto tile_coords = [&](int it, int& tx, int& ty, int& img) {
        int tile = DP_TILE_OF(it);
        if (tile >= total_tiles) tile = 0;           // the odd CTA of the last pair may run a dead tile
        tx = tile % tiles_x; ty = (tile / tiles_x) % tiles_y; img = tile / (tiles_x * tiles_y);
    };

    if (warp == 0) {
        // ================= B-side producer: this CTA's half of the DCN weights, one stage per (chunk, tap), and the sampling
        // windows (one per tile and chunk, two buffers).  The window of chunk G+1 is issued after the weights of stage
        // (G, DP_A_STAGES) were allowed in, i.e. after the MMAs of (G, 0) completed: the gather of chunk G-1 has left its buffer.
        if (lane == 0) {
            uint32_t wit = 0;
            long long pt_total = clock64(), pt_wempty = 0, pt_winempty = 0;
            const int w_rows = static_cast<int>(w_stage_bytes / 512u);                 // rows of 256 fp16 per stage
            const int w_row0 = static_cast<int>(rank) * nc * 9 * w_rows;
            auto issue_window = [&](int it, int c) {
                int tx, ty, img;
                tile_coords(it, tx, ty, img);
                const uint32_t G = static_cast<uint32_t>(it) * nc + c, wb = G & 1u;
                DP_TIMED(pt_winempty, mbar_wait_hint(&win_empty[wb], ((G >> 1) & 1u) ^ 1u, hi));
                mbar_arrive_expect_tx(&win_full[wb], (dbg & 16) ? 0u : DP_WIN_TX);
                if (!(dbg & 16)) tma_load_4d(win_smem + wb * DP_WIN_BYTES, &PP.tmap_x, &win_full[wb], P.x_ch_off + c * 32,
                            tx * DC_TILE_W - 1 - DP_XL, ty * DC_TILE_H - 1 - DP_RY, img);
            };
            if (n_iter > 0) issue_window(0, 0);
            for (int it = 0; it < n_iter; ++it)
                for (int c = 0; c < nc; ++c)
                    for (int t = 0; t < 9; ++t, ++wit) {
                        const uint32_t s = wit % DP_A_STAGES, ph = (wit / DP_A_STAGES) & 1u;
                        DP_TIMED(pt_wempty, mbar_wait_hint(&empty[s], ph ^ 1u, hi));
                        if (t == DP_A_STAGES) {
                            if (c + 1 < nc) issue_window(it, c + 1);
                            else if (it + 1 < n_iter) issue_window(it + 1, 0);
                        }
                        expect_on_leader(&full[s], (dbg & 8) ? 0u : w_stage_bytes);
                        if (!(dbg & 8)) tma_load_2d_pair(w_smem + s * DP_W_STAGE, &PP.tmap_w, &full[s], 0, w_row0 + (c * 9 + t) * w_rows);
                    }
            DP_FLUSH(15, pt_wempty); DP_FLUSH(16, pt_winempty); DP_FLUSH(17, clock64() - pt_total);
        }
    } else if (warp == 1) {
        // ================= B issuer (even CTA): D[256 px, BN] += gathered columns x W, one (chunk, tap) stage = 2 x K16
        if (leader) {
            const uint32_t idesc = umma_idesc_f16(256, P.BN);
            const uint32_t lbo_b = static_cast<uint32_t>(halfN) * 16u;
            const uint32_t a_hi = umma_desc_hi(128), b_hi = umma_desc_hi(128);
            uint32_t git = 0;
            long long pt_total = clock64(), pt_acc = 0, pt_full = 0;
            for (int it = 0; it < n_iter; ++it) {
                const uint32_t ab = it & 1u;
                DP_TIMED(pt_acc, mbar_wait_hint(&acc_empty[ab], ((it >> 1) & 1u) ^ 1u, hc));
                tc_fence_after_sync();
                const uint32_t d = tmem_base + ab * 128u;
                for (int st = 0; st < nstages; ++st, ++git) {
                    const uint32_t as = git % DP_A_STAGES;
                    DP_TIMED(pt_full, mbar_wait_hint(&full[as], (git / DP_A_STAGES) & 1u, hc));
                    tc_fence_after_sync();
                    const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_smem + as * DP_A_STAGE), DP_A_LBO);
                    const uint32_t b_lo0 = umma_desc_lo(smem_u32(w_smem + as * DP_W_STAGE), lbo_b);
                    if (elect_one()) {
#pragma unroll
                        for (int k16 = 0; k16 < 2; ++k16)
                            if (!(dbg & 32)) umma_f16_lohi<2>(d, a_lo0 + k16 * (2 * DP_A_LBO / 16), a_hi, b_lo0 + k16 * ((2u * lbo_b) >> 4), b_hi, idesc,
                                             (st | k16) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty[as], 3);
                        if (st == nstages - 1) umma_commit_pair(&acc_full[ab], 3);
                    }
                    __syncwarp();
                }
            }
            DP_FLUSH(7, pt_acc); DP_FLUSH(9, pt_full); DP_FLUSH(10, clock64() - pt_total);
        }
    } else if (warp < 6) {
        // ================= epilogue: warps 2..5 -> TMEM lane quarters 2,3,0,1; each CTA drains and stores its own tile
        const int q = warp & 3;
        long long pt_total = clock64(), pt_accfull = 0;
        for (int it = 0; it < n_iter; ++it) {
            int tx, ty, img;
            tile_coords(it, tx, ty, img);
            const uint32_t ab = it & 1u;
            DP_TIMED(pt_accfull, mbar_wait_hint(&acc_full[ab], (it >> 1) & 1u, hi));
            tc_fence_after_sync();
            const int y = ty * DC_TILE_H + 4 * q + (lane >> 3);
            const int x = tx * DC_TILE_W + (lane & 7);
            const bool valid = DP_LIVE(it) && (y < P.Ho) && (x < P.Wo);
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + ab * 128u;
#pragma unroll 1
            for (int cc = 0; cc < P.BN; cc += 32) {
                float v[32];
                tmem_ld32(t0 + cc, v);
                epi_store32<EK>(P.epi, has_bias ? bias_s : nullptr, v, img, y, x, cc, valid);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&acc_empty[ab]); else mbar_arrive_remote_cta(&acc_empty[ab], 0); }
        }
        DP_FLUSH(21, pt_accfull); DP_FLUSH(22, clock64() - pt_total);
    } else if (warp == 22) {
        // ================= forwarder: stage gathered by this CTA's 8 warps -> generic -> async proxy fence -> one arrival on the
        // EVEN CTA's full[s]
        if (lane == 0) {
            uint32_t git = 0;
            long long pt_total = clock64(), pt_gathered = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int st = 0; st < nstages; ++st, ++git) {
                    const uint32_t as = git % DP_A_STAGES;
                    DP_TIMED(pt_gathered, mbar_wait_hint(&gathered[as], (git / DP_A_STAGES) & 1u, hc));
                    fence_proxy_async_smem();
                    if (leader) mbar_arrive(&full[as]); else mbar_arrive_remote_cta(&full[as], 0);
                }
            DP_FLUSH(5, pt_gathered); DP_FLUSH(6, clock64() - pt_total);
        }
    } else if (warp == 23) {
        // ================= A-side producer: halo stages of the offset features and this CTA's 56 rows of every conv_offset
        // weight stage, for the half steps (tile 0, half 0), (tile 0, half 1), (tile 1, half 0), ...
        if (lane == 0) {
            uint32_t wo_it = 0, fl = 0;
            long long pt_total = clock64(), pt_woempty = 0, pt_fempty = 0;
            const uint32_t f_total = static_cast<uint32_t>(n_iter) * 2u * nc;
            auto issue_f = [&](uint32_t l) {
                const int it = static_cast<int>(l / (2u * nc)), c = static_cast<int>(l % nc);
                int tx, ty, img;
                tile_coords(it, tx, ty, img);
                const uint32_t fs = l % DP_F_STAGES;
                DP_TIMED(pt_fempty, mbar_wait_hint(&f_empty[fs], ((l / DP_F_STAGES) & 1u) ^ 1u, hi));
                expect_on_leader(&f_full[fs], (dbg & 64) ? 0u : DP_F_STAGE);
                if (!(dbg & 64)) tma_load_5d_pair(f_smem + fs * DP_F_STAGE, &PP.tmap_f, &f_full[fs], 0, tx * DC_TILE_W - 1, ty * DC_TILE_H - 1,
                                 (PP.f_ch_off + c * 32) >> 3, img);
            };
            if (f_total > 0) issue_f(0);
            for (int it = 0; it < n_iter; ++it)
                for (int h = 0; h < 2; ++h) {
                    const int wo_row0 = (h * 2 + static_cast<int>(rank)) * nc * 3 * (DP_WO_STAGE / 512);
                    for (int c = 0; c < nc; ++c, ++fl)
                        for (int t3 = 0; t3 < 3; ++t3, ++wo_it) {
                            const uint32_t s = wo_it % DP_WO_STAGES, ph = (wo_it / DP_WO_STAGES) & 1u;
                            DP_TIMED(pt_woempty, mbar_wait_hint(&wo_empty[s], ph ^ 1u, hi));
                            if (t3 == 1 && fl + 1 < f_total) issue_f(fl + 1);     // halo of the next chunk: its stage is free by now
                            expect_on_leader(&wo_full[s], (dbg & 4) ? 0u : DP_WO_STAGE);
                            if (!(dbg & 4)) tma_load_2d_pair(wo_smem + s * DP_WO_STAGE, &PP.tmap_wo, &wo_full[s], 0, wo_row0 + (c * 3 + t3) * (DP_WO_STAGE / 512));
                        }
                }
            DP_FLUSH(18, pt_woempty); DP_FLUSH(19, pt_fempty); DP_FLUSH(20, clock64() - pt_total);
        }
    } else if (warp == 24) {
        if (leader) {
            // ================= A issuer (even CTA): D_off[256 px, 112] = halo(feat) x Wo[half h], K = C x 9, into TMEM columns
            // [256 + 112 h, +112) of both CTAs.  Half h of tile T may be overwritten once both CTAs' gather warps have read
            // half h of tile T-1 (off_empty), which happens in the middle of tile T-1: this GEMM overlaps the gather.
            const uint32_t idesc_off = umma_idesc_f16(256, DP_OFF_HALF);
            const uint32_t f_hi = umma_desc_hi(DS_F_RP_X * 16), b_hi = umma_desc_hi(128);
            uint32_t f_it = 0, wo_it = 0;
            long long pt_total = clock64(), pt_offempty = 0, pt_ffull = 0, pt_wofull = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int h = 0; h < 2; ++h) {
                    if (it > 0) DP_TIMED(pt_offempty, mbar_wait_hint(&off_empty[h], (it - 1) & 1u, hi));
                    tc_fence_after_sync();
                    const uint32_t d_off = tmem_base + TM_OFF + h * DP_OFF_HALF;
                    for (int c = 0; c < nc; ++c, ++f_it) {
                        const uint32_t fs = f_it % DP_F_STAGES, fph = (f_it / DP_F_STAGES) & 1u;
                        DP_TIMED(pt_ffull, mbar_wait_hint(&f_full[fs], fph, hc));
                        const uint32_t f_lo0 = umma_desc_lo(smem_u32(f_smem + fs * DP_F_STAGE), DS_F_PLANE);
                        for (int ki = 0; ki < 3; ++ki, ++wo_it) {                  // one stage = the three taps of kernel row ki
                            const uint32_t ws = wo_it % DP_WO_STAGES, wph = (wo_it / DP_WO_STAGES) & 1u;
                            DP_TIMED(pt_wofull, mbar_wait_hint(&wo_full[ws], wph, hc));
                            tc_fence_after_sync();
                            const uint32_t w_lo0 = umma_desc_lo(smem_u32(wo_smem + ws * DP_WO_STAGE), DP_WO_ROWS * 16);
                            if (elect_one()) {
#pragma unroll
                                for (int kj = 0; kj < 3; ++kj)
#pragma unroll
                                    for (int k16 = 0; k16 < 2; ++k16)
                                        if (!(dbg & 2)) umma_f16_lohi<2>(d_off, f_lo0 + (ki * DS_F_RP_X + kj) + k16 * (2 * DS_F_PLANE / 16), f_hi,
                                                         w_lo0 + kj * (DP_WO_TAP / 16) + k16 * (2 * DP_WO_ROWS * 16 / 16), b_hi, idesc_off,
                                                         (c | ki | kj | k16) != 0 ? 1u : 0u);
                                umma_commit_pair(&wo_empty[ws], 3);
                                if (ki == 2) umma_commit_pair(&f_empty[fs], 3);
                                if (ki == 2 && c == nc - 1) umma_commit_pair(&off_full[h], 3);
                            }
                            __syncwarp();
                        }
                    }
                }
            DP_FLUSH(11, pt_offempty); DP_FLUSH(12, pt_ffull); DP_FLUSH(13, pt_wofull); DP_FLUSH(14, clock64() - pt_total);
        }
    } else {
        // ================= gather warps: lane = output pixel m = 32 q + lane (q = TMEM lane quarter of this warp); K-atom pair kp
        // covers channels [chunk * 32 + 16 kp, +16): one deformable group when C / dg >= 16, two when it is 8.  Warp group wg
        // (0 / 1) takes the stages of its parity: stage st = 9 * chunk + tap of the tile, st % 2 == wg.
        const int q = warp & 3, j4 = (warp - 6) >> 2, kp = j4 & 1, wg = j4 >> 1;
        const int m = 32 * q + lane;
        const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo, cpg = P.cpg;
        const float fH = static_cast<float>(H), fW = static_cast<float>(W);
        const int ixps = P.x_pix_stride, ixrow = W * ixps;
        const __half* const xview = P.x + P.x_ch_off;
        const bool wide = P.x_wide != 0;             // every (pixel, 16-channel pair) of the x view is 32-byte aligned
        const __half* const zbuf = reinterpret_cast<const __half*>(dcn_zero32);
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + TM_OFF;
        const uint32_t bo_sa = smem_u32(bo_s);
        const uint32_t a_dst0 = smem_u32(a_smem) + (2 * kp) * DP_A_LBO + m * 16;
        const uint32_t atom0 = 2u * kp;
        const int gph = P.dg >> 1;                    // deformable groups per offset half
        float abs_sum = 0.f;

        struct Raw { float dh, dw, mk; };
        // sampling geometry of one (pixel, group, tap): shared-memory addresses of the two upper corners (K atom 2 kp), the four
        // bilinear x mask weights as packed halves, floor(h_im) / floor(w_im) packed 16:16 for the global fallback
        struct Geo { uint32_t a0, a1; __half2 w[4]; int hw; bool slow; };

        long long pt_total = clock64(), pt_offfull = 0, pt_winfull = 0, pt_empty = 0, pt_fetch = 0;
        uint32_t git0 = 0;                            // global stage index of this tile's stage 0
        for (int it = 0; it < n_iter; ++it, git0 += nstages) {
            int tx, ty, img;
            tile_coords(it, tx, ty, img);
            const __half* const ximg = xview + static_cast<long long>(img) * H * ixrow;
            const int ho = ty * DC_TILE_H + (m >> 3), wo = tx * DC_TILE_W + (m & 7);
            const bool ok = DP_LIVE(it) && (ho < Ho) && (wo < Wo);
            const float hbf = static_cast<float>(ho - 1), wbf = static_cast<float>(wo - 1);
            const int wy0 = ty * DC_TILE_H - 1 - DP_RY, wx0 = tx * DC_TILE_W - 1 - DP_XL;
            const bool count_tile = PP.absmean != nullptr && ok;

            // (dh, dw, mask logit) of the TMEM column triple at `col` (relative to TM_OFF): issue, wait, add the bias
            auto fetch_issue = [&](int col, uint32_t (&v)[4]) { tmem_ld4_nowait(tm_lane + col, v); };
            auto fetch_wait = [&](uint32_t (&a)[4], uint32_t (&b)[4]) {
                // the registers are operands so that no use of them can be scheduled above the wait
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]) :: "memory");
            };
            auto fetch_finish = [&](int col, const uint32_t (&v)[4]) -> Raw {
                Raw r;
                r.dh = __uint_as_float(v[0]) + __uint_as_float(lds_u32(bo_sa + col * 4));
                r.dw = __uint_as_float(v[1]) + __uint_as_float(lds_u32(bo_sa + col * 4 + 4));
                r.mk = __uint_as_float(v[2]) + __uint_as_float(lds_u32(bo_sa + col * 4 + 8));
                return r;
            };
            // reference semantics of deform_conv_cuda_kernel.cu:467-497,614-628
            auto geometry = [&](const Raw& r, int tap, uint32_t win) -> Geo {
                Geo gq;
                const int ki = tap / 3, kj = tap - 3 * ki;
                const float h_im = hbf + static_cast<float>(ki) + r.dh;
                const float w_im = wbf + static_cast<float>(kj) + r.dw;
                const bool valid = ok && h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
                const float hf = floorf(h_im), wf = floorf(w_im);
                const float lh = h_im - hf, lw = w_im - wf;
                const float mk = valid ? sigmoidf_fast(r.mk) : 0.f;
                const float a = (1.f - lh) * mk, b = lh * mk, hw = 1.f - lw;
                gq.w[0] = __float2half2_rn(a * hw); gq.w[1] = __float2half2_rn(a * lw);
                gq.w[2] = __float2half2_rn(b * hw); gq.w[3] = __float2half2_rn(b * lw);
                const int hl = valid ? static_cast<int>(hf) : 0, wl = valid ? static_cast<int>(wf) : 0;
                gq.hw = (hl << 16) | (wl & 0xffff);
                const int ry = hl - wy0, rx = wl - wx0;
                const bool inwin = static_cast<unsigned>(ry) <= static_cast<unsigned>(DP_WH - 2) &&
                                   static_cast<unsigned>(rx) <= static_cast<unsigned>(DP_WW - 2);
                gq.slow = valid && !inwin;
                const uint32_t p0 = inwin ? static_cast<uint32_t>(ry * DP_WW + rx) : 0u;        // invalid samples: weights are 0
                // window pixel p sits at win + p * 64, its 16-byte K atom c at ((c ^ ((p >> 1) & 3)) << 4) (64-byte TMA swizzle);
                // the row below (p + DP_WW, DP_WW % 8 == 0) has the same swizzle phase
                gq.a0 = win + (p0 << 6) + ((atom0 ^ ((p0 >> 1) & 3u)) << 4);
                gq.a1 = win + ((p0 + 1) << 6) + ((atom0 ^ (((p0 + 1) >> 1) & 3u)) << 4);
                return gq;
            };
            auto blend = [&](const Geo& gq, const uint4& u0, const uint4& u1, const uint4& u2, const uint4& u3) -> uint4 {
                uint4 out;
                const uint32_t* pa = reinterpret_cast<const uint32_t*>(&u0);
                const uint32_t* pb = reinterpret_cast<const uint32_t*>(&u1);
                const uint32_t* pc = reinterpret_cast<const uint32_t*>(&u2);
                const uint32_t* pd = reinterpret_cast<const uint32_t*>(&u3);
                uint32_t* po = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __half2 acc = __hmul2(gq.w[0], *reinterpret_cast<const __half2*>(&pa[i]));
                    acc = __hfma2(gq.w[1], *reinterpret_cast<const __half2*>(&pb[i]), acc);
                    acc = __hfma2(gq.w[2], *reinterpret_cast<const __half2*>(&pc[i]), acc);
                    acc = __hfma2(gq.w[3], *reinterpret_cast<const __half2*>(&pd[i]), acc);
                    po[i] = *reinterpret_cast<uint32_t*>(&acc);
                }
                return out;
            };
            // Far samples (offset beyond the staged window): the lane replaces the four corners it read from shared memory by
            // global loads (zero buffer for corners outside the image - the reference's border rule).  Only the far lanes take
            // the branch; the shared-memory reads and the blend stay warp-uniform.
            auto far_corners = [&](const Geo& gq, int ch, uint4& u0, uint4& u1, uint4& u2, uint4& u3, uint4& t0, uint4& t1, uint4& t2,
                                   uint4& t3, bool both) {
                const int hl = gq.hw >> 16, wl = static_cast<int>(static_cast<short>(gq.hw & 0xffff));
                const bool t = hl >= 0, b = hl + 1 <= H - 1, l = wl >= 0, r = wl + 1 <= W - 1;
                const __half* base = ximg + hl * ixrow + wl * ixps + ch;
                const __half* p0 = (t && l) ? base : zbuf;
                const __half* p1 = (t && r) ? base + ixps : zbuf;
                const __half* p2 = (b && l) ? base + ixrow : zbuf;
                const __half* p3 = (b && r) ? base + ixrow + ixps : zbuf;
                if (both && wide) {          // the K-atom pair of a corner is 32 contiguous, aligned bytes
                    ldg_nc_v8(p0, u0, t0); ldg_nc_v8(p1, u1, t1); ldg_nc_v8(p2, u2, t2); ldg_nc_v8(p3, u3, t3);
                } else {
                    u0 = ldg_nc_v4(p0); u1 = ldg_nc_v4(p1); u2 = ldg_nc_v4(p2); u3 = ldg_nc_v4(p3);
                    if (both) {
                        t0 = ldg_nc_v4(p0 == zbuf ? zbuf : p0 + 8); t1 = ldg_nc_v4(p1 == zbuf ? zbuf : p1 + 8);
                        t2 = ldg_nc_v4(p2 == zbuf ? zbuf : p2 + 8); t3 = ldg_nc_v4(p3 == zbuf ? zbuf : p3 + 8);
                    }
                }
            };

            // The warp's stages of this tile, chunk by chunk: st = 9 * chunk + tap with st % 2 == wg, i.e. inside a chunk the taps
            // t0, t0 + 2, ... with t0 = (wg + chunk) & 1.  The loop is software pipelined inside a chunk: while the corners of tap t
            // are in flight (LDS) the offsets of tap t + 2 - fetched from tensor memory one stage ahead - are turned into its
            // geometry, so a warp's dependent chain per stage is the sampling only (this role is bound by per-warp latency: 184
            // instructions per stage at ~7 cycles each with 4 gather warps per scheduler, profiles/r02_ncu_dcn_pair_v2).
            for (int chunk = 0; chunk < nc; ++chunk) {
                const uint32_t G = static_cast<uint32_t>(it) * nc + chunk, wb = G & 1u;
                const int h = chunk >= nch ? 1 : 0;
                const int ch0 = chunk * 32 + kp * 16;                       // first channel of this warp's K-atom pair
                const int g0 = ch0 / cpg;
                const int col0 = h * DP_OFF_HALF + (g0 - h * gph) * 27;
                const bool count_abs = count_tile && (ch0 % cpg) == 0;      // each (pixel, group, tap) offset exactly once
                const uint32_t win = smem_u32(win_smem + wb * DP_WIN_BYTES);
                if (chunk == 0 || chunk == nch) { DP_TIMED(pt_offfull, mbar_wait_hint(&off_full[h], it & 1u, hc)); tc_fence_after_sync(); }
                DP_TIMED(pt_winfull, mbar_wait_hint(&win_full[wb], (G >> 1) & 1u, hc));

                int tap = (wg + chunk) & 1;
                uint32_t pv0[4], pv1[4] = {0u, 0u, 0u, 0u};
                fetch_issue(col0 + 3 * tap, pv0);
                if (TWO) fetch_issue(col0 + 27 + 3 * tap, pv1);
                fetch_wait(pv0, pv1);
                Raw r0 = fetch_finish(col0 + 3 * tap, pv0);
                Raw r1 = TWO ? fetch_finish(col0 + 27 + 3 * tap, pv1) : r0;
                if (tap + 2 < 9) { fetch_issue(col0 + 3 * (tap + 2), pv0); if (TWO) fetch_issue(col0 + 27 + 3 * (tap + 2), pv1); }
                if (count_abs) abs_sum += fabsf(r0.dh) + fabsf(r0.dw) + (TWO ? fabsf(r1.dh) + fabsf(r1.dw) : 0.f);
                Geo q0 = geometry(r0, tap, win);
                Geo q1 = TWO ? geometry(r1, tap, win) : q0;
                for (; tap < 9; tap += 2) {
                    const uint32_t git = git0 + 9 * chunk + tap, as = git % DP_A_STAGES, pa = (git / DP_A_STAGES) & 1u;
                    const bool more = tap + 2 < 9;
                    const bool any_slow = __any_sync(0xffffffffu, q0.slow || (TWO && q1.slow));
                    uint4 v0, v1;
                    Geo n0 = q0, n1 = q1;
                    {
                        // eight LDS.128 per lane (window pixel 0 for the lanes whose sample is far or invalid: harmless), conflict
                        // free; far lanes then swap in their global corners
                        const uint32_t b0 = q0.a0, b1 = q0.a1, c0 = (TWO ? q1.a0 : q0.a0) ^ 16u, c1 = (TWO ? q1.a1 : q0.a1) ^ 16u;
                        uint4 u0 = lds_v4(b0), u1 = lds_v4(b1), u2 = lds_v4(b0 + DP_WW * 64), u3 = lds_v4(b1 + DP_WW * 64);
                        uint4 t0 = lds_v4(c0), t1 = lds_v4(c1), t2 = lds_v4(c0 + DP_WW * 64), t3 = lds_v4(c1 + DP_WW * 64);
                        if (more) {          // geometry of the warp's next stage while the corners are in flight
                            fetch_wait(pv0, pv1);
                            r0 = fetch_finish(col0 + 3 * (tap + 2), pv0);
                            if (TWO) r1 = fetch_finish(col0 + 27 + 3 * (tap + 2), pv1);
                            if (tap + 4 < 9) { fetch_issue(col0 + 3 * (tap + 4), pv0); if (TWO) fetch_issue(col0 + 27 + 3 * (tap + 4), pv1); }
                            n0 = geometry(r0, tap + 2, win);
                            if (TWO) n1 = geometry(r1, tap + 2, win);
                        }
                        if (any_slow) {
                            if (TWO) {
                                uint4 d0, d1, d2, d3;
                                if (q0.slow) far_corners(q0, ch0, u0, u1, u2, u3, d0, d1, d2, d3, false);
                                if (q1.slow) far_corners(q1, ch0 + 8, t0, t1, t2, t3, d0, d1, d2, d3, false);
                            } else if (q0.slow) {
                                far_corners(q0, ch0, u0, u1, u2, u3, t0, t1, t2, t3, true);
                            }
                        }
                        v0 = blend(q0, u0, u1, u2, u3);
                        v1 = blend(TWO ? q1 : q0, t0, t1, t2, t3);
                    }
                    if (more && count_abs) abs_sum += fabsf(r0.dh) + fabsf(r0.dw) + (TWO ? fabsf(r1.dh) + fabsf(r1.dw) : 0.f);
                    DP_TIMED(pt_empty, mbar_wait_hint(&empty[as], pa ^ 1u, hg));
                    sts_v4(a_dst0 + as * DP_A_STAGE, v0);
                    sts_v4(a_dst0 + as * DP_A_STAGE + DP_A_LBO, v1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&gathered[as]);
                    q0 = n0; q1 = n1;
                }
                // leaving the chunk: its window buffer is free; after the last chunk of an offset half, so is the half
                __syncwarp();
                if (lane == 0) mbar_arrive(&win_empty[wb]);
                if (chunk == nch - 1 || chunk == nc - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) { if (leader) mbar_arrive(&off_empty[h]); else mbar_arrive_remote_cta(&off_empty[h], 0); }
                }
            }
        }
        if (PP.absmean != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, o);
            if (lane == 0) atomicAdd(PP.absmean, abs_sum);
        }
        DP_FLUSH(0, clock64() - pt_total); DP_FLUSH(1, pt_offfull); DP_FLUSH(2, pt_winfull); DP_FLUSH(3, pt_empty); DP_FLUSH(4, pt_fetch);
    }

#undef DP_TILE_OF
#undef DP_LIVE
    // ---- teardown: nobody leaves while the peer may still read this CTA's shared memory or signal its barriers
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace eb
