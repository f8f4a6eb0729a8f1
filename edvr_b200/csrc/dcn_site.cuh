// dcn_site.cuh — DCNv2 forward for 3x3 / stride 1 / pad 1 / dilation 1 (every DCN of EDVR) with the sampled input
// STAGED IN SHARED MEMORY, optionally with the site's conv_offset convolution fused in front (arch_util.py:243-257).
//
// Replaces (paths under /root/reference/basicsr/models/):
//   DCNv2Pack.forward ........................ archs/arch_util.py:243-257   (FUSED: conv_offset + chunk/cat + sigmoid + dcn)
//   modulated_deform_conv_cuda_forward ....... ops/dcn/src/deform_conv_cuda.cpp:490-569
//   modulated_deformable_im2col_gpu_kernel ... ops/dcn/src/deform_conv_cuda_kernel.cu:570-633 (+ bilinear :467-497)
//
// Why a second DCN kernel (dcn_fused.cuh stays for general stride / dilation / kernel size):
// profiles/r01_ncu_dcn_v5_summary.csv showed the first kernel bound by L1 wavefronts of the per-corner GLOBAL gathers
// (25.5 K of 30 K LSU wavefronts per 128-pixel tile).  Here, per 16x8-pixel output tile and 64-channel chunk, ONE TMA box
// {64 ch, 17 x, 25 y} (the tile + the 3x3 reach + DS_R pixels of offset on every side, zero-filled outside the image by
// the copy engine, 128-byte swizzle) lands in shared memory, and the 4 x 9 bilinear corners of every (pixel, group) are
// served from there with conflict-free LDS.128; samples whose offset leaves the window fall back to global loads lane by
// lane.  The zero fill IS the reference's border rule (corners outside [0,H-1]x[0,W-1] contribute 0,
// deform_conv_cuda_kernel.cu:480-491), so the fast path carries no per-corner validity logic.
//
// Thread mapping: lane = output pixel (TMEM lane = accumulator row), 16 gather warps = 4 TMEM lane quarters x 4 K-atom
// pairs of the chunk.  That is what lets the FUSED variant read the offsets straight out of TENSOR MEMORY: phase A
// computes conv_offset for the tile on the tensor cores (A = halo tile of the offset features by TMA, B = conv_offset
// weights [g][tap][dh, dw, mask] streamed by bulk copies) into 224 TMEM columns; the gather lane of pixel m reads its
// (dh, dw, mask logit) triple with one tcgen05.ld, adds the bias, applies the sigmoid and samples.  Offsets and masks
// never exist in HBM (the fp16 record of the first design cost 826 MB of traffic per N=28 L1 call and the precision of
// multi-pixel offsets), and one launch replaces two.
// Sampling arithmetic is fp32 throughout (coordinates, bilinear weights, blend); only the gathered column is rounded to
// fp16 as the tensor-core operand.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dcn_fused.cuh"
#include "epilogue.cuh"

namespace eb {

constexpr int DS_R = 3;                                   // window margin (pixels of offset served from shared memory)
constexpr int DS_WH = DC_TILE_H + 3 + 2 * DS_R;           // 25 rows: floor(h_im) in [y0-1-R, y0+16+R], +1 for the low corner
constexpr int DS_WW = DC_TILE_W + 3 + 2 * DS_R;           // 17
constexpr int DS_WIN_BYTES = ((DS_WH * DS_WW * 128 + 1023) / 1024) * 1024;       // 55296: 1024-byte aligned for the swizzle
constexpr int DS_WIN_TX = DS_WH * DS_WW * 128;            // bytes the TMA box delivers
constexpr int DS_STAGES = 3;
constexpr int DS_GATHER_WARPS = 16;
constexpr int DS_THREADS = 32 * (6 + DS_GATHER_WARPS + 2);    // warps: 0 weights, 1 MMA, 2-5 epilogue, 6-21 gather, 22 forwarder, 23 windows
constexpr int DS_AB_BYTES = DS_STAGES * (DC_A_BYTES + DC_B_BYTES);               // 98688: gather/MMA stages
constexpr int DS_SMEM_BYTES = 2 * DS_WIN_BYTES + DS_AB_BYTES + DC_MAX_COUT * 4 + 256 * 4 + 512;

// ---- phase A (fused conv_offset) staging, aliased with the gather/MMA stages (the phases alternate per tile)
constexpr int DS_OFF_N = 224;                             // conv_offset rows padded to the MMA N granularity (dg * 27 <= 224)
constexpr int DS_F_RP_Y = DC_TILE_H + 2, DS_F_RP_X = DC_TILE_W + 2;              // 18 x 10 halo of the offset features
constexpr int DS_F_PLANE = DS_F_RP_Y * DS_F_RP_X * 16;    // 2880: one K atom of every halo pixel
constexpr int DS_F_STAGE = 4 * DS_F_PLANE;                // 32 channels per stage = 11520
constexpr int DS_F_STAGES = 3;
constexpr int DS_WO_STAGE = DS_OFF_N * 64;                // (32 channels, one tap): 224 rows x 64 B = 14336
constexpr int DS_WO_STAGES_A = 4;                         // conv_offset weight stages under the gather/MMA stages ...
constexpr int DS_WO_STAGES_B = 3;                         // ... and in window buffer 1, idle until the tile's second chunk
constexpr int DS_WO_STAGES = DS_WO_STAGES_A + DS_WO_STAGES_B;    // 100 KB of weights in flight: phase A streams 516 KB per
                                                          // tile from L2 and was latency-bound with 57 KB (r02_dcn_site_v2)
static_assert(DS_F_STAGES * DS_F_STAGE + DS_WO_STAGES_A * DS_WO_STAGE <= DS_AB_BYTES, "phase A staging must fit under phase B");
static_assert(DS_WO_STAGES_B * DS_WO_STAGE <= DS_WIN_BYTES, "extra weight stages must fit in the second window buffer");

enum : int { DS_OFF_GLOBAL = 0, DS_OFF_TMEM = 1 };

struct DsParams {
    DcnParams d;               // x view, shapes, dg/cpg, offset/mask pointers (DS_OFF_GLOBAL), packed DCN weights, epilogue
    CUtensorMap tmap_x;        // x as {pix_stride, W, H, N} fp16, box {64, DS_WW, DS_WH, 1}, 128-byte swizzle, zero fill
    long long off_img_stride, mask_img_stride;     // DS_OFF_GLOBAL: elements between images of d.offset / d.mask
    int mask_logit;            // DS_OFF_GLOBAL: 1 = d.mask holds logits (sigmoid applied here), 0 = probabilities
    float* absmean;            // optional: += sum |offset| over the call (n-tile 0 only), for the ">50" warning
    int multicast;             // 1: launched as clusters of 2 CTAs that share ONE L2 read of every weight stage (bulk-copy multicast)
    // ---- DS_OFF_TMEM
    CUtensorMap tmap_f;        // offset features as {8, W, H, pix_stride/8, N}, box {8, 10, 18, 4, 1} (K-atom planes)
    int f_ch_off;              // first channel of the offset-feature view
    const __half* wo_pack;     // conv_offset weights [chunk32][tap][k16 2][plane 2][224][8] fp16
    const float* bo;           // conv_offset bias in column order [g][tap][dh, dw, mask], DS_OFF_N entries
};

__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst_smem, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "r"(c4)
        : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}
// this thread's TMEM lane, 4 consecutive fp32 columns; the caller waits (tmem_ld_wait) before using the values
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


template <int OFFSRC, int EK, bool TWO>
__global__ void __launch_bounds__(DS_THREADS, 1) dcn_site_kernel(const __grid_constant__ DsParams PP) {
    const DcnParams& P = PP.d;
    extern __shared__ __align__(1024) uint8_t smem[];
    // the 128-byte swizzle of the window boxes needs a 1024-byte aligned destination: dynamic shared memory starts at the CTA's
    // window base (no static shared memory in this kernel); checked, not assumed
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* win_smem = smem;                                   // [2][DS_WIN_BYTES]
    uint8_t* a_smem = smem + 2 * DS_WIN_BYTES;                  // [DS_STAGES][DC_A_BYTES]
    uint8_t* b_smem = a_smem + DS_STAGES * DC_A_BYTES;          // [DS_STAGES][DC_B_BYTES]
    uint8_t* f_smem = a_smem;                                   // phase A: [DS_F_STAGES][DS_F_STAGE] (aliases a/b stages)
    uint8_t* wo_smem = a_smem + DS_F_STAGES * DS_F_STAGE;       // phase A: [DS_WO_STAGES_A][DS_WO_STAGE] + [DS_WO_STAGES_B] in window 1
    auto wo_stage = [&](uint32_t st) -> uint8_t* {
        return st < DS_WO_STAGES_A ? wo_smem + st * DS_WO_STAGE : win_smem + DS_WIN_BYTES + (st - DS_WO_STAGES_A) * DS_WO_STAGE;
    };
    float* bias_s = reinterpret_cast<float*>(a_smem + DS_AB_BYTES);
    float* bo_s = bias_s + DC_MAX_COUT;                         // [256] conv_offset bias (fused)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bo_s + 256);
    uint64_t* full = bars;                        // [S] forwarder arrival + weights (expect_tx)
    uint64_t* empty = full + DS_STAGES;           // [S] MMA commit
    uint64_t* gathered = empty + DS_STAGES;       // [S] one arrival per gather warp
    uint64_t* win_full = gathered + DS_STAGES;    // [2] TMA bytes
    uint64_t* win_empty = win_full + 2;           // [2] one arrival per gather warp
    uint64_t* acc_full = win_empty + 2;           // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2] 4 epilogue warps
    uint64_t* f_full = acc_empty + 2;             // [DS_F_STAGES]   phase A: feature halo stage landed
    uint64_t* f_empty = f_full + DS_F_STAGES;     // [DS_F_STAGES]
    uint64_t* wo_full = f_empty + DS_F_STAGES;    // [DS_WO_STAGES]
    uint64_t* wo_empty = wo_full + DS_WO_STAGES;  // [DS_WO_STAGES]
    uint64_t* off_full = wo_empty + DS_WO_STAGES; // [1] phase A accumulator complete (commit)
    uint64_t* off_empty = off_full + 1;           // [1] every gather warp has read its offsets of this tile
    uint64_t* ab_free = off_empty + 1;            // [1] phase B of the previous tile no longer touches the aliased stages
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ab_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (P.Wo + DC_TILE_W - 1) / DC_TILE_W;
    const int tiles_y = (P.Ho + DC_TILE_H - 1) / DC_TILE_H;
    const int total_tiles = P.N * tiles_y * tiles_x * P.n_tiles_n;
    const int nchunks = P.C / 64;
    const int nstages = 9 * nchunks;
    const uint32_t b_bytes = static_cast<uint32_t>(P.BN) * 128u;
    constexpr bool FUSED = OFFSRC == DS_OFF_TMEM;
    constexpr uint32_t TM_OFF = 256;              // TMEM columns: [0,128) [128,256) DCN accumulators, [256,480) conv_offset

    const int cout_packed = P.BN * P.n_tiles_n;
    const bool has_bias = P.epi.bias != nullptr;
    if (has_bias)
        for (int i = threadIdx.x; i < cout_packed; i += blockDim.x) bias_s[i] = P.epi.bias[i];
    if (FUSED)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) bo_s[i] = i < DS_OFF_N ? PP.bo[i] : 0.f;
    // Weight multicast (clusters of 2): both CTAs walk the same weight-stage sequence; the leader's producer writes every stage
    // into BOTH CTAs with one multicast bulk copy, so a stage is re-used only when both CTAs' MMAs have consumed it: every
    // "stage consumed" / "aliased region free" commit is delivered to both CTAs and those barriers expect two arrivals.
    const bool mc = PP.multicast != 0;
    const uint32_t crank = mc ? cluster_ctarank() : 0u;
    const uint32_t both = mc ? 2u : 1u;
    // tiles per CTA: in a cluster both CTAs must run the same number of rounds (a CTA without a tile runs a dead one)
    const int n_iter = mc ? (total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                          : (total_tiles > static_cast<int>(blockIdx.x)
                                 ? (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0);
    auto commit = [&](uint64_t* bar) { if (mc) umma_commit_multicast(bar, 3); else umma_commit(bar); };
    if (threadIdx.x == 0) {
        for (int i = 0; i < DS_STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], both); mbar_init(&gathered[i], DS_GATHER_WARPS); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&win_full[i], 1); mbar_init(&win_empty[i], DS_GATHER_WARPS);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < DS_F_STAGES; ++i) { mbar_init(&f_full[i], 1); mbar_init(&f_empty[i], 1); }
        for (int i = 0; i < DS_WO_STAGES; ++i) { mbar_init(&wo_full[i], 1); mbar_init(&wo_empty[i], both); }
        mbar_init(off_full, both); mbar_init(off_empty, DS_GATHER_WARPS); mbar_init(ab_free, both);
        fence_barrier_init();
        tma_prefetch_desc(&PP.tmap_x);
        if (FUSED) tma_prefetch_desc(&PP.tmap_f);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    if (mc) cluster_sync_all();           // both CTAs' barriers exist before any multicast copy or commit targets them
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
#define DS_TILE_LOOP(IT_)                                                                                          \
    for (int IT_ = 0; IT_ < n_iter; ++IT_)
#define DS_TILE_OF(IT_) (blockIdx.x + (IT_) * gridDim.x < static_cast<unsigned>(total_tiles) ? static_cast<int>(blockIdx.x + (IT_) * gridDim.x) : 0)
#define DS_LIVE(IT_) (blockIdx.x + (IT_) * gridDim.x < static_cast<unsigned>(total_tiles))

    if (warp == 0) {
        // ================= weight producer.  Phase A (fused): conv_offset weight stages, one per (32-channel chunk, tap);
        // phase B: DCN weight stages, one per (64-channel chunk, tap).  Both rings alias the same shared memory, so phase A
        // of a tile starts only when the MMAs of the previous tile's phase B have drained (ab_free).
        if (lane == 0) {
            uint32_t it = 0, wo_it = 0;
            auto copy = [&](void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
                if (!mc) bulk_g2s(dst, src, bytes, bar);
                else if (crank == 0) bulk_g2s_multicast(dst, src, bytes, bar, 3);      // the peer only arms its barrier
            };
            DS_TILE_LOOP(tile_it) {
                const int tile = DS_TILE_OF(tile_it);
                if (FUSED) {
                    if (tile_it > 0) mbar_wait_t<64>(ab_free, (tile_it - 1) & 1u);
                    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(PP.wo_pack);
                    for (int st = 0; st < 2 * nstages; ++st, ++wo_it) {
                        const uint32_t s = wo_it % DS_WO_STAGES, ph = (wo_it / DS_WO_STAGES) & 1u;
                        mbar_wait_t<64>(&wo_empty[s], ph ^ 1u);
                        mbar_arrive_expect_tx(&wo_full[s], DS_WO_STAGE);
                        copy(wo_stage(s), wsrc + static_cast<size_t>(st) * DS_WO_STAGE, DS_WO_STAGE, &wo_full[s]);
                    }
                    // phase B stages may be refilled once phase A's MMAs have read the aliased bytes
                    mbar_wait_t<64>(off_full, tile_it & 1u);
                }
                const int nt = tile % P.n_tiles_n;
                const uint8_t* w = reinterpret_cast<const uint8_t*>(P.wpack) + static_cast<size_t>(nt) * nstages * b_bytes;
                for (int st = 0; st < nstages; ++st, ++it) {
                    const uint32_t s = it % DS_STAGES, ph = (it / DS_STAGES) & 1u;
                    mbar_wait_t<64>(&empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[s], b_bytes);
                    copy(b_smem + s * DC_B_BYTES, w + static_cast<size_t>(st) * b_bytes, b_bytes, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (converged warp, elected lane: descriptors stay in uniform registers)
        const uint32_t idesc = umma_idesc_f16(128, P.BN);
        const uint32_t idesc_off = umma_idesc_f16(128, DS_OFF_N);
        const uint32_t lbo_b = static_cast<uint32_t>(P.BN) * 16u;
        const uint32_t a_hi = umma_desc_hi(128), b_hi = umma_desc_hi(128), f_hi = umma_desc_hi(DS_F_RP_X * 16);
        uint32_t it = 0, f_it = 0, wo_it = 0;
        DS_TILE_LOOP(acc_it_) {
            const uint32_t acc_it = acc_it_;
            if (FUSED) {
                // ---- phase A: D_off[128 px, 224] = halo(feat) x Wo, K = C x 9, into TMEM columns [256, 480)
                if (acc_it > 0) mbar_wait_t<64>(off_empty, (acc_it - 1) & 1u);       // the gather has read the previous tile's offsets
                tc_fence_after_sync();
                const uint32_t d_off = tmem_base + TM_OFF;
                for (int c = 0; c < 2 * nchunks; ++c, ++f_it) {
                    const uint32_t fs = f_it % DS_F_STAGES;
                    mbar_wait_t<64>(&f_full[fs], (f_it / DS_F_STAGES) & 1u);
                    tc_fence_after_sync();
                    const uint32_t f_lo0 = umma_desc_lo(smem_u32(f_smem + fs * DS_F_STAGE), DS_F_PLANE);
                    for (int t = 0; t < 9; ++t, ++wo_it) {
                        const uint32_t ws = wo_it % DS_WO_STAGES;
                        mbar_wait_t<64>(&wo_full[ws], (wo_it / DS_WO_STAGES) & 1u);
                        tc_fence_after_sync();
                        const uint32_t w_lo0 = umma_desc_lo(smem_u32(wo_stage(ws)), DS_OFF_N * 16);
                        const int ki = t / 3, kj = t % 3;
                        if (elect_one()) {
#pragma unroll
                            for (int k16 = 0; k16 < 2; ++k16)
                                umma_f16_lohi<1>(d_off, f_lo0 + (ki * DS_F_RP_X + kj) + k16 * (2 * DS_F_PLANE / 16), f_hi,
                                                 w_lo0 + k16 * (2 * DS_OFF_N * 16 / 16), b_hi, idesc_off, (c | t | k16) != 0 ? 1u : 0u);
                            commit(&wo_empty[ws]);
                            if (t == 8) umma_commit(&f_empty[fs]);
                            if (t == 8 && c == 2 * nchunks - 1) commit(off_full);
                        }
                        __syncwarp();
                    }
                }
            }
            // ---- phase B: D[128 px, BN] = gathered columns x W, K = C x 9
            const uint32_t ab = acc_it & 1u;
            mbar_wait_t<64>(&acc_empty[ab], ((acc_it >> 1) & 1u) ^ 1u);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + ab * 128u;
            for (int st = 0; st < nstages; ++st, ++it) {
                const uint32_t s = it % DS_STAGES, ph = (it / DS_STAGES) & 1u;
                mbar_wait_t<64>(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_smem + s * DC_A_BYTES), DC_A_LBO);
                const uint32_t b_lo0 = umma_desc_lo(smem_u32(b_smem + s * DC_B_BYTES), lbo_b);
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        umma_f16_lohi<1>(d, a_lo0 + k16 * (2 * DC_A_LBO / 16), a_hi, b_lo0 + k16 * ((2u * lbo_b) >> 4), b_hi, idesc,
                                         (st | k16) != 0 ? 1u : 0u);
                    commit(&empty[s]);
                    if (st == nstages - 1) { umma_commit(&acc_full[ab]); if (FUSED) commit(ab_free); }
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ================= epilogue: warps 2..5 -> TMEM lane quarters 2,3,0,1
        const int q = warp & 3;
        DS_TILE_LOOP(acc_it_) {
            const uint32_t acc_it = acc_it_;
            const int tile = DS_TILE_OF(acc_it_);
            const int nt = tile % P.n_tiles_n, pt = tile / P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            const uint32_t ab = acc_it & 1u;
            mbar_wait_t<256>(&acc_full[ab], (acc_it >> 1) & 1u);
            tc_fence_after_sync();
            const int y = ty * DC_TILE_H + 4 * q + (lane >> 3);
            const int x = tx * DC_TILE_W + (lane & 7);
            const bool valid = DS_LIVE(acc_it_) && (y < P.Ho) && (x < P.Wo);
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + ab * 128u;
#pragma unroll 1
            for (int cc = 0; cc < P.BN; cc += 32) {
                float v[32];
                tmem_ld32(t0 + cc, v);
                epi_store32<EK>(P.epi, has_bias ? bias_s : nullptr, v, img, y, x, nt * P.BN + cc, valid);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
    } else if (warp == 22) {
        // ================= forwarder: every gather warp has written stage s -> generic->async proxy fence -> full[s]
        if (lane == 0) {
            uint32_t it = 0;
            DS_TILE_LOOP(k_)
                for (int st = 0; st < nstages; ++st, ++it) {
                    const uint32_t s = it % DS_STAGES, ph = (it / DS_STAGES) & 1u;
                    mbar_wait_t<64>(&gathered[s], ph);
                    fence_proxy_async_smem();
                    mbar_arrive(&full[s]);
                }
        }
    } else if (warp == 23) {
        // ================= TMA producer: sampling windows of x (one per tile and 64-channel chunk, two buffers) and, fused,
        // the halo stages of the offset features for phase A
        if (lane == 0) {
            uint32_t used0 = 0, used1 = 0, f_it = 0;  // window buffer = chunk & 1 (buffer 1 doubles as weight staging in phase A)
            DS_TILE_LOOP(tile_it) {
                const int tile = DS_TILE_OF(tile_it);
                const int pt = tile / P.n_tiles_n;
                const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
                const int wx0 = tx * DC_TILE_W - 1 - DS_R, wy0 = ty * DC_TILE_H - 1 - DS_R;
                int c_win = 0;
                auto issue_window = [&]() {
                    const uint32_t wb = c_win & 1;
                    mbar_wait_t<64>(&win_empty[wb], ((wb ? used1 : used0) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&win_full[wb], DS_WIN_TX);
                    tma_load_4d(win_smem + wb * DS_WIN_BYTES, &PP.tmap_x, &win_full[wb], P.x_ch_off + c_win * 64, wx0, wy0, img);
                    if (wb) ++used1; else ++used0;
                    ++c_win;
                };
                if (FUSED) {
                    // the first window of the tile does not alias anything: fetch it while phase A runs
                    issue_window();
                    if (tile_it > 0) mbar_wait_t<64>(ab_free, (tile_it - 1) & 1u);
                    for (int c = 0; c < 2 * nchunks; ++c, ++f_it) {
                        const uint32_t fs = f_it % DS_F_STAGES;
                        mbar_wait_t<64>(&f_empty[fs], ((f_it / DS_F_STAGES) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&f_full[fs], DS_F_STAGE);
                        tma_load_5d(f_smem + fs * DS_F_STAGE, &PP.tmap_f, &f_full[fs], 0, tx * DC_TILE_W - 1, ty * DC_TILE_H - 1,
                                    (PP.f_ch_off + c * 32) >> 3, img);
                    }
                    // window buffer 1 held conv_offset weight stages: free once phase A's MMAs have completed
                    if (nchunks > 1) mbar_wait_t<64>(off_full, tile_it & 1u);
                }
                while (c_win < nchunks) issue_window();
            }
        }
    } else {
        // ================= gather warps: lane = output pixel m = 32 q + lane (q = TMEM lane quarter of this warp), the warp's
        // K-atom pair kp covers channels [chunk * 64 + 16 kp, +16): one deformable group when C / dg >= 16, two when it is 8.
        // The loop over the nine taps is fully unrolled: stage index (tap % 3), tap coordinates and TMEM columns are
        // immediates.  This role is ISSUE bound (profiles/r02_dcn_site_v1: 384 instructions per warp and tap with an fp32
        // blend), hence the packed-half blend below: 32 HFMA2 instead of 64 conversions + 64 FFMA per 16 channels.
        static_assert(DS_STAGES == 3, "stage index of tap t is t % 3 only because 9 % DS_STAGES == 0");
        const int q = warp & 3, kp = (warp - 6) >> 2;
        const int m = 32 * q + lane;
        const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo, cpg = P.cpg;
        const float fH = static_cast<float>(H), fW = static_cast<float>(W);
        const int ixps = P.x_pix_stride, ixrow = W * ixps;
        const __half* const xview = P.x + P.x_ch_off;
        const long long plane = static_cast<long long>(Ho) * Wo;
        const bool logit = FUSED || PP.mask_logit != 0;
        const bool has_mask = FUSED || P.mask != nullptr;
        const bool wide = P.x_wide != 0;             // every (pixel, 16-channel pair) of the x view is 32-byte aligned
        const __half* const zbuf = reinterpret_cast<const __half*>(dcn_zero32);
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + TM_OFF;
        const uint32_t bo_sa = smem_u32(bo_s);
        const uint32_t a_dst0 = smem_u32(a_smem) + (2 * kp) * DC_A_LBO + m * 16;
        const uint32_t atom0 = 2u * kp;
        float abs_sum = 0.f;

        struct Raw { float dh, dw, mk; };
        struct Geo { uint32_t a0, a1, a2; __half2 w[4]; int hl, wl; bool slow; };

        uint32_t chunk_ctr = 0, used0 = 0, used1 = 0;
        DS_TILE_LOOP(tile_it) {
            const int tile = DS_TILE_OF(tile_it);
            const int pt = tile / P.n_tiles_n, nt = tile % P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            const __half* const ximg = xview + static_cast<long long>(img) * H * ixrow;
            const int ho = ty * DC_TILE_H + (m >> 3), wo = tx * DC_TILE_W + (m & 7);
            const bool ok = DS_LIVE(tile_it) && (ho < Ho) && (wo < Wo);
            const float hbf = static_cast<float>(ho - 1), wbf = static_cast<float>(wo - 1);
            const int wy0 = ty * DC_TILE_H - 1 - DS_R, wx0 = tx * DC_TILE_W - 1 - DS_R;
            const long long pix = static_cast<long long>(ho) * Wo + wo;
            const float* const offb = FUSED ? nullptr : P.offset + static_cast<long long>(img) * PP.off_img_stride + pix;
            const float* const mskb = (FUSED || !has_mask) ? nullptr : P.mask + static_cast<long long>(img) * PP.mask_img_stride + pix;
            const bool count_tile = PP.absmean != nullptr && nt == 0 && ok;

            if (FUSED) { mbar_wait_warp(off_full, tile_it & 1u); tc_fence_after_sync(); }

            // (dh, dw, mask logit / mask) of group g, tap t: tensor memory (fused) or the reference-layout tensors
            auto fetch = [&](int g, int t) -> Raw {
                Raw r;
                r.dh = r.dw = 0.f; r.mk = logit ? 0.f : 1.f;
                if (FUSED) {
                    uint32_t v[4];
                    const int col = g * 27 + 3 * t;
                    tmem_ld4_nowait(tm_lane + col, v);
                    const float b0 = __uint_as_float(lds_u32(bo_sa + col * 4)), b1 = __uint_as_float(lds_u32(bo_sa + col * 4 + 4)),
                                b2 = __uint_as_float(lds_u32(bo_sa + col * 4 + 8));
                    tmem_ld_wait();
                    r.dh = __uint_as_float(v[0]) + b0;
                    r.dw = __uint_as_float(v[1]) + b1;
                    r.mk = __uint_as_float(v[2]) + b2;
                } else if (ok) {
                    const float* ob = offb + (static_cast<long long>(g) * 18 + 2 * t) * plane;
                    r.dh = __ldg(ob);
                    r.dw = __ldg(ob + plane);
                    if (has_mask) r.mk = __ldg(mskb + (static_cast<long long>(g) * 9 + t) * plane);
                }
                return r;
            };
            // sampling geometry of one (pixel, group, tap): reference semantics of deform_conv_cuda_kernel.cu:467-497,614-628
            auto geometry = [&](const Raw& r, int ki, int kj, uint32_t win) -> Geo {
                Geo gq;
                const float h_im = hbf + static_cast<float>(ki) + r.dh;
                const float w_im = wbf + static_cast<float>(kj) + r.dw;
                const bool valid = ok && h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
                const float hf = floorf(h_im), wf = floorf(w_im);
                const float lh = h_im - hf, lw = w_im - wf;
                const float mk = valid ? (logit ? sigmoidf_fast(r.mk) : r.mk) : 0.f;
                const float a = (1.f - lh) * mk, b = lh * mk, hw = 1.f - lw;
                gq.w[0] = __float2half2_rn(a * hw); gq.w[1] = __float2half2_rn(a * lw);
                gq.w[2] = __float2half2_rn(b * hw); gq.w[3] = __float2half2_rn(b * lw);
                gq.hl = valid ? static_cast<int>(hf) : 0;
                gq.wl = valid ? static_cast<int>(wf) : 0;
                const int ry = gq.hl - wy0, rx = gq.wl - wx0;
                const bool inwin = static_cast<unsigned>(ry) <= static_cast<unsigned>(DS_WH - 2) &&
                                   static_cast<unsigned>(rx) <= static_cast<unsigned>(DS_WW - 2);
                gq.slow = valid && !inwin;
                const uint32_t p0 = inwin ? static_cast<uint32_t>(ry * DS_WW + rx) : 0u;        // invalid samples: weights are 0
                // window pixel p sits at win + p * 128, its 16-byte K atom c at ((c ^ (p & 7)) << 4) (128-byte TMA swizzle)
                gq.a0 = win + ((((p0) << 3 | ((p0) & 7u)) ^ atom0) << 4);
                gq.a1 = win + ((((p0 + 1) << 3 | ((p0 + 1) & 7u)) ^ atom0) << 4);
                gq.a2 = win + ((((p0 + 2) << 3 | ((p0 + 2) & 7u)) ^ atom0) << 4);
                return gq;
            };
            // 16-byte K atom `hi` of the warp's pair for one sample: corners p, p+1, p+17, p+18 of the window (17 = 1 mod 8:
            // same swizzle phase as p+1, p+2), packed-half blend
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
            auto sample8 = [&](const Geo& gq, uint32_t hi, int ch) -> uint4 {
                uint4 u0, u1, u2, u3;
                if (!gq.slow) {
                    const uint32_t x16 = hi << 4;
                    u0 = lds_v4(gq.a0 ^ x16);
                    u1 = lds_v4(gq.a1 ^ x16);
                    u2 = lds_v4((gq.a1 ^ x16) + 16 * 128);
                    u3 = lds_v4((gq.a2 ^ x16) + 16 * 128);
                } else {
                    const bool t = gq.hl >= 0, b = gq.hl + 1 <= H - 1, l = gq.wl >= 0, r = gq.wl + 1 <= W - 1;
                    const __half* base = ximg + gq.hl * ixrow + gq.wl * ixps + ch;
                    u0 = ldg_nc_v4((t && l) ? base : zbuf);
                    u1 = ldg_nc_v4((t && r) ? base + ixps : zbuf);
                    u2 = ldg_nc_v4((b && l) ? base + ixrow : zbuf);
                    u3 = ldg_nc_v4((b && r) ? base + ixrow + ixps : zbuf);
                }
                return blend(gq, u0, u1, u2, u3);
            };
            // both K atoms of one sample (one deformable group covers the 16 channels); the out-of-window fallback fetches the
            // 32 contiguous bytes of a corner with ONE load when the view allows it
            auto sample16 = [&](const Geo& gq, int ch, uint4& v0, uint4& v1) {
                if (!gq.slow || !wide) {
                    v0 = sample8(gq, 0u, ch);
                    v1 = sample8(gq, 1u, ch + 8);
                    return;
                }
                const bool t = gq.hl >= 0, b = gq.hl + 1 <= H - 1, l = gq.wl >= 0, r = gq.wl + 1 <= W - 1;
                const __half* base = ximg + gq.hl * ixrow + gq.wl * ixps + ch;
                uint4 a0, a1, b0, b1, c0, c1, d0, d1;
                ldg_nc_v8((t && l) ? base : zbuf, a0, a1);
                ldg_nc_v8((t && r) ? base + ixps : zbuf, b0, b1);
                ldg_nc_v8((b && l) ? base + ixrow : zbuf, c0, c1);
                ldg_nc_v8((b && r) ? base + ixrow + ixps : zbuf, d0, d1);
                v0 = blend(gq, a0, b0, c0, d0);
                v1 = blend(gq, a1, b1, c1, d1);
            };

            for (int chunk = 0; chunk < nchunks; ++chunk, ++chunk_ctr) {
                const int ch0 = chunk * 64 + kp * 16;                  // first channel of this warp's K-atom pair
                const int g0 = ch0 / cpg, g1 = TWO ? g0 + 1 : g0;
                const bool count_abs = count_tile && (ch0 % cpg) == 0;    // each (pixel, group, tap) offset exactly once
                const uint32_t wbuf = chunk & 1;
                const uint32_t win = smem_u32(win_smem + wbuf * DS_WIN_BYTES);
                Raw nx0 = fetch(g0, 0), nx1 = TWO ? fetch(g1, 0) : nx0;
                mbar_wait_warp(&win_full[wbuf], (wbuf ? used1 : used0) & 1u);
                if (wbuf) ++used1; else ++used0;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t s = tap % 3, ph = (chunk_ctr + tap / 3) & 1u;      // == it % 3, (it / 3) & 1 with it = 9 chunk_ctr + tap
                    const Raw r0 = nx0, r1 = nx1;
                    if (tap + 1 < 9) { nx0 = fetch(g0, tap + 1); nx1 = TWO ? fetch(g1, tap + 1) : nx0; }
                    if (count_abs) abs_sum += fabsf(r0.dh) + fabsf(r0.dw) + (TWO ? fabsf(r1.dh) + fabsf(r1.dw) : 0.f);
                    const Geo q0 = geometry(r0, tap / 3, tap % 3, win);
                    uint4 v0, v1;
                    if (TWO) {
                        const Geo q1 = geometry(r1, tap / 3, tap % 3, win);
                        v0 = sample8(q0, 0u, ch0);
                        v1 = sample8(q1, 1u, ch0 + 8);
                    } else {
                        sample16(q0, ch0, v0, v1);
                    }
                    mbar_wait_warp(&empty[s], ph ^ 1u);
                    sts_v4(a_dst0 + s * DC_A_BYTES, v0);
                    sts_v4(a_dst0 + s * DC_A_BYTES + DC_A_LBO, v1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&gathered[s]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&win_empty[wbuf]);
            }
            if (FUSED) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(off_empty);
            }
        }
        if (PP.absmean != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, o);
            if (lane == 0) atomicAdd(PP.absmean, abs_sum);
        }
    }

#undef DS_TILE_LOOP
#undef DS_TILE_OF
#undef DS_LIVE
    tc_fence_before_sync();
    __syncthreads();
    if (mc) cluster_sync_all();           // nobody leaves while the peer may still multicast into this CTA or signal its barriers
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace eb
