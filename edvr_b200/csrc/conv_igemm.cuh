// conv_igemm.cuh — dense 3x3 (pad 1) / 1x1 convolution as an implicit GEMM on tcgen05.
//
// Replaces every cuDNN-dispatched nn.Conv2d on the EDVR path
// (/root/reference/basicsr/models/archs/edvr_arch.py:36-66,138-155,232-244,327-353 and
//  arch_util.py:84-95) for NHWC fp16 activations, fp32 accumulation in TMEM.
//
// CTA tile  : 16x16 output pixels (two M=128 MMAs of 16 rows x 8 pixels) x BN<=128 channels.
// A operand : the (16+2)x(16+2) input halo of one 64-channel chunk is loaded ONCE into
//             shared memory as 8 planes of 16-byte K atoms (common.cuh layout); the nine
//             taps are nine start addresses inside it, so activations cross L2->SM once
//             per chunk instead of nine times.
// B operand : weights pre-packed on the host in consumption order
//             [n_tile][chunk][tap][kc=8][BN][8] fp16, streamed by 1-D bulk async copies.
// Pipeline  : warp 0 B-producer | warp 1 MMA issuer | warps 2-5 A-producers |
//             warps 6-9 epilogue; mbarrier rings; 2 x 256 TMEM columns so that the
//             epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace eb {

constexpr int CV_TILE = 16;
constexpr int CV_A_BUFS = 2;
constexpr int CV_B_STAGES = 6;
constexpr int CV_PLANE_BYTES = 325 * 16;              // >= 18*18*16, odd # of 16B units
constexpr int CV_A_BUF_BYTES = 8 * CV_PLANE_BYTES;    // 41600
constexpr int CV_B_STAGE_BYTES = 128 * 128;           // BN(<=128) rows x 64 ch x 2 B
constexpr int CV_MAX_COUT = 512;
constexpr int CV_THREADS = 320;
constexpr int CV_SMEM_BYTES = CV_A_BUFS * CV_A_BUF_BYTES + CV_B_STAGES * CV_B_STAGE_BYTES +
                              CV_MAX_COUT * 4 + 256;

struct ConvSrc {
    const __half* ptr;
    int C;            // channels consumed from this source (multiple of 64)
    int pix_stride;   // elements between pixels
    int ch_off;       // first channel
    // image index of accumulator image n:  (n / div) * mul + (n % div) * keep + add
    int div, mul, keep, add;
};

struct ConvParams {
    ConvSrc src[2];
    int nsrc;
    int N, H, W;      // accumulator grid (== input grid; stride handled by the epilogue)
    int taps;         // 9 or 1
    int BN;           // channels per n-tile (multiple of 32, <= 128)
    int n_tiles_n;    // number of n-tiles; packed Cout = BN * n_tiles_n
    const __half* wpack;
    EpiParams epi;
    unsigned long long* stats;   // optional [gridDim.x][16] cycle counters (profiling builds of the call only)
    int dbg;                     // profiling only: bit0 skip epilogue global memory, bit1 skip phase 2, bit2 skip tmem loads
};

template <int HALO>
__device__ __forceinline__ void conv_load_halo(const ConvParams& P, int chunk, int img, int ty,
                                               int tx, uint32_t abuf_saddr, int tid) {
    constexpr int RP = CV_TILE + 2 * HALO;
    constexpr int NITEM = RP * RP * 8;
    constexpr int U = HALO ? 7 : 8;
    // which source does this 64-channel chunk come from?
    int s = 0, ch = chunk * 64;
    if (P.nsrc > 1 && ch >= P.src[0].C) { s = 1; ch -= P.src[0].C; }
    const ConvSrc& S = P.src[s];
    const int simg = (img / S.div) * S.mul + (img % S.div) * S.keep + S.add;
    const __half* base = S.ptr + S.ch_off + ch;
    const int y0 = ty * CV_TILE - HALO, x0 = tx * CV_TILE - HALO;
    for (int b = 0; b < NITEM; b += 128 * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = b + u * 128 + tid;
            v[u] = make_uint4(0, 0, 0, 0);
            if (idx < NITEM) {
                const int p = idx >> 3, kc = idx & 7;
                const int y = p / RP, x = p - y * RP;
                const int gy = y0 + y, gx = x0 + x;
                if (gy >= 0 && gy < P.H && gx >= 0 && gx < P.W)
                    v[u] = ldg_nc_v4(base + ((static_cast<size_t>(simg) * P.H + gy) * P.W + gx) *
                                                S.pix_stride + kc * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = b + u * 128 + tid;
            if (idx < NITEM) sts_v4(abuf_saddr + (idx & 7) * CV_PLANE_BYTES + (idx >> 3) * 16, v[u]);
        }
    }
}

template <int HALO, int EK>
__global__ void __launch_bounds__(CV_THREADS, 1) conv_igemm_kernel(const ConvParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + CV_A_BUFS * CV_A_BUF_BYTES;
    float* bias_s = reinterpret_cast<float*>(b_smem + CV_B_STAGES * CV_B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + CV_MAX_COUT);
    uint64_t* a_full = bars;                    // [2]
    uint64_t* a_empty = bars + 2;               // [2]
    uint64_t* b_full = bars + 4;                // [6]
    uint64_t* b_empty = bars + 4 + CV_B_STAGES; // [6]
    uint64_t* acc_full = bars + 4 + 2 * CV_B_STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int RP = CV_TILE + 2 * HALO;

    const int tiles_x = (P.W + CV_TILE - 1) / CV_TILE;
    const int tiles_y = (P.H + CV_TILE - 1) / CV_TILE;
    const int total_tiles = P.N * tiles_y * tiles_x * P.n_tiles_n;
    const int cin = P.src[0].C + (P.nsrc > 1 ? P.src[1].C : 0);
    const int nchunks = cin / 64;
    const uint32_t b_bytes = static_cast<uint32_t>(P.BN) * 128u;

    // ---- one-time setup
    const int cout_packed = P.BN * P.n_tiles_n;
    const bool has_bias = P.epi.bias != nullptr;   // host guarantees cout_packed <= CV_MAX_COUT then
    if (has_bias)
        for (int i = threadIdx.x; i < cout_packed; i += blockDim.x) bias_s[i] = P.epi.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < CV_A_BUFS; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < CV_B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= B producer: stream packed weights, one stage per (chunk, tap)
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % P.n_tiles_n;
                const uint8_t* w = reinterpret_cast<const uint8_t*>(P.wpack) +
                                   static_cast<size_t>(nt) * nchunks * P.taps * b_bytes;
                for (int st = 0; st < nchunks * P.taps; ++st, ++it) {
                    const uint32_t s = it % CV_B_STAGES, ph = (it / CV_B_STAGES) & 1u;
                    mbar_wait(&b_empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&b_full[s], b_bytes);
                    bulk_g2s(b_smem + s * CV_B_STAGE_BYTES, w + static_cast<size_t>(st) * b_bytes,
                             b_bytes, &b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, P.BN);
            const uint32_t lbo_b = static_cast<uint32_t>(P.BN) * 16u;
            uint32_t a_it = 0, b_it = 0, acc_it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
                const uint32_t ab = acc_it & 1u;
                mbar_wait(&acc_empty[ab], ((acc_it >> 1) & 1u) ^ 1u);
                tc_fence_after_sync();
                for (int c = 0; c < nchunks; ++c, ++a_it) {
                    const uint32_t as = a_it % CV_A_BUFS, aph = (a_it / CV_A_BUFS) & 1u;
                    mbar_wait(&a_full[as], aph);
                    tc_fence_after_sync();
                    const uint32_t a_base = smem_u32(a_smem + as * CV_A_BUF_BYTES);
                    for (int t = 0; t < P.taps; ++t, ++b_it) {
                        const uint32_t bs = b_it % CV_B_STAGES, bph = (b_it / CV_B_STAGES) & 1u;
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after_sync();
                        const uint32_t b_base = smem_u32(b_smem + bs * CV_B_STAGE_BYTES);
                        const int ki = HALO ? t / 3 : 0, kj = HALO ? t % 3 : 0;
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            const uint32_t a_tap = a_base + ((ki * RP) + 8 * sub + kj) * 16;
                            const uint32_t d = tmem_base + ab * 256u + sub * 128u;
#pragma unroll
                            for (int k16 = 0; k16 < 4; ++k16) {
                                const uint64_t ad = umma_desc_nosw(a_tap + k16 * 2 * CV_PLANE_BYTES,
                                                                   CV_PLANE_BYTES, RP * 16);
                                const uint64_t bd = umma_desc_nosw(b_base + k16 * 2 * lbo_b, lbo_b, 128);
                                umma_f16(d, ad, bd, idesc, (c | t | k16) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&b_empty[bs]);
                    }
                    umma_commit(&a_empty[as]);
                }
                umma_commit(&acc_full[ab]);
            }
        }
    } else if (warp < 6) {
        // ================= A producers (128 threads): halo tile of one 64-channel chunk
        const int tid = threadIdx.x - 64;
        uint32_t a_it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int pt = tile / P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            for (int c = 0; c < nchunks; ++c, ++a_it) {
                const uint32_t as = a_it % CV_A_BUFS, aph = (a_it / CV_A_BUFS) & 1u;
                mbar_wait_warp(&a_empty[as], aph ^ 1u);
                conv_load_halo<HALO>(P, c, img, ty, tx, smem_u32(a_smem + as * CV_A_BUF_BYTES), tid);
                fence_proxy_async_smem();   // generic-proxy stores -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[as]);
            }
        }
    } else {
        // ================= epilogue (128 threads == 128 TMEM lanes)
        const int q = warp & 3;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
            const int nt = tile % P.n_tiles_n, pt = tile / P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            const uint32_t ab = acc_it & 1u;
            mbar_wait_warp(&acc_full[ab], (acc_it >> 1) & 1u);
            tc_fence_after_sync();
            const int y = ty * CV_TILE + 4 * q + (lane >> 3);
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                const int x = tx * CV_TILE + 8 * sub + (lane & 7);
                const bool valid = (y < P.H) && (x < P.W);
                const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + ab * 256u + sub * 128u;
#pragma unroll 1
                for (int cc = 0; cc < P.BN; cc += 32) {
                    float v[32];
                    tmem_ld32(t0 + cc, v);
                    epi_store32<EK>(P.epi, has_bias ? bias_s : nullptr, v, img, y, x, nt * P.BN + cc, valid);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
    }

    // ---- teardown
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace eb
