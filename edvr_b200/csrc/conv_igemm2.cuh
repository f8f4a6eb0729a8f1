// conv_igemm2.cuh — second-generation dense conv (3x3 pad 1 / 1x1) for Cout tiles of exactly 128:
// the GEMM is transposed with respect to conv_igemm.cuh:  D[Cout=128][pixels=256] = W · X^T.
//
//   * M = 128 output channels (TMEM lanes), N = 256 pixels per MMA (a 32-row x 8-pixel image tile), K = 16.
//     One tcgen05.mma now reads 4 KB of weights + 8 KB of pixels per 128 cycles (96 B/cycle of shared-memory
//     bandwidth instead of the 128 B/cycle an M128 x N128 pair of MMAs needs), and half as many instructions.
//   * The pixel operand is the same halo-tile trick: one (32+2)x(8+2) halo of a 64-channel chunk is loaded once,
//     every 3x3 tap is a start-address offset into it (rows of the canonical no-swizzle layout = pixels).
//   * The accumulator is channel-major, so the epilogue transposes through a small shared-memory slab: phase 1,
//     each thread (= one TMEM lane = one channel) adds bias, applies the activation and writes 32 pixels of its
//     channel as fp32 (conflict-free: lanes are consecutive channels); phase 2, each WARP owns one pixel and each
//     lane 4 consecutive channels: residual loads and NHWC stores are 256/512 contiguous bytes per instruction
//     (2-4 L1 wavefronts instead of the 32 a pixel-per-lane epilogue costs).  PixelShuffle / stride-2 /
//     DCN-record variants included.
// Same warp roles, mbarrier rings, bulk-copied pre-packed weights and persistent scheduling as conv_igemm.cuh.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"
#include "conv_igemm.cuh"

namespace eb {

constexpr int C2_TH = 32, C2_TW = 8;                   // pixel tile: 32 rows x 8 columns = 256 = MMA N
constexpr int C2_A_BUFS = 2;
constexpr int C2_W_STAGES = 5;
constexpr int C2_PLANE_BYTES = 341 * 16;               // >= 34*10*16, odd number of 16-byte units
constexpr int C2_A_BUF_BYTES = 8 * C2_PLANE_BYTES;     // 43648
constexpr int C2_W_STAGE_BYTES = 128 * 128;            // 128 channels x 64 k x 2 B
constexpr int C2_THREADS = 320;
constexpr int C2_SLAB_FLOATS = 32 * 128;                // one 32-pixel x 128-channel fp32 slab
constexpr int C2_SMEM_BYTES = C2_A_BUFS * C2_A_BUF_BYTES + C2_W_STAGES * C2_W_STAGE_BYTES +
                              2 * C2_SLAB_FLOATS * 4 + 256;

template <int HALO>
__device__ __forceinline__ void conv2_load_halo(const ConvParams& P, int chunk, int img, int ty, int tx,
                                                uint32_t abuf_saddr, int tid) {
    constexpr int RPX = C2_TW + 2 * HALO;               // row pitch in pixels
    constexpr int NPIX = (C2_TH + 2 * HALO) * RPX;
    constexpr int NITEM = NPIX * 8;
    constexpr int U = 8;
    int s = 0, ch = chunk * 64;
    if (P.nsrc > 1 && ch >= P.src[0].C) { s = 1; ch -= P.src[0].C; }
    const ConvSrc& S = P.src[s];
    const bool img_ok = img < P.N;                      // ghost tiles (odd tail) read nothing
    const int simg = img_ok ? (img / S.div) * S.mul + (img % S.div) * S.keep + S.add : 0;
    const __half* base = S.ptr + S.ch_off + ch;
    const int y0 = ty * C2_TH - HALO, x0 = tx * C2_TW - HALO;
    for (int b = 0; b < NITEM; b += 128 * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = b + u * 128 + tid;
            v[u] = make_uint4(0, 0, 0, 0);
            if (idx < NITEM) {
                const int p = idx >> 3, kc = idx & 7;
                const int y = p / RPX, x = p - y * RPX;
                const int gy = y0 + y, gx = x0 + x;
                if (img_ok && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W)
                    v[u] = ldg_nc_v4(base + ((static_cast<size_t>(simg) * P.H + gy) * P.W + gx) * S.pix_stride + kc * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = b + u * 128 + tid;
            if (idx < NITEM) sts_v4(abuf_saddr + (idx & 7) * C2_PLANE_BYTES + (idx >> 3) * 16, v[u]);
        }
    }
}

// Epilogue phase 2: this lane owns channels c0..c0+3 (packed index) of output pixel (img, y, x); the whole warp
// shares the pixel, so every validity test is warp-uniform.  v already holds bias + activation.
__device__ __forceinline__ void conv2_store4(const EpiParams& p, float4 v, int img, int y, int x, int c0, bool valid) {
    if (p.act == ACT_DCN_PACK) {
        // packed record: channel j = c % 32 of each group: [0,18) offsets, [18,27) mask logits, rest pad
        const int j0 = c0 & 31;
        float* e = reinterpret_cast<float*>(&v);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j0 + i;
            if (j < 18) s += fabsf(e[i]);
            else if (j < 27) e[i] = sigmoidf_fast(e[i]);
        }
        if (p.absmean_acc != nullptr) {
            s = valid ? s : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane_id() == 0 && valid) atomicAdd(p.absmean_acc, s);
        }
    }
    if (!valid) return;
    if (p.out_mode == OUT_SAME) {
        const size_t pix = (static_cast<size_t>(img) * p.H + y) * p.W + x;
        if (p.res16 != nullptr) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.res16 + pix * p.res_pix_stride + p.res_ch_off + c0));
            const float2 a = unpack_h2(u.x), b = unpack_h2(u.y);
            v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
        }
        if (p.res32 != nullptr) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res32 + pix * p.res_pix_stride + p.res_ch_off + c0));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (p.out16 != nullptr)
            *reinterpret_cast<uint2*>(p.out16 + pix * p.out16_pix_stride + p.out16_ch_off + c0) =
                make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
        if (p.out32 != nullptr)
            *reinterpret_cast<float4*>(p.out32 + pix * p.out32_pix_stride + p.out32_ch_off + c0) = v;
    } else if (p.out_mode == OUT_PIXSHUF2) {
        // out[b, c, 2y+i, 2x+j] = in[b, 4c + 2i + j, y, x]; this lane owns output channel c0/4
        const int H2 = 2 * p.H, W2 = 2 * p.W;
        __half* o = p.out16 + ((static_cast<size_t>(img) * H2 + 2 * y) * W2 + 2 * x) * p.out16_pix_stride +
                    p.out16_ch_off + (c0 >> 2);
        const size_t rs = static_cast<size_t>(W2) * p.out16_pix_stride;
        o[0] = __float2half_rn(v.x);
        o[p.out16_pix_stride] = __float2half_rn(v.y);
        o[rs] = __float2half_rn(v.z);
        o[rs + p.out16_pix_stride] = __float2half_rn(v.w);
    } else {   // OUT_STRIDE2
        if ((y | x) & 1) return;
        const int Ho = (p.H + 1) >> 1, Wo = (p.W + 1) >> 1;
        const size_t opix = (static_cast<size_t>(img) * Ho + (y >> 1)) * Wo + (x >> 1);
        *reinterpret_cast<uint2*>(p.out16 + opix * p.out16_pix_stride + p.out16_ch_off + c0) =
            make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
    }
}

// cycle-counter slots of ConvParams.stats (per CTA): who waited on what
enum : int { ST_MMA_TOTAL = 0, ST_MMA_WAIT_ACC = 1, ST_MMA_WAIT_A = 2, ST_MMA_WAIT_W = 3, ST_A_TOTAL = 4,
             ST_A_WAIT_EMPTY = 5, ST_W_TOTAL = 6, ST_W_WAIT_EMPTY = 7, ST_E_TOTAL = 8, ST_E_WAIT_ACC = 9, ST_TILES = 10, ST_E_TMEM = 11, ST_E_P1 = 12, ST_E_BAR = 13, ST_E_P2 = 14 };

#define C2_TIMED_WAIT_(WAITFN, bar, parity, slot)                                  \
    do {                                                                   \
        if (P.stats != nullptr) {                                          \
            const long long t_ = clock64();                                \
            WAITFN(bar, parity);                                           \
            st_acc[slot] += static_cast<unsigned long long>(clock64() - t_); \
        } else {                                                           \
            WAITFN(bar, parity);                                           \
        }                                                                  \
    } while (0)
#define C2_TIMED_WAIT(bar, parity, slot) C2_TIMED_WAIT_(mbar_wait, bar, parity, slot)
#define C2_TIMED_WAIT_WARP(bar, parity, slot) C2_TIMED_WAIT_(mbar_wait_warp, bar, parity, slot)

template <int HALO>
__global__ void __launch_bounds__(C2_THREADS, 1) conv_igemm2_kernel(const ConvParams P) {
    unsigned long long st_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long st_t0 = clock64();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_smem = smem;                                         // pixel halo chunks
    uint8_t* w_smem = smem + C2_A_BUFS * C2_A_BUF_BYTES;            // weight stages
    float* slab = reinterpret_cast<float*>(w_smem + C2_W_STAGES * C2_W_STAGE_BYTES);      // 2 x [32 px][128 ch]
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab + 2 * C2_SLAB_FLOATS);
    uint64_t* a_full = bars;                          // [2]
    uint64_t* a_empty = bars + 2;                     // [2]
    uint64_t* w_full = bars + 4;                      // [6]
    uint64_t* w_empty = bars + 4 + C2_W_STAGES;       // [6]
    uint64_t* acc_full = bars + 4 + 2 * C2_W_STAGES;  // [2]
    uint64_t* acc_empty = acc_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int RPX = C2_TW + 2 * HALO;
    const int tiles_x = (P.W + C2_TW - 1) / C2_TW;
    const int tiles_y = (P.H + C2_TH - 1) / C2_TH;
    const int total_tiles = P.N * tiles_y * tiles_x * P.n_tiles_n;
    const int cin = P.src[0].C + (P.nsrc > 1 ? P.src[1].C : 0);
    const int nchunks = cin / 64;
    const bool has_bias = P.epi.bias != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C2_A_BUFS; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C2_W_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= weight producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % P.n_tiles_n;
                const uint8_t* w = reinterpret_cast<const uint8_t*>(P.wpack) +
                                   static_cast<size_t>(nt) * nchunks * P.taps * C2_W_STAGE_BYTES;
                for (int st = 0; st < nchunks * P.taps; ++st, ++it) {
                    const uint32_t s = it % C2_W_STAGES, ph = (it / C2_W_STAGES) & 1u;
                    C2_TIMED_WAIT(&w_empty[s], ph ^ 1u, ST_W_WAIT_EMPTY);
                    mbar_arrive_expect_tx(&w_full[s], C2_W_STAGE_BYTES);
                    bulk_g2s(w_smem + s * C2_W_STAGE_BYTES, w + static_cast<size_t>(st) * C2_W_STAGE_BYTES,
                             C2_W_STAGE_BYTES, &w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D[128 ch][256 px] += W[128 x 16] . X[256 x 16]^T
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 256);
            uint32_t a_it = 0, w_it = 0, acc_it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
                const uint32_t ab = acc_it & 1u;
                C2_TIMED_WAIT(&acc_empty[ab], ((acc_it >> 1) & 1u) ^ 1u, ST_MMA_WAIT_ACC);
                tc_fence_after_sync();
                const uint32_t d = tmem_base + ab * 256u;
                for (int c = 0; c < nchunks; ++c, ++a_it) {
                    const uint32_t as = a_it % C2_A_BUFS, aph = (a_it / C2_A_BUFS) & 1u;
                    C2_TIMED_WAIT(&a_full[as], aph, ST_MMA_WAIT_A);
                    tc_fence_after_sync();
                    const uint32_t x_base = smem_u32(a_smem + as * C2_A_BUF_BYTES);
                    for (int t = 0; t < P.taps; ++t, ++w_it) {
                        const uint32_t ws = w_it % C2_W_STAGES, wph = (w_it / C2_W_STAGES) & 1u;
                        C2_TIMED_WAIT(&w_full[ws], wph, ST_MMA_WAIT_W);
                        tc_fence_after_sync();
                        const uint32_t w_base = smem_u32(w_smem + ws * C2_W_STAGE_BYTES);
                        const int ki = HALO ? t / 3 : 0, kj = HALO ? t % 3 : 0;
                        const uint32_t x_tap = x_base + (ki * RPX + kj) * 16;
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
                            const uint64_t wd = umma_desc_nosw(w_base + k16 * 2 * 2048, 2048, 128);
                            const uint64_t xd = umma_desc_nosw(x_tap + k16 * 2 * C2_PLANE_BYTES, C2_PLANE_BYTES, RPX * 16);
                            umma_f16(d, wd, xd, idesc, (c | t | k16) != 0 ? 1u : 0u);
                        }
                        umma_commit(&w_empty[ws]);
                    }
                    umma_commit(&a_empty[as]);
                }
                umma_commit(&acc_full[ab]);
            }
        }
    } else if (warp < 6) {
        // ================= pixel (halo) producers
        const int tid = threadIdx.x - 64;
        uint32_t a_it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int pt = tile / P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            for (int c = 0; c < nchunks; ++c, ++a_it) {
                const uint32_t as = a_it % C2_A_BUFS, aph = (a_it / C2_A_BUFS) & 1u;
                C2_TIMED_WAIT_WARP(&a_empty[as], aph ^ 1u, ST_A_WAIT_EMPTY);
                conv2_load_halo<HALO>(P, c, img, ty, tx, smem_u32(a_smem + as * C2_A_BUF_BYTES), tid);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[as]);
            }
        }
    } else {
        // ================= epilogue: 4 warps, warp q owns TMEM lanes (= channels) 32q..32q+31
        const int q = warp & 3;
        // every parameter the inner loops need lives in a register: with one epilogue warp per scheduler the
        // latency of a constant-bank load or a 64-bit multiply inside the loop is fully exposed (r01_conv_stats)
        const EpiParams E = P.epi;
        const int H = P.H, W = P.W, NIMG = P.N, dbg = P.dbg;
        const bool same = E.out_mode == OUT_SAME;
        const bool pack = E.act == ACT_DCN_PACK;
        const int act1 = pack ? ACT_NONE : E.act;
        const long long ps16 = E.out16_pix_stride, ps32 = E.out32_pix_stride, psr = E.res_pix_stride;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
            const int nt = tile % P.n_tiles_n, pt = tile / P.n_tiles_n;
            const int tx = pt % tiles_x, ty = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
            const uint32_t ab = acc_it & 1u;
            const float bias_c = has_bias ? __ldg(E.bias + nt * 128 + 32 * q + lane) : 0.f;
            // phase-2 ownership: this lane = channels c0..c0+3, this warp = pixel column xq of every slab row pair
            const int c0 = nt * 128 + lane * 4;
            const int x0 = tx * C2_TW, y0 = ty * C2_TH;
            const bool img_ok = img < NIMG;
            const long long pix0 = (static_cast<long long>(img) * H + y0) * W + x0;      // pixel index of the tile origin
            __half* o16 = (E.out16 != nullptr) ? E.out16 + E.out16_ch_off + c0 + pix0 * ps16 : nullptr;
            float* o32 = (E.out32 != nullptr) ? E.out32 + E.out32_ch_off + c0 + pix0 * ps32 : nullptr;
            const __half* r16 = (E.res16 != nullptr) ? E.res16 + E.res_ch_off + c0 + pix0 * psr : nullptr;
            const float* r32 = (E.res32 != nullptr) ? E.res32 + E.res_ch_off + c0 + pix0 * psr : nullptr;
            C2_TIMED_WAIT_WARP(&acc_full[ab], (acc_it >> 1) & 1u, ST_E_WAIT_ACC);
            tc_fence_after_sync();
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + ab * 256u;
#pragma unroll 1
            for (int col = 0; col < 256; col += 32) {
                float* sl = slab + ((col >> 5) & 1) * C2_SLAB_FLOATS;
                long long tq = P.stats ? clock64() : 0;
                {
                    float v[32];
                    if (dbg & 4) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    } else {
                        tmem_ld32(t0 + col, v);
                    }
                    if (P.stats) { const long long t = clock64(); st_acc[ST_E_TMEM] += t - tq; tq = t; }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bias_c;
                    act_inplace<32>(v, act1);
                    // phase 1: channel (32q + lane) of pixels col..col+31 -> slab[pixel][channel]
#pragma unroll
                    for (int j = 0; j < 32; ++j) sl[j * 128 + 32 * q + lane] = v[j];
                }
                if (P.stats) { const long long t = clock64(); st_acc[ST_E_P1] += t - tq; tq = t; }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (P.stats) { const long long t = clock64(); st_acc[ST_E_BAR] += t - tq; tq = t; }
                if (dbg & 2) continue;
                // phase 2: warp q handles pixels q, q+4, ... of the slab: row = i/2, column = 4*(i&1) + q
                const int yb = (col >> 3);                         // first tile row of this slab
                if (same) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ry = yb + (i >> 1), rx = 4 * (i & 1) + q;
                        const bool valid = img_ok && (y0 + ry < H) && (x0 + rx < W) && !(dbg & 1);
                        float4 v = *reinterpret_cast<const float4*>(sl + (q + 4 * i) * 128 + lane * 4);
                        if (pack) {
                            float* e = reinterpret_cast<float*>(&v);
                            const int j0 = c0 & 31;
                            float s = 0.f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int j = j0 + k;
                                if (j < 18) s += fabsf(e[k]);
                                else if (j < 27) e[k] = sigmoidf_fast(e[k]);
                            }
                            if (E.absmean_acc != nullptr) {
                                s = valid ? s : 0.f;
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                                if (lane == 0 && valid) atomicAdd(E.absmean_acc, s);
                            }
                        }
                        if (valid) {
                            const long long pofs = static_cast<long long>(ry) * W + rx;
                            if (r16 != nullptr) {
                                const uint2 u = __ldg(reinterpret_cast<const uint2*>(r16 + pofs * psr));
                                const float2 a = unpack_h2(u.x), b = unpack_h2(u.y);
                                v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
                            }
                            if (r32 != nullptr) {
                                const float4 r = __ldg(reinterpret_cast<const float4*>(r32 + pofs * psr));
                                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                            }
                            if (o16 != nullptr)
                                *reinterpret_cast<uint2*>(o16 + pofs * ps16) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
                            if (o32 != nullptr) *reinterpret_cast<float4*>(o32 + pofs * ps32) = v;
                        }
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {
                        const int y = y0 + yb + (i >> 1), x = x0 + 4 * (i & 1) + q;
                        const bool valid = img_ok && (y < H) && (x < W) && !(dbg & 1);
                        const float4 v4 = *reinterpret_cast<const float4*>(sl + (q + 4 * i) * 128 + lane * 4);
                        conv2_store4(E, v4, img, y, x, c0, valid);
                    }
                }
                if (P.stats) st_acc[ST_E_P2] += clock64() - tq;
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
    }

    if (P.stats != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6)) {
        unsigned long long* o = P.stats + static_cast<size_t>(blockIdx.x) * 16;
        const unsigned long long tot = static_cast<unsigned long long>(clock64() - st_t0);
        if (warp == 1) { o[ST_MMA_TOTAL] = tot; o[ST_MMA_WAIT_ACC] = st_acc[ST_MMA_WAIT_ACC]; o[ST_MMA_WAIT_A] = st_acc[ST_MMA_WAIT_A]; o[ST_MMA_WAIT_W] = st_acc[ST_MMA_WAIT_W]; }
        if (warp == 2) { o[ST_A_TOTAL] = tot; o[ST_A_WAIT_EMPTY] = st_acc[ST_A_WAIT_EMPTY]; }
        if (warp == 0) { o[ST_W_TOTAL] = tot; o[ST_W_WAIT_EMPTY] = st_acc[ST_W_WAIT_EMPTY]; }
        if (warp == 6) { o[ST_E_TMEM] = st_acc[ST_E_TMEM]; o[ST_E_P1] = st_acc[ST_E_P1]; o[ST_E_BAR] = st_acc[ST_E_BAR]; o[ST_E_P2] = st_acc[ST_E_P2]; o[ST_E_TOTAL] = tot; o[ST_E_WAIT_ACC] = st_acc[ST_E_WAIT_ACC]; o[ST_TILES] = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace eb
