// conv_train.cuh — weight gradient of the dense 3x3 (pad 1) / 1x1 convolutions on tcgen05, for the training step
// (BASELINE cfg 5; reference side: autograd of nn.Conv2d = cuDNN wgrad, options/train/EDVR/train_EDVR_L_x4_SR_REDS.yml,
// basicsr/models/base_model.py:62-69 wraps the net in DDP).  The data gradient needs no kernel of its own: dgrad of a
// stride-1 convolution is the convolution of grad_out with the transposed, spatially flipped weights, i.e. the CTA-pair
// forward kernel (conv_pair.cuh) on re-packed weights.
//
//   gW[co][ci][tap] = sum_{n,y,x} gy[n,y,x,co] * x[n, y+dy(tap), x+dx(tap), ci]
//
// is a GEMM with K = pixels.  Both operands are NHWC (pixel-major) in HBM, i.e. M/N-major for this GEMM, so they are first
// transposed into channel-major rows over a ZERO-PADDED pixel grid (n, y+1, x+1) of row pitch Wp (a multiple of 8):
//   gyT[co][p]          p = margin + (n * Hp + y + 1) * Wp + x + 1
//   xT_dx[ci][p] = xpad[ci][p + dx],  dx = -1, 0, +1   (three copies, so that every tap is a 16-byte ALIGNED row offset:
//                                                        tap (dy, dx) reads xT_dx at p + dy * Wp)
// The zero border makes the flattened shift exact (out-of-image taps multiply zeros), and the GEMM kernel is the split-K
// tcgen05 kernel of the DCN weight gradient (dcn_backward.cuh) with a per-tap operand base.  2-byte elements are moved
// as raw bits, so fp16 and bf16 share the transposes; the MMA operand format is selected by the instruction descriptor.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eb {

// src: NHWC 2-byte elements (view: pix_stride, ch_off, C channels) -> dst[dx][c][p] for dx in [0, ncopies): ncopies == 3
// writes the three shifted copies (dst copy i holds xpad shifted by dx = i - 1), ncopies == 1 the unshifted one.
// grid: (ceil(W/32), ceil(H/rows), N * ceil(C/32)), block (32, 8); dst must be zero-filled (borders, padding, margins).
// colsum (optional, fp32 [C]): += sum over all pixels of src[.., c] - the bias gradient when src is grad_out, taken while
// the tile is in flight (a separate ATen reduction over NHWC bf16 cost 37 us per layer, 11 % of the training step).
constexpr int CT_ROWS = 8;          // image rows per block when the column sum is taken (one atomic per channel and block)
template <bool BF16>
__global__ void nhwc_to_cmajor_pad_kernel(const uint16_t* __restrict__ src, int pix_stride, int ch_off, int C, int H, int W,
                                          uint16_t* __restrict__ dst, long long copy_stride, long long Ppad, int Hp, int Wp,
                                          int margin, int ncopies, float* __restrict__ colsum, int rows) {
    __shared__ uint16_t tile[32][34];
    __shared__ float red[8][33];
    const int cblocks = (C + 31) / 32;
    const int n = blockIdx.z / cblocks, c0 = (blockIdx.z % cblocks) * 32;
    const int x0 = blockIdx.x * 32;
    float acc = 0.f;
    for (int y = blockIdx.y * rows; y < H && y < (blockIdx.y + 1) * rows; ++y) {
        for (int i = threadIdx.y; i < 32; i += 8) {                    // i: pixel within the tile, threadIdx.x: channel
            const int x = x0 + i, c = c0 + threadIdx.x;
            const uint16_t v = (x < W && c < C)
                ? src[((static_cast<size_t>(n) * H + y) * W + x) * pix_stride + ch_off + c] : static_cast<uint16_t>(0);
            tile[i][threadIdx.x] = v;
            if (colsum != nullptr)
                acc += BF16 ? __uint_as_float(static_cast<uint32_t>(v) << 16) : __half2float(__ushort_as_half(v));
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += 8) {                    // i: channel, threadIdx.x: pixel
            const int c = c0 + i, x = x0 + threadIdx.x;
            if (c < C && x < W) {
                const long long p = margin + (static_cast<long long>(n) * Hp + y + 1) * Wp + x + 1;
                const uint16_t v = tile[threadIdx.x][i];
                if (ncopies == 1) {
                    dst[static_cast<size_t>(c) * Ppad + p] = v;
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d)         // copy d holds xpad[p + (d - 1)]  =>  x lands at p - (d - 1)
                        dst[d * copy_stride + static_cast<size_t>(c) * Ppad + p - (d - 1)] = v;
                }
            }
        }
        __syncthreads();
    }
    if (colsum != nullptr) {
        red[threadIdx.y][threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.y == 0 && c0 + threadIdx.x < C) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x];
            atomicAdd(colsum + c0 + threadIdx.x, s);
        }
    }
}

constexpr int CW_STAGES = 3;                               // 96 KB per CTA: two CTAs per SM keep 192 KB in flight
constexpr int CW_STAGE_BYTES = 128 * 128 + 128 * 128;      // A: 128 rows x 64 K x 2 B as 8 K-atom planes, B: BN <= 128 rows
constexpr int CW_SMEM_BYTES = CW_STAGES * CW_STAGE_BYTES + 256;
constexpr int CW_THREADS = 192;                            // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

struct WgradParams {
    CUtensorMap tmap_a;        // gyT as {8, rowsA, Ppad / 8}: box {8, 128, 8} lands as [K atom][row][16 B] (no-swizzle K-major)
    CUtensorMap tmap_b;        // xT copies as {8, copies * Cin, Ppad / 8}: box {8, BN, 8}
    float* partial;            // [split][m tile][tile][128][BN] fp32
    int Cin, taps, BN, Wp, steps_per_split, ab_fmt;
    long long k_begin, k_steps;
};

// partial[split][mt][tile][co 128][ci BN] = sum over this split's pixels of A[co][p] * B_tap[ci][p]; A = gyT (rows padded to a
// multiple of 128), B_tap = xT_dx(tap) at row offset dy(tap) * Wp.  grid: (taps * (Cin / BN), ceil(Cout / 128), splits).
// Warp-specialised: one thread streams both operand tiles of a 64-pixel K step with two TMA boxes per stage (3 stages,
// two CTAs per SM = 192 KB in flight), one thread issues 4 tcgen05.mma per stage, four warps drain the 128 x BN accumulator.  The first version
// (128 threads loading with LDG -> STS -> __syncthreads per step) ran at 95 TFLOP/s and was a third of the training step
// (profiles/r02_train_profile_v2.txt).  Split-K partial tiles go to a workspace and are summed in a fixed order by
// conv_wgrad_reduce_kernel: deterministic, no atomics.
__global__ void __launch_bounds__(CW_THREADS, 2) conv_wgrad_kernel(const __grid_constant__ WgradParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CW_STAGES * CW_STAGE_BYTES);
    uint64_t* full = bars;                    // [CW_STAGES]
    uint64_t* empty = bars + CW_STAGES;       // [CW_STAGES]
    uint64_t* done_bar = empty + CW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = P.BN, ntiles_c = P.Cin / BN;
    const int tap = blockIdx.x / ntiles_c, c0 = (blockIdx.x % ntiles_c) * BN, mt = blockIdx.y;
    const long long s0 = static_cast<long long>(blockIdx.z) * P.steps_per_split;
    long long s1 = s0 + P.steps_per_split;
    if (s1 > P.k_steps) s1 = P.k_steps;
    const int nsteps = s1 > s0 ? static_cast<int>(s1 - s0) : 0;
    const int dy = P.taps == 9 ? tap / 3 - 1 : 0, dxi = P.taps == 9 ? tap % 3 : 0;      // copy index dxi holds dx = dxi - 1

    if (tid == 0) {
        for (int i = 0; i < CW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&P.tmap_a);
        tma_prefetch_desc(&P.tmap_b);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b_bytes = static_cast<uint32_t>(BN) * 128u;

    if (warp == 0) {
        if (lane == 0) {
            const int a_atom0 = static_cast<int>((P.k_begin + s0 * 64) >> 3);
            const int b_atom0 = static_cast<int>((P.k_begin + s0 * 64 + static_cast<long long>(dy) * P.Wp) >> 3);
            for (int i = 0; i < nsteps; ++i) {
                const int s = i % CW_STAGES;
                mbar_wait_t<32>(&empty[s], ((i / CW_STAGES) & 1) ^ 1);
                uint8_t* a_s = smem + s * CW_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], 128u * 128u + b_bytes);
                tma_load_3d(a_s, &P.tmap_a, &full[s], 0, mt * 128, a_atom0 + i * 8);
                tma_load_3d(a_s + 128 * 128, &P.tmap_b, &full[s], 0, dxi * P.Cin + c0, b_atom0 + i * 8);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_f16(128, BN, static_cast<uint32_t>(P.ab_fmt));
        const uint32_t lbo_b = static_cast<uint32_t>(BN) * 16u;
        const uint32_t hi = umma_desc_hi(128);
        for (int i = 0; i < nsteps; ++i) {
            const int s = i % CW_STAGES;
            mbar_wait(&full[s], (i / CW_STAGES) & 1);
            tc_fence_after_sync();
            const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem + s * CW_STAGE_BYTES), 2048);
            const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem + s * CW_STAGE_BYTES + 128 * 128), lbo_b);
            if (elect_one()) {
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16)
                    umma_f16_lohi<1>(tmem_base, a_lo0 + k16 * (2 * 2048 / 16), hi, b_lo0 + k16 * ((2u * lbo_b) >> 4), hi, idesc,
                                     (i | k16) != 0 ? 1u : 0u);
                umma_commit(&empty[s]);
                if (i == nsteps - 1) umma_commit(done_bar);
            }
            __syncwarp();
        }
    } else {
        // this CTA's tile of the partial sums: row co = 32 q + lane (q = TMEM lane quarter of the warp), BN contiguous floats
        const int q = warp & 3;
        float* out = P.partial + ((static_cast<size_t>(blockIdx.z) * gridDim.y + mt) * gridDim.x + blockIdx.x) * (128 * static_cast<size_t>(BN)) +
                     static_cast<size_t>(32 * q + lane) * BN;
        if (nsteps > 0) {
            mbar_wait_t<64>(done_bar, 0);
            tc_fence_after_sync();
        }
        for (int cc = 0; cc < BN; cc += 32) {
            float v[32];
            if (nsteps > 0) tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + cc, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// grad_weight[co][ci][tap] += scale * sum_split partial[split][mt][tile = tap * (Cin / BN) + ci / BN][co % 128][ci % BN]
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gW, int Cout, int Cin, int taps,
                                         int BN, int m_tiles, int splits, float scale) {
    const int ntc = Cin / BN, tiles = taps * ntc;
    const long long total = static_cast<long long>(taps) * Cout * Cin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = i % Cin;
        const int co = (i / Cin) % Cout;
        const int tap = i / (static_cast<long long>(Cin) * Cout);
        const int mt = co / 128, tile = tap * ntc + ci / BN;
        const float* src = partial + (static_cast<size_t>(mt) * tiles + tile) * (128 * static_cast<size_t>(BN)) +
                           static_cast<size_t>(co % 128) * BN + ci % BN;
        const size_t split_stride = static_cast<size_t>(m_tiles) * tiles * 128 * BN;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += src[k * split_stride];
        gW[(static_cast<size_t>(co) * Cin + ci) * taps + tap] += s * scale;
    }
}

}  // namespace eb
