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
#include "common.cuh"

namespace eb {

// src: NHWC 2-byte elements (view: pix_stride, ch_off, C channels) -> dst[dx][c][p] for dx in [0, ncopies): ncopies == 3
// writes the three shifted copies (dst copy i holds xpad shifted by dx = i - 1), ncopies == 1 the unshifted one.
// grid: (ceil(W/32), ceil(H/CT_ROWS), N * ceil(C/32)), block (32, 8); dst must be zero-filled (borders, padding, margins).
// colsum (optional, fp32 [C]): += sum over all pixels of src[.., c] - the bias gradient when src is grad_out, taken while
// the tile is in flight (a separate ATen reduction over NHWC bf16 cost 37 us per layer, 11 % of the training step).
constexpr int CT_ROWS = 8;
template <bool BF16>
__global__ void nhwc_to_cmajor_pad_kernel(const uint16_t* __restrict__ src, int pix_stride, int ch_off, int C, int H, int W,
                                          uint16_t* __restrict__ dst, long long copy_stride, long long Ppad, int Hp, int Wp,
                                          int margin, int ncopies, float* __restrict__ colsum) {
    __shared__ uint16_t tile[32][34];
    __shared__ float red[8][33];
    const int cblocks = (C + 31) / 32;
    const int n = blockIdx.z / cblocks, c0 = (blockIdx.z % cblocks) * 32;
    const int x0 = blockIdx.x * 32;
    float acc = 0.f;
    for (int y = blockIdx.y * CT_ROWS; y < H && y < (blockIdx.y + 1) * CT_ROWS; ++y) {
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

constexpr int CW_STAGES = 3;
constexpr int CW_SMEM_BYTES = CW_STAGES * (128 * 128 + 128 * 128);

// partial[split][tile][co 128][ci BN] = sum over this split's pixels of A[co][p] * B_tap[ci][p]; A = gyT (rows padded to a
// multiple of 128), B_tap = xT_dx(tap) at row offset dy(tap) * Wp.  grid: (taps * (Cin / BN), ceil(Cout / 128), splits),
// 128 threads.  Split-K partial tiles go to a workspace and are summed by conv_wgrad_reduce_kernel in a fixed order:
// deterministic, and 30x fewer memory operations than fp32 atomics on [co][ci][tap] (4.7 M atomics per layer made this
// kernel 37 % of the first training step profile, profiles/r02_train_profile_v1.txt).
__global__ void __launch_bounds__(128, 1)
conv_wgrad_kernel(const __half* __restrict__ A, const __half* __restrict__ B, long long b_copy_stride, float* __restrict__ partial,
                  int Cout, int Cin, int taps, long long Ppad, int Wp, int BN, int steps_per_split, long long k_begin,
                  long long k_steps, int ab_fmt) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t stage_free[CW_STAGES];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ntiles_c = Cin / BN;
    const int tap = blockIdx.x / ntiles_c, c0 = (blockIdx.x % ntiles_c) * BN, mt = blockIdx.y;
    const long long s0 = static_cast<long long>(blockIdx.z) * steps_per_split;
    long long s1 = s0 + steps_per_split;
    if (s1 > k_steps) s1 = k_steps;
    const int nsteps = s1 > s0 ? static_cast<int>(s1 - s0) : 0;
    const int dy = taps == 9 ? tap / 3 - 1 : 0, dxi = taps == 9 ? tap % 3 : 1;      // copy index dxi holds dx = dxi - 1

    if (tid == 0) {
        for (int i = 0; i < CW_STAGES; ++i) mbar_init(&stage_free[i], 1);
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = umma_idesc_f16(128, BN, static_cast<uint32_t>(ab_fmt));
    const uint32_t lbo_b = static_cast<uint32_t>(BN) * 16u;

    const __half* Arow = A + (static_cast<size_t>(mt) * 128 + tid) * Ppad + k_begin;                 // one A row per thread
    const __half* Brow = B + dxi * b_copy_stride + (static_cast<size_t>(c0) + (tid < BN ? tid : 0)) * Ppad + k_begin +
                         static_cast<long long>(dy) * Wp;
    for (int i = 0; i < nsteps; ++i) {
        const int s = i % CW_STAGES;
        if (i >= CW_STAGES) mbar_wait(&stage_free[s], ((i / CW_STAGES) - 1) & 1);
        uint8_t* a_s = smem + s * (2 * 128 * 128);
        uint8_t* b_s = a_s + 128 * 128;
        const long long p0 = (s0 + i) * 64;
        uint4 va[8];
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) va[kc] = ldg_nc_v4(Arow + p0 + kc * 8);
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) sts_v4(smem_u32(a_s) + kc * 2048 + tid * 16, va[kc]);
        if (tid < BN) {
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) va[kc] = ldg_nc_v4(Brow + p0 + kc * 8);
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) sts_v4(smem_u32(b_s) + kc * lbo_b + tid * 16, va[kc]);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t ad = umma_desc_nosw(smem_u32(a_s) + k16 * 2 * 2048, 2048, 128);
                const uint64_t bd = umma_desc_nosw(smem_u32(b_s) + k16 * 2 * lbo_b, lbo_b, 128);
                umma_f16(tmem_base, ad, bd, idesc, (i | k16) != 0 ? 1u : 0u);
            }
            umma_commit(&stage_free[s]);
        }
    }
    if (tid == 0) umma_commit(&done_bar);
    {
        // this CTA's tile of the partial sums: [co = tid][BN] contiguous floats (zeros when the split got no K steps)
        float* out = partial + ((static_cast<size_t>(blockIdx.z) * gridDim.y + mt) * gridDim.x + blockIdx.x) * (128 * static_cast<size_t>(BN)) +
                     static_cast<size_t>(tid) * BN;
        if (nsteps > 0) {
            mbar_wait(&done_bar, 0);
            tc_fence_after_sync();
        }
        for (int cc = 0; cc < BN; cc += 32) {
            float v[32];
            if (nsteps > 0) tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + cc, v);
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
