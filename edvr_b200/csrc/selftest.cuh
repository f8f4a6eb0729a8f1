// selftest.cuh — minimal tcgen05 GEMM used to pin the shared-memory descriptor convention on
// real hardware: D[128, N] = A[128, K] * B[N, K]^T, fp16 operands (row-major, K contiguous in
// global memory), fp32 result.  One CTA, 128 threads; operands are laid out in shared memory
// exactly like the production kernels do (common.cuh planes).
#pragma once
#include "common.cuh"

namespace eb {

// variant bit 0: swap the LBO/SBO fields of both descriptors
// variant bit 1: leave the descriptor "version" bits [46,48) at 0
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                     int N, int K, int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    // A planes: [K/8][128 rows][16 B]; B planes: [K/8][N rows][16 B]
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + (K / 8) * 128 * 16;
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 128 * (K / 8); i += 128) {
        const int r = i % 128, kc = i / 128;
        *reinterpret_cast<uint4*>(a_s + (kc * 128 + r) * 16) =
            *reinterpret_cast<const uint4*>(A + static_cast<size_t>(r) * K + kc * 8);
    }
    for (int i = tid; i < N * (K / 8); i += 128) {
        const int r = i % N, kc = i / N;
        *reinterpret_cast<uint4*>(b_s + (kc * N + r) * 16) =
            *reinterpret_cast<const uint4*>(B + static_cast<size_t>(r) * K + kc * 8);
    }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t ad, bd;
            if (variant & 1) {
                ad = umma_desc_nosw(smem_u32(a_s) + k16 * 2 * lbo_a, sbo, lbo_a);
                bd = umma_desc_nosw(smem_u32(b_s) + k16 * 2 * lbo_b, sbo, lbo_b);
            } else {
                ad = umma_desc_nosw(smem_u32(a_s) + k16 * 2 * lbo_a, lbo_a, sbo);
                bd = umma_desc_nosw(smem_u32(b_s) + k16 * 2 * lbo_b, lbo_b, sbo);
            }
            if (variant & 2) { ad &= ~(3ull << 46); bd &= ~(3ull << 46); }
            umma_f16(tmem_base, ad, bd, idesc, k16 != 0 ? 1u : 0u);
        }
        umma_commit(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after_sync();
    const int row = tid;  // warp w owns TMEM lanes [32w, 32w+32)
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + c, v);
        for (int j = 0; j < 32; ++j) D[static_cast<size_t>(row) * N + c + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---- MMA issue-rate probe: R back-to-back MMAs on zeroed operands, cycles per MMA reported per CTA.
// layout: 0 = no-swizzle planes (LBO = plane pitch), 2/4/6 = 128/64/32-byte swizzled rows.  CG = 1 or 2 (CTA pair).
struct MmaRateParams {
    int M, N, layout, reps;
    int a_lbo, a_sbo, b_lbo, b_sbo;     // bytes
    int kstep_bytes;                    // start-address advance between the 4 K steps of a stage
    unsigned long long* cycles;         // [gridDim.x]
};

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const MmaRateParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 0) { if (CG == 2) tmem_alloc_pair(&tmem_slot, 256); else tmem_alloc(&tmem_slot, 256); }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const bool leader = CG == 1 || cluster_ctarank() == 0;
    long long t0 = 0, t1 = 0;
    if (tid == 0 && leader) {
        const uint32_t idesc = umma_idesc_f16(P.M, P.N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32 * 1024;
        t0 = clock64();
        for (int i = 0; i < P.reps; ++i) {
            const uint32_t off = (i & 3) * P.kstep_bytes;
            uint64_t ad, bd;
            if (P.layout == 0) {
                ad = umma_desc_nosw(a0 + off, P.a_lbo, P.a_sbo);
                bd = umma_desc_nosw(b0 + off, P.b_lbo, P.b_sbo);
            } else {
                ad = umma_desc_sw(a0 + off, P.a_sbo, P.layout);
                bd = umma_desc_sw(b0 + off, P.b_sbo, P.layout);
            }
            if (CG == 2) umma_f16_pair(tmem_base, ad, bd, idesc, i != 0);
            else umma_f16(tmem_base, ad, bd, idesc, i != 0);
        }
        if (CG == 2) umma_commit_pair(&done_bar, 3); else umma_commit(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    if (tid == 0 && leader) {
        t1 = clock64();
        P.cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { if (CG == 2) tmem_dealloc_pair(tmem_base, 256); else tmem_dealloc(tmem_base, 256); }
}

}  // namespace eb
