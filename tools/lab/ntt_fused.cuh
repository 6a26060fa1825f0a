// ntt_fused.cuh - experiment kept for the record (tools/lab/ntt_lab.cu): one-launch N = 2^16 transforms on a cluster of
// 8 CTAs exchanging through distributed shared memory.  Bit-identical to the two-pass kernels, but SLOWER on B200 at
// every job count (profiles/r2_ntt_lab.md): one limb-polynomial on 8 SMs is issue-bound at ~18 us, the two-pass
// kernels spread it over 32-64 SMs and need ~6 us per pass.  Not part of the product.
#pragma once
#include "ntt.cuh"

// ============================================================================================
// Whole-transform kernels for launches of a few limb-polynomials at N = 2^16 (the special limbs of a ModDown, the
// last limb of a rescale, everything a ResNet does below ~18 limbs).  The two-pass scheme above would occupy 16-64
// small CTAs per limb-polynomial twice, with the intermediate going through L2 and a second launch in between.
// Here one thread-block cluster of 8 CTAs (512 threads, 64 KiB of shared memory each: together the 512 KiB of one
// limb-polynomial) does all 16 stages:
//   forward:  CTA j runs the column pass on columns [32j, 32j+32) (all 256 rows), then every thread writes its 16
//             results straight into the shared memory of the CTA that owns their row (distributed shared memory: a
//             warp writes 256 contiguous bytes of one peer), and CTA j runs the block pass on rows [32j, 32j+32);
//   inverse:  the mirror image (block pass on rows, exchange, column pass on columns, N^-1 folded in).
// Two cluster barriers per transform: one before the exchange (every CTA is done reading its own tile), one after.
// Load / Store are the functors of the two-pass kernels, so every fusion (Galois gather, base conversion, rescale and
// ModDown tails ...) is available unchanged.
// ============================================================================================
#include <cooperative_groups.h>

constexpr int FUSED_CLUSTER = 8;
constexpr int FUSED_THREADS = 512;
constexpr int FUSED_SMEM = 8192 * 8;

template <class Load, class Store, bool WIDE>
__global__ void __cluster_dims__(FUSED_CLUSTER, 1, 1) __launch_bounds__(FUSED_THREADS, 1)
    k_fwd_fused(Load ld, Store st, NttTables T)
{
    extern __shared__ __align__(16) u64 fsm[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int job = blockIdx.y;
    if (ld.skip(job) || st.skip(job)) // uniform over the cluster
        return;
    const int j = blockIdx.x; // rank in the cluster
    const int pi = st.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << 16;
    const ulonglong2 *tw = T.tw + (size_t)pi * n;
    const u64 neg_q = 0ull - pd.q, four_q = 2 * pd.two_q;
    u64 x[16];
    {
        // column pass: thread (c, t) owns column 32j + c, rows t + 16k, then rows 16t + k
        const int c = threadIdx.x & 31, t = threadIdx.x >> 5;
        const int col = j * 32 + c;
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = ld.load(job, (t + 16 * k) * 256 + col, pd);
        if (WIDE)
            fwd_radix_wide<4>(x, tw, 1u, neg_q, four_q);
        else
            fwd_radix<4>(x, tw, 1u, pd.q, pd.two_q);
#pragma unroll
        for (int k = 0; k < 16; k++)
            fsm[(t + 16 * k) * 32 + c] = x[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = fsm[(16 * t + k) * 32 + c];
        if (WIDE)
            fwd_radix_wide<4>(x, tw, 16u + (unsigned)t, neg_q, four_q);
        else
            fwd_radix<4>(x, tw, 16u + (unsigned)t, pd.q, pd.two_q);
        cluster.sync();
        // row 16t + k belongs to CTA t >> 1, where it is local row 16 (t & 1) + k
        u64 *peer = cluster.map_shared_rank(fsm, (unsigned)(t >> 1));
#pragma unroll
        for (int k = 0; k < 16; k++)
            peer[(16 * (t & 1) + k) * 256 + swz(col)] = x[k];
        cluster.sync();
    }
    {
        const int t = threadIdx.x & 15, lb = threadIdx.x >> 4;
        const int blk = j * 32 + lb;
        const unsigned B = 256u + (unsigned)blk;
        u64 *s = fsm + lb * 256;
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = s[swz(t + 16 * k)];
        if (WIDE)
            fwd_radix_wide<4>(x, tw, B, neg_q, four_q);
        else
            fwd_radix<4>(x, tw, B, pd.q, pd.two_q);
#pragma unroll
        for (int k = 0; k < 16; k++)
            s[swz(t + 16 * k)] = x[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = s[swz(16 * t + k)];
        if (WIDE)
        {
            fwd_radix_wide<4>(x, tw, 16u * B + (unsigned)t, neg_q, four_q);
            if (!Store::RAW)
            {
#pragma unroll
                for (int k = 0; k < 16; k++)
                    x[k] = barrett64_r32(x[k], pd);
            }
        }
        else
            fwd_radix<4>(x, tw, 16u * B + (unsigned)t, pd.q, pd.two_q);
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = st.pre(job, blk, t, k, x[k], pd);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++)
            s[swz(16 * t + k)] = x[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++)
            st.post(job, blk * 256 + t + 16 * k, s[swz(t + 16 * k)], pd);
    }
}

template <class Load, class Store>
__global__ void __cluster_dims__(FUSED_CLUSTER, 1, 1) __launch_bounds__(FUSED_THREADS, 1)
    k_inv_fused(Load ld, Store st, NttTables T)
{
    extern __shared__ __align__(16) u64 fsm[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int job = blockIdx.y;
    const int j = blockIdx.x;
    const int pi = ld.prime(job);
    const PrimeDev pd = T.primes[pi];
    const size_t n = size_t(1) << 16;
    const ulonglong2 *itw = T.itw + (size_t)pi * n;
    u64 x[16];
    {
        const int t = threadIdx.x & 15, lb = threadIdx.x >> 4;
        const int blk = j * 32 + lb;
        const unsigned B = 256u + (unsigned)blk;
        u64 *s = fsm + lb * 256;
        if (Load::TLAYOUT)
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
                x[k] = ld.load_t(job, blk, k * 16 + t);
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
                s[swz(t + 16 * k)] = ld.load(job, blk * 256 + t + 16 * k, pd);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 16; k++)
                x[k] = s[swz(16 * t + k)];
            __syncwarp();
        }
        inv_radix<4, false>(x, itw, 16u * B + (unsigned)t, pd);
#pragma unroll
        for (int k = 0; k < 16; k++)
            s[swz(16 * t + k)] = x[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = s[swz(t + 16 * k)];
        inv_radix<4, false>(x, itw, B, pd);
        cluster.sync();
        // element e = t + 16k of row blk is column e: CTA e >> 5 = k >> 1, tile position [row blk][e & 31]
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            u64 *peer = cluster.map_shared_rank(fsm, (unsigned)(k >> 1));
            peer[blk * 32 + 16 * (k & 1) + t] = x[k];
        }
        cluster.sync();
    }
    {
        const int c = threadIdx.x & 31, t = threadIdx.x >> 5;
        const int col = j * 32 + c;
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = fsm[(16 * t + k) * 32 + c];
        inv_radix<4, false>(x, itw, 16u + (unsigned)t, pd);
#pragma unroll
        for (int k = 0; k < 16; k++)
            fsm[(16 * t + k) * 32 + c] = x[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++)
            x[k] = fsm[(t + 16 * k) * 32 + c];
        inv_radix<4, true>(x, itw, 1u, pd);
#pragma unroll
        for (int k = 0; k < 16; k++)
            st.store(job, (t + 16 * k) * 256 + col, x[k], pd);
    }
}
