// ntt_lab.cu - stand-alone timing / equivalence bench of the NTT kernels of fhe-gpt-2_b200/csrc/ntt.cuh at N = 2^16.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I fhe-gpt-2_b200/csrc tools/lab/ntt_lab.cu -o /tmp/ntt_lab
// Random twiddles with correct Shoup companions: every kernel computes the same butterfly network on them, so
// variants are compared bit for bit with the two-pass kernels (whose equality with the reference's NTT is the
// business of tests/test_gpu_parity.py).
#include "ntt_fused.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct LabLd
{
    const u64 *src;
    int np = 1;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return job % np; }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const { return src[((size_t)job << 16) + idx]; }
};
struct LabSt
{
    static constexpr bool RAW = false;
    static constexpr bool BATCH = false;
    u64 *dst;
    int np = 1;
    __device__ __forceinline__ bool skip(int) const { return false; }
    __device__ __forceinline__ int prime(int job) const { return job % np; }
    __device__ __forceinline__ u64 pre(int, int, int, int, u64 v, const PrimeDev &) const { return v; }
    __device__ __forceinline__ void post(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        v = v >= pd.two_q ? v - pd.two_q : v;
        dst[((size_t)job << 16) + idx] = v >= pd.q ? v - pd.q : v;
    }
};
struct LabInvLd
{
    static constexpr bool TLAYOUT = false;
    const u64 *src;
    int np = 1;
    __device__ __forceinline__ int prime(int job) const { return job % np; }
    __device__ __forceinline__ u64 load(int job, int idx, const PrimeDev &) const { return src[((size_t)job << 16) + idx]; }
    __device__ __forceinline__ u64 load_t(int, int, int) const { return 0; }
};
struct LabInvSt
{
    u64 *dst;
    int np = 1;
    __device__ __forceinline__ int prime(int job) const { return job % np; }
    __device__ __forceinline__ void store(int job, int idx, u64 v, const PrimeDev &pd) const
    {
        dst[((size_t)job << 16) + idx] = v >= pd.q ? v - pd.q : v;
    }
};

typedef unsigned __int128 u128;

template <class F>
static float time_us(F f, int reps = 30)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 5; i++)
        f();
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++)
        f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms * 1000.f / reps;
}

int main(int argc, char **argv)
{
    const u64 q = 2251799780917249ull; // 51-bit prime of the CNN chain
    const size_t n = 1 << 16;
    const int max_jobs = 124;
    std::mt19937_64 rng(1);
    const int NP = 31; // distinct twiddle tables (same modulus: only the footprint matters here)
    std::vector<ulonglong2> tw(n * NP), itw(n * NP);
    for (size_t i = 0; i < n * NP; i++)
    {
        u64 w = rng() % q, v = rng() % q;
        tw[i] = make_ulonglong2(w, (u64)(((u128)w << 64) / q));
        itw[i] = make_ulonglong2(v, (u64)(((u128)v << 64) / q));
    }
    PrimeDev pd;
    pd.q = q;
    pd.two_q = 2 * q;
    u128 ratio = (~(u128)0) / q; // floor((2^128 - 1) / q) == floor(2^128 / q) for odd q
    pd.r0 = (u64)ratio;
    pd.r1 = (u64)(ratio >> 64);
    pd.ninv = rng() % q;
    pd.ninv_s = (u64)(((u128)pd.ninv << 64) / q);
    pd.ninvw = rng() % q;
    pd.ninvw_s = (u64)(((u128)pd.ninvw << 64) / q);
    std::vector<u64> h((size_t)max_jobs * n);
    for (auto &v : h)
        v = rng() % q;

    ulonglong2 *d_tw, *d_itw;
    PrimeDev *d_pd;
    u64 *d_in, *d_tmp, *d_out, *d_out2;
    CK(cudaMalloc(&d_tw, n * 16 * NP));
    CK(cudaMalloc(&d_itw, n * 16 * NP));
    CK(cudaMalloc(&d_pd, sizeof(pd) * NP));
    CK(cudaMalloc(&d_in, h.size() * 8));
    CK(cudaMalloc(&d_tmp, h.size() * 8));
    CK(cudaMalloc(&d_out, h.size() * 8));
    CK(cudaMalloc(&d_out2, h.size() * 8));
    CK(cudaMemcpy(d_tw, tw.data(), n * 16 * NP, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_itw, itw.data(), n * 16 * NP, cudaMemcpyHostToDevice));
    for (int i = 0; i < NP; i++)
        CK(cudaMemcpy(d_pd + i, &pd, sizeof(pd), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    NttTables T{ d_tw, d_itw, d_pd, 16, 1 };
    LabLd ld{ d_in };
    LabSt st{ d_out }, st2{ d_out2 };
    LabInvLd ild{ d_in };
    LabInvSt ist{ d_out }, ist2{ d_out2 };

    CK(cudaFuncSetAttribute(k_fwd_fused<LabLd, LabSt, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
    CK(cudaFuncSetAttribute(k_fwd_fused<LabLd, LabSt, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));
    CK(cudaFuncSetAttribute(k_inv_fused<LabInvLd, LabInvSt>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM));

    // ---- equivalence -----------------------------------------------------------------------
    auto compare = [&](const char *what, int jobs) {
        std::vector<u64> a((size_t)jobs * n), b((size_t)jobs * n);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(a.data(), d_out, a.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_out2, b.size() * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); i++)
            bad += a[i] != b[i];
        printf("%-40s %s (%zu of %zu words differ)\n", what, bad ? "DIFFERENT" : "identical", bad, a.size());
        return bad == 0;
    };
    bool ok = true;
    for (int wide = 0; wide < 2; wide++)
    {
        const int jobs = 5;
        CK(cudaMemset(d_out, 0, h.size() * 8));
        CK(cudaMemset(d_out2, 0xff, h.size() * 8));
        if (wide)
        {
            k_fwd_cols<8, LabLd, true, 16><<<dim3(16, jobs), 256>>>(ld, d_tmp, T);
            k_fwd_blocks<LabSt, true><<<dim3(16, jobs), 256, 32768>>>(d_tmp, st, T);
            k_fwd_fused<LabLd, LabSt, true><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ld, st2, T);
        }
        else
        {
            k_fwd_cols<8, LabLd, false, 16><<<dim3(16, jobs), 256>>>(ld, d_tmp, T);
            k_fwd_blocks<LabSt, false><<<dim3(16, jobs), 256, 32768>>>(d_tmp, st, T);
            k_fwd_fused<LabLd, LabSt, false><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ld, st2, T);
        }
        CK(cudaGetLastError());
        ok &= compare(wide ? "forward fused vs two-pass (wide)" : "forward fused vs two-pass (lazy)", jobs);
    }
    {
        const int jobs = 5;
        CK(cudaMemset(d_out, 0, h.size() * 8));
        CK(cudaMemset(d_out2, 0xff, h.size() * 8));
        k_inv_blocks<LabInvLd><<<dim3(16, jobs), 256, 32768>>>(ild, d_tmp, T);
        k_inv_cols<8, LabInvSt, 16><<<dim3(16, jobs), 256>>>(d_tmp, ist, T);
        k_inv_fused<LabInvLd, LabInvSt><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ild, ist2, T);
        CK(cudaGetLastError());
        ok &= compare("inverse fused vs two-pass", jobs);
    }

    if (argc > 1 && !strcmp(argv[1], "ncu"))
    {
        // one launch of each pass at 124 limb-polynomials (a 31-limb key-switch chunk of 4 output moduli) for ncu
        const int jobs = 124;
        for (int rep = 0; rep < 2; rep++)
        {
            k_fwd_cols<8, LabLd, true, 16><<<dim3(16, jobs), 256>>>(ld, d_tmp, T);
            k_fwd_blocks<LabSt, true><<<dim3(32, jobs), 128, 16384>>>(d_tmp, st, T);
            k_inv_blocks<LabInvLd><<<dim3(16, jobs), 256, 32768>>>(ild, d_tmp, T);
            k_inv_cols<8, LabInvSt, 16><<<dim3(16, jobs), 256>>>(d_tmp, ist, T);
        }
        CK(cudaDeviceSynchronize());
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "cold"))
    {
        // the situation inside a key switch: 31 different twiddle tables (31 MiB per direction) and an L2 that other
        // kernels have swept since the table was last used.  Each launch is timed alone after a 512 MiB memset.
        const bool persist = argc > 2 && !strcmp(argv[2], "persist");
        cudaStream_t st_;
        CK(cudaStreamCreate(&st_));
        if (persist)
        {
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)40 << 20));
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = d_tw;
            av.accessPolicyWindow.num_bytes = n * 16 * NP;
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st_, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        void *flush;
        CK(cudaMalloc(&flush, (size_t)512 << 20));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        LabLd ldp{ d_in, NP };
        LabSt stp{ d_out, NP };
        auto cold = [&](auto f, bool do_flush) {
            float tot = 0;
            for (int rep = 0; rep < 6; rep++)
            {
                if (do_flush)
                    CK(cudaMemsetAsync(flush, rep, (size_t)512 << 20, st_));
                CK(cudaEventRecord(e0, st_));
                f();
                CK(cudaEventRecord(e1, st_));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep)
                    tot += ms;
            }
            return tot * 1000.f / 5;
        };
        printf("%s\n%5s | %10s %10s | %10s %10s   (us per launch, %d twiddle tables)\n", persist ? "twiddles persisting in L2" : "default L2 policy", "jobs", "fcols warm", "fcols cold", "fblk warm", "fblk cold", NP);
        for (int jobs : { 8, 31, 62, 124 })
        {
            float a = cold([&] { k_fwd_cols<8, LabLd, true, 8><<<dim3(32, jobs), 128, 0, st_>>>(ldp, d_tmp, T); }, false);
            float b = cold([&] { k_fwd_cols<8, LabLd, true, 8><<<dim3(32, jobs), 128, 0, st_>>>(ldp, d_tmp, T); }, true);
            float c = cold([&] { k_fwd_blocks<LabSt, true><<<dim3(32, jobs), 128, 16384, st_>>>(d_tmp, stp, T); }, false);
            float d = cold([&] { k_fwd_blocks<LabSt, true><<<dim3(32, jobs), 128, 16384, st_>>>(d_tmp, stp, T); }, true);
            printf("%5d | %10.2f %10.2f | %10.2f %10.2f\n", jobs, a, b, c, d);
        }
        return 0;
    }
    // ---- timing -----------------------------------------------------------------------------
    printf("\nmicroseconds per launch (warm L2 where the working set fits), N = 2^16, one 51-bit prime\n");
    printf("%5s | %8s %8s %8s %8s %8s | %8s | %8s %8s %8s %8s | %8s\n", "jobs", "fcols16", "fcols8", "fblk256", "fblk128", "fblk64",
           "ffused", "iblk256", "iblk128", "icols16", "icols8", "ifused");
    const int js[] = { 1, 2, 3, 4, 6, 8, 12, 16, 18, 24, 31, 36, 48, 62, 93, 124 };
    for (int jobs : js)
    {
        float a = time_us([&] { k_fwd_cols<8, LabLd, true, 16><<<dim3(16, jobs), 256>>>(ld, d_tmp, T); });
        float b = time_us([&] { k_fwd_cols<8, LabLd, true, 8><<<dim3(32, jobs), 128>>>(ld, d_tmp, T); });
        float c = time_us([&] { k_fwd_blocks<LabSt, true><<<dim3(16, jobs), 256, 32768>>>(d_tmp, st, T); });
        float c2 = time_us([&] { k_fwd_blocks<LabSt, true><<<dim3(32, jobs), 128, 16384>>>(d_tmp, st, T); });
        float d = time_us([&] { k_fwd_blocks<LabSt, true><<<dim3(64, jobs), 64, 8192>>>(d_tmp, st, T); });
        float e = jobs <= 36 ? time_us([&] { k_fwd_fused<LabLd, LabSt, true><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ld, st2, T); }) : 0.f;
        float f = time_us([&] { k_inv_blocks<LabInvLd><<<dim3(16, jobs), 256, 32768>>>(ild, d_tmp, T); });
        float g = time_us([&] { k_inv_blocks<LabInvLd><<<dim3(32, jobs), 128, 16384>>>(ild, d_tmp, T); });
        float hh = time_us([&] { k_inv_cols<8, LabInvSt, 16><<<dim3(16, jobs), 256>>>(d_tmp, ist, T); });
        float i = time_us([&] { k_inv_cols<8, LabInvSt, 8><<<dim3(32, jobs), 128>>>(d_tmp, ist, T); });
        float k = jobs <= 36 ? time_us([&] { k_inv_fused<LabInvLd, LabInvSt><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ild, ist2, T); }) : 0.f;
        printf("%5d | %8.2f %8.2f %8.2f %8.2f %8.2f | %8.2f | %8.2f %8.2f %8.2f %8.2f | %8.2f\n", jobs, a, b, c, c2, d, e, f, g, hh, i, k);
    }
    // programmatic dependent launch: the same pair with the launch of the second pass (and of the next pair) set up
    // while the previous kernel runs
    {
        auto pdl = [&](auto kern, dim3 grid, unsigned threads, size_t smem, auto... args) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = 0;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, kern, args...));
        };
        printf("\npairs back to back, plain launches against programmatic dependent launches, microseconds\n%5s | %10s %10s | %10s %10s\n", "jobs", "fwd", "fwd pdl", "inv", "inv pdl");
        for (int jobs : { 1, 2, 4, 8, 16, 31, 62 })
        {
            const unsigned bt = jobs * 32 < 296 ? 64u : 128u;
            float a = time_us([&] {
                k_fwd_cols<8, LabLd, true, 8><<<dim3(32, jobs), 128>>>(ld, d_tmp, T);
                k_fwd_blocks<LabSt, true><<<dim3(4096 / bt, jobs), bt, bt * 128>>>(d_tmp, st, T);
            });
            float b = time_us([&] {
                pdl(k_fwd_cols<8, LabLd, true, 8>, dim3(32, jobs), 128u, (size_t)0, ld, d_tmp, T);
                pdl(k_fwd_blocks<LabSt, true>, dim3(4096 / bt, jobs), bt, (size_t)bt * 128, (const u64 *)d_tmp, st, T);
            });
            float c = time_us([&] {
                k_inv_blocks<LabInvLd><<<dim3(4096 / bt, jobs), bt, bt * 128>>>(ild, d_tmp, T);
                k_inv_cols<8, LabInvSt, 8><<<dim3(32, jobs), 128>>>(d_tmp, ist, T);
            });
            float d = time_us([&] {
                pdl(k_inv_blocks<LabInvLd>, dim3(4096 / bt, jobs), bt, (size_t)bt * 128, ild, d_tmp, T);
                pdl(k_inv_cols<8, LabInvSt, 8>, dim3(32, jobs), 128u, (size_t)0, (const u64 *)d_tmp, ist, T);
            });
            printf("%5d | %10.2f %10.2f | %10.2f %10.2f\n", jobs, a, b, c, d);
        }
        // equality of the PDL chain with the plain one
        CK(cudaMemset(d_out, 0, h.size() * 8));
        CK(cudaMemset(d_out2, 0xff, h.size() * 8));
        k_fwd_cols<8, LabLd, true, 8><<<dim3(32, 5), 128>>>(ld, d_tmp, T);
        k_fwd_blocks<LabSt, true><<<dim3(32, 5), 128, 16384>>>(d_tmp, st, T);
        CK(cudaDeviceSynchronize());
        for (int rep = 0; rep < 20; rep++)
        {
            pdl(k_fwd_cols<8, LabLd, true, 8>, dim3(32, 5), 128u, (size_t)0, ld, d_tmp, T);
            pdl(k_fwd_blocks<LabSt, true>, dim3(32, 5), 128u, (size_t)16384, (const u64 *)d_tmp, st2, T);
        }
        ok &= compare("PDL chain vs plain launches", 5);
    }
    // pairs back to back (what the engine does today) against the fused kernel
    printf("\npairs back to back (two launches) against one fused launch, microseconds\n%5s | %10s %10s | %10s %10s\n", "jobs", "fwd 2pass", "fwd fused", "inv 2pass", "inv fused");
    for (int jobs : { 1, 2, 4, 8, 12, 16, 18, 24, 31 })
    {
        const bool small = jobs * 16 < 296;
        float a = time_us([&] {
            if (small)
            {
                k_fwd_cols<8, LabLd, true, 8><<<dim3(32, jobs), 128>>>(ld, d_tmp, T);
                k_fwd_blocks<LabSt, true><<<dim3(64, jobs), 64, 8192>>>(d_tmp, st, T);
            }
            else
            {
                k_fwd_cols<8, LabLd, true, 16><<<dim3(16, jobs), 256>>>(ld, d_tmp, T);
                k_fwd_blocks<LabSt, true><<<dim3(16, jobs), 256, 32768>>>(d_tmp, st, T);
            }
        });
        float b = time_us([&] { k_fwd_fused<LabLd, LabSt, true><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ld, st2, T); });
        float c = time_us([&] {
            if (small)
            {
                k_inv_blocks<LabInvLd><<<dim3(64, jobs), 64, 8192>>>(ild, d_tmp, T);
                k_inv_cols<8, LabInvSt, 8><<<dim3(32, jobs), 128>>>(d_tmp, ist, T);
            }
            else
            {
                k_inv_blocks<LabInvLd><<<dim3(16, jobs), 256, 32768>>>(ild, d_tmp, T);
                k_inv_cols<8, LabInvSt, 16><<<dim3(16, jobs), 256>>>(d_tmp, ist, T);
            }
        });
        float d = time_us([&] { k_inv_fused<LabInvLd, LabInvSt><<<dim3(8, jobs), FUSED_THREADS, FUSED_SMEM>>>(ild, ist2, T); });
        printf("%5d | %10.2f %10.2f | %10.2f %10.2f\n", jobs, a, b, c, d);
    }
    printf("%s\n", ok ? "LAB OK" : "LAB FAILED");
    return ok ? 0 : 1;
}
