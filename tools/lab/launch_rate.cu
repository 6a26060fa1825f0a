// Host-side launch throughput of programmatic dependent launches: T host threads, each on its own stream, each issuing
// `count` launches of a kernel that does nothing (grid of `ctas` CTAs of 128 threads), as launch_pdl (csrc/engine.h) does.
// Answers whether the ~140 k launches per second of four images in flight (84 k launches per ResNet-20 image at 1.66
// images/s) sit near what the driver can issue from four threads.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bin/launch_rate launch_rate.cu && bin/launch_rate
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_nothing(int *p)
{
    asm volatile("griddepcontrol.wait;");
    asm volatile("griddepcontrol.launch_dependents;");
    if (p && threadIdx.x == 9999)
        *p = 1;
}

static void issue(cudaStream_t s, int count, int ctas, bool pdl)
{
    for (int i = 0; i < count; i++)
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(128);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, k_nothing, (int *)nullptr);
    }
}

int main(int argc, char **argv)
{
    const int count = argc > 1 ? atoi(argv[1]) : 200000;
    for (int pdl = 1; pdl >= 0; pdl--)
        for (int ctas : { 1, 64, 640 })
            for (int T : { 1, 2, 4, 8 })
            {
                std::vector<cudaStream_t> st(T);
                for (auto &s : st)
                    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                issue(st[0], 1000, ctas, pdl);
                cudaDeviceSynchronize();
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back(issue, st[t], count / T, ctas, (bool)pdl);
                for (auto &t : th)
                    t.join();
                auto t1 = std::chrono::steady_clock::now();
                cudaDeviceSynchronize();
                auto t2 = std::chrono::steady_clock::now();
                const double issue_s = std::chrono::duration<double>(t1 - t0).count(), all_s = std::chrono::duration<double>(t2 - t0).count();
                printf("pdl %d  ctas %4d  threads %d: issued %7.0f k launches/s (host), completed %7.0f k/s\n", pdl, ctas, T,
                       count / issue_s / 1e3, count / all_s / 1e3);
                for (auto &s : st)
                    cudaStreamDestroy(s);
            }
    return 0;
}
