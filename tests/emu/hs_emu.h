// hs_emu.h -- CPU emulation of the CUDA execution model for the kernel bodies in csrc/*.cuh.
// TEST INFRASTRUCTURE ONLY: one OS thread per CUDA thread of a CTA, std::barrier for
// __syncthreads, CTAs run one after another.  Lets tests/ check index arithmetic, barrier
// placement and numerics of the real kernel source without a GPU.
#pragma once
#define HS_EMU 1
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include "hs_common.cuh"

extern std::barrier<> *hs_emu_bar, *hs_emu_bar2;

#include <mutex>
extern std::mutex hs_emu_launch_mu; // the launch state below is global: launches from several host threads take turns

template <class K, class... Args>
void hs_emu_launch(K kernel, long long grid, int block, size_t smem_bytes, Args... args)
{
    std::lock_guard<std::mutex> launch_guard(hs_emu_launch_mu);
    hs_emu_blockDim.x = block;
    hs_emu_gridDim.x = (unsigned)grid;
    std::vector<unsigned char> smem(smem_bytes + 64);
    for (long long bid = 0; bid < grid; ++bid) {
        std::barrier<> bar(block), bar2(block > 64 ? block - 64 : block); // bar2: the compute warps of a warp-specialised kernel (two producer warps on top)
        hs_emu_bar = &bar;
        hs_emu_bar2 = &bar2;
        std::vector<std::thread> th;
        th.reserve(block);
        for (int t = 0; t < block; ++t)
            th.emplace_back([&, t, bid]() {
                hs_emu_threadIdx.x = t;
                hs_emu_blockIdx.x = (unsigned)bid;
                hs_emu_smem = smem.data();
                kernel(args...);
            });
        for (auto &x : th) x.join();
    }
}
