#include "hs_emu.h"
thread_local hs_dim3 hs_emu_threadIdx;
thread_local hs_dim3 hs_emu_blockIdx;
hs_dim3 hs_emu_blockDim, hs_emu_gridDim;
thread_local unsigned char *hs_emu_smem;
std::barrier<> *hs_emu_bar;
void hs_emu_barrier() { hs_emu_bar->arrive_and_wait(); }
std::mutex hs_emu_mbar_mu;
std::barrier<> *hs_emu_bar2;
void hs_emu_barrier2() { hs_emu_bar2->arrive_and_wait(); }
std::mutex hs_emu_launch_mu;
thread_local double hs_emu_tmem[128];
