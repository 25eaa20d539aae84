#include "hs_emu.h"
thread_local hs_dim3 hs_emu_threadIdx;
thread_local hs_dim3 hs_emu_blockIdx;
hs_dim3 hs_emu_blockDim, hs_emu_gridDim;
thread_local unsigned char *hs_emu_smem;
std::barrier<> *hs_emu_bar;
void hs_emu_barrier() { hs_emu_bar->arrive_and_wait(); }
