// occupancy probe: how many CTAs of (threads, regs via launch bounds, dynamic smem) fit one SM
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ unsigned char sm[];
__global__ void __launch_bounds__(192, 2) k192(double *p) { sm[threadIdx.x] = 1; p[threadIdx.x] = sm[0]; }
__global__ void __launch_bounds__(128, 2) k128(double *p) { sm[threadIdx.x] = 1; p[threadIdx.x] = sm[0]; }
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: smem/SM %zu, smem/block optin %zu, reserved/block %zu, regs/SM %d\n", pr.name, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.reservedSharedMemPerBlock, pr.regsPerMultiprocessor);
    for (int carve = 0; carve < 2; ++carve)
        for (int smem = 96 * 1024; smem <= 116 * 1024; smem += 1024) {
            cudaFuncSetAttribute(k192, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (carve) cudaFuncSetAttribute(k192, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k192, 192, smem);
            printf("carve=%d smem=%d (%d KB) -> %d CTAs/SM\n", carve, smem, smem / 1024, n);
        }
    return 0;
}
