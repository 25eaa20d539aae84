// hs_pipe_inst.cu -- the persistent pipeline kernels of ONE two-pass size (compiled once per size with
// -DHS_PIPE_ID=CT_Txx so that the sizes build in parallel).  sm_100a only.
#include <cstdio>
#include <cstdlib>

#include "hs_abi.h"
#include "hs_cuda_util.h"

typedef cudaStream_t hs_stream_t;

#include "hs_pipe.cuh"

template <class PH0, class PH1, class PH2, int MINB>
static void launch_pipe(hs_pipe_args &a, cudaStream_t st)
{
    using SH = HsPipeShape<PH0, PH1, PH2>;
    static_assert(SH::SMEM <= 227 * 1024, "tile buffers do not fit");
    auto kern = hs_pipe_kernel<PH0, PH1, PH2, MINB>;
    static int per_sm = 0;
    if (!per_sm) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SH::SMEM));
        HS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SH::NT, SH::SMEM));
        if (per_sm < 1) per_sm = 1;
        if (const char *e = getenv("HSFFT_PIPE_CTAS")) { const int v = atoi(e); if (v > 0 && v < per_sm) per_sm = v; }
    }
    const long long ngroups = (a.howmany + a.G - 1) / a.G;
    long long items = 0;
    for (int j = 0; j < a.nphase; ++j) items += (long long)a.ph[j].tiles * a.howmany;
    long long blocks = (long long)hs_cu_sm_count() * per_sm;
    if (blocks > items) blocks = items;
    HS_CUDA_CHECK(cudaMemsetAsync(a.ctl, 0, HS_PIPE_CTL_BYTES(ngroups), st));
    kern<<<(unsigned)blocks, SH::NT, SH::SMEM, st>>>(a);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_cu_count_launch();
}

#include "hs_pipe_table.cuh"

#define HS_PIPE_FN2(id) hs_cu_pipe_##id
#define HS_PIPE_FN(id) HS_PIPE_FN2(id)

// HS_PIPE_NAME is the size's suffix (14..20): one entry point per translation unit
extern "C" int HS_PIPE_FN(HS_PIPE_NAME)(int variant, int sgn, hs_pipe_args *a, void *stream, int query)
{
    return pipe_dispatch<HS_PIPE_ID>(variant, sgn, *a, (cudaStream_t)stream, query != 0);
}
