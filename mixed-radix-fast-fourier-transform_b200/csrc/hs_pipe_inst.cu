// hs_pipe_inst.cu -- the persistent pipeline kernels of ONE two-pass size (compiled once per size with
// -DHS_PIPE_ID=CT_Txx so that the sizes build in parallel).  sm_100a only.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hs_abi.h"
#include "hs_cuda_util.h"

typedef cudaStream_t hs_stream_t;

#include "hs_pipe.cuh"

// tensor map of a staged phase's input: [transforms][rows][2*ncols doubles]; an unstaged phase gets an empty one
template <class PH>
static CUtensorMap phase_tmap(const hs_pipe_args &a, int j)
{
    CUtensorMap m;
    memset(&m, 0, sizeof m);
    if (PH::STAGED) {
        const void *base = j == 0 ? a.in : a.ring[j - 1];
        const long long count = j == 0 ? a.howmany : (long long)a.nslot * a.G;
        if (a.ph[j].g.in_c != 1 || a.ph[j].g.blk_log != 0) { fprintf(stderr, "Error: a staged pipeline phase needs adjacent columns\n"); exit(EXIT_FAILURE); }
        m = make_tmap(base, a.ph[j].g, PH::ROWS, PH::C, count);
    }
    return m;
}

template <class PH0, class PH1, class PH2>
static void launch_pipe(hs_pipe_args &a, cudaStream_t st)
{
    using SH = HsPipeShape<PH0, PH1, PH2>;
    static_assert(SH::SMEM <= 227 * 1024, "staging and exchange buffers do not fit");
    auto kern = hs_pipe_kernel<PH0, PH1, PH2>;
    static int per_sm = 0;
    if (!per_sm) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SH::SMEM));
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SH::NTHREADS, SH::SMEM));
        if (per_sm < 1) per_sm = 1;
        if (const char *e = getenv("HSFFT_PIPE_CTAS")) { const int v = atoi(e); if (v > 0 && v < per_sm) per_sm = v; }
        if (getenv("HSFFT_VERBOSE")) fprintf(stderr, "hsfft: pipeline kernel: %d CTAs/SM of %d threads, %d bytes smem\n", per_sm, SH::NTHREADS, SH::SMEM);
    }
    long long items = 0;
    for (int j = 0; j < a.nphase; ++j) items += (long long)a.ph[j].tiles * a.howmany;
    long long blocks = (long long)hs_cu_sm_count() * per_sm;
    if (blocks > items) blocks = items;
    const CUtensorMap tm0 = phase_tmap<PH0>(a, 0), tm1 = phase_tmap<PH1>(a, 1), tm2 = phase_tmap<PH2>(a, 2);
    HS_CUDA_CHECK(cudaMemsetAsync(a.ctl, 0, HS_PIPE_CTL_BYTES(a.ngroups), st));
    kern<<<(unsigned)blocks, SH::NTHREADS, SH::SMEM, st>>>(a, tm0, tm1, tm2);
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
