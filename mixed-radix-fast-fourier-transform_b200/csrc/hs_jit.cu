// hs_jit.cu -- plan-time specialisation: any length whose stages fit one SM gets its own compile-time kernel.
//
// The compile-time plans of hs_plans.cuh cover the sizes BASELINE.json names; every other length used to run
// through the run-time planned engine (hs_generic.cuh: radices, strides and table offsets as kernel arguments,
// ~10 % of the HBM roofline).  The kernels of hs_smem_fft.cuh are templates over (N, threads per transform,
// transforms per CTA, padding, stage radices, pass fusion), so here the plan builder picks those parameters for the
// length at hand, writes a ten-line translation unit that instantiates the transform for them and compiles it
// with NVRTC for sm_100a when the plan is created (about a second, once per (N, direction) and process).  The
// headers NVRTC needs are embedded in the library (build/hs_jit_sources.inc, written by build.py from csrc/*.cuh),
// NVRTC itself is dlopen'ed: where libnvrtc.so is missing, or HSFFT_JIT=0, the run-time engine still serves the
// length.  Stage order, butterflies, literals and twiddle values are exactly those of the static plans, so parity
// with the reference is unchanged.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include "hs_abi.h"
#include "hs_cuda_util.h"
#include "hs_jit_sources.inc" // hs_jit_src_names[], hs_jit_src_texts[], hs_jit_src_count

// ------------------------------------------------------------------------------------------------ NVRTC, dlopen'ed
typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char *);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char *, const char **);
    bool ok;
};
static Nvrtc *nvrtc()
{
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n.ok ? &n : nullptr;
    tried = true;
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    void *h = nullptr;
    for (const char *nm : names) if ((h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) return nullptr;
#define HS_SYM(field, sym) *(void **)(&n.field) = dlsym(h, sym); if (!n.field) return nullptr
    HS_SYM(CreateProgram, "nvrtcCreateProgram");
    HS_SYM(CompileProgram, "nvrtcCompileProgram");
    HS_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    HS_SYM(GetCUBIN, "nvrtcGetCUBIN");
    HS_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    HS_SYM(GetProgramLog, "nvrtcGetProgramLog");
    HS_SYM(DestroyProgram, "nvrtcDestroyProgram");
    HS_SYM(AddNameExpression, "nvrtcAddNameExpression");
    HS_SYM(GetLoweredName, "nvrtcGetLoweredName");
#undef HS_SYM
    n.ok = true;
    return &n;
}

// ------------------------------------------------------------------------------------------------ driver entry points
typedef CUresult (*fn_cuModuleLoadData)(CUmodule *, const void *);
typedef CUresult (*fn_cuModuleGetFunction)(CUfunction *, CUmodule, const char *);
typedef CUresult (*fn_cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int);
typedef CUresult (*fn_cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
typedef CUresult (*fn_cuModuleUnload)(CUmodule);
typedef CUresult (*fn_cuFuncGetAttribute)(int *, CUfunction_attribute, CUfunction);
struct Driver {
    fn_cuModuleLoadData ModuleLoadData;
    fn_cuModuleGetFunction ModuleGetFunction;
    fn_cuFuncSetAttribute FuncSetAttribute;
    fn_cuLaunchKernel LaunchKernel;
    fn_cuModuleUnload ModuleUnload;
    fn_cuFuncGetAttribute FuncGetAttribute;
    bool ok;
};
static Driver *driver()
{
    static Driver d;
    static bool tried = false;
    if (tried) return d.ok ? &d : nullptr;
    tried = true;
    cudaDriverEntryPointQueryResult q;
#define HS_DRV(field, sym)                                                                                             \
    { void *p = nullptr; if (cudaGetDriverEntryPoint(sym, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr; \
      *(void **)(&d.field) = p; }
    HS_DRV(ModuleLoadData, "cuModuleLoadData");
    HS_DRV(ModuleGetFunction, "cuModuleGetFunction");
    HS_DRV(FuncSetAttribute, "cuFuncSetAttribute");
    HS_DRV(LaunchKernel, "cuLaunchKernel");
    HS_DRV(ModuleUnload, "cuModuleUnload");
    HS_DRV(FuncGetAttribute, "cuFuncGetAttribute");
#undef HS_DRV
    d.ok = true;
    return &d;
}

// ------------------------------------------------------------------------------------------------ choosing the plan
static int pad_of(int p, int logpad) { return logpad ? p + (p >> logpad) : p; }

// shared-memory wavefronts per ideal wavefront over all exchanges (1.0 = conflict free); same model as tests/emu/bank_check.py:
// a fused pass looks like one stage of radix RA*RB to shared memory
static double conflict_ratio(int N, int T, int logpad, const std::vector<int> &pass_radix)
{
    long long total = 0, ideal = 0;
    int m = 1;
    const int np = (int)pass_radix.size();
    for (int s = 0; s < np; ++s) {
        const int R = pass_radix[s], QP = N / (m * R), NB = N / R, B = (NB + T - 1) / T;
        for (int w0 = 0; w0 < T; w0 += 32) {
            for (int u = 0; u < B; ++u) {
                for (int dir = 0; dir < 2; ++dir) { // 0: reads of this pass (s > 0), 1: writes (s < np-1)
                    if ((dir == 0 && s == 0) || (dir == 1 && s == np - 1)) continue;
                    for (int i = 0; i < R; ++i) {
                        for (int q0 = w0; q0 < w0 + 32 && q0 < T; q0 += 8) {
                            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[8][8], lanes = 0;
                            for (int t = q0; t < q0 + 8 && t < T && t < w0 + 32; ++t) {
                                const int b = t + T * u;
                                if (b >= NB) continue;
                                const int q = b % QP, k = b / QP;
                                const int a = dir == 0 ? pad_of(q + QP * i + QP * R * k, logpad) : pad_of(b + NB * i, logpad);
                                const int g = a & 7;
                                bool dup = false;
                                for (int z = 0; z < cnt[g]; ++z) dup = dup || seen[g][z] == a;
                                if (!dup && cnt[g] < 8) seen[g][cnt[g]++] = a;
                                ++lanes;
                            }
                            if (!lanes) continue;
                            int worst = 1;
                            for (int g = 0; g < 8; ++g) worst = cnt[g] > worst ? cnt[g] : worst;
                            total += worst;
                            ideal += 1;
                        }
                    }
                }
            }
        }
        m *= R;
    }
    return ideal ? (double)total / (double)ideal : 1.0;
}

struct JitChoice {
    int T, X, logpad, groups, minb, npass;
    bool fat;
    std::vector<int> pass_radix;
};

static int env_or(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static bool choose(int n, const int *rx, int nst, JitChoice &c, bool allow_fat = true)
{
    if (nst < 2 || nst > 8 || n > 8192 || n < 16) return false;
    for (int s = 0; s < nst; ++s) {
        const int r = rx[s];
        if (!(r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16 || r == 11 || r == 13 || r == 17)) return false;
    }
    // fuse neighbouring stages while the pair stays within 16 points per butterfly (one exchange less each)
    // (HSFFT_JIT_FUSE / _T / _X / _MINB / _LOGPAD override the choices below: tuning aids)
    const int fuse_max = env_or("HSFFT_JIT_FUSE", 16);
    c.pass_radix.clear();
    c.groups = 0;
    int maxpp = 1;
    for (int s = 0; s < nst;) {
        int pp = rx[s], sz = 1;
        if (s + 1 < nst && rx[s] * rx[s + 1] <= fuse_max) { pp = rx[s] * rx[s + 1]; sz = 2; }
        c.pass_radix.push_back(pp);
        c.groups = c.groups * 10 + sz;
        maxpp = pp > maxpp ? pp : maxpp;
        s += sz;
    }
    c.npass = (int)c.pass_radix.size();
    if (c.npass < 2) { // everything fused into one pass would leave no exchange to switch buffers on: keep the stages apart
        c.pass_radix.assign(rx, rx + nst);
        c.groups = 0;
        c.npass = nst;
        maxpp = 1;
        for (int s = 0; s < nst; ++s) maxpp = rx[s] > maxpp ? rx[s] : maxpp;
    }
    // Threads per transform: the widest pass has exactly one butterfly per thread; two when the transform would otherwise
    // not fit the register file of one SM.  Registers: up to 10 points per thread run without spills in 80, 16 want 128.
    c.T = n / maxpp;
    int pts = maxpp;
    if (((c.T + 31) / 32 * 32) * (pts <= 10 ? 80 : 128) > 65536 && maxpp <= 8) { c.T = (c.T + 1) / 2; pts *= 2; }
    while (pts * 2 <= 10 && c.T > 128) { c.T = (c.T + 1) / 2; pts *= 2; } // radix 2..5 only: too little work per thread otherwise
    if (env_or("HSFFT_JIT_T", 0) > 0) { c.T = env_or("HSFFT_JIT_T", 0); pts = maxpp * ((n / maxpp + c.T - 1) / c.T); }
    // Fewer, fatter threads for the longer lengths (n >= 1800): pairs fused up to 25 points, and a team so small that a thread
    // carries 16-32 points per pass -- 8-10 warps with that much independent work keep an SM busier than 16-24 warps of 5-10
    // points, and the smaller lengths fit three CTAs per SM at 168 registers.  Among the team sizes T = ceil(NB_p / k) (NB_p:
    // butterflies of pass p) with n/26 <= T <= n/15 and at most 32 points per thread, take the one for which the passes with
    // exactly NB_p = k T (no guard on their loads and stores) carry the most points; none exact -> the rule above stays.
    // Measured on B200 (profiles/r2_plan_time_fat_threads.txt), % of the HBM peak, rule above -> this one: N = 2000 72.8 -> 79.0,
    // 2400 64.6 -> 86.3, 3125 54.0 -> 88.4, 4000 51.2 -> 63.3, 4500 52.2 -> 70.6, 5000 43.4 -> 69.6, 5120 51.7 -> 69.9,
    // 6000 44.5 -> 63.9, 6250 56.5 -> 70.8, 6400 47.5 -> 56.8; N = 1000 (82 %), 1536 (95.7 %, 85.7 with this rule) and 6561 (65 %) keep the rule above.
    bool fat = false;
    if (allow_fat && n >= 1800 && !getenv("HSFFT_JIT_T") && !getenv("HSFFT_JIT_FUSE") && env_or("HSFFT_JIT_FAT", 1) != 0) {
        std::vector<int> pr;
        int groups = 0;
        for (int s = 0; s < nst;) {
            int pp = rx[s], sz = 1;
            if (s + 1 < nst && rx[s] * rx[s + 1] <= 25) { pp = rx[s] * rx[s + 1]; sz = 2; }
            pr.push_back(pp);
            groups = groups * 10 + sz;
            s += sz;
        }
        int bestT = 0, bestW = 0, bestPts = 0;
        if (pr.size() >= 2)
            for (size_t p = 0; p < pr.size(); ++p)
                for (int k = 1; k <= 8; ++k) {
                    const int T = (n / pr[p] + k - 1) / k;
                    if (T * 26 < n || T * 15 > n || T > 1024) continue;
                    int mp = 0, w = 0;
                    for (size_t q = 0; q < pr.size(); ++q) {
                        const int nb = n / pr[q], b = (nb + T - 1) / T;
                        mp = pr[q] * b > mp ? pr[q] * b : mp;
                        if (nb % T == 0) w += pr[q];
                    }
                    if (mp > 32 || w == 0) continue;
                    // CTAs per SM at 168 registers (registers come in units of four warps) and n + n/8 points of shared memory:
                    // three of them (T <= 128) beat one (N = 2400, 3000, 3125: 86-88 % against 63-70 % for N = 4000 ... 6250)
                    const int by_regs = 65536 / ((T + 127) / 128 * 128 * 168), by_smem = 227 * 1024 / ((n + n / 8 + 1) * 16 + 1024);
                    const int ctas = by_regs < by_smem ? by_regs : by_smem;
                    if (ctas < 1 || ctas * ((T + 31) / 32) < 7) continue; // fewer than 7 warps per SM: too little to hide anything
                    w += ctas >= 3 ? 1000 : 0;
                    if (w > bestW || (w == bestW && T < bestT)) { bestW = w; bestT = T; bestPts = mp; }
                }
        if (bestW >= 1000 && bestT <= 128) {
            // three-CTA plans: a last warp that is mostly idle costs more than the exact passes win -- round the team up to whole
            // warps when more than a tenth of its lanes would idle (N = 2400: T = 100 -> 128 87.0 -> 92.6 % of the HBM peak, N = 2000:
            // 80 -> 96 78.0 -> 84.6 %; N = 3125 / 2500 at T = 125 lose 20 points at 128 and stay)
            const int Tw = (bestT + 31) / 32 * 32;
            int mp = 0;
            for (size_t q = 0; q < pr.size(); ++q) {
                const int nb = n / pr[q], b = (nb + Tw - 1) / Tw;
                mp = pr[q] * b > mp ? pr[q] * b : mp;
            }
            if ((Tw - bestT) * 10 > Tw && mp <= 32) { bestT = Tw; bestPts = mp; }
        }
        if (bestW > 0) {
            fat = true;
            c.pass_radix = pr;
            c.groups = groups;
            c.npass = (int)pr.size();
            c.T = bestT;
            pts = bestPts;
        }
    }
    if (c.T < 1 || c.T > 1024) return false;
    int best = 0;
    double best_ratio = 1e30;
    const int cand[5] = {3, 4, 2, 5, 0};
    for (int i = 0; i < 5; ++i) {
        const double r = conflict_ratio(n, c.T, cand[i], c.pass_radix);
        if (r < best_ratio - 1e-9) { best_ratio = r; best = cand[i]; }
    }
    c.logpad = env_or("HSFFT_JIT_LOGPAD", best);
    best = c.logpad;
    const int npad = best ? n + (n >> best) + 1 : n;
    // Small CTAs, many of them: independent CTAs are in different phases (load / exchange / arithmetic), which is what keeps
    // the SM's pipes busy at once (measured: N=1000 68 % of the HBM peak with 2 x 125 threads x 3 CTAs, 81 % with 125 x 6).
    int X = fat ? 1 : env_or("HSFFT_JIT_THREADS", 128) / c.T;
    if (env_or("HSFFT_JIT_X", 0) > 0) X = env_or("HSFFT_JIT_X", 0);
    if (X < 1) X = 1;
    while (X > 1 && (long long)X * npad * 16 > 100 * 1024) --X;
    if ((long long)X * npad * 16 > 200 * 1024) return false;
    c.X = X;
    const int threads = c.T * X, smem = X * npad * 16;
    // CTAs per SM: as many as the registers (see above), shared memory and 2048 threads allow, at most 8
    // (a CTA is allocated registers for a multiple of four warps: 200 threads count as 256; up to 20 points per thread run in
    // 128 registers -- N = 3600 at 240 threads of 18 points: two CTAs per SM 71.8 %, one 66.0 %)
    const int regs = fat ? (pts <= 20 ? 128 : 168) : pts <= 10 ? 80 : 128, warps32 = fat ? (threads + 127) / 128 * 128 : (threads + 31) / 32 * 32;
    c.fat = fat;
    if ((long long)warps32 * regs > 65536) return false;
    int minb = (int)(65536 / ((long long)warps32 * regs));
    if (minb > 8) minb = 8;
    while (minb > 1 && ((long long)minb * (smem + 1024) > 227 * 1024 || minb * threads > 2048)) --minb;
    c.minb = env_or("HSFFT_JIT_MINB", minb);
    return true;
}

// ------------------------------------------------------------------------------------------------ the plans
struct JitPlan {
    int n, sgn, dev; // a module belongs to the context of the device that was current when it was loaded
    hs_ct_desc desc;
    CUmodule mod;
    CUfunction fn, fn_r2c, fn_c2r, fn_staged;
    int threads, smem, X, minb, smem_staged;
    int threads_c2r, smem_c2r, X_c2r; // hs_jit_c2r may run on its own plan
    std::string text;
};
static std::mutex g_mu;
static std::vector<JitPlan *> g_plans;

static std::string make_source(int n, int sgn, const int *rx, int nst, const JitChoice &c, const JitChoice &cr)
{
    char buf[8192];
    std::string rad;
    for (int s = 0; s < nst; ++s) { snprintf(buf, sizeof buf, ", %d", rx[s]); rad += buf; }
    // three entry points over one plan: the complex transform and the two real transforms that ride on it (the bodies of
    // hs_c2c_smem_kernel / hs_r2c_smem_kernel / hs_c2r_smem_kernel, spelled out: wrapped in a device function the same
    // code ran at 43 % instead of 78 % of the HBM peak for N = 3000 (measured), so each kernel keeps its loads and stores
    // in its own scope)
    snprintf(buf, sizeof buf,
             "#include \"hs_smem_fft.cuh\"\n"
             "using PJ = HsPlan<%d, %d, %d, %d, -2, %d%s>;\n"
             "using PR = HsPlan<%d, %d, %d, %d, -2, %d%s>; // the plan of hs_jit_c2r (see hs_cu_jit_c2c_create)\n"
             "extern \"C\" __global__ void __launch_bounds__(PJ::threads, %d)\n"
             "hs_jit_c2c(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, long long howmany, int pf_dist)\n"
             "{\n"
             "    HS_DYN_SMEM(cplx, smem);\n"
             "    const int team = HS_TID / PJ::T, t = HS_TID %% PJ::T;\n"
             "    const long long xf = HS_BID * PJ::X + team;\n"
             "    const bool active = xf < howmany;\n"
             "    HsLoadGlobal ld{in + (active ? xf : 0) * PJ::N};\n"
             "    HsStoreGlobal st{out + (active ? xf : 0) * PJ::N};\n"
             "    hs_fft_team<PJ, %d>(ld, st, smem + team * PJ::NPAD, tw, t, active);\n"
             "}\n"
             "extern \"C\" __global__ void __launch_bounds__(PJ::threads, %d)\n"
             "hs_jit_r2c(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, const cplx *__restrict__ tw2,\n"
             "           long long out_stride, int full, long long howmany)\n"
             "{\n"
             "    HS_DYN_SMEM(cplx, smem);\n"
             "    const int team = HS_TID / PJ::T, t = HS_TID %% PJ::T;\n"
             "    const long long xf = HS_BID * PJ::X + team;\n"
             "    const bool active = xf < howmany;\n"
             "    cplx *sm = smem + team * PJ::NPAD;\n"
             "    HsLoadGlobal ld{in + (active ? xf : 0) * PJ::N};\n"
             "    HsStoreSmemNatural<PJ> st{sm};\n"
             "    hs_fft_team<PJ, %d>(ld, st, sm, tw, t, active);\n"
             "    HsR2CSplitTw<PJ> w2;\n"
             "    w2.load(tw2, t); // in flight across the barrier\n"
             "    HS_SYNC();\n"
             "    if (active) hs_r2c_split_rows_w<PJ>(sm, out + xf * out_stride, w2, full, t);\n"
             "}\n"
             "extern \"C\" __global__ void __launch_bounds__(PR::threads, %d)\n"
             "hs_jit_c2r(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, const cplx *__restrict__ tw2,\n"
             "           long long in_stride, long long howmany)\n"
             "{\n"
             "    HS_DYN_SMEM(cplx, smem);\n"
             "    const int team = HS_TID / PR::T, t = HS_TID %% PR::T;\n"
             "    const long long xf = HS_BID * PR::X + team;\n"
             "    const bool active = xf < howmany;\n"
             "    HsStoreGlobal st{out + (active ? xf : 0) * PR::N};\n"
             "    if (%d) { // merge from a copy of the row staged in the team's buffer (see hs_c2r_smem_kernel)\n"
             "        cplx *sm = smem + team * PR::NPAD;\n"
             "        const cplx *x = in + (active ? xf : 0) * in_stride;\n"
             "        if (active)\n"
             "            for (int k = t; k < PR::N; k += PR::T) hs_cp_async16(sm + k, x + k);\n"
             "        hs_cp_async_wait_all();\n"
             "        HS_SYNC();\n"
             "        HsLoadC2RMergeSmem ld{sm, x + PR::N, tw2, PR::N};\n"
             "        hs_fft_team<PR, %d>(ld, st, sm, tw, t, active);\n"
             "    } else {\n"
             "        HsLoadC2RMerge ld{in + (active ? xf : 0) * in_stride, tw2, PR::N};\n"
             "        hs_fft_team<PR, %d>(ld, st, smem + team * PR::NPAD, tw, t, active);\n"
             "    }\n"
             "}\n",
             n, c.T, c.X, c.logpad, c.groups, rad.c_str(), n, cr.T, cr.X, cr.logpad, cr.groups, rad.c_str(), c.minb, sgn, c.minb, sgn, cr.minb,
             env_or("HSFFT_C2R_STAGE", 1) == 2 ? 1 : 0, sgn, sgn); // plan-time kernels: the staged merge measured 2-6 % slower (N = 2000, 2002, 4000, 4374)
    std::string src(buf);
    if (c.X == 1 && c.npass > 1) { // persistent twin with the next input staged by the copy engine (see HsStageRefill)
        snprintf(buf, sizeof buf,
                 "extern \"C\" __global__ void __launch_bounds__(PJ::threads, %d)\n"
                 "hs_jit_c2c_staged(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, long long howmany)\n"
                 "{\n"
                 "    HS_DYN_SMEM(cplx, smem);\n"
                 "    cplx *stage = smem + HsStagedSmem<PJ>::stage_off;\n"
                 "    HsMbar *bar = reinterpret_cast<HsMbar *>(stage + PJ::N);\n"
                 "    if (HS_TID == 0) hs_mbar_init(bar);\n"
                 "    hs_fence_async_proxy();\n"
                 "    HS_SYNC();\n"
                 "    if (HS_TID == 0 && (long long)HS_BID < howmany) {\n"
                 "        hs_mbar_expect(bar, (unsigned)(PJ::N * 16));\n"
                 "        hs_bulk_g2s(stage, in + (long long)HS_BID * PJ::N, (unsigned)(PJ::N * 16), bar);\n"
                 "    }\n"
                 "    for (int j = 0; HS_BID + (long long)j * HS_NBLOCKS < howmany; ++j) {\n"
                 "        hs_mbar_wait(bar, (unsigned)(j & 1));\n"
                 "        HsStageRefill hk{bar, stage, in, howmany, j, PJ::N};\n"
                 "        HsLoadStage ld{stage};\n"
                 "        HsStoreGlobal st{out + (HS_BID + (long long)j * HS_NBLOCKS) * PJ::N};\n"
                 "        hs_fft_team_h<PJ, %d>(ld, st, smem, tw, HS_TID, true, hk);\n"
                 "        HS_SYNC();\n"
                 "    }\n"
                 "}\n",
                 c.minb, sgn);
        src += buf;
    }
    return src;
}

static bool jit_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("HSFFT_JIT"); v = e ? atoi(e) : 1; }
    return v != 0;
}

// the generated translation unit for (n, sgn, radices in execution order), without compiling it (tests; returns its length or 0)
extern "C" int hs_cu_jit_source(int n, int sgn, const int *rx, int nst, char *out, int cap)
{
    JitChoice c, cr;
    if (!choose(n, rx, nst, c)) return 0;
    cr = c;
    if (c.fat && c.minb == 1 && !choose(n, rx, nst, cr, false)) cr = c;
    const std::string s = make_source(n, sgn, rx, nst, c, cr);
    if ((int)s.size() + 1 > cap) return 0;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// compile (or find) the kernel for (n, sgn, radices in execution order); returns a handle >= 0 and fills desc, or -1
extern "C" int hs_cu_jit_c2c_create(int n, int sgn, const int *rx, int nst, hs_ct_desc *desc)
{
    if (!jit_enabled()) return -1;
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    for (size_t i = 0; i < g_plans.size(); ++i)
        if (g_plans[i]->n == n && g_plans[i]->sgn == sgn && g_plans[i]->dev == dev) { *desc = g_plans[i]->desc; return (int)i; }
    // The c2r kernel has no persistent twin to make up for a single small CTA per SM, so where the fat-thread rule ends at one
    // CTA per SM it keeps the thin-thread plan (c2r N = 10000: 35.6 % of the HBM peak against 21.3 %; r2c gains either way)
    JitChoice c, cr;
    if (!choose(n, rx, nst, c)) return -1;
    cr = c;
    if (c.fat && c.minb == 1 && !choose(n, rx, nst, cr, false)) cr = c;
    Nvrtc *rt = nvrtc();
    Driver *drv = driver();
    if (!rt || !drv) return -1;
    const std::string src = make_source(n, sgn, rx, nst, c, cr);
    nvrtcProgram prog = nullptr;
    if (rt->CreateProgram(&prog, src.c_str(), "hs_jit_plan.cu", hs_jit_src_count, hs_jit_src_texts, hs_jit_src_names) != 0) return -1;
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-DHS_JIT=1"};
    const nvrtcResult rc = rt->CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t ls = 0;
        rt->GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) rt->GetProgramLog(prog, &log[0]);
        fprintf(stderr, "hsfft: plan-time compilation for N=%d failed (falling back to the run-time engine):\n%s\n", n, log.c_str());
        rt->DestroyProgram(&prog);
        return -1;
    }
    size_t cs = 0;
    rt->GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    rt->GetCUBIN(prog, cubin.data());
    rt->DestroyProgram(&prog);

    JitPlan *p = new JitPlan();
    p->n = n; p->sgn = sgn; p->dev = dev; p->text = src;
    HS_CUDA_CHECK(cudaFree(0)); // make sure the primary context is current
    if (drv->ModuleLoadData(&p->mod, cubin.data()) != CUDA_SUCCESS || drv->ModuleGetFunction(&p->fn, p->mod, "hs_jit_c2c") != CUDA_SUCCESS) {
        fprintf(stderr, "hsfft: loading the plan-time kernel for N=%d failed (falling back to the run-time engine)\n", n);
        delete p;
        return -1;
    }
    const int npad = c.logpad ? n + (n >> c.logpad) + 1 : n;
    p->threads = c.T * c.X; p->X = c.X; p->smem = c.X * npad * 16; p->minb = c.minb;
    const int npad_r = cr.logpad ? n + (n >> cr.logpad) + 1 : n;
    p->threads_c2r = cr.T * cr.X; p->X_c2r = cr.X; p->smem_c2r = cr.X * npad_r * 16;
    drv->FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->smem);
    int pct = (int)(((long long)(p->smem + 1024) * c.minb * 100 + 228 * 1024 - 1) / (228 * 1024));
    drv->FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct > 100 ? 100 : pct);
    p->fn_r2c = p->fn_c2r = nullptr;
    if (drv->ModuleGetFunction(&p->fn_r2c, p->mod, "hs_jit_r2c") == CUDA_SUCCESS && drv->ModuleGetFunction(&p->fn_c2r, p->mod, "hs_jit_c2r") == CUDA_SUCCESS) {
        drv->FuncSetAttribute(p->fn_r2c, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->smem);
        drv->FuncSetAttribute(p->fn_r2c, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct > 100 ? 100 : pct);
        const int pctr = (int)(((long long)(p->smem_c2r + 1024) * cr.minb * 100 + 228 * 1024 - 1) / (228 * 1024));
        drv->FuncSetAttribute(p->fn_c2r, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->smem_c2r);
        drv->FuncSetAttribute(p->fn_c2r, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pctr > 100 ? 100 : pctr);
    } else p->fn_r2c = p->fn_c2r = nullptr;
    // HSFFT_JIT_STAGED: 0 never, 1 plans that run one CTA per SM, 2 (default) also two CTAs per SM where both buffers fit -- each
    // measured against the one-shot kernel when the plan is made; 3 always (no measurement)
    p->fn_staged = nullptr;
    p->smem_staged = (((npad + 7) / 8 * 8) + n) * 16 + 16;
    const int want_staged = env_or("HSFFT_JIT_STAGED", 2);
    if (c.X == 1 && c.npass > 1 && want_staged > 0 && c.minb <= (want_staged >= 2 ? 2 : 1) && (long long)c.minb * (p->smem_staged + 1024) <= 227 * 1024 &&
        drv->ModuleGetFunction(&p->fn_staged, p->mod, "hs_jit_c2c_staged") == CUDA_SUCCESS) {
        drv->FuncSetAttribute(p->fn_staged, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->smem_staged);
        int pcts = (int)(((long long)(p->smem_staged + 1024) * c.minb * 100 + 228 * 1024 - 1) / (228 * 1024));
        drv->FuncSetAttribute(p->fn_staged, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pcts > 100 ? 100 : pcts);
    } else p->fn_staged = nullptr;
    if (p->fn_staged && want_staged < 3) {
        // Which of the two wins depends on what the persistent loop does to the register allocation (N = 4500: 29.7 -> 52.2 % of
        // the HBM peak, N = 6000: 46.3 -> 28.6 %; ptxas' local-memory size does not separate the cases), so the plan measures
        // both once, on zero-filled scratch rows, and keeps the staged kernel only where it is faster.
        const long long resident = (long long)hs_cu_sm_count() * c.minb, rows = 8 * resident;
        const size_t bytes = (size_t)rows * n * 16;
        void *d_a = nullptr, *d_b = nullptr, *d_t = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        float ms_plain = 0.f, ms_staged = 0.f;
        bool ok = cudaMalloc(&d_a, bytes) == cudaSuccess && cudaMalloc(&d_b, bytes) == cudaSuccess && cudaMalloc(&d_t, (size_t)2 * n * 16) == cudaSuccess &&
                  cudaMemset(d_a, 0, bytes) == cudaSuccess && cudaMemset(d_t, 0, (size_t)2 * n * 16) == cudaSuccess &&
                  cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
        if (ok) {
            long long hm = rows;
            int pf = 0;
            void *a_plain[] = {(void *)&d_a, (void *)&d_b, (void *)&d_t, (void *)&hm, (void *)&pf};
            void *a_staged[] = {(void *)&d_a, (void *)&d_b, (void *)&d_t, (void *)&hm};
            for (int which = 0; which < 2 && ok; ++which) {
                float best = 1e30f;
                for (int rep = 0; rep < 4 && ok; ++rep) { // the first repetition warms up
                    cudaEventRecord(e0, 0);
                    const CUresult r = which == 0
                        ? drv->LaunchKernel(p->fn, (unsigned)rows, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem, 0, a_plain, nullptr)
                        : drv->LaunchKernel(p->fn_staged, (unsigned)resident, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem_staged, 0, a_staged, nullptr);
                    cudaEventRecord(e1, 0);
                    float ms = 0.f;
                    ok = r == CUDA_SUCCESS && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess;
                    if (rep > 0 && ms < best) best = ms;
                }
                (which == 0 ? ms_plain : ms_staged) = best;
            }
        }
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        cudaFree(d_a); cudaFree(d_b); cudaFree(d_t);
        if (!ok) cudaGetLastError();
        if (!ok || !(ms_staged < 0.97f * ms_plain)) p->fn_staged = nullptr;
        if (getenv("HSFFT_VERBOSE"))
            fprintf(stderr, "hsfft: plan-time kernel N=%d: one-shot %.3f ms, staged %.3f ms per %lld transforms -> %s\n", n, ms_plain, ms_staged, rows,
                    p->fn_staged ? "staged" : "one-shot");
    }
    memset(&p->desc, 0, sizeof p->desc);
    p->desc.id = (int)g_plans.size();
    p->desc.family = HS_CT_C2C;
    p->desc.n = n;
    p->desc.nst = nst;
    for (int s = 0; s < 8 && s < nst; ++s) p->desc.radix[s] = rx[s];
    p->desc.split_min = -2;
    p->desc.threads = p->threads;
    p->desc.smem_bytes = p->smem;
    if (getenv("HSFFT_VERBOSE"))
        fprintf(stderr, "hsfft: plan-time kernel N=%d: T=%d X=%d LOGPAD=%d GROUPS=%d, %d threads, %d bytes smem, %d CTAs/SM asked\n", n, c.T, c.X, c.logpad,
                c.groups, p->threads, p->smem, c.minb);
    g_plans.push_back(p);
    *desc = p->desc;
    return p->desc.id;
}

extern "C" void hs_cu_jit_c2c_launch(int handle, const void *d_in, void *d_out, const void *d_tw, long long howmany, void *stream)
{
    JitPlan *p;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (handle < 0 || handle >= (int)g_plans.size()) { fprintf(stderr, "Error: unknown plan-time kernel %d\n", handle); exit(EXIT_FAILURE); }
        p = g_plans[handle];
    }
    const long long blocks = (howmany + p->X - 1) / p->X;
    if (blocks <= 0 || blocks > 2147483647LL) { fprintf(stderr, "Error: batch too large for one launch (%lld CTAs)\n", blocks); exit(EXIT_FAILURE); }
    const long long resident = (long long)hs_cu_sm_count() * p->minb;
    if (p->fn_staged && howmany >= 4 * resident) { // persistent CTAs, next input staged by the copy engine
        void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw, (void *)&howmany};
        const CUresult r = driver()->LaunchKernel(p->fn_staged, (unsigned)resident, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem_staged, (CUstream)stream, args, nullptr);
        if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: CUDA failure %d launching the staged plan-time kernel for N=%d\n", (int)r, p->n); exit(EXIT_FAILURE); }
        hs_cu_count_launch();
        hs_cu_timing_mark("c2c (plan-time kernel, staged)", stream);
        return;
    }
    int pf = 0;
    void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw, (void *)&howmany, (void *)&pf};
    const CUresult r = driver()->LaunchKernel(p->fn, (unsigned)blocks, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: CUDA failure %d launching the plan-time kernel for N=%d\n", (int)r, p->n); exit(EXIT_FAILURE); }
    hs_cu_count_launch();
    hs_cu_timing_mark("c2c (plan-time kernel)", stream);
}

extern "C" void hs_cu_jit_real_launch(int handle, int mode, const void *d_in, void *d_out, const void *d_tw, const void *d_tw2, long long stride,
                                      int full, long long howmany, void *stream)
{
    JitPlan *p;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (handle < 0 || handle >= (int)g_plans.size()) { fprintf(stderr, "Error: unknown plan-time kernel %d\n", handle); exit(EXIT_FAILURE); }
        p = g_plans[handle];
    }
    if (!p->fn_r2c) { fprintf(stderr, "Error: the plan-time kernel for N=%d has no real-transform entry points\n", p->n); exit(EXIT_FAILURE); }
    const int X = mode == 1 ? p->X : p->X_c2r;
    const long long blocks = (howmany + X - 1) / X;
    if (blocks <= 0 || blocks > 2147483647LL) { fprintf(stderr, "Error: batch too large for one launch (%lld CTAs)\n", blocks); exit(EXIT_FAILURE); }
    CUresult r;
    if (mode == 1) {
        void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw, (void *)&d_tw2, (void *)&stride, (void *)&full, (void *)&howmany};
        r = driver()->LaunchKernel(p->fn_r2c, (unsigned)blocks, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem, (CUstream)stream, args, nullptr);
    } else {
        void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw, (void *)&d_tw2, (void *)&stride, (void *)&howmany};
        r = driver()->LaunchKernel(p->fn_c2r, (unsigned)blocks, 1, 1, (unsigned)p->threads_c2r, 1, 1, (unsigned)p->smem_c2r, (CUstream)stream, args, nullptr);
    }
    if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: CUDA failure %d launching the plan-time real kernel for N=%d\n", (int)r, 2 * p->n); exit(EXIT_FAILURE); }
    hs_cu_count_launch();
    hs_cu_timing_mark(mode == 1 ? "r2c (plan-time kernel)" : "c2r (plan-time kernel)", stream);
}

extern "C" int hs_cu_jit_is_staged(int handle)
{
    std::lock_guard<std::mutex> lock(g_mu);
    return handle >= 0 && handle < (int)g_plans.size() && g_plans[handle]->fn_staged != nullptr;
}

extern "C" int hs_cu_jit_has_real(int handle)
{
    std::lock_guard<std::mutex> lock(g_mu);
    return handle >= 0 && handle < (int)g_plans.size() && g_plans[handle]->fn_r2c != nullptr;
}

// ------------------------------------------------------------------------------------------------ two tiled passes
// Lengths above one SM's shared memory whose stages are all compile-time radices (N = 10000, 3*2^14, 6561 ...): the factor
// list, in execution order, is cut in two; pass 1 runs the first stages as strided sub-transforms of length n1, pass 2 the
// rest (length n2) with the inter-pass twiddle on its loads -- the kernels of hs_tiled.cuh instantiated for these two
// sub-plans.  Same stages, butterflies and literals in the same order as the reference's recursion, so parity is unchanged;
// the run-time engine (hs_generic.cuh) served these lengths at about a tenth of the HBM roofline.
struct SubChoice { int T, C, logpad, minb, threads, smem; };
static bool choose_sub(int L, const int *rx, int nst, SubChoice &c, int want_c = 0)
{
    if (nst < 2 || nst > 8 || L < 16 || L > 2048) return false;
    int maxr = 1;
    for (int s = 0; s < nst; ++s) {
        const int r = rx[s];
        if (!(r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16 || r == 11 || r == 13 || r == 17)) return false;
        maxr = r > maxr ? r : maxr;
    }
    c.T = L / maxr;                       // one butterfly per thread in the widest stage
    int pts = maxr;
    while (c.T > 64 && pts * 2 <= 16) { c.T = (c.T + 1) / 2; pts *= 2; }
    if (c.T < 1 || c.T > 256) return false;
    std::vector<int> pr(rx, rx + nst);
    int best = 0;
    double best_ratio = 1e30;
    const int cand[5] = {3, 4, 2, 5, 0};
    for (int i = 0; i < 5; ++i) {
        const double r = conflict_ratio(L, c.T, cand[i], pr);
        if (r < best_ratio - 1e-9) { best_ratio = r; best = cand[i]; }
    }
    c.logpad = best;
    c.C = c.T <= 16 ? 8 : (c.T <= 64 ? 4 : 2);
    if (want_c > 0) c.C = want_c;
    c.threads = c.T * c.C;
    if (c.threads > 512) return false;
    const int npad = best ? L + (L >> best) + 1 : L;
    // same rule as hs_col_stride (hs_tiled.cuh)
    int stride = npad;
    const int want = c.C >= 8 ? 1 : (c.C == 4 ? 2 : 4);
    while (c.C >= 8 ? (stride % 2 == 0) : (stride % (2 * want) != want)) ++stride;
    c.smem = c.C * stride * 16;
    if (c.smem > 200 * 1024) return false;
    const int regs = pts <= 10 ? 80 : 128, w32 = (c.threads + 31) / 32 * 32;
    int minb = (int)(65536 / ((long long)w32 * regs));
    if (minb > 6) minb = 6;
    while (minb > 1 && (long long)minb * (c.smem + 1024) > 227 * 1024) --minb;
    c.minb = minb < 1 ? 1 : minb;
    return true;
}

struct JitTwoPass {
    long long n;
    int sgn, dev, n1, n2;
    hs_ct_desc desc;
    CUmodule mod;
    CUfunction fn[2];
    SubChoice sc[2];
};
static std::vector<JitTwoPass *> g_tp;

static std::string radix_list(const int *rx, int nst)
{
    std::string r;
    char b[16];
    for (int s = 0; s < nst; ++s) { snprintf(b, sizeof b, ", %d", rx[s]); r += b; }
    return r;
}

// rx[0..nst): stage radices in execution order; the first `split` stages form pass 1.  Fills desc like a static two-pass
// descriptor (n1, n2, radix / radix2) and returns a handle >= 0, or -1.
extern "C" int hs_cu_jit_twopass_create(long long n, int sgn, const int *rx, int nst, int split, hs_ct_desc *desc)
{
    if (!jit_enabled() || split < 1 || split >= nst || nst - split > 8 || split > 8) return -1;
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    long long n1 = 1, n2 = 1;
    for (int s = 0; s < split; ++s) n1 *= rx[s];
    for (int s = split; s < nst; ++s) n2 *= rx[s];
    if (n1 * n2 != n) return -1;
    for (size_t i = 0; i < g_tp.size(); ++i)
        if (g_tp[i]->n == n && g_tp[i]->sgn == sgn && g_tp[i]->dev == dev && g_tp[i]->n1 == n1) { *desc = g_tp[i]->desc; return (int)i; }
    SubChoice c1, c2;
    // HSFFT_JIT_C1 / _C2: columns per tile of the first / second pass (tuning aids)
    if (!choose_sub((int)n1, rx, split, c1, env_or("HSFFT_JIT_C1", 0)) || !choose_sub((int)n2, rx + split, nst - split, c2, env_or("HSFFT_JIT_C2", 0))) return -1;
    if (getenv("HSFFT_VERBOSE"))
        fprintf(stderr, "hsfft: plan-time two-pass N=%lld = %lld x %lld: pass 1 T=%d C=%d (%d threads, %d B, %d CTAs/SM), pass 2 T=%d C=%d (%d threads, %d B, %d CTAs/SM)\n",
                n, n1, n2, c1.T, c1.C, c1.threads, c1.smem, c1.minb, c2.T, c2.C, c2.threads, c2.smem, c2.minb);
    Nvrtc *rt = nvrtc();
    Driver *drv = driver();
    if (!rt || !drv) return -1;
    char buf[4096];
    snprintf(buf, sizeof buf,
             "#define HS_TILE_PARTIAL(IO) %d\n" // a run-time column mask only where a pass has a ragged last tile (it costs 8-14 %)
             "#include \"hs_tiled.cuh\"\n"
             "using P1 = HsPlan<%d, %d, 1, %d, -2, 0%s>;\n"
             "using P2 = HsPlan<%d, %d, 1, %d, -2, 0%s>;\n"
             "extern \"C\" __global__ void __launch_bounds__(%d, %d)\n"
             "hs_jit_pass1(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, const HsTileGeom g, long long howmany)\n"
             "{\n"
             "    HS_DYN_SMEM(cplx, smem);\n"
             "    const int tiles = (g.ncols + %d - 1) / %d;\n"
             "    const long long xf = HS_BID / tiles;\n"
             "    const int c0 = (int)(HS_BID %% tiles) * %d;\n"
             "    HsNoHooks hk;\n"
             "    hs_tile_run<P1, %d, %d, 0, false, 0, 1, 0, 0>(in + xf * g.in_xf, out + xf * g.out_xf, tw, g, xf, c0, smem, HS_TID, true, hk);\n"
             "}\n"
             "extern \"C\" __global__ void __launch_bounds__(%d, %d)\n"
             "hs_jit_pass2(const cplx *__restrict__ in, cplx *__restrict__ out, const cplx *__restrict__ tw, const HsTileGeom g, long long howmany)\n"
             "{\n"
             "    HS_DYN_SMEM(cplx, smem);\n"
             "    const int tiles = (g.ncols + %d - 1) / %d;\n"
             "    const long long xf = HS_BID / tiles;\n"
             "    const int c0 = (int)(HS_BID %% tiles) * %d;\n"
             "    HsNoHooks hk;\n"
             "    hs_tile_run<P2, %d, %d, 1, true, 0, 1, 0, 0>(in + xf * g.in_xf, out + xf * g.out_xf, tw, g, xf, c0, smem, HS_TID, true, hk);\n"
             "}\n",
             (n2 % c1.C != 0 || n1 % c2.C != 0) ? 1 : 0,
             (int)n1, c1.T, c1.logpad, radix_list(rx, split).c_str(), (int)n2, c2.T, c2.logpad, radix_list(rx + split, nst - split).c_str(),
             c1.threads, c1.minb, c1.C, c1.C, c1.C, sgn, c1.C, c2.threads, c2.minb, c2.C, c2.C, c2.C, sgn, c2.C);
    nvrtcProgram prog = nullptr;
    if (rt->CreateProgram(&prog, buf, "hs_jit_twopass.cu", hs_jit_src_count, hs_jit_src_texts, hs_jit_src_names) != 0) return -1;
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-DHS_JIT=1"};
    if (rt->CompileProgram(prog, 4, opts) != 0) {
        size_t ls = 0;
        rt->GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) rt->GetProgramLog(prog, &log[0]);
        fprintf(stderr, "hsfft: plan-time compilation of the two passes for N=%lld failed (falling back to the run-time engine):\n%s\n", n, log.c_str());
        rt->DestroyProgram(&prog);
        return -1;
    }
    size_t cs = 0;
    rt->GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    rt->GetCUBIN(prog, cubin.data());
    rt->DestroyProgram(&prog);
    JitTwoPass *p = new JitTwoPass();
    p->n = n; p->sgn = sgn; p->dev = dev; p->n1 = (int)n1; p->n2 = (int)n2; p->sc[0] = c1; p->sc[1] = c2;
    HS_CUDA_CHECK(cudaFree(0));
    if (drv->ModuleLoadData(&p->mod, cubin.data()) != CUDA_SUCCESS || drv->ModuleGetFunction(&p->fn[0], p->mod, "hs_jit_pass1") != CUDA_SUCCESS ||
        drv->ModuleGetFunction(&p->fn[1], p->mod, "hs_jit_pass2") != CUDA_SUCCESS) {
        fprintf(stderr, "hsfft: loading the plan-time passes for N=%lld failed (falling back to the run-time engine)\n", n);
        delete p;
        return -1;
    }
    for (int k = 0; k < 2; ++k) drv->FuncSetAttribute(p->fn[k], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, p->sc[k].smem);
    memset(&p->desc, 0, sizeof p->desc);
    p->desc.id = (int)g_tp.size();
    p->desc.family = HS_CT_TWOPASS;
    p->desc.n = (int)n;
    p->desc.n1 = (int)n1; p->desc.n2 = (int)n2;
    p->desc.nst = split; p->desc.nst2 = nst - split;
    for (int s = 0; s < split; ++s) p->desc.radix[s] = rx[s];
    for (int s = split; s < nst; ++s) p->desc.radix2[s - split] = rx[s];
    p->desc.split_min = p->desc.split_min2 = -2;
    p->desc.threads = c1.threads; p->desc.smem_bytes = c1.smem;
    if (getenv("HSFFT_VERBOSE"))
        fprintf(stderr, "hsfft: plan-time passes N=%lld = %lld x %lld: pass 1 T=%d C=%d LOGPAD=%d %d CTAs/SM, pass 2 T=%d C=%d LOGPAD=%d %d CTAs/SM\n", n, n1, n2,
                c1.T, c1.C, c1.logpad, c1.minb, c2.T, c2.C, c2.logpad, c2.minb);
    g_tp.push_back(p);
    *desc = p->desc;
    return p->desc.id;
}

// pass 1 or 2 with the geometry of a contiguous batch (as launch_twopass in hs_ct_table.cuh)
extern "C" void hs_cu_jit_twopass_launch(int handle, int pass, const void *d_in, void *d_out, const void *d_tw, const void *d_twb, const void *d_tws,
                                         long long howmany, void *stream)
{
    JitTwoPass *p;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (handle < 0 || handle >= (int)g_tp.size()) { fprintf(stderr, "Error: unknown plan-time two-pass plan %d\n", handle); exit(EXIT_FAILURE); }
        p = g_tp[handle];
    }
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.in_xf = g.out_xf = p->n;
    const SubChoice &c = p->sc[pass - 1];
    if (pass == 1) { g.ncols = p->n2; g.in_c = 1; g.in_e = p->n2; g.out_c = p->n1; g.out_e = 1; }
    else { g.ncols = p->n1; g.in_c = 1; g.in_e = p->n1; g.out_c = 1; g.out_e = p->n1; g.twb = d_twb; g.tws = d_tws; g.tw_ld = p->n1; }
    const long long blocks = howmany * ((g.ncols + c.C - 1) / c.C);
    if (blocks <= 0 || blocks > 2147483647LL) { fprintf(stderr, "Error: batch too large for one launch (%lld CTAs)\n", blocks); exit(EXIT_FAILURE); }
    void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw, (void *)&g, (void *)&howmany};
    const CUresult r = driver()->LaunchKernel(p->fn[pass - 1], (unsigned)blocks, 1, 1, (unsigned)c.threads, 1, 1, (unsigned)c.smem, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: CUDA failure %d launching plan-time pass %d for N=%lld\n", (int)r, pass, p->n); exit(EXIT_FAILURE); }
    hs_cu_count_launch();
    hs_cu_timing_mark(pass == 1 ? "tiled pass1 (plan-time kernel)" : "tiled pass2 (plan-time kernel)", stream);
}

// ------------------------------------------------------------------------------------------------ Rader kernels
// A prime N whose N-1 is a product of compile-time radices and fits one SM (N <= 4097): the cyclic convolution of length
// L = N-1 of hs_fused.cuh's hs_rader_kernel, instantiated for this L (forward plan = the radix list, inverse plan = the
// list reversed, full-precision constants).  Half the arithmetic of the padded chirp-z the reference's route implies.
struct JitRader {
    int L, dev;
    hs_ct_desc desc;
    CUmodule mod;
    CUfunction fn;
    int threads, smem, minb;
};
static std::vector<JitRader *> g_rader;

extern "C" int hs_cu_jit_rader_create(int L, const int *rx, int nst, hs_ct_desc *desc)
{
    if (!jit_enabled() || nst < 2 || nst > 8 || L < 16 || L > 4096) return -1;
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    for (size_t i = 0; i < g_rader.size(); ++i)
        if (g_rader[i]->L == L && g_rader[i]->dev == dev) { *desc = g_rader[i]->desc; return (int)i; }
    int maxr = 1, inv[8];
    for (int s = 0; s < nst; ++s) {
        const int r = rx[s];
        if (!(r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 16 || r == 11 || r == 13 || r == 17)) return -1;
        maxr = r > maxr ? r : maxr;
        inv[s] = rx[nst - 1 - s];
    }
    int T = L / maxr, pts = maxr;
    while (T > 128 && pts * 2 <= 16) { T = (T + 1) / 2; pts *= 2; }
    if (T < 1 || T > 512) return -1;
    std::vector<int> pf(rx, rx + nst), pi(inv, inv + nst);
    int lpf = 0, lpi = 0;
    double bf = 1e30, bi = 1e30;
    const int cand[5] = {3, 4, 2, 5, 0};
    for (int i = 0; i < 5; ++i) {
        const double a = conflict_ratio(L, T, cand[i], pf), b = conflict_ratio(L, T, cand[i], pi);
        if (a < bf - 1e-9) { bf = a; lpf = cand[i]; }
        if (b < bi - 1e-9) { bi = b; lpi = cand[i]; }
    }
    const int npf = lpf ? L + (L >> lpf) + 1 : L, npi = lpi ? L + (L >> lpi) + 1 : L;
    int npad = npf > npi ? npf : npi;
    if (npad < L + 2) npad = L + 2;        // HsRaderSmem: room for the N = L+1 staged outputs
    const int smem = npad * 16;
    const int regs = pts <= 10 ? 80 : 128, w32 = (T + 31) / 32 * 32;
    int minb = (int)(65536 / ((long long)w32 * regs));
    if (minb > 8) minb = 8;
    while (minb > 1 && ((long long)minb * (smem + 1024) > 227 * 1024 || minb * T > 2048)) --minb;
    Nvrtc *rt = nvrtc();
    Driver *drv = driver();
    if (!rt || !drv) return -1;
    // the kernel template of hs_fused.cuh itself, instantiated by name (no wrapper around its body)
    char buf[1024], expr[128];
    snprintf(buf, sizeof buf,
             "#include \"hs_fused.cuh\"\n"
             "using PF = HsExact<HsPlan<%d, %d, 1, %d, -2, 0%s>>;\n"
             "using PI = HsExact<HsPlan<%d, %d, 1, %d, -2, 0%s>>;\n",
             L, T, lpf, radix_list(rx, nst).c_str(), L, T, lpi, radix_list(inv, nst).c_str());
    snprintf(expr, sizeof expr, "hs_rader_kernel<PF, PI, %d, 2>", minb);
    nvrtcProgram prog = nullptr;
    if (rt->CreateProgram(&prog, buf, "hs_jit_rader.cu", hs_jit_src_count, hs_jit_src_texts, hs_jit_src_names) != 0) return -1;
    if (rt->AddNameExpression(prog, expr) != 0) { rt->DestroyProgram(&prog); return -1; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-DHS_JIT=1"};
    if (rt->CompileProgram(prog, 4, opts) != 0) {
        size_t ls = 0;
        rt->GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) rt->GetProgramLog(prog, &log[0]);
        fprintf(stderr, "hsfft: plan-time compilation of the Rader kernel for N=%d failed (falling back to chirp-z):\n%s\n", L + 1, log.c_str());
        rt->DestroyProgram(&prog);
        return -1;
    }
    const char *lowered = nullptr;
    if (rt->GetLoweredName(prog, expr, &lowered) != 0 || !lowered) { rt->DestroyProgram(&prog); return -1; }
    const std::string kname(lowered);
    size_t cs = 0;
    rt->GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    rt->GetCUBIN(prog, cubin.data());
    rt->DestroyProgram(&prog);
    JitRader *p = new JitRader();
    p->L = L; p->dev = dev; p->threads = T; p->smem = smem; p->minb = minb;
    HS_CUDA_CHECK(cudaFree(0));
    if (drv->ModuleLoadData(&p->mod, cubin.data()) != CUDA_SUCCESS || drv->ModuleGetFunction(&p->fn, p->mod, kname.c_str()) != CUDA_SUCCESS) {
        delete p;
        return -1;
    }
    drv->FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
    int pct = (int)(((long long)(smem + 1024) * minb * 100 + 228 * 1024 - 1) / (228 * 1024));
    drv->FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct > 100 ? 100 : pct);
    memset(&p->desc, 0, sizeof p->desc);
    p->desc.id = (int)g_rader.size();
    p->desc.family = HS_CT_RADER;
    p->desc.n = L;
    p->desc.nst = nst;
    for (int s = 0; s < nst; ++s) p->desc.radix[s] = rx[s];
    p->desc.split_min = -2;
    p->desc.threads = T; p->desc.smem_bytes = smem; p->desc.exact = 1;
    if (getenv("HSFFT_VERBOSE"))
        fprintf(stderr, "hsfft: plan-time Rader kernel N=%d: L=%d T=%d LOGPAD=%d/%d, %d bytes smem, %d CTAs/SM asked\n", L + 1, L, T, lpf, lpi, smem, minb);
    g_rader.push_back(p);
    *desc = p->desc;
    return p->desc.id;
}

extern "C" void hs_cu_jit_rader_launch(int handle, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i, const void *d_perm,
                                       const void *d_bk, long long howmany, void *stream)
{
    JitRader *p;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (handle < 0 || handle >= (int)g_rader.size()) { fprintf(stderr, "Error: unknown plan-time Rader kernel %d\n", handle); exit(EXIT_FAILURE); }
        p = g_rader[handle];
    }
    if (howmany <= 0 || howmany > 2147483647LL) { fprintf(stderr, "Error: batch too large for one launch (%lld CTAs)\n", howmany); exit(EXIT_FAILURE); }
    int pf = 0;
    void *args[] = {(void *)&d_in, (void *)&d_out, (void *)&d_tw_f, (void *)&d_tw_i, (void *)&d_perm, (void *)&d_bk, (void *)&howmany, (void *)&pf};
    const CUresult r = driver()->LaunchKernel(p->fn, (unsigned)howmany, 1, 1, (unsigned)p->threads, 1, 1, (unsigned)p->smem, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: CUDA failure %d launching the plan-time Rader kernel for N=%d\n", (int)r, p->L + 1); exit(EXIT_FAILURE); }
    hs_cu_count_launch();
    hs_cu_timing_mark("rader (plan-time kernel)", stream);
}
