// hs_plans.cuh -- the compile-time plans (stage radices, team sizes, padding) and their ids, shared by the
// dispatch tables hs_ct_table.cuh (one-launch kernels and stand-alone passes) and hs_pipe_table.cuh
// (persistent pipelines).
#pragma once
#include "hs_smem_fft.cuh"
#include "hs_conv_fft.cuh"

// ------------------------------------------------------------------------ compile-time plan table
// Stage radices in execution order (the reference's factor list reversed; runs of power-of-two
// factors regrouped, which only exact constants allow -- SURVEY.md section 8c).  LOGPAD values come
// from tests/emu/bank_check.py (all stages conflict free).
//                      N     T   X  LOGPAD SPLIT GROUPS radices...
using P4096a = HsPlan<4096, 256, 1, 4, -1, 0, 16, 16, 16>;       // reference factors [8,8,8,8]
using P4096b = HsPlan<4096, 512, 1, 3, -1, 0, 8, 8, 8, 8>;
// Twiddle mode of the plans that are bound by the L1/LSU pipe plus FP64 rather than by HBM (measured on B200, fraction of
// the HBM peak with SPLIT 0 / -1 / -2 = load every row / rows 1,2,4,8 / row 1 only and form the rest by products):
// N = 3000 61.4 / 75.6 / 78.4 %, fused Bluestein N = 1009 23.8 / 26.1 / 26.9 %, long Bluestein N = 10007 (sub-plans)
// 11.6 / 12.5 / 13.0 %; the two-pass c2c sizes do not care.  Fewer loads win even at the price of more multiplies.
#ifndef HS_3000_SPLIT
#define HS_3000_SPLIT -2
#endif
#ifndef HS_B2048_SPLIT
#define HS_B2048_SPLIT -2
#endif
// N = 3000: reference factors [8,5,5,5,3] verbatim, passes {3,5}{5,5}{8}.  Fewer, fatter threads: 128 threads (4 whole warps)
// hold 30 / 25 / 24 points each in the three passes, 158 registers without a spill, THREE CTAs per SM.  With 200 threads (15 / 25 /
// 16 points) a CTA is allocated 8 warps, so two CTAs per SM cap it at 128 registers and a third at 80 (spills, 59 %).
// Measured on B200, 65536 transforms (profiles/r2_fat_threads_ab.txt): T = 128 x 3 CTAs 1.092 ms = 88.1 % of the HBM peak; T = 125
// (the radix-8 pass in exactly three rounds, but the second round of a pass starts at butterfly 125: its 8-lane groups straddle
// two 128-byte lines, 16 % more shared-memory wavefronts) 1.108 ms = 86.7 %; T = 120 (25-point pass exact) 1.134 ms; T = 200 x 2
// CTAs 1.227 ms = 78.3 %; T = 125 x 4 CTAs (128 registers, 192 B of spills) 1.587 ms.
using P3000  = HsPlan<3000, 128, 1, 3, HS_3000_SPLIT, 221, 3, 5, 5, 5, 8>;
using P3000t = HsPlan<3000, 200, 1, 3, HS_3000_SPLIT, 221, 3, 5, 5, 5, 8>; // round 1 / early round 2 (A/B: HSFFT_PREFER_CT)
using P1024  = HsPlan<1024, 64, 4, 3, -1, 0, 16, 8, 8>;          // reference factors [8,8,8,2]
using P2048  = HsPlan<2048, 128, 2, 3, -1, 0, 16, 16, 8>;        // reference factors [8,8,8,4]
using P512   = HsPlan<512, 64, 4, 3, -1, 0, 8, 8, 8>;
using P256   = HsPlan<256, 16, 16, 4, -1, 0, 16, 16>;
// N = 8192 in THREE passes (two shared-memory exchanges instead of three): 256 threads of 32 points, the first pass a fused
// {4,8}; 230 registers, no spill, one CTA per SM either way (139 KB).  Measured: 1.820 ms per 32768 transforms = 72.1 % of the
// HBM peak against 2.026 ms = 64.8 % for 512 threads in four passes 8.8.8.16 (P8192q); {16}{16}{8,4} the same (1.822 ms).
using P8192  = HsPlan<8192, 256, 1, 4, -1, 211, 4, 8, 16, 16>;
using P8192q = HsPlan<8192, 512, 1, 4, -1, 0, 8, 8, 8, 16>;
// (8192 points with 1024 threads -- 8 points per thread, five passes, 64 registers, twice the warps -- measured slower: 56 % and
// 53 % of the HBM peak for 4.8.8.8.4 / 8.8.8.8.2 against 65 %; N = 3000 at three CTAs per SM spills and drops from 78 % to 59 %)
// fused Bluestein, M = 2048 (N = 513..1024, e.g. the N = 1009 config): forward / inverse pair
using PB2048f = HsPlan<2048, 128, 1, 4, HS_B2048_SPLIT, 0, 8, 16, 16>;
using PB2048i = HsPlan<2048, 128, 1, 4, HS_B2048_SPLIT, 0, 16, 16, 8>;
// the same fused kernel for the other pad lengths one SM can hold (N = 65..2048 all run as one kernel); the radix lists
// are palindromes, so the inverse plan is the forward plan and the spectrum is handed over in registers
// (all conflict free, tests/emu/bank_check.py).  One transform per CTA (four for M = 256) and many CTAs per SM: measured
// against 2-4 transforms per CTA, M = 2048 0.300 -> 0.267 ms, 1024 0.283 -> 0.235, 512 0.237 -> 0.208, 256 0.183 -> 0.162
// (profiles/r1b_small_cta_variants.txt)
using PB256  = HsPlan<256, 16, 4, 4, HS_B2048_SPLIT, 0, 16, 16>;
using PB512  = HsPlan<512, 64, 1, 3, HS_B2048_SPLIT, 0, 8, 8, 8>;
using PB1024 = HsPlan<1024, 64, 1, 3, HS_B2048_SPLIT, 0, 8, 16, 8>;
using PB4096 = HsPlan<4096, 256, 1, 4, HS_B2048_SPLIT, 0, 16, 16, 16>;
using PB8192f = HsPlan<8192, 512, 1, 3, HS_B2048_SPLIT, 0, 8, 16, 8, 8>; // one CTA per SM (147 KB); N = 2049..4096
using PB8192i = HsPlan<8192, 512, 1, 3, HS_B2048_SPLIT, 0, 8, 8, 16, 8>;
// split chirp-z (hs_bluestein_split_kernel): two cyclic convolutions of length M' >= N per transform, M' = the plan length.
// 10240 = 16*8*5*16 with 640 threads: every radix-16 stage has exactly one butterfly per thread, the radix-8 stage two,
// the radix-5 stage 3.2 (four rounds); both orders are conflict free with one pad slot per 16 points (bank_check.py).
// Exact radix-5 constants: the reference runs this convolution with power-of-two stages only.
// (for 10240 = 16*8*5*16 with 640 threads this pair needs all 16-20 points of a pass in 96 registers: 2 KB of spill stores
// per thread, 8.7 ms against 5.0 ms for the in-place transforms below, so it is not built)
// the same split chirp-z on the in-place transforms of hs_conv_fft.cuh (digit-reversed spectrum, one butterfly at a time
// per thread): 16 points in registers instead of the 16-20 of a whole pass, so 640 threads fit 96 registers without a spill
// (the auto-sort pair above: 2 KB of spill stores per thread).  All conflict free with one pad slot per 16 points
// (tests/emu/bank_check.py inplace ...).
// Twiddle mode -3 (row 1 and successive powers, two live twiddles): 96 registers, 160 B of spill stores per thread; the
// product tree (-2) spills 380 B, loading rows 1,2,4,8 (-1) 400 B (ptxas -v).  With the tensor-memory kernel: -3 4.33 ms,
// -2 4.59, -1 4.77, all rows loaded 5.77 (N = 10007 x 16384).
#ifndef HS_CONV_TWM
#define HS_CONV_TWM -3
#endif
// Measured on B200, N = 10007 x 16384 (profiles/r2_sweep_split_variants.txt): this plan 5.11 ms; product tree 4.96, rows 1,2,4,8
// loaded 5.07, 320 threads with twice the butterflies each (no spill at 168 registers) 5.00 -- all within run-to-run noise of
// each other, so only one is built.  For M' = 8192 / 4096 the auto-sort pair above is faster (4.10 vs 4.60 ms, 1.61 vs 4.67 ms).
#ifndef HS_CONV_T
#define HS_CONV_T 640
#endif
using PC10240  = HsConvPlan<10240, HS_CONV_T, 4, HS_CONV_TWM, 16, 8, 5, 16>;
// M' = 8192 the same way: N = 8191 3.81 ms against 4.11 ms for the auto-sort pair, but N = 4099 3.41 against 3.28 ms, so the
// planner takes it for the upper part of the range only; M' = 4096 in place measured no faster (1.61 ms both, N = 2053 3.02
// against 2.58 ms) and is not built (profiles/r2_split_inplace_tmem_8192_4096.txt)
using PC8192   = HsConvPlan<8192, 512, 4, HS_CONV_TWM, 16, 8, 4, 16>;
// sub-plans of the two-pass transforms (N = n1 * n2 > 8192); X is unused there
#ifndef HS_SUB_SPLIT
#define HS_SUB_SPLIT -2
#endif
using S128  = HsPlan<128, 8, 1, 4, HS_SUB_SPLIT, 0, 8, 16>;
using S256  = HsPlan<256, 16, 1, 4, HS_SUB_SPLIT, 0, 16, 16>;
using S512  = HsPlan<512, 64, 1, 3, HS_SUB_SPLIT, 0, 8, 8, 8>;
using S1024 = HsPlan<1024, 64, 1, 3, HS_SUB_SPLIT, 0, 16, 8, 8>;      // (T=128, 8.8.8.2 measured slower: 29% vs 36% at 2^20;
                                                                      //  T=128, 8.8.4.4 in round 2: 2^20 29.0 vs 39.0 %)
// 512 points as a SECOND pass (family HS_CT_SUB2): 32 threads of 16 points, hence 8 columns per tile -- 128-byte instead of
// 64-byte runs on the strided stores.  Measured on B200 (profiles/r2_subplan_variants.txt): second pass 1.44 vs 1.60 ms at
// 2^18 x 1024, the real four-step's mirrored pass 1.34 vs 1.83 ms; as a FIRST pass it loses (2.54 vs 1.82 ms), so S512 stays.
using S512b = HsPlan<512, 32, 1, 3, HS_SUB_SPLIT, 0, 16, 8, 4>;
// Which split a two-pass length takes is decided by its SECOND pass (strided stores): per 8.6 GB moved the 256-point plan
// needs 1.23 ms, the second-pass 512-point plan 1.54-1.68, the 1024-point plan 1.87, while first passes run at 1.25-1.41 ms
// whatever their length (profiles/r2_subplan_second_pass_and_r4_splits.txt).  So 2^17 = 512 x 256 and 2^18 = 1024 x 256
// (before: 256 x 512 and 512 x 512).
#ifndef HS_T17_P1
#define HS_T17_P1 S512
#define HS_T17_P2 S256
#endif
#ifndef HS_T18_P1
#define HS_T18_P1 S1024
#define HS_T18_P2 S256
#endif
// profiling-only ablations (never selected unless HSFFT_PREFER_CT names them)
using P4096x = HsPlan<4096, 256, 1, 4, 0, 0, 16, 16, 16>;
using P3000x = HsPlan<3000, 200, 1, 3, -1, 0, 3, 5, 5, 5, 8>;
using P4096y = HsPlan<4096, 256, 1, 4, -2, 0, 16, 16, 16>;
using P3000y = HsPlan<3000, 200, 1, 3, 0, 0, 3, 5, 5, 5, 8>;       // one stage per pass, load-all twiddles (round-1 first contact)
using P3000z = HsPlan<3000, 200, 1, 3, -1, 2111, 3, 5, 5, 5, 8>;  // passes {3,5} {5} {5} {8}
using P3000v = HsPlan<3000, 125, 1, 3, HS_3000_SPLIT, 221, 3, 5, 5, 5, 8>; // 125 threads: the radix-8 pass exact, rounds not 8-aligned
using P3000u = HsPlan<3000, 120, 1, 3, HS_3000_SPLIT, 221, 3, 5, 5, 5, 8>; // the 25-point pass exact (120 butterflies), the radix-8 pass in four rounds
using P8192b = HsPlan<8192, 256, 1, 5, -1, 112, 16, 16, 8, 4>;    // the fused 32-point pass last instead of first

// inverse-order partners of the sub-plans (first radix = the forward plan's last radix)
using S128i  = HsPlan<128, 8, 1, 4, HS_SUB_SPLIT, 0, 16, 8>;
using S256i  = S256;
using S512i  = S512;
using S1024i = HsPlan<1024, 64, 1, 3, HS_SUB_SPLIT, 0, 8, 8, 16>;

// Bluestein-only sub-plans with a factor 5 (exact constants): pad lengths 5*2^k sit between the powers of two, e.g.
// N = 10007 needs M >= 20013: 20480 = 160 x 128 instead of 32768 (1.6 x fewer points in every pass)
using S160x  = HsExact<HsPlan<160, 20, 1, 3, HS_SUB_SPLIT, 0, 5, 4, 8>>;
using S160xi = HsExact<HsPlan<160, 20, 1, 3, HS_SUB_SPLIT, 0, 8, 4, 5>>;

// Rader (prime N, cyclic convolution of length N-1 held in shared memory; hs_rader_kernel): N = 1009 -> 1008 = 7*9*16,
// full-precision radix-3/7 constants.  The forward plan ends and the inverse plan starts with the same pass, so the
// spectrum is handed over in registers; pads differ because the strided reads of the two directions differ
// (forward: last pass reads stride 16 -> one pad per 16; inverse: last pass reads stride 7, conflict free unpadded).
// Measured on B200, 16384 transforms (profiles/r1b_rader_variants.txt): one transform per CTA and 5-6 CTAs per SM
// 0.194-0.200 ms whatever the pass split (T = 112 / 63 / 126); two or four per CTA 0.22-0.25 ms; the reference's
// chirp-z route (two 2048-point transforms) 0.300 ms.
using PR1008f  = HsExact<HsPlan<1008, 112, 1, 4, HS_B2048_SPLIT, 121, 7, 3, 3, 16>>;
using PR1008i  = HsExact<HsPlan<1008, 112, 1, 0, HS_B2048_SPLIT, 121, 16, 3, 3, 7>>;
using PR1008af = HsExact<HsPlan<1008, 63, 1, 4, HS_B2048_SPLIT, 121, 7, 3, 3, 16>>;      // 3 | 2 | 1 butterflies per thread
using PR1008ai = HsExact<HsPlan<1008, 63, 1, 0, HS_B2048_SPLIT, 121, 16, 3, 3, 7>>;
using PR1008gf = HsExact<HsPlan<1008, 112, 1, 0, HS_B2048_SPLIT, 112, 7, 16, 3, 3>>;     // 9-point hand-over: 36 registers instead of 64
using PR1008gi = HsExact<HsPlan<1008, 112, 1, 0, HS_B2048_SPLIT, 211, 3, 3, 16, 7>>;
// the round-1 forms with several transforms per CTA, kept for A/B runs (HSFFT_PREFER_CT)
using PB256x16 = HsPlan<256, 16, 16, 4, HS_B2048_SPLIT, 0, 16, 16>;
using PB2048x2f = HsPlan<2048, 128, 2, 4, HS_B2048_SPLIT, 0, 8, 16, 16>;
using PB2048x2i = HsPlan<2048, 128, 2, 4, HS_B2048_SPLIT, 0, 16, 16, 8>;

enum { CT_4096A = 0, CT_4096B, CT_3000, CT_1024, CT_2048, CT_512, CT_256, CT_8192, CT_B2048, CT_4096X, CT_3000X, CT_4096Y, CT_3000Y, CT_3000Z,
       CT_T14, CT_T15, CT_T16, CT_T17, CT_T18, CT_T19, CT_T20,
       CT_S128, CT_S256, CT_S512, CT_S1024,
       CT_TX20480, CT_TX40960, CT_TX81920, CT_TX163840, CT_S160X,
       CT_B256, CT_B512, CT_B1024, CT_B4096, CT_B8192,
       CT_R1008, CT_R1008A, CT_R1008G, CT_R256, CT_B2048X2, CT_B256X16,
       CT_H4096, CT_H8192, CT_C10240, CT_C8192, CT_S512B, CT_3000T, CT_3000V, CT_8192Q, CT_8192B, CT_3000U, CT_COUNT };
