// Gram build on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM.
//
// Same contraction as the fp64 kernel in gram.cu -- M[(qa,qb),qc] = sum_rows U[row,(qa,qb)] V[row,qc]
// with U = w*pair(fa)*pair(fb), V = pair(fc) -- but the operand tiles are synthesised by producer warps
// straight into the UMMA canonical shared-memory layout (K-major, no swizzle: 8-row x 16-byte core
// matrices, SBO = 128 B between 8-row groups, LBO = rows*16 B between the two 4-sample halves of an
// MMA-K step) and multiplied by tcgen05.mma into fp32 TMEM accumulators.
//
//   mode 1 (TF32)   : one MMA per K step on the tf32-rounded operands.
//   mode 2 (3xTF32) : operands split x = hi + lo (hi = tf32(x), lo = x - hi); three MMAs per K step
//                     (hi*hi, hi*lo, lo*hi) recover ~fp32 products.
// Every FLUSH_ROWS rows the accumulator is drained TMEM -> registers -> fp64 and added to M in
// global memory, which bounds the fp32 accumulation length.
//
// A pre-pass (tc_stage_kernel) evaluates the feature maps / row divisors once and writes the raw factors
// feature-major in fp32, Z[(w*fa | fa | fb | fc)][rows], so a 16-sample piece of a factor row is 64
// contiguous bytes.  Warp roles in a CTA of 288 threads: warp 0 allocates TMEM and issues the MMAs (one
// elected lane); warps 1..8 cp.async the next chunks of Z into a 3-slot shared ring, synthesise the operand
// tiles, and drain the accumulator at flush points.  Operand stages are handed over with mbarriers:
// full[s] (producers -> MMA, after fence.proxy.async) and empty[s] (tcgen05.commit -> producers).
#include <stdlib.h>
#include <atomic>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"

namespace tn {

constexpr int TC_M = 128;          // rows of one U tile = MMA M
constexpr int TC_KC = 16;          // samples per pipeline stage (two MMA K steps of 8)
constexpr int TC_KCP = TC_KC + 4;  // padded row of the transposed raw-factor staging (floats)
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_PROD = TC_PROD_WARPS * 32;
constexpr int TC_THREADS = 32 + TC_PROD;
constexpr int TC_RAW_SLOTS = 3;    // shared ring of raw-factor chunks filled by cp.async
constexpr int TC_FLUSH_ROWS_DEFAULT = 2048;

struct TcFactor {
    const double* ptr;
    int64_t ld;
    int m;
    int div;
    int map_kind;
};

struct TcParams {
    const unsigned long long* amax;   // max |fa|, |fb|, |fc|, |w| (bit patterns), see tc_absmax_kernel
    const float* Z;      // staged factors, feature-major: rows [w*fa (mA) | fa (mA) | fb (mB) | fc (mC)], each zpitch floats
    int64_t zpitch;      // floats per row of Z (rows rounded up to TC_KC, zero padded)
    int mA, mB, mC;
    int64_t rows;
    int64_t rows_per_split;
    double* M;
    int nA, nB, nC;   // pair counts
    int BN;           // V tile width (multiple of 16, <= 256)
    int T;            // U tiles per CTA (1 or 2)
    int nstages;
    int split;        // 1 = 3xTF32, 0 = TF32
    int flush_rows;   // rows accumulated in fp32 before a drain to fp64 (multiple of TC_KC)
    int dbg;          // TN_TC16_DBG (measurement only, wrong results): bit 0 = producers skip the synthesis, bit 1 = no MMAs are issued,
                      // bit 2 = no fence.proxy.async, bit 3 = no raw-factor ring, bit 4 = no producer barrier, bit 5 = no drain
    int planar;       // 1: Z is staged piece-planar, Z[((s/16)*4 + (s%16)/4) * z_rows + row][s%4] (see gram_tc_kernel<.., PLANAR>)
};

// ---- range scaling.  The operands are staged in fp32; factors of long chains span hundreds of binades (environments of a
//      normalised 90-site train are ~1e-30), so each factor (and the weights) is scaled by a power of two to magnitude
//      O(1) before the conversion, and the drain multiplies the result back: M = 2^(2(ea+eb+ec)+ew) * M_scaled, exactly.
//      amax[0..3] = max |fa|, |fb|, |fc|, |w| (as ordered uint bit patterns of non-negative doubles).
__global__ void __launch_bounds__(256)
tc_absmax_kernel(TcFactor fa, TcFactor fb, TcFactor fc, const double* __restrict__ w, int64_t rows, unsigned long long* __restrict__ amax) {
    const int mA = fa.m, mB = fb.m, mC = fc.m;
    const int msum = mA + mB + mC + 1;
    double loc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t total = rows * msum;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx / msum;
        const int i = (int)(idx - s * msum);
        double v;
        int which;
        if (i < mA) { v = map_eval(fa.map_kind, fa.ptr + (fa.div == 1 ? s : s / fa.div) * fa.ld, i); which = 0; }
        else if (i < mA + mB) { v = map_eval(fb.map_kind, fb.ptr + (fb.div == 1 ? s : s / fb.div) * fb.ld, i - mA); which = 1; }
        else if (i < mA + mB + mC) { v = map_eval(fc.map_kind, fc.ptr + (fc.div == 1 ? s : s / fc.div) * fc.ld, i - mA - mB); which = 2; }
        else { v = w ? w[s] : 1.0; which = 3; }
        loc[which] = fmax(loc[which], fabs(v));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double v = loc[q];
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(&amax[q], (unsigned long long)__double_as_longlong(v));
    }
}

// exponent e such that x * 2^-e is in [0.5, 1) (0 for x == 0, inf or nan)
__device__ __forceinline__ int tc_exponent(unsigned long long bits) {
    const double x = __longlong_as_double((long long)bits);
    if (!(x > 0.0) || isinf(x)) return 0;
    int e;
    frexp(x, &e);
    return e;
}

// ---- pre-pass: factors (fp64, sample-major, possibly mapped / shared by V rows) -> Z (fp32, feature-major)
__global__ void __launch_bounds__(256)
tc_stage_kernel(TcFactor fa, TcFactor fb, TcFactor fc, const double* __restrict__ w, int64_t rows, float* __restrict__ Z,
                int64_t zpitch, const unsigned long long* __restrict__ amax, int planar = 0, int64_t plane_halves = 0) {
    // planar: 0 = row-major fp32, 1 = planes of 4 fp32 samples (16 B per Z row), 2 = planes of 8 fp16 samples (16 B per Z row),
    // 3 = as 2 with every plane padded to plane_halves fp16 (the shared-memory plane stride of gram_tc16_run_kernel): a 64-sample
    //     chunk is one contiguous image of a raw-factor slot (8 planes; the zero row and the padding come from a memset of Z)
    __half* Zh = reinterpret_cast<__half*>(Z);
    __shared__ float tile[32][33];
    const double sa = ldexp(1.0, -tc_exponent(amax[0])), sb = ldexp(1.0, -tc_exponent(amax[1]));
    const double sc = ldexp(1.0, -tc_exponent(amax[2])), sw = ldexp(1.0, -tc_exponent(amax[3]));
    const int mA = fa.m, mB = fb.m, mC = fc.m;
    const int msum = mA + mB + mC;
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    const int i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    // read: lane = feature (coalesced along a sample's row), 8 samples per pass
    for (int r = ty; r < 32; r += 8) {
        const int64_t s = s0 + r;
        const int i = i0 + tx;
        float v = 0.f;
        if (s < rows && i < msum) {
            const TcFactor& f = (i < mA) ? fa : ((i < mA + mB) ? fb : fc);
            const int il = (i < mA) ? i : ((i < mA + mB) ? i - mA : i - mA - mB);
            const int64_t ri = (f.div == 1) ? s : s / f.div;
            v = (float)(map_eval(f.map_kind, f.ptr + ri * f.ld, il) * ((i < mA) ? sa : ((i < mA + mB) ? sb : sc)));
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    // write: lane = sample (coalesced along Z rows); rows past `rows` are written as zeros up to zpitch
    for (int c = ty; c < 32; c += 8) {
        const int i = i0 + c;
        const int64_t s = s0 + tx;
        if (i < msum && s < zpitch) {
            const float v = tile[tx][c];
            const int64_t zr = 2 * mA + mB + mC;                         // Z rows
            // row-major: Z[row][s];  piece-planar: the 16 bytes (4 samples) of every row of one piece are contiguous
            auto at = [&](int64_t row) -> int64_t {
                if (planar == 3) return (s >> 3) * plane_halves + row * 8 + (s & 7);
                if (planar == 2) return (((s >> 5) * 4 + ((s >> 3) & 3)) * zr + row) * 8 + (s & 7);
                return planar ? ((((s >> 4) * 4 + ((s >> 2) & 3)) * zr + row) * 4 + (s & 3)) : (row * zpitch + s);
            };
            const float vw = (i < mA && s < rows) ? v * (float)((w ? w[s] : 1.0) * sw) : 0.f;
            if (planar >= 2) {
                Zh[at(mA + i)] = __float2half_rn(v);
                if (i < mA) Zh[at(i)] = __float2half_rn(vw);
            } else {
                Z[at(mA + i)] = v;
                if (i < mA) Z[at(i)] = vw;
            }
        }
    }
}

// TMEM accumulator -> registers -> (transpose through shared memory) -> fp64 atomic adds into M.
// Warp w may touch TMEM lanes 32*(w%4)..+31 (= 32 rows of a U tile); the two producer warps that share a lane quarter
// split the columns.  tcgen05.ld hands every lane one ROW; the 32x32 block is transposed through a padded shared
// scratch so that each RED instruction adds 32 consecutive doubles of one row of M (coalesced) instead of 32 rows.
// fp64 reduction into M with an L2 evict_last hint: the tile of M a CTA owns is flushed again a few milliseconds later, and
// the tiles of all resident CTAs together (148 x 512 KB) fit the L2 -- keep them there instead of round-tripping through HBM.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// The LAST flush of a CTA: its tile of M is not touched again by this CTA, so it should leave the L2 first and not compete
// (as an evict_last line) with the tiles of the CTAs that are still accumulating.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_f64_keep(double* addr, double v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(pol) : "memory");
}

__device__ __noinline__ void drain_accumulator(uint32_t tmem_base, int cols_total, int BN, int warp, int lane, int64_t u0,
                                               int64_t nU, int v0, int nC, double* __restrict__ M, float* __restrict__ scratch,
                                               double unscale, int tile_stride = TC_M, bool final_flush = false, int nparts = 2,
                                               int red_policy = 1) {
    const int q = warp & 3;
    const int half = (warp - 1) >> 2;                  // which part of the columns (two warps per lane quarter by default, four with 16 producer warps)
    const int ngroups = (cols_total + 31) / 32;
    const int per_part = (ngroups + nparts - 1) / nparts;
    const int g_lo = half * per_part;
    const int g_hi = min(ngroups, g_lo + per_part);
    float* sc = scratch + (size_t)(warp - 1) * (32 * 33);
    // red_policy 1: keep the tile in L2 between flushes (short windows); 0: let it leave first (long windows: the tile is not touched
    // again for hundreds of stages, and the L2 is better spent on the operand streams all CTAs of a wave share)
    const uint64_t l2_keep = (final_flush || red_policy == 0) ? l2_evict_first_policy() : l2_evict_last_policy();
    for (int g = g_lo; g < g_hi; ++g) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), r);
#pragma unroll
        for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(r[e]);
        __syncwarp();
        const int col = g * 32 + lane;                 // this lane's accumulator column
        const int t = col / BN;
        const int gv = v0 + (col - t * BN);
        const bool col_ok = (col < cols_total) && (gv < nC);
        const int64_t gu0 = u0 + (int64_t)t * tile_stride + q * 32;
        double* dst = M + gu0 * nC + gv;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            if (col_ok && gu0 + rr < nU) red_add_f64_keep(dst + (int64_t)rr * nC, unscale * (double)sc[rr * 33 + lane], l2_keep);
        }
        __syncwarp();
    }
}

// PLANAR selects the layout of a raw-factor slot.  false: row-major, one Z row = TC_KC floats + 4 of padding (80 B), the layout all
// measurements of round 1 were taken with.  true (the default since round 2; TN_TC_RAW_ROWMAJOR=1 selects false): four planes, one per
// 4-sample piece, plane p holding 16 B per Z row at p * plane_stride + row * 16 with plane_stride = 32 (mod 128), filled by FOUR
// BULK COPIES (cp.async.bulk, one per plane, issued by one thread, completion on an mbarrier) from a staging buffer that the
// pre-pass writes piece-planar.  Why: the round-1 ncu capture (profiles/r1_ncu_gram_tc_v6_hotspots.txt) attributes 6.1e9 of the
// kernel's 8.2e9 excessive shared-memory wavefronts to the cp.async (LDGSTS) writes of the ring -- 3.7 wavefronts per ideal one
// here and 8 per ideal one in syrk_tc_kernel, whose LDS of the very same addresses is conflict free: LDGSTS spends about one
// shared-memory wavefront per 32-byte global sector, whatever the lanes' bank pattern.  The bulk-copy engine writes whole lines,
// takes the fill off the LSU pipe and off the producers' instruction stream, and the operand loads of consecutive Z rows become
// contiguous in the planar slot.
template <int SPLIT, int T, bool PLANAR = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gram_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, NS = p.nstages;
    const int mA = p.mA, mB = p.mB, mC = p.mC;

    // ---- shared memory carve-up: [NS operand stages][TC_RAW_SLOTS raw-factor slots][mbarriers][tmem slot]
    constexpr uint32_t a_tile_bytes = TC_M * TC_KC * 4;      // one U tile, hi or lo, one stage
    const uint32_t b_tile_bytes = (uint32_t)BN * TC_KC * 4;  // V tile, hi or lo
    const uint32_t stage_bytes = 2 * T * a_tile_bytes + 2 * b_tile_bytes;
    const uint32_t z_rows = (uint32_t)(2 * mA + mB + mC);
    const uint32_t raw_rows = z_rows + 1;                      // + an all-zero row for the padding rows of the tiles
    const uint32_t raw_bytes = raw_rows * TC_KCP * 4;          // slot pitch (the planar layout needs less and keeps the pitch)
    const uint32_t plane_stride = ((raw_rows * 16 + 95) / 128) * 128 + 32;     // >= raw_rows * 16, = 32 (mod 128)
    const uint32_t row_pitch = PLANAR ? 16u : (uint32_t)(TC_KCP * 4);          // bytes between two Z rows of a slot
    const uint32_t piece_pitch = PLANAR ? plane_stride : 16u;                  // bytes between two 4-sample pieces of a Z row
    uint8_t* stage_base = smem_raw;
    uint8_t* raw_base = smem_raw + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)TC_RAW_SLOTS * raw_bytes);
    uint64_t* full = bars;              // [NS]
    uint64_t* empty = bars + NS;        // [NS]
    uint64_t* acc_full = bars + 2 * NS;
    uint64_t* acc_empty = bars + 2 * NS + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
    uint64_t* raw_full = bars + 2 * NS + 3;      // [TC_RAW_SLOTS], PLANAR only (the launcher adds their bytes)

    const uint32_t tmem_cols_needed = (uint32_t)(T * BN);
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], TC_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, TC_PROD_WARPS);
        if (PLANAR)
            for (int i = 0; i < TC_RAW_SLOTS; ++i) mbar_init(&raw_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (PLANAR) {
        for (int i = tid; i < TC_RAW_SLOTS * TC_KC; i += TC_THREADS) {   // the zero row of every raw slot: 4 floats in each plane
            const int slot = i / TC_KC, e = i % TC_KC;
            reinterpret_cast<float*>(raw_base + (size_t)slot * raw_bytes + (size_t)(e >> 2) * plane_stride + (size_t)z_rows * 16)[e & 3] = 0.f;
        }
    } else {
        for (int i = tid; i < TC_RAW_SLOTS * TC_KCP; i += TC_THREADS)   // the zero row of every raw slot
            reinterpret_cast<float*>(raw_base + (size_t)(i / TC_KCP) * raw_bytes)[z_rows * TC_KCP + (i % TC_KCP)] = 0.f;
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t k_begin = (int64_t)blockIdx.z * p.rows_per_split;          // multiples of TC_KC; Z is zero padded
    const int64_t k_end = min(p.zpitch, k_begin + p.rows_per_split);
    const int64_t nchunks = (k_end > k_begin) ? (k_end - k_begin) / TC_KC : 0;
    const int64_t chunks_per_flush = p.flush_rows / TC_KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t u0 = (int64_t)blockIdx.x * (TC_M * T);
    const int v0 = blockIdx.y * BN;

    if (warp == 0) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc(TC_M, BN);
            const uint32_t lbo_a = TC_M * 16, lbo_b = (uint32_t)BN * 16, sbo = 128;
            uint32_t acc_phase = 0;
            int s = 0;
            uint32_t ph = 0;
            int64_t in_window = 0;
            for (int64_t c = 0; c < nchunks; ++c) {
                const bool first_of_window = in_window == 0;
                if (first_of_window && c > 0) {
                    mbar_wait(acc_empty, acc_phase);   // accumulator drained by the producers
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * stage_bytes);
                const uint32_t b_hi = sb + 2 * T * a_tile_bytes;
                const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t a_hi = sb + (uint32_t)t * 2 * a_tile_bytes;
                    const uint32_t a_lo = a_hi + a_tile_bytes;
                    const uint32_t d = tmem_base + (uint32_t)(t * BN);
#pragma unroll
                    for (int j = 0; j < TC_KC / 8; ++j) {
                        const uint32_t ao = (uint32_t)(2 * j) * lbo_a, bo = (uint32_t)(2 * j) * lbo_b;
                        const uint32_t acc0 = (first_of_window && j == 0) ? 0u : 1u;
                        umma_tf32(d, make_desc(a_hi + ao, lbo_a, sbo), make_desc(b_hi + bo, lbo_b, sbo), idesc, acc0);
                        if (SPLIT) {
                            umma_tf32(d, make_desc(a_hi + ao, lbo_a, sbo), make_desc(b_lo + bo, lbo_b, sbo), idesc, 1u);
                            umma_tf32(d, make_desc(a_lo + ao, lbo_a, sbo), make_desc(b_hi + bo, lbo_b, sbo), idesc, 1u);
                        }
                    }
                }
                umma_commit(&empty[s]);                          // smem stage free once these MMAs have read it
                const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
                if (last_of_window) {
                    umma_commit(acc_full);                       // accumulator complete for this window
                    in_window = 0;
                }
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================== producers / epilogue ===============================
        const int pt = tid - 32;  // 0..255
        const double unscale = ldexp(1.0, 2 * (tc_exponent(p.amax[0]) + tc_exponent(p.amax[1]) + tc_exponent(p.amax[2])) +
                                              tc_exponent(p.amax[3]));
        const uint32_t raw_s = smem_u32(raw_base);
        const uint32_t zero_row = z_rows * row_pitch;                  // byte offset of the zero row in a raw slot
        // -- rows of the operand tiles this thread synthesises (fixed for the kernel)
        uint32_t usrc[4] = {zero_row, zero_row, zero_row, zero_row};   // byte offsets of w*fa[ia], fa[ja], fb[ib], fb[jb]
        int u_tile = 0, u_row = pt & 127, u_c0 = 0;
        constexpr int U_NC = (T == 2) ? TC_KC / 4 : TC_KC / 8;          // k-chunks of 4 samples per thread and stage
        if (T == 2) u_tile = pt >> 7;
        else u_c0 = (pt >> 7) * U_NC;
        {
            const int64_t gu = u0 + (int64_t)u_tile * TC_M + u_row;
            if (gu < nU) {
                const int qa = (int)(gu / p.nB), qb = (int)(gu - (int64_t)qa * p.nB);
                int ia, ja, ib, jb;
                pair_decode(qa, mA, ia, ja);
                pair_decode(qb, mB, ib, jb);
                usrc[0] = (uint32_t)ia * row_pitch;
                usrc[1] = (uint32_t)(mA + ja) * row_pitch;
                usrc[2] = (uint32_t)(2 * mA + ib) * row_pitch;
                usrc[3] = (uint32_t)(2 * mA + jb) * row_pitch;
            }
        }
        uint32_t vsrc[2] = {zero_row, zero_row};
        if (pt < BN && v0 + pt < p.nC) {
            int ic, jc;
            pair_decode(v0 + pt, mC, ic, jc);
            vsrc[0] = (uint32_t)(2 * mA + mB + ic) * row_pitch;
            vsrc[1] = (uint32_t)(2 * mA + mB + jc) * row_pitch;
        }
        const uint32_t udst = (uint32_t)u_tile * 2 * a_tile_bytes + (uint32_t)u_row * 16 + (uint32_t)u_c0 * (TC_M * 16);
        const uint32_t vdst = 2 * T * a_tile_bytes + (uint32_t)pt * 16;
        const uint32_t lbo_b = (uint32_t)BN * 16;
        const uint32_t stage_s = smem_u32(stage_base);

        // -- cp.async plan: 16-byte pieces (4 samples of one Z row) of a chunk, spread over the producer threads
        const int npieces = (int)z_rows * (TC_KC / 4);
        auto issue_chunk = [&](int64_t chunk) {
            if (PLANAR) {
                // one thread: expect the bytes of the four planes on the slot's barrier, then one bulk copy per plane
                if (chunk < nchunks && pt == 0) {
                    const int slot = (int)(chunk % TC_RAW_SLOTS);
                    const uint32_t bar = smem_u32(&raw_full[slot]);
                    const uint32_t plane_bytes = z_rows * 16;
                    const uint32_t dst0 = raw_s + (uint32_t)slot * raw_bytes;
                    const float* src0 = p.Z + ((k_begin / TC_KC + chunk) * 4) * (int64_t)z_rows * 4;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4 * plane_bytes) : "memory");
#pragma unroll
                    for (int part = 0; part < 4; ++part)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst0 + (uint32_t)part * plane_stride), "l"(src0 + (int64_t)part * z_rows * 4), "r"(plane_bytes), "r"(bar)
                                     : "memory");
                }
                return;
            }
            if (chunk < nchunks) {
                const float* src0 = p.Z + k_begin + chunk * TC_KC;
                const uint32_t dst0 = raw_s + (uint32_t)(chunk % TC_RAW_SLOTS) * raw_bytes;
                for (int pc = pt; pc < npieces; pc += TC_PROD) {
                    const int row = pc >> 2, part = pc & 3;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)row * row_pitch + (uint32_t)part * piece_pitch),
                                 "l"(src0 + (int64_t)row * p.zpitch + part * 4) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue_chunk(0);
        issue_chunk(1);

        uint32_t acc_phase = 0;
        int s = 0, rs = 0;
        uint32_t ph = 0;
        int64_t in_window = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            if (!PLANAR) asm volatile("cp.async.wait_group 1;" ::: "memory");       // this thread's pieces of chunk c have landed
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");  // everyone's have; chunk c-1 is fully consumed
            issue_chunk(c + 2);                                          // reuses the slot chunk c-1 occupied
            if (PLANAR) mbar_wait(&raw_full[rs], (uint32_t)((c / TC_RAW_SLOTS) & 1));   // every reader waits: the bulk copies of chunk c landed
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);                 // first pass over the ring returns immediately
            __syncwarp();
            const uint32_t rb = raw_s + (uint32_t)rs * raw_bytes;
            const uint32_t sb = stage_s + (uint32_t)s * stage_bytes;
            {   // ---- U rows: all loads, then the products, then the stores (no load waits behind a store)
                float4 x0[U_NC], x1[U_NC], x2[U_NC], x3[U_NC];
#pragma unroll
                for (int cc = 0; cc < U_NC; ++cc) {
                    const uint32_t o = (uint32_t)(u_c0 + cc) * piece_pitch;
                    x0[cc] = lds128(rb + usrc[0] + o);
                    x1[cc] = lds128(rb + usrc[1] + o);
                    x2[cc] = lds128(rb + usrc[2] + o);
                    x3[cc] = lds128(rb + usrc[3] + o);
                }
#pragma unroll
                for (int cc = 0; cc < U_NC; ++cc) {
                    const float4 v = mul4(mul4(x0[cc], x1[cc]), mul4(x2[cc], x3[cc]));
                    const uint32_t d = sb + udst + (uint32_t)cc * (TC_M * 16);
                    store_split<SPLIT>(v, d, d + a_tile_bytes);
                }
            }
            if (pt < BN) {   // ---- V rows
                float4 y0[TC_KC / 4], y1[TC_KC / 4];
#pragma unroll
                for (int cc = 0; cc < TC_KC / 4; ++cc) {
                    y0[cc] = lds128(rb + vsrc[0] + (uint32_t)cc * piece_pitch);
                    y1[cc] = lds128(rb + vsrc[1] + (uint32_t)cc * piece_pitch);
                }
#pragma unroll
                for (int cc = 0; cc < TC_KC / 4; ++cc) {
                    const uint32_t d = sb + vdst + (uint32_t)cc * lbo_b;
                    store_split<SPLIT>(mul4(y0[cc], y1[cc]), d, d + b_tile_bytes);
                }
            }
            fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++rs == TC_RAW_SLOTS) rs = 0;
            // ---- flush: drain the accumulator into fp64 M
            const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
            if (last_of_window) {
                in_window = 0;
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                drain_accumulator(tmem_base, T * BN, BN, warp, lane, u0, nU, v0, p.nC, p.M, reinterpret_cast<float*>(stage_base), unscale, TC_M, (c + 1) == nchunks);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}


// ---- FP16 operand variant (mode 3).  Same pipeline as gram_tc_kernel<0, T, true>, with the operand tiles in fp16:
// a 16-byte piece of a tile row holds 8 samples instead of 4, a stage covers 32 samples (two tcgen05.mma kind::f16, K = 16), the
// raw-factor ring holds fp16 rows and the producers multiply with packed half2 arithmetic.  Per sample this halves every
// shared-memory transaction of the kernel -- raw-factor loads, tile stores, the tensor core's operand fetch -- which is what
// bounds the one-pass TF32 kernel (its shared-memory data pipe is ~90 % busy at 43 % of the TF32 MMA rate).  fp16 has the
// mantissa of tf32 (11 bits) but a narrow exponent: the factors are scaled to [0.5, 1) per factor as in the other modes, so
// entries below 6e-5 of a factor's largest lose relative precision and entries below 6e-8 vanish.  The mode exists for the
// sweep's exact refinement (TensorNetwork.refine = 'exact'), where the Gram is only a preconditioner; measured as one
// (tools/precond_experiment.py): the same conjugate-gradient iteration counts as TF32 operands.
constexpr int H_KC_DEFAULT = 64;    // samples per stage (four MMAs of K = 16 per tile); measured on the config-5a middle site, 262 144
                                    // rows: 701 TF/s against 590 with 32-sample stages (TN_TC16_KC=32), same bits -- the per-stage
                                    // barriers, fences and bulk-copy bookkeeping are paid half as often

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t hmul2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint4 hmul8(uint4 a, uint4 b) {
    return make_uint4(hmul2u(a.x, b.x), hmul2u(a.y, b.y), hmul2u(a.z, b.z), hmul2u(a.w, b.w));
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Instruction descriptor for kind::f16 with fp16 operands, fp32 accumulate, both operands K-major.
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int T, int H_KC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gram_tc16_kernel(TcParams p) {
    constexpr int H_NP = H_KC / 8;      // 16-byte pieces (8 samples) of a row per stage
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, NS = p.nstages;
    const int mA = p.mA, mB = p.mB, mC = p.mC;

    // ---- shared memory: [NS operand stages][TC_RAW_SLOTS raw-factor slots of four planes][mbarriers][tmem slot]
    constexpr uint32_t a_tile_bytes = TC_M * H_KC * 2;         // one U tile of one stage: [piece][row][16 B]
    const uint32_t b_tile_bytes = (uint32_t)BN * H_KC * 2;
    const uint32_t stage_bytes = T * a_tile_bytes + b_tile_bytes;
    const uint32_t z_rows = (uint32_t)(2 * mA + mB + mC);
    const uint32_t raw_rows = z_rows + 1;                       // + an all-zero row for the padding rows of the tiles
    const uint32_t plane_stride = ((raw_rows * 16 + 95) / 128) * 128 + 32;
    const uint32_t raw_bytes = H_NP * plane_stride;
    uint8_t* stage_base = smem_raw;
    uint8_t* raw_base = smem_raw + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)TC_RAW_SLOTS * raw_bytes);
    uint64_t* full = bars;              // [NS]
    uint64_t* empty = bars + NS;        // [NS]
    uint64_t* acc_full = bars + 2 * NS;
    uint64_t* acc_empty = bars + 2 * NS + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
    uint64_t* raw_full = bars + 2 * NS + 3;      // [TC_RAW_SLOTS]

    const uint32_t tmem_cols_needed = (uint32_t)(T * BN);
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], TC_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, TC_PROD_WARPS);
        for (int i = 0; i < TC_RAW_SLOTS; ++i) mbar_init(&raw_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TC_RAW_SLOTS * H_NP * 4; i += TC_THREADS) {      // the zero row of every plane of every slot (16 B each)
        const int slot = i / (H_NP * 4), e = i % (H_NP * 4);
        reinterpret_cast<uint32_t*>(raw_base + (size_t)slot * raw_bytes + (size_t)(e >> 2) * plane_stride + (size_t)z_rows * 16)[e & 3] = 0u;
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t k_begin = (int64_t)blockIdx.z * p.rows_per_split;          // multiples of H_KC; Z is zero padded
    const int64_t k_end = min(p.zpitch, k_begin + p.rows_per_split);
    const int64_t nchunks = (k_end > k_begin) ? (k_end - k_begin) / H_KC : 0;
    const int64_t chunks_per_flush = p.flush_rows / H_KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t u0 = (int64_t)blockIdx.x * (TC_M * T);
    const int v0 = blockIdx.y * BN;

    if (warp == 0) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc_f16(TC_M, BN);
            const uint32_t lbo_a = TC_M * 16, lbo_b = (uint32_t)BN * 16, sbo = 128;
            uint32_t acc_phase = 0;
            int s = 0;
            uint32_t ph = 0;
            int64_t in_window = 0;
            for (int64_t c = 0; c < nchunks; ++c) {
                const bool first_of_window = in_window == 0;
                if (first_of_window && c > 0) {
                    mbar_wait(acc_empty, acc_phase);   // accumulator drained by the producers
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * stage_bytes);
                const uint32_t b_base = sb + T * a_tile_bytes;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t a_base = sb + (uint32_t)t * a_tile_bytes;
                    const uint32_t d = tmem_base + (uint32_t)(t * BN);
#pragma unroll
                    for (int j = 0; j < H_KC / 16; ++j) {          // one MMA = 16 samples = two 16-byte pieces
                        const uint32_t ao = (uint32_t)(2 * j) * lbo_a, bo = (uint32_t)(2 * j) * lbo_b;
                        const uint32_t acc0 = (first_of_window && j == 0) ? 0u : 1u;
                        umma_f16(d, make_desc(a_base + ao, lbo_a, sbo), make_desc(b_base + bo, lbo_b, sbo), idesc, acc0);
                    }
                }
                umma_commit(&empty[s]);
                const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
                if (last_of_window) {
                    umma_commit(acc_full);
                    in_window = 0;
                }
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================== producers / epilogue ===============================
        const int pt = tid - 32;  // 0..255
        const double unscale = ldexp(1.0, 2 * (tc_exponent(p.amax[0]) + tc_exponent(p.amax[1]) + tc_exponent(p.amax[2])) +
                                              tc_exponent(p.amax[3]));
        const uint32_t raw_s = smem_u32(raw_base);
        const uint32_t zero_row = z_rows * 16;
        uint32_t usrc[4] = {zero_row, zero_row, zero_row, zero_row};   // byte offsets of w*fa[ia], fa[ja], fb[ib], fb[jb] in a plane
        int u_tile = 0, u_row = pt & 127, u_c0 = 0;
        constexpr int U_NC = (T == 2) ? H_NP : H_NP / 2;                // pieces per thread and stage
        if (T == 2) u_tile = pt >> 7;
        else u_c0 = (pt >> 7) * U_NC;
        {
            const int64_t gu = u0 + (int64_t)u_tile * TC_M + u_row;
            if (gu < nU) {
                const int qa = (int)(gu / p.nB), qb = (int)(gu - (int64_t)qa * p.nB);
                int ia, ja, ib, jb;
                pair_decode(qa, mA, ia, ja);
                pair_decode(qb, mB, ib, jb);
                usrc[0] = (uint32_t)ia * 16;
                usrc[1] = (uint32_t)(mA + ja) * 16;
                usrc[2] = (uint32_t)(2 * mA + ib) * 16;
                usrc[3] = (uint32_t)(2 * mA + jb) * 16;
            }
        }
        uint32_t vsrc[2] = {zero_row, zero_row};
        if (pt < BN && v0 + pt < p.nC) {
            int ic, jc;
            pair_decode(v0 + pt, mC, ic, jc);
            vsrc[0] = (uint32_t)(2 * mA + mB + ic) * 16;
            vsrc[1] = (uint32_t)(2 * mA + mB + jc) * 16;
        }
        const uint32_t udst = (uint32_t)u_tile * a_tile_bytes + (uint32_t)u_row * 16 + (uint32_t)u_c0 * (TC_M * 16);
        const uint32_t vdst = T * a_tile_bytes + (uint32_t)pt * 16;
        const uint32_t lbo_b = (uint32_t)BN * 16;
        const uint32_t stage_s = smem_u32(stage_base);
        const __half* Zh = reinterpret_cast<const __half*>(p.Z);

        auto issue_chunk = [&](int64_t chunk) {
            // one thread: expect the bytes of the four planes on the slot's barrier, then one bulk copy per plane
            if (chunk < nchunks && pt == 0) {
                const int slot = (int)(chunk % TC_RAW_SLOTS);
                const uint32_t bar = smem_u32(&raw_full[slot]);
                const uint32_t plane_bytes = z_rows * 16;
                const uint32_t dst0 = raw_s + (uint32_t)slot * raw_bytes;
                const __half* src0 = Zh + ((k_begin / H_KC + chunk) * H_NP) * (int64_t)z_rows * 8;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(H_NP * plane_bytes) : "memory");
#pragma unroll
                for (int part = 0; part < H_NP; ++part)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst0 + (uint32_t)part * plane_stride), "l"(src0 + (int64_t)part * z_rows * 8), "r"(plane_bytes), "r"(bar)
                                 : "memory");
            }
        };
        issue_chunk(0);
        issue_chunk(1);

        uint32_t acc_phase = 0;
        int s = 0, rs = 0;
        uint32_t ph = 0;
        int64_t in_window = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");  // chunk c-1 is fully consumed by every producer
            issue_chunk(c + 2);                                          // reuses the slot chunk c-1 occupied
            mbar_wait(&raw_full[rs], (uint32_t)((c / TC_RAW_SLOTS) & 1));   // the bulk copies of chunk c have landed
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);                 // first pass over the ring returns immediately
            __syncwarp();
            const uint32_t rb = raw_s + (uint32_t)rs * raw_bytes;
            const uint32_t sb = stage_s + (uint32_t)s * stage_bytes;
            // ---- U rows, four pieces at a time: all loads, then the products, then the stores
#pragma unroll
            for (int g0 = 0; g0 < U_NC; g0 += 4) {
                constexpr int G = (U_NC < 4) ? U_NC : 4;
                uint4 x0[G], x1[G], x2[G], x3[G];
#pragma unroll
                for (int cc = 0; cc < G; ++cc) {
                    const uint32_t o = (uint32_t)(u_c0 + g0 + cc) * plane_stride;
                    x0[cc] = lds128u(rb + usrc[0] + o);
                    x1[cc] = lds128u(rb + usrc[1] + o);
                    x2[cc] = lds128u(rb + usrc[2] + o);
                    x3[cc] = lds128u(rb + usrc[3] + o);
                }
#pragma unroll
                for (int cc = 0; cc < G; ++cc)
                    sts128u(sb + udst + (uint32_t)(g0 + cc) * (TC_M * 16), hmul8(hmul8(x0[cc], x1[cc]), hmul8(x2[cc], x3[cc])));
            }
            if (pt < BN) {   // ---- V rows, four pieces at a time
#pragma unroll
                for (int g0 = 0; g0 < H_NP; g0 += 4) {
                    uint4 y0[4], y1[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        y0[cc] = lds128u(rb + vsrc[0] + (uint32_t)(g0 + cc) * plane_stride);
                        y1[cc] = lds128u(rb + vsrc[1] + (uint32_t)(g0 + cc) * plane_stride);
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) sts128u(sb + vdst + (uint32_t)(g0 + cc) * lbo_b, hmul8(y0[cc], y1[cc]));
                }
            }
            fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++rs == TC_RAW_SLOTS) rs = 0;
            const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
            if (last_of_window) {
                in_window = 0;
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                drain_accumulator(tmem_base, T * BN, BN, warp, lane, u0, nU, v0, p.nC, p.M, reinterpret_cast<float*>(stage_base), unscale, TC_M, (c + 1) == nchunks);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---- fp16 kernel, run-ordered producers (T = 2 U tiles, 64-sample stages).
// The capture of gram_tc16_kernel (profiles/r2_ncu_gram_tc16_v1.txt) shows the shared-memory data pipe 73 % busy at 27 % tensor
// activity: per 64-sample stage 1409 wavefronts of producer loads, 598 of producer stores, 768 of tensor-core operand fetch.  The
// loads are the part that can shrink: a U entry is w fa[ia] fa[ja] fb[ib] * fb[jb], and along consecutive rows of a tile only jb
// changes (runs of mB - ib rows share the first three factors); a V entry is fc[ic] * fc[jc] with runs of mC - ic.  Here a producer
// thread owns ONE 16-byte piece (8 samples) of EIGHT consecutive tile rows (= one 8-row core matrix of the UMMA layout) instead
// of all pieces of one row: it builds the (at most two, else a slow path reloads per row) run prefixes of its rows once per stage
// and then spends one LDS.128, four HMUL2 and one STS.128 per tile piece -- 24 loads per thread and stage instead of 48.
// Bank conflicts (first version, profiles/r2_ncu_gram_tc16_run_v1.txt: lane = core matrix with rotated row order; the loads of a
// quarter warp crossed run boundaries and collided, 1281 load wavefronts instead of 768): the eight lanes of a quarter warp now
// own the EIGHT PIECES of the same rows, so every load instruction reads one raw-factor row in eight planes and every store one
// tile row in eight piece slabs; with the plane stride and the piece stride (the descriptors' LBO) both == 16 (mod 128) bytes
// the eight 16-byte accesses fall into eight different bank groups whatever the row pattern.
struct RunRows {            // per thread: its eight rows
    uint32_t key[8];        // the three prefix rows of a U row (10 bits each) / the fc[ic] row of a V row
    uint32_t last[8];       // byte offset of the row of the last factor (fb[jb] / fc[jc]) in a plane
    uint32_t keyA, keyB;    // the (at most) two distinct prefixes handled without reloading
    uint32_t selB;          // bit i: row i uses prefix B
    uint32_t slow;          // bit i: row i has a third prefix: reload it in the loop
};
__device__ __forceinline__ void run_rows_finish(RunRows& r) {
    r.keyA = r.key[0];
    r.keyB = r.key[0];
    r.selB = 0u;
    r.slow = 0u;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (r.key[i] != r.keyA && r.keyB == r.keyA) r.keyB = r.key[i];
        if (r.key[i] != r.keyA) {
            if (r.key[i] == r.keyB) r.selB |= 1u << i;
            else r.slow |= 1u << i;
        }
    }
}
__device__ __forceinline__ uint4 sel8(bool b, uint4 x, uint4 y) { return b ? x : y; }

__host__ __device__ __forceinline__ uint32_t run_plane_stride(uint32_t raw_rows) { return ((raw_rows * 16 + 127) / 128) * 128 + 16; }
constexpr uint32_t RUN_LBO_A = TC_M * 16 + 16;                   // bytes between two pieces of a U tile
constexpr uint32_t RUN_A_TILE = 8 * RUN_LBO_A;                   // one U tile of one stage

__global__ void __launch_bounds__(TC_THREADS, 1)
gram_tc16_run_kernel(TcParams p) {
    constexpr int T = 2, H_KC = 64, H_NP = 8;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, NS = p.nstages;
    const int mA = p.mA, mB = p.mB, mC = p.mC;

    const uint32_t lbo_b = (uint32_t)BN * 16 + 16;
    const uint32_t b_tile_bytes = 8 * lbo_b;
    const uint32_t stage_bytes = T * RUN_A_TILE + b_tile_bytes;
    const uint32_t z_rows = (uint32_t)(2 * mA + mB + mC);
    const uint32_t raw_rows = z_rows + 1;
    const uint32_t plane_stride = run_plane_stride(raw_rows);
    const uint32_t raw_bytes = H_NP * plane_stride;
    uint8_t* stage_base = smem_raw;
    uint8_t* raw_base = smem_raw + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)TC_RAW_SLOTS * raw_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + NS;
    uint64_t* acc_full = bars + 2 * NS;
    uint64_t* acc_empty = bars + 2 * NS + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
    uint64_t* raw_full = bars + 2 * NS + 3;

    const uint32_t tmem_cols_needed = (uint32_t)(T * BN);
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], TC_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, TC_PROD_WARPS);
        for (int i = 0; i < TC_RAW_SLOTS; ++i) mbar_init(&raw_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TC_RAW_SLOTS * H_NP * 4; i += TC_THREADS) {
        const int slot = i / (H_NP * 4), e = i % (H_NP * 4);
        reinterpret_cast<uint32_t*>(raw_base + (size_t)slot * raw_bytes + (size_t)(e >> 2) * plane_stride + (size_t)z_rows * 16)[e & 3] = 0u;
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t k_begin = (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t k_end = min(p.zpitch, k_begin + p.rows_per_split);
    const int64_t nchunks = (k_end > k_begin) ? (k_end - k_begin) / H_KC : 0;
    const int64_t chunks_per_flush = p.flush_rows / H_KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t u0 = (int64_t)blockIdx.x * (TC_M * T);
    const int v0 = blockIdx.y * BN;

    if (warp == 0) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc_f16(TC_M, BN);
            const uint32_t sbo = 128;
            uint32_t acc_phase = 0;
            int s = 0;
            uint32_t ph = 0;
            int64_t in_window = 0;
            for (int64_t c = 0; c < nchunks; ++c) {
                const bool first_of_window = in_window == 0;
                if (first_of_window && c > 0) {
                    mbar_wait(acc_empty, acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * stage_bytes);
                const uint32_t b_base = sb + T * RUN_A_TILE;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t a_base = sb + (uint32_t)t * RUN_A_TILE;
                    const uint32_t d = tmem_base + (uint32_t)(t * BN);
#pragma unroll
                    for (int j = 0; j < H_KC / 16; ++j) {
                        const uint32_t ao = (uint32_t)(2 * j) * RUN_LBO_A, bo = (uint32_t)(2 * j) * lbo_b;
                        const uint32_t acc0 = (first_of_window && j == 0) ? 0u : 1u;
                        if (!(p.dbg & 2)) umma_f16(d, make_desc(a_base + ao, RUN_LBO_A, sbo), make_desc(b_base + bo, lbo_b, sbo), idesc, acc0);
                    }
                }
                umma_commit(&empty[s]);
                const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
                if (last_of_window) {
                    umma_commit(acc_full);
                    in_window = 0;
                }
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================== producers / epilogue ===============================
        const int pt = tid - 32;              // 0..255
        const int pc = lane & 7;              // this lane's piece of the stage (8 samples): the 8 lanes of a quarter warp = the 8 pieces of one row
        const int grp = (pt >> 5) * 4 + (lane >> 3);     // 8-row core matrix: U rows 8 grp .. 8 grp + 7 of the CTA's 256, V rows likewise
        const double unscale = ldexp(1.0, 2 * (tc_exponent(p.amax[0]) + tc_exponent(p.amax[1]) + tc_exponent(p.amax[2])) +
                                              tc_exponent(p.amax[3]));
        const uint32_t raw_s = smem_u32(raw_base);
        RunRows ur, vr;
        const bool v_active = 8 * grp < BN;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gu = u0 + 8 * grp + i;
            uint32_t key = z_rows | (z_rows << 10) | (z_rows << 20), last = z_rows;      // the zero row
            if (gu < nU) {
                const int qa = (int)(gu / p.nB), qb = (int)(gu - (int64_t)qa * p.nB);
                int ia, ja, ib, jb;
                pair_decode(qa, mA, ia, ja);
                pair_decode(qb, mB, ib, jb);
                key = (uint32_t)ia | ((uint32_t)(mA + ja) << 10) | ((uint32_t)(2 * mA + ib) << 20);
                last = (uint32_t)(2 * mA + jb);
            }
            ur.key[i] = key;
            ur.last[i] = last * 16u;
            const int gv = v0 + 8 * grp + i;
            uint32_t vkey = z_rows, vlast = z_rows;
            if (v_active && gv < p.nC) {
                int ic, jc;
                pair_decode(gv, mC, ic, jc);
                vkey = (uint32_t)(2 * mA + mB + ic);
                vlast = (uint32_t)(2 * mA + mB + jc);
            }
            vr.key[i] = vkey;
            vr.last[i] = vlast * 16u;
        }
        run_rows_finish(ur);
        run_rows_finish(vr);
        // destinations: tile (8 grp) / 128, row (8 grp) % 128 + i; piece pc
        const uint32_t udst = (uint32_t)(grp >> 4) * RUN_A_TILE + (uint32_t)pc * RUN_LBO_A + (uint32_t)((8 * grp) & 127) * 16;
        const uint32_t vdst = T * RUN_A_TILE + (uint32_t)pc * lbo_b + (uint32_t)(8 * grp) * 16;
        const uint32_t stage_s = smem_u32(stage_base);
        const __half* Zh = reinterpret_cast<const __half*>(p.Z);

        auto issue_chunk = [&](int64_t chunk) {
            // one thread, ONE bulk copy: the pre-pass stores a chunk as the image of a slot (planes padded to the slot's plane stride)
            if (chunk < nchunks && pt == 0) {
                const int slot = (int)(chunk % TC_RAW_SLOTS);
                const uint32_t bar = smem_u32(&raw_full[slot]);
                const uint32_t dst0 = raw_s + (uint32_t)slot * raw_bytes;
                const __half* src0 = Zh + (k_begin / H_KC + chunk) * (int64_t)(raw_bytes / 2);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(raw_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst0), "l"(src0), "r"(raw_bytes), "r"(bar) : "memory");
            }
        };
        if (!(p.dbg & 8)) {
            issue_chunk(0);
            issue_chunk(1);
        }

        auto u_prefix = [&](uint32_t rbp, uint32_t key) -> uint4 {
            const uint4 a = lds128u(rbp + (key & 1023u) * 16u);
            const uint4 b = lds128u(rbp + ((key >> 10) & 1023u) * 16u);
            const uint4 c = lds128u(rbp + (key >> 20) * 16u);
            return hmul8(hmul8(a, b), c);
        };

        uint32_t acc_phase = 0;
        int s = 0, rs = 0;
        uint32_t ph = 0;
        int64_t in_window = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            if (!(p.dbg & 16)) asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");
            if (!(p.dbg & 8)) {
                issue_chunk(c + 2);
                mbar_wait(&raw_full[rs], (uint32_t)((c / TC_RAW_SLOTS) & 1));
            }
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
            __syncwarp();
            const uint32_t rbp = raw_s + (uint32_t)rs * raw_bytes + (uint32_t)pc * plane_stride;     // this lane's plane of the slot
            const uint32_t sb = stage_s + (uint32_t)s * stage_bytes;
            if (!(p.dbg & 1)) {
            {   // ---- U: prefixes, then eight rows
                const uint4 preA = u_prefix(rbp, ur.keyA);
                const uint4 preB = u_prefix(rbp, ur.keyB);
                uint4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = lds128u(rbp + ur.last[i]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 pre = sel8((ur.selB >> i) & 1u, preB, preA);
                    if ((ur.slow >> i) & 1u) pre = u_prefix(rbp, ur.key[i]);
                    sts128u(sb + udst + (uint32_t)i * 16u, hmul8(pre, x[i]));
                }
            }
            if (v_active) {   // ---- V
                const uint4 preA = lds128u(rbp + vr.keyA * 16u);
                const uint4 preB = lds128u(rbp + vr.keyB * 16u);
                uint4 y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = lds128u(rbp + vr.last[i]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 pre = sel8((vr.selB >> i) & 1u, preB, preA);
                    if ((vr.slow >> i) & 1u) pre = lds128u(rbp + vr.key[i] * 16u);
                    sts128u(sb + vdst + (uint32_t)i * 16u, hmul8(pre, y[i]));
                }
            }
            }
            if (!(p.dbg & 4)) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++rs == TC_RAW_SLOTS) rs = 0;
            const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
            if (last_of_window) {
                in_window = 0;
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                if (!(p.dbg & 32)) drain_accumulator(tmem_base, T * BN, BN, warp, lane, u0, nU, v0, p.nC, p.M, reinterpret_cast<float*>(stage_base), unscale, TC_M, (c + 1) == nchunks);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static size_t tc16_run_smem_bytes(int mA, int mB, int mC, int BN, int NS) {
    const size_t stage = 2 * (size_t)RUN_A_TILE + 8 * ((size_t)BN * 16 + 16);
    const size_t plane = run_plane_stride((uint32_t)(2 * mA + mB + mC + 1));
    return NS * stage + TC_RAW_SLOTS * 8 * plane + (2 * NS + 2) * 8 + 16 + 32;
}

// ---- fp16 kernel with a PRE-SYNTHESISED V operand and pair products of the left factor (T = 2 U tiles, 64-sample stages).
// Measured on gram_tc16_run_kernel (tools/tc16_probe.py with TN_TC16_DBG, profiles/r2_gram_tc16_where_the_time_goes.txt): the
// producers' synthesis alone takes 86 % of the kernel's time, and it is bound by shared-memory wavefronts (845 load + 600 store
// per stage at 90 % of the pipe); the MMAs cost 15 % on top (their operand fetch shares the pipe), the barriers 18 %.  40 % of the
// synthesis is the V tile pair(fc) -- which is THE SAME for all 1260 CTAs of a tile column.  So:
//  * a pre-pass writes pair(fc) once, in fp16, as ready images of the V region of a stage (K-major core matrices, piece slabs
//    padded like the shared-memory layout), and a loader warp drops a stage's image into place with ONE bulk copy that completes
//    on the stage's `full` barrier (arrive.expect_tx): no producer instruction and no LSU wavefront is spent on V;
//  * the same pre-pass writes pa[qa] = w fa[ia] fa[ja] per sample (fp32 product, one rounding), so the raw-factor ring of a CTA
//    holds only the fb rows and the (at most 256 / nB + 2) pa rows of its own tile rows: 4.5 KB per stage instead of 18.5 KB,
//    which frees the shared memory for a third operand stage (the V image needs about a microsecond from L2);
//  * a U entry is pa[qa] fb[ib] * fb[jb]: the run prefix is two loads and one product.
// Warp roles (352 threads): 0 = MMA issuer, 1..8 = U producers / drain, 9 = raw-ring loader, 10 = V-image loader.
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), the 16-byte chunk index of
// a row XORed with row % 8 by the hardware; LBO is not used by this mode (encoded as 1).  Tile bases are 1024-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct VimgParams {
    const unsigned long long* amax;
    const __half* Zc;        // [chunk][8 planes][psc / 2]: rows fb[0..mB) and the zero row, 8 samples per row and plane
    const __half* Zpa;       // [chunk][nA][64]
    const __half* Vimg;      // [tile column][chunk][8 pieces][lbo_b / 2]
    int64_t zchunks;         // chunks of 64 samples in the staged data
    int64_t rows_per_split;  // multiple of 64
    double* M;
    int mA, mB, mC;
    int nA, nB, nC;
    int BN, nstages, flush_rows, nq_max;
    int red_policy;          // L2 policy of the flush's reductions: 1 = evict_last, 0 = evict_first
    int stream_keep;         // 1: L2 evict_last hint on the V-image copies
    int sw128;               // 1: operand tiles in the 128-byte-swizzled K-major layout (rows of 128 B = the 64 samples of a stage, 16-byte chunk index XOR row % 8)
    int dbg;                 // TN_TC16_DBG (measurement only, wrong results): 1 = no U synthesis, 2 = no MMAs, 8 = no raw ring, 32 = no drain, 64 = no V copies
};
constexpr int VI_THREADS = 32 * 11;

// one block = one 64-sample chunk
__global__ void __launch_bounds__(256)
tc16_vimg_stage_kernel(TcFactor fa, TcFactor fb, TcFactor fc, const double* __restrict__ w, int64_t rows, const unsigned long long* __restrict__ amax,
                       __half* __restrict__ Zc, __half* __restrict__ Zpa, __half* __restrict__ Vimg, int nA, int nC, int BN, int ytiles,
                       int64_t zchunks, uint32_t psc, uint32_t lbo_b, int sw128) {
    extern __shared__ float vs[];
    const int mA = fa.m, mB = fb.m, mC = fc.m;
    constexpr int LD = 65;
    float* sWA = vs;                       // [mA][LD]  w * fa (scaled)
    float* sA = sWA + mA * LD;             // [mA][LD]
    float* sB = sA + mA * LD;              // [mB][LD]
    float* sC = sB + mB * LD;              // [mC][LD]
    short* tA = reinterpret_cast<short*>(sC + mC * LD);      // [nA][2]
    short* tC = tA + 2 * nA;                                  // [nC][2]
    const double sa = ldexp(1.0, -tc_exponent(amax[0])), sb = ldexp(1.0, -tc_exponent(amax[1]));
    const double sc = ldexp(1.0, -tc_exponent(amax[2])), sw = ldexp(1.0, -tc_exponent(amax[3]));
    const int tid = threadIdx.x;
    const int64_t chunk = blockIdx.x;
    const int64_t s0 = chunk * 64;
    const int msum = mA + mB + mC;
    for (int idx = tid; idx < 64 * msum; idx += 256) {        // lane = feature: coalesced along a sample's row
        const int s = idx / msum, i = idx - s * msum;
        const int64_t row = s0 + s;
        if (i < mA) {
            float v = 0.f, vw = 0.f;
            if (row < rows) {
                const double x = map_eval(fa.map_kind, fa.ptr + (fa.div == 1 ? row : row / fa.div) * fa.ld, i) * sa;
                v = (float)x;
                vw = (float)(x * ((w ? w[row] : 1.0) * sw));
            }
            sA[i * LD + s] = v;
            sWA[i * LD + s] = vw;
        } else if (i < mA + mB) {
            const int il = i - mA;
            sB[il * LD + s] = (row < rows) ? (float)(map_eval(fb.map_kind, fb.ptr + (fb.div == 1 ? row : row / fb.div) * fb.ld, il) * sb) : 0.f;
        } else {
            const int il = i - mA - mB;
            sC[il * LD + s] = (row < rows) ? (float)(map_eval(fc.map_kind, fc.ptr + (fc.div == 1 ? row : row / fc.div) * fc.ld, il) * sc) : 0.f;
        }
    }
    for (int q = tid; q < nA + nC; q += 256) {
        int i, j;
        if (q < nA) { pair_decode(q, mA, i, j); tA[2 * q] = (short)i; tA[2 * q + 1] = (short)j; }
        else { pair_decode(q - nA, mC, i, j); tC[2 * (q - nA)] = (short)i; tC[2 * (q - nA) + 1] = (short)j; }
    }
    __syncthreads();
    auto pack8 = [](const float* x, const float* y) -> uint4 {       // eight products, one rounding each
        uint4 o;
        __half2 h0 = __floats2half2_rn(x[0] * y[0], x[1] * y[1]), h1 = __floats2half2_rn(x[2] * y[2], x[3] * y[3]);
        __half2 h2 = __floats2half2_rn(x[4] * y[4], x[5] * y[5]), h3 = __floats2half2_rn(x[6] * y[6], x[7] * y[7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
        return o;
    };
    // common rows: fb as fp16, 8 samples per plane
    for (int it = tid; it < mB * 8; it += 256) {
        const int r = it >> 3, pc = it & 7;
        const float* x = sB + r * LD + pc * 8;
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        *reinterpret_cast<uint4*>(Zc + (chunk * 8 + pc) * (int64_t)(psc / 2) + r * 8) = pack8(x, one);
    }
    // pa[qa] = (w fa[ia]) fa[ja]
    for (int it = tid; it < nA * 8; it += 256) {
        const int qa = it >> 3, pc = it & 7;
        *reinterpret_cast<uint4*>(Zpa + (chunk * nA + qa) * 64 + pc * 8) = pack8(sWA + tA[2 * qa] * LD + pc * 8, sA + tA[2 * qa + 1] * LD + pc * 8);
    }
    // V images: pair(fc) in the layout of a stage's V region
    for (int it = tid; it < ytiles * BN * 8; it += 256) {
        const int pc = it & 7, rr = it >> 3;
        const int y = rr / BN, r = rr - y * BN;
        const int gv = y * BN + r;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (gv < nC) o = pack8(sC + tC[2 * gv] * LD + pc * 8, sC + tC[2 * gv + 1] * LD + pc * 8);
        if (sw128)      // image of BN rows x 128 B, 8-row groups of 1024 B, chunk ^= row % 8
            *reinterpret_cast<uint4*>(Vimg + ((int64_t)y * zchunks + chunk) * (int64_t)(BN * 64) + (r >> 3) * 512 + (r & 7) * 64 + ((pc ^ (r & 7)) * 8)) = o;
        else
            *reinterpret_cast<uint4*>(Vimg + (((int64_t)y * zchunks + chunk) * 8 + pc) * (int64_t)(lbo_b / 2) + r * 8) = o;
    }
}

__global__ void __launch_bounds__(VI_THREADS, 1)
gram_tc16_vimg_kernel(VimgParams p) {
    constexpr int T = 2, H_KC = 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, NS = p.nstages;
    const int mB = p.mB;

    const bool sw = p.sw128 != 0;
    const uint32_t lbo_b = (uint32_t)BN * 16 + 16;
    const uint32_t v_bytes = sw ? (uint32_t)BN * 128 : 8 * lbo_b;
    const uint32_t a_tile = sw ? (uint32_t)TC_M * 128 : RUN_A_TILE;
    const uint32_t stage_bytes = T * a_tile + v_bytes;
    const uint32_t psc = run_plane_stride((uint32_t)mB + 1);       // plane of the common rows (fb + the zero row)
    const uint32_t common_bytes = 8 * psc;
    const uint32_t slot_bytes = common_bytes + (uint32_t)p.nq_max * 128;
    uint8_t* stage_base = smem_raw;
    uint8_t* raw_base = smem_raw + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)TC_RAW_SLOTS * slot_bytes);
    uint64_t* full = bars;              // [NS]  8 producer warps + the V loader's expect_tx
    uint64_t* empty = bars + NS;        // [NS]
    uint64_t* acc_full = bars + 2 * NS;
    uint64_t* acc_empty = bars + 2 * NS + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
    uint64_t* raw_full = bars + 2 * NS + 3;            // [TC_RAW_SLOTS]
    uint64_t* raw_empty = raw_full + TC_RAW_SLOTS;     // [TC_RAW_SLOTS]

    const uint32_t tmem_cols_needed = (uint32_t)(T * BN);
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], TC_PROD_WARPS + 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, TC_PROD_WARPS);
        for (int i = 0; i < TC_RAW_SLOTS; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], TC_PROD_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t k_begin = (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t k_end = min(p.zchunks * H_KC, k_begin + p.rows_per_split);
    const int64_t nchunks = (k_end > k_begin) ? (k_end - k_begin) / H_KC : 0;
    const int64_t chunk0 = k_begin / H_KC;
    const int64_t chunks_per_flush = p.flush_rows / H_KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t u0 = (int64_t)blockIdx.x * (TC_M * T);
    const int v0 = blockIdx.y * BN;
    const int qa_lo = (int)(u0 / p.nB);
    const int64_t u_last = min(u0 + TC_M * T, nU) - 1;
    const int nq = (int)(u_last / p.nB) - qa_lo + 1;               // pa rows this CTA needs (<= nq_max)

    if (warp == 0) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc_f16(TC_M, BN);
            const uint32_t sbo = 128;
            uint32_t acc_phase = 0;
            int s = 0;
            uint32_t ph = 0;
            int64_t in_window = 0;
            for (int64_t c = 0; c < nchunks; ++c) {
                const bool first_of_window = in_window == 0;
                if (first_of_window && c > 0) {
                    mbar_wait(acc_empty, acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * stage_bytes);
                const uint32_t b_base = sb + T * a_tile;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t a_base = sb + (uint32_t)t * a_tile;
                    const uint32_t d = tmem_base + (uint32_t)(t * BN);
#pragma unroll
                    for (int j = 0; j < H_KC / 16; ++j) {
                        const uint32_t acc0 = (first_of_window && j == 0) ? 0u : 1u;
                        if (p.dbg & 2) continue;
                        if (sw) {       // K step j = 32 bytes further inside the 128-byte rows
                            umma_f16(d, make_desc_sw128(a_base + 32u * j), make_desc_sw128(b_base + 32u * j), idesc, acc0);
                        } else {
                            const uint32_t ao = (uint32_t)(2 * j) * RUN_LBO_A, bo = (uint32_t)(2 * j) * lbo_b;
                            umma_f16(d, make_desc(a_base + ao, RUN_LBO_A, sbo), make_desc(b_base + bo, lbo_b, sbo), idesc, acc0);
                        }
                    }
                }
                umma_commit(&empty[s]);
                const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
                if (last_of_window) {
                    umma_commit(acc_full);
                    in_window = 0;
                }
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 + TC_PROD_WARPS) {
        // =============================== raw-ring loader: fb rows + this CTA's pa rows of every chunk ===============================
        if (lane == 0 && !(p.dbg & 8)) {
            const uint32_t raw_s = smem_u32(raw_base);
            const uint32_t pa_bytes = (uint32_t)nq * 128;
            for (int64_t c = 0; c < nchunks; ++c) {
                const int slot = (int)(c % TC_RAW_SLOTS);
                if (c >= TC_RAW_SLOTS) mbar_wait(&raw_empty[slot], (uint32_t)((c / TC_RAW_SLOTS - 1) & 1));
                const uint32_t bar = smem_u32(&raw_full[slot]);
                const uint32_t dst0 = raw_s + (uint32_t)slot * slot_bytes;
                const __half* srcc = p.Zc + (chunk0 + c) * (int64_t)(common_bytes / 2);
                const __half* srcp = p.Zpa + ((chunk0 + c) * p.nA + qa_lo) * 64;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(common_bytes + pa_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst0), "l"(srcc), "r"(common_bytes), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst0 + common_bytes), "l"(srcp), "r"(pa_bytes), "r"(bar) : "memory");
            }
        }
    } else if (warp == 2 + TC_PROD_WARPS) {
        // =============================== V loader: the ready image of pair(fc) for every chunk, straight into the stage ===============================
        if (lane == 0) {
            const uint32_t stage_s = smem_u32(stage_base);
            const uint64_t keep_pol = l2_evict_last_policy();
            for (int64_t c = 0; c < nchunks; ++c) {
                const int s = (int)(c % NS);
                if (c >= NS) mbar_wait(&empty[s], (uint32_t)((c / NS - 1) & 1));      // the MMAs of chunk c - NS have read the stage
                const uint32_t bar = smem_u32(&full[s]);
                const __half* src = p.Vimg + ((int64_t)blockIdx.y * p.zchunks + chunk0 + c) * (int64_t)(v_bytes / 2);
                if (p.dbg & 64) { mbar_arrive(&full[s]); continue; }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(v_bytes) : "memory");
                if (p.stream_keep)      // the image is read by every CTA of the tile column: ask the L2 to keep it
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                 ::"r"(stage_s + (uint32_t)s * stage_bytes + T * a_tile), "l"(src), "r"(v_bytes), "r"(bar), "l"(keep_pol) : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(stage_s + (uint32_t)s * stage_bytes + T * a_tile), "l"(src), "r"(v_bytes), "r"(bar) : "memory");
            }
        }
    } else {
        // =============================== U producers / epilogue ===============================
        const int pt = tid - 32;              // 0..255
        const int pc = lane & 7;              // this lane's piece of the stage (8 samples)
        const int grp = (pt >> 5) * 4 + (lane >> 3);     // 8-row core matrix: U rows 8 grp .. 8 grp + 7 of the CTA's 256
        const double unscale = ldexp(1.0, 2 * (tc_exponent(p.amax[0]) + tc_exponent(p.amax[1]) + tc_exponent(p.amax[2])) +
                                              tc_exponent(p.amax[3]));
        const uint32_t raw_s = smem_u32(raw_base);
        RunRows ur;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gu = u0 + 8 * grp + i;
            uint32_t key = (uint32_t)mB, last = (uint32_t)mB;      // padding row: pa row 0 times the zero row
            if (gu < nU) {
                const int qa = (int)(gu / p.nB), qb = (int)(gu - (int64_t)qa * p.nB);
                int ib, jb;
                pair_decode(qb, mB, ib, jb);
                key = ((uint32_t)(qa - qa_lo) << 10) | (uint32_t)ib;
                last = (uint32_t)jb;
            }
            ur.key[i] = key;
            ur.last[i] = last * 16u;
        }
        run_rows_finish(ur);
        // padded slabs: tile, piece slab, row;  swizzled: tile, 8-row group of 1024 B (row i of it at i * 128, chunk pc ^ i)
        const uint32_t udst = sw ? (uint32_t)(grp >> 4) * a_tile + (uint32_t)(grp & 15) * 1024
                                 : (uint32_t)(grp >> 4) * RUN_A_TILE + (uint32_t)pc * RUN_LBO_A + (uint32_t)((8 * grp) & 127) * 16;
        const uint32_t stage_s = smem_u32(stage_base);
        auto u_prefix = [&](uint32_t slot, uint32_t key) -> uint4 {
            const uint4 a = lds128u(slot + common_bytes + (key >> 10) * 128u + (uint32_t)pc * 16u);      // pa[qa]
            const uint4 b = lds128u(slot + (uint32_t)pc * psc + (key & 1023u) * 16u);                    // fb[ib]
            return hmul8(a, b);
        };

        uint32_t acc_phase = 0;
        int s = 0, rs = 0;
        uint32_t ph = 0;
        int64_t in_window = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            if (!(p.dbg & 8)) mbar_wait(&raw_full[rs], (uint32_t)((c / TC_RAW_SLOTS) & 1));
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);                 // first pass over the ring returns immediately
            __syncwarp();
            const uint32_t slot = raw_s + (uint32_t)rs * slot_bytes;
            const uint32_t rbp = slot + (uint32_t)pc * psc;              // this lane's plane of the common rows
            const uint32_t sb = stage_s + (uint32_t)s * stage_bytes;
            if (!(p.dbg & 1)) {
                uint4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = lds128u(rbp + ur.last[i]);
                const uint4 preA = u_prefix(slot, ur.keyA);
                const uint4 preB = u_prefix(slot, ur.keyB);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 pre = sel8((ur.selB >> i) & 1u, preB, preA);
                    if ((ur.slow >> i) & 1u) pre = u_prefix(slot, ur.key[i]);
                    const uint32_t off = sw ? (uint32_t)i * 128u + (uint32_t)((pc ^ i) * 16) : (uint32_t)i * 16u;
                    sts128u(sb + udst + off, hmul8(pre, x[i]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&full[s]);
                mbar_arrive(&raw_empty[rs]);
            }
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++rs == TC_RAW_SLOTS) rs = 0;
            const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
            if (last_of_window) {
                in_window = 0;
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                // transpose scratch: the U regions of stages 0 and 1 (the V regions belong to the loader, which may already be filling them)
                float* scratch = reinterpret_cast<float*>(stage_base + (size_t)((warp - 1) >> 2) * stage_bytes) - (size_t)(((warp - 1) >> 2) * 4) * (32 * 33);
                if (!(p.dbg & 32)) drain_accumulator(tmem_base, T * BN, BN, warp, lane, u0, nU, v0, p.nC, p.M, scratch, unscale, TC_M, (c + 1) == nchunks, 2, p.red_policy);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
                asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");      // nobody writes the next U tiles over a scratch still in use
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static size_t tc16_vimg_smem_bytes(int mB, int BN, int nq_max, int NS) {
    const size_t stage = 2 * (size_t)RUN_A_TILE + 8 * ((size_t)BN * 16 + 16);
    const size_t slot = 8 * (size_t)run_plane_stride((uint32_t)mB + 1) + (size_t)nq_max * 128;
    return NS * stage + TC_RAW_SLOTS * slot + (2 * NS + 2) * 8 + 16 + 2 * TC_RAW_SLOTS * 8 + 16;
}

static size_t tc16_smem_bytes(int mA, int mB, int mC, int BN, int T, int NS, int kc) {
    const size_t stage = (size_t)T * TC_M * kc * 2 + (size_t)BN * kc * 2;
    const size_t plane = (((size_t)(2 * mA + mB + mC + 1) * 16 + 95) / 128) * 128 + 32;
    return NS * stage + TC_RAW_SLOTS * (kc / 8) * plane + (2 * NS + 2) * 8 + 16 + 32;
}

// ---- CTA-pair variant (cta_group::2) for the large shapes (T == 2, BN == 256).
// Two CTAs of a cluster (the two SMs of a TPC) execute every MMA together as M = 256 x N = 256: each CTA synthesises its own
// 128 rows of the two U tiles and only HALF of the V tile (128 of the 256 rows of pair(fc)); the tensor cores of both SMs read
// the two halves of V from both shared memories.  Per SM this removes a third of the operand bytes the MMAs fetch from shared
// memory and a quarter of what the producers write -- the 1-CTA kernel is shared-memory-bandwidth bound (each K = 8 step moves
// ~104 KB through a 128 B/clk port in the time the six MMAs need).  The leader CTA (cluster rank 0) issues all MMAs; the
// producers of both CTAs arrive on the leader's full / acc_empty barriers (cluster-scope release), tcgen05.commit multicasts
// the empty / acc_full arrivals to both CTAs.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    mbar_timeout();
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int TP_BN = 256;          // V tile of the pair
constexpr int TP_BH = TP_BN / 2;    // rows of it each CTA synthesises
constexpr int TP_T = 2;             // U tiles (M = 256 each: 128 rows per CTA) per pair

template <int SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gram_tc_pair_kernel(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int NS = p.nstages;
    const int mA = p.mA, mB = p.mB, mC = p.mC;
    const uint32_t rank = cluster_rank();

    constexpr uint32_t a_tile_bytes = TC_M * TC_KC * 4;
    constexpr uint32_t b_tile_bytes = TP_BH * TC_KC * 4;
    constexpr uint32_t stage_bytes = 2 * TP_T * a_tile_bytes + 2 * b_tile_bytes;
    const uint32_t z_rows = (uint32_t)(2 * mA + mB + mC);
    const uint32_t raw_rows = z_rows + 1;
    const uint32_t raw_bytes = raw_rows * TC_KCP * 4;
    uint8_t* stage_base = smem_raw;
    uint8_t* raw_base = smem_raw + (size_t)NS * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)TC_RAW_SLOTS * raw_bytes);
    uint64_t* full = bars;              // [NS]  used in the leader: producers of both CTAs -> MMA
    uint64_t* empty = bars + NS;        // [NS]  per CTA: multicast tcgen05.commit -> producers
    uint64_t* acc_full = bars + 2 * NS;     // per CTA (multicast commit)
    uint64_t* acc_empty = bars + 2 * NS + 1;   // used in the leader: drainers of both CTAs -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 2 * TC_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 2 * TC_PROD_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TC_RAW_SLOTS * TC_KCP; i += TC_THREADS)
        reinterpret_cast<float*>(raw_base + (size_t)(i / TC_KCP) * raw_bytes)[z_rows * TC_KCP + (i % TC_KCP)] = 0.f;
    cluster_sync_all();                  // barrier inits visible to the peer before anything arrives remotely
    if (warp == 0) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t k_begin = (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t k_end = min(p.zpitch, k_begin + p.rows_per_split);
    const int64_t nchunks = (k_end > k_begin) ? (k_end - k_begin) / TC_KC : 0;
    const int64_t chunks_per_flush = p.flush_rows / TC_KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t u0_pair = (int64_t)(blockIdx.x >> 1) * (2 * TC_M * TP_T);     // 512 U rows per pair
    const int v0 = blockIdx.y * TP_BN;

    if (warp == 0) {
        // =============================== MMA issuer (leader CTA only) ===============================
        if (rank == 0 && lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc(2 * TC_M, TP_BN);
            constexpr uint32_t lbo_a = TC_M * 16, lbo_b = TP_BH * 16, sbo = 128;
            uint32_t acc_phase = 0;
            int s = 0;
            uint32_t ph = 0;
            int64_t in_window = 0;
            for (int64_t c = 0; c < nchunks; ++c) {
                const bool first_of_window = in_window == 0;
                if (first_of_window && c > 0) {
                    mbar_wait_cluster(acc_empty, acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                mbar_wait_cluster(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * stage_bytes);
                const uint32_t b_hi = sb + 2 * TP_T * a_tile_bytes;
                const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
                for (int t = 0; t < TP_T; ++t) {
                    const uint32_t a_hi = sb + (uint32_t)t * 2 * a_tile_bytes;
                    const uint32_t a_lo = a_hi + a_tile_bytes;
                    const uint32_t d = tmem_base + (uint32_t)(t * TP_BN);
#pragma unroll
                    for (int j = 0; j < TC_KC / 8; ++j) {
                        const uint32_t ao = (uint32_t)(2 * j) * lbo_a, bo = (uint32_t)(2 * j) * lbo_b;
                        const uint32_t acc0 = (first_of_window && j == 0) ? 0u : 1u;
                        umma2_tf32(d, make_desc(a_hi + ao, lbo_a, sbo), make_desc(b_hi + bo, lbo_b, sbo), idesc, acc0);
                        if (SPLIT) {
                            umma2_tf32(d, make_desc(a_hi + ao, lbo_a, sbo), make_desc(b_lo + bo, lbo_b, sbo), idesc, 1u);
                            umma2_tf32(d, make_desc(a_lo + ao, lbo_a, sbo), make_desc(b_hi + bo, lbo_b, sbo), idesc, 1u);
                        }
                    }
                }
                umma_commit2(&empty[s]);
                const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
                if (last_of_window) {
                    umma_commit2(acc_full);
                    in_window = 0;
                }
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================== producers / epilogue (both CTAs) ===============================
        const int pt = tid - 32;
        const double unscale = ldexp(1.0, 2 * (tc_exponent(p.amax[0]) + tc_exponent(p.amax[1]) + tc_exponent(p.amax[2])) +
                                              tc_exponent(p.amax[3]));
        const uint32_t raw_s = smem_u32(raw_base);
        const uint32_t zero_row = z_rows * TC_KCP * 4;
        uint32_t usrc[4] = {zero_row, zero_row, zero_row, zero_row};
        const int u_tile = pt >> 7, u_row = pt & 127;
        constexpr int U_NC = TC_KC / 4;
        const int64_t u0_cta = u0_pair + (int64_t)rank * TC_M;        // tile t of this CTA covers rows u0_cta + t*256 + [0, 128)
        {
            const int64_t gu = u0_cta + (int64_t)u_tile * (2 * TC_M) + u_row;
            if (gu < nU) {
                const int qa = (int)(gu / p.nB), qb = (int)(gu - (int64_t)qa * p.nB);
                int ia, ja, ib, jb;
                pair_decode(qa, mA, ia, ja);
                pair_decode(qb, mB, ib, jb);
                usrc[0] = (uint32_t)(ia * TC_KCP * 4);
                usrc[1] = (uint32_t)((mA + ja) * TC_KCP * 4);
                usrc[2] = (uint32_t)((2 * mA + ib) * TC_KCP * 4);
                usrc[3] = (uint32_t)((2 * mA + jb) * TC_KCP * 4);
            }
        }
        uint32_t vsrc[2] = {zero_row, zero_row};
        const int gv = v0 + (int)rank * TP_BH + pt;                    // this CTA's half of the V tile
        if (pt < TP_BH && gv < p.nC) {
            int ic, jc;
            pair_decode(gv, mC, ic, jc);
            vsrc[0] = (uint32_t)((2 * mA + mB + ic) * TC_KCP * 4);
            vsrc[1] = (uint32_t)((2 * mA + mB + jc) * TC_KCP * 4);
        }
        const uint32_t udst = (uint32_t)u_tile * 2 * a_tile_bytes + (uint32_t)u_row * 16;
        const uint32_t vdst = 2 * TP_T * a_tile_bytes + (uint32_t)pt * 16;
        constexpr uint32_t lbo_b = TP_BH * 16;
        const uint32_t stage_s = smem_u32(stage_base);

        const int npieces = (int)z_rows * (TC_KC / 4);
        auto issue_chunk = [&](int64_t chunk) {
            if (chunk < nchunks) {
                const float* src0 = p.Z + k_begin + chunk * TC_KC;
                const uint32_t dst0 = raw_s + (uint32_t)(chunk % TC_RAW_SLOTS) * raw_bytes;
                for (int pc = pt; pc < npieces; pc += TC_PROD) {
                    const int row = pc >> 2, part = pc & 3;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)row * (TC_KCP * 4) + part * 16),
                                 "l"(src0 + (int64_t)row * p.zpitch + part * 4) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue_chunk(0);
        issue_chunk(1);

        uint32_t acc_phase = 0;
        int s = 0, rs = 0;
        uint32_t ph = 0;
        int64_t in_window = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");
            issue_chunk(c + 2);
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
            __syncwarp();
            const uint32_t rb = raw_s + (uint32_t)rs * raw_bytes;
            const uint32_t sb = stage_s + (uint32_t)s * stage_bytes;
            {
                float4 x0[U_NC], x1[U_NC], x2[U_NC], x3[U_NC];
#pragma unroll
                for (int cc = 0; cc < U_NC; ++cc) {
                    const uint32_t o = (uint32_t)cc * 16;
                    x0[cc] = lds128(rb + usrc[0] + o);
                    x1[cc] = lds128(rb + usrc[1] + o);
                    x2[cc] = lds128(rb + usrc[2] + o);
                    x3[cc] = lds128(rb + usrc[3] + o);
                }
#pragma unroll
                for (int cc = 0; cc < U_NC; ++cc) {
                    const float4 v = mul4(mul4(x0[cc], x1[cc]), mul4(x2[cc], x3[cc]));
                    const uint32_t d = sb + udst + (uint32_t)cc * (TC_M * 16);
                    store_split<SPLIT>(v, d, d + a_tile_bytes);
                }
            }
            if (pt < TP_BH) {
                float4 y0[TC_KC / 4], y1[TC_KC / 4];
#pragma unroll
                for (int cc = 0; cc < TC_KC / 4; ++cc) {
                    y0[cc] = lds128(rb + vsrc[0] + cc * 16);
                    y1[cc] = lds128(rb + vsrc[1] + cc * 16);
                }
#pragma unroll
                for (int cc = 0; cc < TC_KC / 4; ++cc) {
                    const uint32_t d = sb + vdst + (uint32_t)cc * lbo_b;
                    store_split<SPLIT>(mul4(y0[cc], y1[cc]), d, d + b_tile_bytes);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&full[s], 0);           // on the leader's barrier
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++rs == TC_RAW_SLOTS) rs = 0;
            const bool last_of_window = (++in_window == chunks_per_flush) || (c + 1) == nchunks;
            if (last_of_window) {
                in_window = 0;
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                drain_accumulator(tmem_base, TP_T * TP_BN, TP_BN, warp, lane, u0_cta, nU, v0, p.nC, p.M,
                                  reinterpret_cast<float*>(stage_base), unscale, 2 * TC_M, (c + 1) == nchunks);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty, 0);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();                  // the peer may still be reading this CTA's shared / tensor memory until here
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

static size_t tc_pair_smem_bytes(int mA, int mB, int mC, int NS) {
    const size_t stage = 2 * (size_t)TP_T * TC_M * TC_KC * 4 + 2 * (size_t)TP_BH * TC_KC * 4;
    const size_t raw = (size_t)(2 * mA + mB + mC + 1) * TC_KCP * 4;
    return NS * stage + TC_RAW_SLOTS * raw + (2 * NS + 2) * 8 + 16;
}

static size_t tc_smem_bytes(int mA, int mB, int mC, int BN, int T, int NS) {
    const size_t stage = 2 * (size_t)T * TC_M * TC_KC * 4 + 2 * (size_t)BN * TC_KC * 4;
    const size_t raw = (size_t)(2 * mA + mB + mC + 1) * TC_KCP * 4;
    return NS * stage + TC_RAW_SLOTS * raw + (2 * NS + 2) * 8 + 16;
}

}  // namespace tn

namespace tn { static thread_local int g_tc_flush_rows = 0; }

// Rows accumulated in the fp32 TMEM accumulators between two fp64 flushes, for the calling thread's later tn_gram_kr3 calls
// (0 = the default, 2048).  Longer windows are faster and less accurate (relative error of M ~ 6e-9 per row of the window): the
// exact refinement of the sweep (tn_cg on the fp64 operator) only needs M as a preconditioner and uses 8192.
extern "C" int tn_gram_tc_flush_rows(int rows) {
    const int prev = tn::g_tc_flush_rows;
    if (rows >= 0) tn::g_tc_flush_rows = (rows / tn::TC_KC) * tn::TC_KC;
    return prev;
}

int tn_gram_kr3_tc(int mode, const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows,
                   double* M, int accumulate, void* stream) {
    using namespace tn;
    cudaStream_t st = as_stream(stream);
    const TcFactor A{fa->ptr, fa->ld, fa->m, fa->div < 1 ? 1 : fa->div, fa->map_kind};
    const TcFactor B{fb->ptr, fb->ld, fb->m, fb->div < 1 ? 1 : fb->div, fb->map_kind};
    const TcFactor C{fc->ptr, fc->ld, fc->m, fc->div < 1 ? 1 : fc->div, fc->map_kind};
    TcParams p;
    p.mA = A.m; p.mB = B.m; p.mC = C.m;
    p.rows = rows;
    p.M = M;
    p.nA = npairs(A.m);
    p.nB = npairs(B.m);
    p.nC = npairs(C.m);
    const bool f16 = (mode == 3);
    int kc16 = H_KC_DEFAULT;
    if (const char* e = getenv("TN_TC16_KC")) kc16 = (atoi(e) == 32) ? 32 : 64;
    const int KC = f16 ? kc16 : TC_KC;             // samples per pipeline stage
    p.split = (mode == 2) ? 1 : 0;
    p.dbg = getenv("TN_TC16_DBG") ? atoi(getenv("TN_TC16_DBG")) : 0;
    p.planar = 0;
    p.flush_rows = (tn::g_tc_flush_rows > 0) ? tn::g_tc_flush_rows : TC_FLUSH_ROWS_DEFAULT;
    if (const char* e = getenv("TN_TC_FLUSH_ROWS")) {
        const int v = atoi(e);
        if (v >= KC) p.flush_rows = v;
    }
    p.flush_rows = (p.flush_rows / KC) * KC;
    if (p.flush_rows < KC) p.flush_rows = KC;
    const int64_t nU = (int64_t)p.nA * p.nB;
    const int64_t n = nU * p.nC;
    {   // V tiles of equal width: nC = 300 (config 3) becomes 2 x 160 columns instead of 2 x 256 (41 % of them padding)
        const int ntile = (p.nC + 255) / 256;
        p.BN = (((p.nC + ntile - 1) / ntile + 15) / 16) * 16;
    }
    p.T = (nU > TC_M && p.BN * 2 <= 512) ? 2 : 1;
    auto smem_of = [&](int T_, int NS_) { return f16 ? tc16_smem_bytes(A.m, B.m, C.m, p.BN, T_, NS_, kc16) : tc_smem_bytes(A.m, B.m, C.m, p.BN, T_, NS_); };
    int NS = 4;
    while (NS >= 2 && smem_of(p.T, NS) > 226 * 1024) --NS;
    if (NS < 2 && p.T == 2) {
        p.T = 1;
        NS = 4;
        while (NS >= 2 && smem_of(p.T, NS) > 226 * 1024) --NS;
    }
    TN_CHECK_ARG(NS >= 2, "tn_gram_kr3 (tensor-core modes): factor sizes %d+%d+%d do not fit the shared-memory pipeline", A.m, B.m, C.m);
    p.nstages = NS;
    const size_t smem = smem_of(p.T, NS);
    if (!accumulate) TN_CUDA(cudaMemsetAsync(M, 0, (size_t)n * sizeof(double), st));
    if (rows == 0) return TN_OK;

    // ---- pre-pass: Z = [w*fa | fa | fb | fc] feature-major fp32, rows zero padded to a multiple of TC_KC
    p.zpitch = ceil_div64(rows, KC) * KC;
    const int z_rows = 2 * A.m + B.m + C.m;
    {   // keep freed scratch in the stream-ordered pool: the default threshold (0) returns it to the OS at every sync,
        // which made each call pay a fresh multi-hundred-MB allocation
        static std::atomic<unsigned long long> pool_ready{0};     // one bit per device
        int dev = 0;
        TN_CUDA(cudaGetDevice(&dev));
        if (!((pool_ready.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
            cudaMemPool_t pool;
            TN_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            unsigned long long keep = ~0ull;
            TN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_ready.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
        }
    }
    // fp16, two U tiles, 64-sample stages: pre-synthesised V images + pa rows (gram_tc16_vimg_kernel; TN_TC16_VIMG=0 selects the older kernels)
    if (f16 && kc16 == 64 && p.T == 2 && B.m + 1 < 1023 && !(getenv("TN_TC16_VIMG") && atoi(getenv("TN_TC16_VIMG")) == 0)) {
        const int nq_max = 255 / p.nB + 2;
        int NSV = 4;
        while (NSV >= 2 && tc16_vimg_smem_bytes(B.m, p.BN, nq_max, NSV) > 226 * 1024) --NSV;
        // (A CTA-pair variant of this kernel -- cta_group::2, half of the V image per CTA, four stages -- produced bit-identical M and
        // was 10 % slower on the config-5a site, 23 % on config 5b: profiles/r2_gram_tc16_vimg_probe.jsonl; removed.)
        if (NSV >= 2) {
            const int64_t zchunks = p.zpitch / 64;
            const int ytiles = (int)ceil_div64(p.nC, p.BN);
            const uint32_t psc = run_plane_stride((uint32_t)B.m + 1);
            const uint32_t lbo_b = (uint32_t)p.BN * 16 + 16;
            const size_t zc_bytes = (size_t)zchunks * 8 * psc;
            const size_t zpa_bytes = (size_t)zchunks * p.nA * 128;
            const size_t vimg_bytes = (size_t)ytiles * zchunks * 8 * lbo_b;
            AsyncScratch scratch;
            TN_CUDA(scratch.alloc(zc_bytes + zpa_bytes + vimg_bytes + 64, st));
            char* buf = static_cast<char*>(scratch.ptr);
            TN_CUDA(cudaMemsetAsync(buf, 0, zc_bytes, st));
            unsigned long long* amaxv = reinterpret_cast<unsigned long long*>(buf + zc_bytes + zpa_bytes + vimg_bytes);
            TN_CUDA(cudaMemsetAsync(amaxv, 0, 4 * sizeof(unsigned long long), st));
            int64_t blocks = ceil_div64(rows * (A.m + B.m + C.m + 1), 256 * 8);
            if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
            if (blocks < 1) blocks = 1;
            tc_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, B, C, w, rows, amaxv);
            TN_LAUNCH_CHECK();
            VimgParams v;
            v.amax = amaxv;
            v.Zc = reinterpret_cast<const __half*>(buf);
            v.Zpa = reinterpret_cast<const __half*>(buf + zc_bytes);
            v.Vimg = reinterpret_cast<const __half*>(buf + zc_bytes + zpa_bytes);
            v.zchunks = zchunks;
            v.M = M;
            v.mA = A.m; v.mB = B.m; v.mC = C.m;
            v.nA = p.nA; v.nB = p.nB; v.nC = p.nC;
            v.BN = p.BN; v.nstages = NSV; v.flush_rows = p.flush_rows; v.nq_max = nq_max; v.dbg = p.dbg;
            v.sw128 = (getenv("TN_TC16_SW128") && atoi(getenv("TN_TC16_SW128")) != 0) ? 1 : 0;
            v.red_policy = getenv("TN_TC16_RED_POLICY") ? atoi(getenv("TN_TC16_RED_POLICY")) : 1;
            v.stream_keep = getenv("TN_TC16_STREAM_KEEP") ? atoi(getenv("TN_TC16_STREAM_KEEP")) : 1;      // measured: DRAM reads 53.7 -> 39.2 GB per 131 072 rows (profiles/r2_traffic.json)
            const size_t ssmem = (size_t)(2 * A.m + B.m + C.m) * 65 * sizeof(float) + (size_t)(p.nA + p.nC) * 2 * sizeof(short);
            TN_CHECK_ARG(ssmem <= 200 * 1024, "tn_gram_kr3 (fp16): factors %d+%d+%d too wide for the staging pass", A.m, B.m, C.m);
            TN_SMEM(tc16_vimg_stage_kernel, ssmem);
            tc16_vimg_stage_kernel<<<(unsigned)zchunks, 256, ssmem, st>>>(A, B, C, w, rows, amaxv, reinterpret_cast<__half*>(buf),
                                                                       reinterpret_cast<__half*>(buf + zc_bytes), reinterpret_cast<__half*>(buf + zc_bytes + zpa_bytes),
                                                                       p.nA, p.nC, p.BN, ytiles, zchunks, psc, lbo_b, v.sw128);
            TN_LAUNCH_CHECK();
            const int64_t gxv = ceil_div64(nU, (int64_t)TC_M * 2), gyv = ytiles;
            int64_t ksv = ceil_div64((int64_t)sm_count(), gxv * gyv);
            const int64_t max_ksv = ceil_div64(p.zpitch, 4 * 64);
            if (ksv > max_ksv) ksv = max_ksv;
            if (ksv < 1) ksv = 1;
            if (ksv > 65535) ksv = 65535;
            v.rows_per_split = ceil_div64(ceil_div64(p.zpitch, ksv), 64) * 64;
            ksv = ceil_div64(p.zpitch, v.rows_per_split);
            TN_CHECK_ARG(gyv <= 65535 && gxv <= 0x7fffffff, "tn_gram_kr3: grid too large");
            {
                const size_t vsmem = tc16_vimg_smem_bytes(B.m, p.BN, nq_max, NSV);
                TN_SMEM(gram_tc16_vimg_kernel, vsmem);
                dim3 gridv((unsigned)gxv, (unsigned)gyv, (unsigned)ksv);
                gram_tc16_vimg_kernel<<<gridv, VI_THREADS, vsmem, st>>>(v);
            }
            TN_LAUNCH_CHECK();
            return TN_OK;
        }
    }
    // fp16, two U tiles, 64-sample stages: the run-ordered kernel (TN_TC16_RUN=0 keeps the row-per-thread kernel)
    int NSR = 4;
    while (NSR >= 2 && tc16_run_smem_bytes(A.m, B.m, C.m, p.BN, NSR) > 226 * 1024) --NSR;
    const bool use_run = f16 && kc16 == 64 && p.T == 2 && z_rows < 1023 && NSR >= 2 && getenv("TN_TC16_RUN") && atoi(getenv("TN_TC16_RUN")) != 0;
    const size_t run_plane = run_plane_stride((uint32_t)(z_rows + 1));
    AsyncScratch zscratch;
    float* Z = nullptr;
    const size_t z_bytes = use_run ? (size_t)(p.zpitch / 8) * run_plane : (size_t)z_rows * p.zpitch * (f16 ? sizeof(__half) : sizeof(float));
    TN_CUDA(zscratch.alloc(z_bytes + 64, st));
    Z = static_cast<float*>(zscratch.ptr);
    if (use_run) TN_CUDA(cudaMemsetAsync(Z, 0, z_bytes, st));
    unsigned long long* amax = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(Z) + z_bytes);
    TN_CUDA(cudaMemsetAsync(amax, 0, 4 * sizeof(unsigned long long), st));
    p.Z = Z;
    p.amax = amax;
    {
        int64_t blocks = ceil_div64(rows * (A.m + B.m + C.m + 1), 256 * 8);
        if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
        if (blocks < 1) blocks = 1;
        tc_absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, B, C, w, rows, amax);
        TN_LAUNCH_CHECK();
        dim3 grid((unsigned)ceil_div64(p.zpitch, 32), (unsigned)ceil_div64(A.m + B.m + C.m, 32));
        // planar raw slots + bulk-copy fill: the default since round 2 (measured on the config-5a middle site, 131 072 rows:
        // 588 vs 575 TF/s issued, M bit-identical; profiles/r2_tc_variants.txt); TN_TC_RAW_ROWMAJOR=1 selects the cp.async ring.
        // Needs the four planes to fit the slot pitch of the row-major layout (80 B per Z row) and 32 more bytes of shared
        // memory for the slots' barriers; the CTA-pair kernel stays row-major.
        p.planar = (!getenv("TN_TC_RAW_ROWMAJOR") && !getenv("TN_TC_PAIR") &&
                    4 * ((((size_t)(z_rows + 1) * 16 + 95) / 128) * 128 + 32) <= (size_t)(z_rows + 1) * TC_KCP * 4 &&
                    smem + 32 <= 227 * 1024) ? 1 : 0;
        if (f16) p.planar = use_run ? 3 : 2;       // planes of eight fp16 samples (3: padded to the run kernel's slot image)
        tc_stage_kernel<<<grid, 256, 0, st>>>(A, B, C, w, rows, Z, p.zpitch, amax, p.planar, (int64_t)(run_plane / 2));
        TN_LAUNCH_CHECK();
    }
    // CTA-pair kernel (opt-in, TN_TC_PAIR=1).  Measured on the config-5a middle site (131 072 rows, tools/tc_pair_probe.py):
    // bit-identical M, but 530 vs 557 TF/s issued in 3xTF32 and 217 vs 268 in TF32 -- the 1-CTA kernel already runs at ~0.9 of
    // the sustained (power-capped) tensor rate, so halving the operand traffic buys nothing and the cluster-scope barriers cost.
    bool pair = (!f16 && p.T == 2 && p.BN == TP_BN && nU >= 8LL * 512 && getenv("TN_TC_PAIR") && !getenv("TN_TC_NO_PAIR"));
    int NSP = 5;
    if (pair) {
        while (NSP >= 2 && tc_pair_smem_bytes(A.m, B.m, C.m, NSP) > 226 * 1024) --NSP;
        if (NSP < 2) pair = false;
    }
    if (pair) {
        p.nstages = NSP;
        const size_t psmem = tc_pair_smem_bytes(A.m, B.m, C.m, NSP);
        const int64_t gxp = 2 * ceil_div64(nU, 2LL * TC_M * TP_T), gyp = ceil_div64(p.nC, TP_BN);
        TN_CHECK_ARG(gyp <= 65535 && gxp <= 0x7fffffff, "tn_gram_kr3: grid too large");
        int64_t ksp = ceil_div64((int64_t)sm_count(), gxp * gyp);
        const int64_t max_ksp = ceil_div64(p.zpitch, 4 * TC_KC);
        if (ksp > max_ksp) ksp = max_ksp;
        if (ksp < 1) ksp = 1;
        if (ksp > 65535) ksp = 65535;
        p.rows_per_split = ceil_div64(ceil_div64(p.zpitch, ksp), TC_KC) * TC_KC;
        ksp = ceil_div64(p.zpitch, p.rows_per_split);
        using KernP = void (*)(TcParams);
        static const KernP pk[2] = {gram_tc_pair_kernel<0>, gram_tc_pair_kernel<1>};
        TN_SMEM(pk[p.split], psmem);
        dim3 pgrid((unsigned)gxp, (unsigned)gyp, (unsigned)ksp);
        pk[p.split]<<<pgrid, TC_THREADS, psmem, st>>>(p);
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    const int64_t gx = ceil_div64(nU, (int64_t)TC_M * p.T), gy = ceil_div64(p.nC, p.BN);
    TN_CHECK_ARG(gy <= 65535 && gx <= 0x7fffffff, "tn_gram_kr3: grid too large");
    // split the rows when there are too few tiles to fill the machine (flushes are atomic adds, so splits compose)
    int64_t ks = ceil_div64((int64_t)sm_count(), gx * gy);
    const int64_t max_ks = ceil_div64(p.zpitch, 4 * KC);
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    if (ks > 65535) ks = 65535;
    p.rows_per_split = ceil_div64(ceil_div64(p.zpitch, ks), KC) * KC;
    ks = ceil_div64(p.zpitch, p.rows_per_split);
    if (f16) {
        using Kern16 = void (*)(TcParams);
        if (use_run) {
            p.nstages = NSR;
            const size_t rsmem = tc16_run_smem_bytes(A.m, B.m, C.m, p.BN, NSR);
            dim3 gridr((unsigned)gx, (unsigned)gy, (unsigned)ks);
            TN_SMEM(gram_tc16_run_kernel, rsmem);
            gram_tc16_run_kernel<<<gridr, TC_THREADS, rsmem, st>>>(p);
            TN_LAUNCH_CHECK();
            return TN_OK;
        }
        Kern16 k16 = (kc16 == 64) ? ((p.T == 2) ? gram_tc16_kernel<2, 64> : gram_tc16_kernel<1, 64>)
                                  : ((p.T == 2) ? gram_tc16_kernel<2, 32> : gram_tc16_kernel<1, 32>);
        TN_SMEM(k16, smem);
        dim3 grid16((unsigned)gx, (unsigned)gy, (unsigned)ks);
        k16<<<grid16, TC_THREADS, smem, st>>>(p);
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    using Kern = void (*)(TcParams);
    static const Kern kerns[2][2][2] = {{{gram_tc_kernel<0, 1, false>, gram_tc_kernel<0, 2, false>},
                                         {gram_tc_kernel<1, 1, false>, gram_tc_kernel<1, 2, false>}},
                                        {{gram_tc_kernel<0, 1, true>, gram_tc_kernel<0, 2, true>},
                                         {gram_tc_kernel<1, 1, true>, gram_tc_kernel<1, 2, true>}}};
    const int planar = p.planar;
    const size_t smem_k = smem + (planar ? 32 : 0);          // + the three barriers of the raw slots
    Kern k = kerns[planar][p.split][p.T - 1];
    TN_SMEM(k, smem_k);
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)ks);
    k<<<grid, TC_THREADS, smem_k, st>>>(p);
    TN_LAUNCH_CHECK();
    return TN_OK;
}
