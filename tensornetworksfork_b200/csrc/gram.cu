// Gram build without the Jacobian: fp64 path (FP64 tensor pipe, DMMA m8n8k4), right-hand side, sigma, expansion.
//
// Replaces TensorNetwork.get_A_b (reference tensor/network.py:174-217).  The local Jacobian
// of a core is a row-wise Kronecker product J[row, (ia,ib,ic)] = fa[ia] fb[ib] fc[ic], so
//   A[(i),(j)] = sum_rows w * fa[ia]fa[ja] * fb[ib]fb[jb] * fc[ic]fc[jc]
// is invariant under ia<->ja, ib<->jb, ic<->jc separately.  Only the unique entries
//   M[qa,qb,qc] = sum_rows w * pair(fa)[qa] * pair(fb)[qb] * pair(fc)[qc]
// are accumulated (about P^2/8 of them instead of P^2/2 for a plain SYRK) as one GEMM
//   M = U^T V,  U[row,(qa,qb)] = w*pair(fa)*pair(fb),  V[row,qc] = pair(fc)
// whose operand tiles are synthesised in shared memory from the raw factors; nothing of
// size rows x P ever exists.  tn_gram_expand scatters M to the dense scaled system.
#include <stdlib.h>
#include "common.cuh"

namespace tn {

constexpr int GR_TU = 128;      // U columns per CTA
constexpr int GR_TV = 64;       // V columns per CTA
constexpr int GR_KC = 32;       // rows per staged chunk
constexpr int GR_THREADS = 256;
constexpr int GR_SU = GR_TU + 4;   // padded tile rows: conflict-free DMMA fragment reads (row offset 4 doubles per k)
constexpr int GR_SV = GR_TV + 4;

struct FactorDev {
    const double* ptr;
    int64_t ld;
    int m;
    int div;
    int map_kind;
};

static inline FactorDev to_dev(const tn_factor* f) { return FactorDev{f->ptr, f->ld, f->m, f->div < 1 ? 1 : f->div, f->map_kind}; }

__device__ __forceinline__ void stage_factor(const FactorDev& f, double* dst, int st, int64_t kb, int64_t k_end, int tid) {
    for (int idx = tid; idx < GR_KC * f.m; idx += GR_THREADS) {
        const int k = idx / f.m, i = idx - k * f.m;
        const int64_t row = kb + k;
        double v = 0.0;
        if (row < k_end) v = map_eval(f.map_kind, f.ptr + (row / f.div) * f.ld, i);
        dst[k * st + i] = v;
    }
}

// MODE 1: Gram of a Kronecker Jacobian (pairs of each factor).  MODE 0: right-hand side (plain products).
// MODE 2: Gram of a Jacobian whose column i is fa[t1[i]]*fb[t2[i]]*fc[t3[i]] (index tables; no Kronecker structure
//         assumed -- the cum-sum train, reference layers.py:408-477): out is the full P x P matrix.
// MODE 3: as 2 with V = 1 (one column): out = J^T w, the right-hand side.
template <int MODE>
__global__ void __launch_bounds__(GR_THREADS)
kr3_f64_kernel(FactorDev fa, FactorDev fb, FactorDev fc, const double* __restrict__ w, int64_t rows,
               double* __restrict__ out, int nA, int nB, int nC, int64_t rows_per_split,
               const int* __restrict__ t1, const int* __restrict__ t2, const int* __restrict__ t3) {
    constexpr bool PAIR = (MODE == 1);
    extern __shared__ double sm[];
    const int stA = fa.m | 1, stB = fb.m | 1, stC = fc.m | 1;
    double* sFA = sm;
    double* sFB = sFA + GR_KC * stA;
    double* sFC = sFB + GR_KC * stB;
    double* sW = sFC + GR_KC * stC;
    double* sU = sW + GR_KC;              // [GR_KC][GR_SU]
    double* sV = sU + GR_KC * GR_SU;      // [GR_KC][GR_SV]
    short* tIA = reinterpret_cast<short*>(sV + GR_KC * GR_SV);
    short* tJA = tIA + GR_TU;
    short* tIB = tJA + GR_TU;
    short* tJB = tIB + GR_TU;
    short* tIC = tJB + GR_TU;
    short* tJC = tIC + GR_TV;
    short* tKC = tJC + GR_TV;

    const int tid = threadIdx.x;
    const int64_t nU = (int64_t)nA * nB;
    const int64_t u0 = (int64_t)blockIdx.x * GR_TU;
    const int v0 = blockIdx.y * GR_TV;
    const int64_t k_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t k_end = min(rows, k_begin + rows_per_split);

    if (tid < GR_TU) {
        const int64_t gu = u0 + tid;
        int ia = -1, ja = 0, ib = 0, jb = 0;
        if (MODE >= 2) {
            if (gu < nU) { ia = t1[gu]; ib = t2[gu]; ja = t3[gu]; }
        } else if (gu < nU) {
            const int qa = (int)(gu / nB), qb = (int)(gu - (int64_t)qa * nB);
            if (PAIR) {
                pair_decode(qa, fa.m, ia, ja);
                pair_decode(qb, fb.m, ib, jb);
            } else {
                ia = ja = qa;
                ib = jb = qb;
            }
        }
        tIA[tid] = (short)ia; tJA[tid] = (short)ja; tIB[tid] = (short)ib; tJB[tid] = (short)jb;
    } else if (tid < GR_TU + GR_TV) {
        const int t = tid - GR_TU;
        const int gv = v0 + t;
        int ic = -1, jc = 0, kc = 0;
        if (MODE == 2) {
            if (gv < nC) { ic = t1[gv]; jc = t2[gv]; kc = t3[gv]; }
        } else if (MODE == 3) {
            ic = (gv == 0) ? 0 : -1;
        } else if (gv < nC) {
            if (PAIR) pair_decode(gv, fc.m, ic, jc);
            else ic = jc = gv;
        }
        tIC[t] = (short)ic; tJC[t] = (short)jc; tKC[t] = (short)kc;
    }

    // 8 warps as 4 (u) x 2 (v): each warp owns a 32 x 32 block of the tile = 4 x 4 DMMA m8n8k4 accumulators
    const int lane = tid & 31, warp = tid >> 5;
    const int wu = (warp >> 1) * 32, wv = (warp & 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int64_t kb = k_begin; kb < k_end; kb += GR_KC) {
        __syncthreads();  // previous chunk's sU/sV fully consumed; tables visible on first pass
        stage_factor(fa, sFA, stA, kb, k_end, tid);
        stage_factor(fb, sFB, stB, kb, k_end, tid);
        stage_factor(fc, sFC, stC, kb, k_end, tid);
        if (tid < GR_KC) {
            const int64_t row = kb + tid;
            sW[tid] = (row < k_end) ? (w ? w[row] : 1.0) : 0.0;
        }
        __syncthreads();
        for (int idx = tid; idx < GR_KC * GR_TU; idx += GR_THREADS) {
            const int k = idx >> 7, u = idx & (GR_TU - 1);
            const int ia = tIA[u];
            double v = 0.0;
            if (ia >= 0) {
                const double* a = sFA + k * stA;
                const double* b = sFB + k * stB;
                if (MODE >= 2) v = sW[k] * a[ia] * b[tIB[u]] * sFC[k * stC + tJA[u]];
                else v = PAIR ? sW[k] * a[ia] * a[tJA[u]] * b[tIB[u]] * b[tJB[u]] : sW[k] * a[ia] * b[tIB[u]];
            }
            sU[k * GR_SU + u] = v;
        }
        for (int idx = tid; idx < GR_KC * GR_TV; idx += GR_THREADS) {
            const int k = idx >> 6, t = idx & (GR_TV - 1);
            const int ic = tIC[t];
            double v = 0.0;
            if (ic >= 0) {
                const double* c = sFC + k * stC;
                if (MODE == 2) v = sFA[k * stA + ic] * sFB[k * stB + tJC[t]] * c[tKC[t]];
                else if (MODE == 3) v = 1.0;
                else v = PAIR ? c[ic] * c[tJC[t]] : c[ic];
            }
            sV[k * GR_SV + t] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < GR_KC; k4 += 4) {
            // A fragment = U^T: a[m = u][k];  B fragment: b[k][n = v]
            double af[4], bf[4];
            const double* up = sU + (k4 + fk) * GR_SU + wu + fr;
            const double* vp = sV + (k4 + fk) * GR_SV + wv + fr;
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = up[i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = vp[j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }

    double* o = out + (int64_t)blockIdx.z * nU * nC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gu = u0 + wu + i * 8 + fr;
        if (gu >= nU) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gv = v0 + wv + j * 8 + 2 * fk + e;
                if (gv < nC) o[gu * nC + gv] = acc[i][j][e];
            }
        }
    }
}

// ---- fp64 Gram of a Kronecker Jacobian with a FACTORED left operand (MODE 1 of the kernel above: same 128 x 64 tile, same warp
// layout, same DMMA fragments; results differ from it by the rounding of the operand products only).
// The capture of kr3_f64_kernel<1> (profiles/r2_ncu_kr3_f64.txt, hot lines by tools/ncu_hotspots.py) has the FP64 tensor pipe 42 %
// active and almost no stall samples in the DMMA loop: 45 % of them sit in the staging loop (one exposed global load and two integer
// divisions per element), 25 % in the synthesis of the 32 x 128 U tile (four gathers and four multiplies per element), 9 % in the
// three block barriers of a chunk.  Here
//  (a) U = (w * pair(fa)) (x) pair(fb) is never formed: the tile keeps the two factors' pair columns ([PA | PB | V], TAW + TBW + 64
//      columns per row instead of 128 + 64) and a DMMA A-fragment element is one product PA[k][la] * PB[k][lb] made in registers.  The
//      128 U columns of a CTA are a 16 x 8 box of (qa, qb) when the middle factor has many pairs (BOX: 24 operand columns), else
//      128 consecutive (qa, qb) indices (all pairs of the middle factor, <= 129 of the left one);
//  (b) the raw factors of chunk c+2 arrive by cp.async (8 bytes per element, 16 where the factor's rows are 16-byte aligned -- measured
//      equal --, zero fill past the end of the split, no divisions, a warp per row) while chunk c+1 is synthesised and chunk c multiplied; a feature-mapped factor (one raw value per row) is
//      loaded into a register at the top of the iteration and expanded at its end by one warp;
//  (c) one block barrier per 32-row chunk: raw buffers and operand tiles are double buffered, so the synthesis of the next chunk by
//      fast warps overlaps the DMMAs of slow ones.
__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(double* dst, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
// identity factor: the m values of one row (zero fill when the row is past the end); a lane per column pair when the factor's rows are
// 16-byte aligned in global memory (vec: base aligned, even row stride -- the environments of even bond dimension), else a lane per column
__device__ __forceinline__ void gf_fill_row(const FactorDev& f, double* dst, int64_t row, bool ok, int lane, bool vec) {
    if (f.map_kind != TN_MAP_IDENTITY) return;
    const double* src = f.ptr + (f.div == 1 ? row : row / f.div) * f.ld;
    if (vec) {
        for (int i = 2 * lane; i + 1 < f.m; i += 64) cp_async16(dst + i, src + i, ok);
        if ((f.m & 1) && lane == 31) cp_async8(dst + f.m - 1, src + f.m - 1, ok);
    } else {
        for (int i = lane; i < f.m; i += 32) cp_async8(dst + i, src + i, ok);
    }
}
// feature-mapped factor: lane k of an owning warp holds the raw value of row k and writes every second of its m features
__device__ __forceinline__ double gf_load_raw(const FactorDev& f, int64_t row) {
    return f.ptr[(f.div == 1 ? row : row / f.div) * f.ld];
}
__device__ __forceinline__ void gf_store_mapped(const FactorDev& f, double* dst, double raw, bool ok, int first) {
    for (int i = first; i < f.m; i += 2) dst[i] = ok ? map_apply(f.map_kind, raw, i) : 0.0;
}

template <bool BOX>
__global__ void __launch_bounds__(GR_THREADS, 2)
gram_f64_fact_kernel(FactorDev fa, FactorDev fb, FactorDev fc, const double* __restrict__ w, int64_t rows, double* __restrict__ out,
                     int nA, int nB, int nC, int64_t rows_per_split, int TAW, int TBW, int stT, int nTB, int swapU, int vecmask) {
    extern __shared__ __align__(16) double smf[];
    // even row strides: every row of a raw factor starts on a 16-byte boundary (16-byte cp.async); the synthesis gathers columns of ONE row
    // per instruction, so the stride does not enter its bank pattern
    const int stA = (fa.m + 1) & ~1, stB = (fb.m + 1) & ~1, stC = (fc.m + 1) & ~1;
    const int rawsz = GR_KC * (stA + stB + stC + 1);                      // one raw buffer: [FA | FB | FC | W]
    double* sRaw = smf;                                                    // [2][rawsz]
    double* sT = sRaw + 2 * rawsz;                                        // [2][GR_KC][stT]: rows of [PA | PB | V]
    unsigned* tCol = reinterpret_cast<unsigned*>(sT + 2 * GR_KC * stT);   // [NCOL]: i | j << 16 of the column's index pair
    const int NCOL = TAW + TBW + GR_TV;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nU = (int64_t)nA * nB;
    const int v0 = blockIdx.y * GR_TV;
    const int64_t k_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t k_end = min(rows, k_begin + rows_per_split);
    int64_t u0 = 0;
    int a0, b0 = 0;
    if (BOX) {
        const int ta = blockIdx.x / nTB;
        a0 = ta * 16;
        b0 = (blockIdx.x - ta * nTB) * 8;
    } else {
        u0 = (int64_t)blockIdx.x * GR_TU;
        a0 = (int)(u0 / nB);
    }

    for (int col = tid; col < NCOL; col += GR_THREADS) {
        int i = 0, j = 0;
        bool valid;
        if (col < TAW) {
            const int q = a0 + col;
            valid = q < nA;
            if (valid) pair_decode(q, fa.m, i, j);
        } else if (col < TAW + TBW) {
            const int q = b0 + col - TAW;
            valid = q < nB;
            if (valid) pair_decode(q, fb.m, i, j);
        } else {
            const int q = v0 + col - TAW - TBW;
            valid = q < nC;
            if (valid) pair_decode(q, fc.m, i, j);
        }
        tCol[col] = valid ? ((unsigned)i | ((unsigned)j << 16)) : 0xFFFFFFFFu;
    }

    // 8 warps as 4 (u) x 2 (v): each warp owns a 32 x 32 block of the tile = 4 x 4 DMMA m8n8k4 accumulators
    const int wu = (warp >> 1) * 32, wv = (warp & 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    int offA[4], offB[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int jc = wu + i * 8 + fr;
        int la = 0, lb = 0;
        if (BOX) {
            la = jc >> 3;
            lb = jc & 7;
        } else {
            const int64_t gu = u0 + jc;
            if (gu < nU) {
                const int qa = (int)(gu / nB);
                la = qa - a0;
                lb = (int)(gu - (int64_t)qa * nB);
            }
        }
        offA[i] = la;
        offB[i] = TAW + lb;
    }
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nch = (k_end > k_begin) ? (int)((k_end - k_begin + GR_KC - 1) / GR_KC) : 0;
    const int64_t last_row = rows - 1;
    // warps 1 / 4, 2 / 5, 3 / 6 expand the even / odd features of a feature-mapped fa, fb, fc (lane = row of the chunk)
    const int mw = (warp >= 1 && warp <= 6) ? (warp - 1) % 3 : -1, m_par = (warp >= 4) ? 1 : 0;
    FactorDev fm;
    fm.ptr = (mw == 0) ? fa.ptr : (mw == 1) ? fb.ptr : fc.ptr;
    fm.ld = (mw == 0) ? fa.ld : (mw == 1) ? fb.ld : fc.ld;
    fm.m = (mw == 0) ? fa.m : (mw == 1) ? fb.m : fc.m;
    fm.div = (mw == 0) ? fa.div : (mw == 1) ? fb.div : fc.div;
    fm.map_kind = (mw == 0) ? fa.map_kind : (mw == 1) ? fb.map_kind : fc.map_kind;
    const bool mapper = mw >= 0 && fm.map_kind != TN_MAP_IDENTITY;
    const int m_off = (mw == 0) ? 0 : (mw == 1) ? GR_KC * stA : GR_KC * (stA + stB);
    const int m_st = (mw == 0) ? stA : (mw == 1) ? stB : stC;

    auto fill = [&](int c, double* raw) {          // asynchronous part of the raw factors of chunk c
        const int64_t kb = k_begin + (int64_t)c * GR_KC;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = warp + 8 * r;
            const int64_t row = kb + k;
            const bool ok = row < k_end;
            const int64_t rc = ok ? row : last_row;
            gf_fill_row(fa, raw + k * stA, rc, ok, lane, vecmask & 1);
            gf_fill_row(fb, raw + GR_KC * stA + k * stB, rc, ok, lane, vecmask & 2);
            gf_fill_row(fc, raw + GR_KC * (stA + stB) + k * stC, rc, ok, lane, vecmask & 4);
        }
        if (warp == 0) {
            const int64_t row = kb + lane;
            const bool ok = row < k_end;
            double* dw = raw + GR_KC * (stA + stB + stC) + lane;
            if (w) cp_async8(dw, w + (ok ? row : last_row), ok);
            else *dw = ok ? 1.0 : 0.0;
        }
    };
    auto load_mapped = [&](int c) -> double {
        const int64_t row = k_begin + (int64_t)c * GR_KC + lane;
        return (mapper && row < k_end) ? gf_load_raw(fm, row) : 0.0;
    };
    auto store_mapped = [&](int c, double* raw, double x) {
        if (!mapper) return;
        const int64_t row = k_begin + (int64_t)c * GR_KC + lane;
        gf_store_mapped(fm, raw + m_off + lane * m_st, x, row < k_end, m_par);
    };
    auto synth = [&](const double* raw, double* tile) {       // operand columns of 4 rows per warp, a lane per column
        const double* rW = raw + GR_KC * (stA + stB + stC);
        for (int col = lane; col < NCOL; col += 32) {
            const unsigned pk = tCol[col];
            const int i = pk & 0xFFFF, j = pk >> 16;
            const bool isA = col < TAW;
            const double* src = raw;
            int st = stA;
            if (!isA) {
                if (col < TAW + TBW) { src = raw + GR_KC * stA; st = stB; }
                else { src = raw + GR_KC * (stA + stB); st = stC; }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int k = warp + 8 * r;
                double v = 0.0;
                if (pk != 0xFFFFFFFFu) {
                    v = src[k * st + i] * src[k * st + j];
                    if (isA) v *= rW[k];
                }
                tile[k * stT + col] = v;
            }
        }
    };
    auto mma_chunk = [&](const double* tile) {
        const double* rowp = tile + fk * stT;
#pragma unroll
        for (int k4 = 0; k4 < GR_KC; k4 += 4, rowp += 4 * stT) {
            double af[4], bf[4];
            if (BOX) {
                const double vb = rowp[offB[0]];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = rowp[offA[i]] * vb;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = rowp[offA[i]] * rowp[offB[i]];
            }
            const double* vp = rowp + TAW + TBW + wv + fr;
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = vp[j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    };

    double* raw0 = sRaw;
    double* raw1 = sRaw + rawsz;
    double* tile0 = sT;
    double* tile1 = sT + GR_KC * stT;
    if (nch > 0) {
        fill(0, raw0);
        store_mapped(0, raw0, load_mapped(0));
    }
    cp_async_wait_all();
    __syncthreads();              // raw factors of chunk 0 and the column table visible
    {
        double xpre = 0.0;
        if (nch > 1) {
            fill(1, raw1);
            xpre = load_mapped(1);
        }
        if (nch > 0) synth(raw0, tile0);
        if (nch > 1) store_mapped(1, raw1, xpre);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int c = 0; c < nch; ++c) {
        double* rawc = (c & 1) ? raw1 : raw0;          // holds chunk c (already synthesised): refilled with chunk c + 2
        double* rawn = (c & 1) ? raw0 : raw1;          // chunk c + 1
        double* tilec = (c & 1) ? tile1 : tile0;
        double* tilen = (c & 1) ? tile0 : tile1;
        double xpre = 0.0;
        if (c + 2 < nch) {
            fill(c + 2, rawc);
            xpre = load_mapped(c + 2);
        }
        if (c + 1 < nch) synth(rawn, tilen);
        mma_chunk(tilec);
        if (c + 2 < nch) store_mapped(c + 2, rawc, xpre);
        cp_async_wait_all();
        __syncthreads();          // the one barrier of a chunk: tiles of c + 1 and raw factors of c + 2 visible, buffers of c free
    }

    double* o = out + (int64_t)blockIdx.z * nU * nC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int jc = wu + i * 8 + fr;
        int64_t gu;        // output row: (qa, qb) in the caller's order (swapU: the launcher exchanged the first two factors)
        if (BOX) {
            const int qa = a0 + (jc >> 3), qb = b0 + (jc & 7);
            gu = (qa < nA && qb < nB) ? (swapU ? (int64_t)qb * nA + qa : (int64_t)qa * nB + qb) : nU;
        } else {
            gu = u0 + jc;
            if (swapU && gu < nU) {
                const int qa = (int)(gu / nB), qb = (int)(gu - (int64_t)qa * nB);
                gu = (int64_t)qb * nA + qa;
            }
        }
        if (gu >= nU) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gv = v0 + wv + j * 8 + 2 * fk + e;
                if (gv < nC) o[gu * nC + gv] = acc[i][j][e];
            }
        }
    }
}

// Right-hand side for small cores (P <= 4096): up to four entries of b per thread, rows streamed through shared memory in tiles.
// The GEMM-shaped kernel above wastes most of its 128 x 64 tile when P is a few hundred.
constexpr int RS_ROWS = 64;
constexpr int RS_PER = 4;     // entries of b per thread
constexpr int RS_MAXP = 1024 * RS_PER;
__global__ void __launch_bounds__(1024)
rhs_small_kernel(FactorDev fa, FactorDev fb, FactorDev fc, const double* __restrict__ w, int64_t rows, double* __restrict__ out,
                 int64_t rows_per_cta) {
    extern __shared__ double sm[];
    const int stA = fa.m | 1, stB = fb.m | 1, stC = fc.m | 1;
    double* sA = sm;
    double* sB = sA + RS_ROWS * stA;
    double* sC = sB + RS_ROWS * stB;
    double* sW = sC + RS_ROWS * stC;
    const int P = fa.m * fb.m * fc.m;
    const int tid = threadIdx.x, nt = blockDim.x;
    int ia[RS_PER], ib[RS_PER], ic[RS_PER];      // up to RS_PER entries of b per thread: i = tid + q * blockDim
    double acc[RS_PER];
#pragma unroll
    for (int q = 0; q < RS_PER; ++q) {
        const int i = min(tid + q * nt, P - 1);
        ic[q] = i % fc.m;
        ib[q] = (i / fc.m) % fb.m;
        ia[q] = i / (fc.m * fb.m);
        acc[q] = 0.0;
    }
    const int64_t k_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t k_end = min(rows, k_begin + rows_per_cta);
    for (int64_t kb = k_begin; kb < k_end; kb += RS_ROWS) {
        __syncthreads();
        for (int idx = tid; idx < RS_ROWS * fa.m; idx += nt) {
            const int k = idx / fa.m, i = idx - k * fa.m;
            const int64_t row = kb + k;
            sA[k * stA + i] = (row < k_end) ? map_eval(fa.map_kind, fa.ptr + (fa.div == 1 ? row : row / fa.div) * fa.ld, i) : 0.0;
        }
        for (int idx = tid; idx < RS_ROWS * fb.m; idx += nt) {
            const int k = idx / fb.m, i = idx - k * fb.m;
            const int64_t row = kb + k;
            sB[k * stB + i] = (row < k_end) ? map_eval(fb.map_kind, fb.ptr + (fb.div == 1 ? row : row / fb.div) * fb.ld, i) : 0.0;
        }
        for (int idx = tid; idx < RS_ROWS * fc.m; idx += nt) {
            const int k = idx / fc.m, i = idx - k * fc.m;
            const int64_t row = kb + k;
            sC[k * stC + i] = (row < k_end) ? map_eval(fc.map_kind, fc.ptr + (fc.div == 1 ? row : row / fc.div) * fc.ld, i) : 0.0;
        }
        for (int k = tid; k < RS_ROWS; k += nt) {
            const int64_t row = kb + k;
            sW[k] = (row < k_end) ? (w ? w[row] : 1.0) : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < RS_ROWS; ++k) {
            const double wk = sW[k];
#pragma unroll
            for (int q = 0; q < RS_PER; ++q)
                acc[q] = fma(wk * sA[k * stA + ia[q]], sB[k * stB + ib[q]] * sC[k * stC + ic[q]], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < RS_PER; ++q)
        if (tid + q * nt < P) out[(int64_t)blockIdx.x * P + tid + q * nt] = acc[q];
}

// Variant for cores whose last factor is at least 8 wide: a thread owns RS_PER entries b[ia, ib, ic0 + q*chunks] of ONE
// (ia, ib) pair (the tail is padded; consecutive lanes own consecutive ic0: conflict-free shared-memory reads), so one product w*A[ia]*B[ib] per sample feeds RS_PER fused multiply-adds
// with consecutive entries of C and no thread of a warp takes a different path.
__global__ void __launch_bounds__(1024)
rhs_small_vec_kernel(FactorDev fa, FactorDev fb, FactorDev fc, const double* __restrict__ w, int64_t rows, double* __restrict__ out,
                     int64_t rows_per_cta) {
    extern __shared__ double sm[];
    const int stA = fa.m | 1, stB = fb.m | 1, stC = (fc.m + RS_PER) | 1;      // C rows padded: the last chunk may read past fc.m
    double* sA = sm;
    double* sB = sA + RS_ROWS * stA;
    double* sC = sB + RS_ROWS * stB;
    double* sW = sC + RS_ROWS * stC;
    const int P = fa.m * fb.m * fc.m;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int chunks = (fc.m + RS_PER - 1) / RS_PER;
    const int slots = fa.m * fb.m * chunks;
    const bool active = tid < slots;
    const int pair = active ? tid / chunks : 0;
    const int ic0 = active ? tid - pair * chunks : 0;        // entries ic0 + q*chunks: consecutive lanes read consecutive words of C
    const int ia = pair / fb.m, ib = pair - ia * fb.m;
    double acc[RS_PER];
#pragma unroll
    for (int q = 0; q < RS_PER; ++q) acc[q] = 0.0;
    const int64_t k_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t k_end = min(rows, k_begin + rows_per_cta);
    for (int i = tid; i < RS_ROWS * stC; i += nt) sC[i] = 0.0;                 // the pad columns stay zero
    for (int64_t kb = k_begin; kb < k_end; kb += RS_ROWS) {
        __syncthreads();
        for (int idx = tid; idx < RS_ROWS * fa.m; idx += nt) {
            const int k = idx / fa.m, i = idx - k * fa.m;
            const int64_t row = kb + k;
            sA[k * stA + i] = (row < k_end) ? map_eval(fa.map_kind, fa.ptr + (fa.div == 1 ? row : row / fa.div) * fa.ld, i) : 0.0;
        }
        for (int idx = tid; idx < RS_ROWS * fb.m; idx += nt) {
            const int k = idx / fb.m, i = idx - k * fb.m;
            const int64_t row = kb + k;
            sB[k * stB + i] = (row < k_end) ? map_eval(fb.map_kind, fb.ptr + (fb.div == 1 ? row : row / fb.div) * fb.ld, i) : 0.0;
        }
        for (int idx = tid; idx < RS_ROWS * fc.m; idx += nt) {
            const int k = idx / fc.m, i = idx - k * fc.m;
            const int64_t row = kb + k;
            sC[k * stC + i] = (row < k_end) ? map_eval(fc.map_kind, fc.ptr + (fc.div == 1 ? row : row / fc.div) * fc.ld, i) : 0.0;
        }
        for (int k = tid; k < RS_ROWS; k += nt) {
            const int64_t row = kb + k;
            sW[k] = (row < k_end) ? (w ? w[row] : 1.0) : 0.0;
        }
        __syncthreads();
        if (active) {
            const double* pa = sA + ia;
            const double* pb = sB + ib;
            const double* pc = sC + ic0;
#pragma unroll 4
            for (int k = 0; k < RS_ROWS; ++k) {
                const double t = (sW[k] * pa[k * stA]) * pb[k * stB];
#pragma unroll
                for (int q = 0; q < RS_PER; ++q) acc[q] = fma(t, pc[k * stC + q * chunks], acc[q]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int q = 0; q < RS_PER; ++q)
            if (ic0 + q * chunks < fc.m) out[(int64_t)blockIdx.x * P + (int64_t)pair * fc.m + ic0 + q * chunks] = acc[q];
    }
}

// Right-hand side / J^T u for LARGE cores whose outer factors are at most 40 wide (config 5a: 38 x 29 x 38, P = 41 876).
//   out[a, p, b] = sum_rows w * fa[a] * fb[p] * fc[b]  =  for every p:  (fa (.) (w fb[p]))^T fc   -- an (mA x rows) x (rows x mC) product
// The GEMM-shaped kr3 kernel synthesises a 128-wide operand tile w*fa*fb per 32-row chunk in shared memory and pads mC = 38 to
// a 64-wide tile; it runs at 4.5 TF/s on this shape (profiles/r2_launches_cfg5a_1M_summary.txt: 18.5 ms per call, and the exact
// refinement calls it once per conjugate-gradient iteration).  Here nothing is synthesised in memory: a warp owns ONE p and the whole
// (8 MT) x (8 NT) output block of that p in registers; per 4-row DMMA step it loads MT fragments of fa, scales them by the per-row
// scalar w * fb[p] in registers (one multiply per fragment), loads NT fragments of fc and issues MT x NT DMMAs.  A CTA = 8 warps =
// 8 consecutive p; rows are split over blockIdx.y.  Chunks of 32 rows are prefetched into registers while the previous chunk is
// multiplied (one block barrier per chunk, two shared buffers).
constexpr int RB_KC = 32;          // rows per chunk
constexpr int RB_PW = 8;           // p values per CTA = warps
constexpr int RB_MAXPRE = 14;      // prefetch registers per thread (>= 32 * (40 + 40 + 8) / 256)

template <int MT, int NT>
__global__ void __launch_bounds__(256, 1)
rhs_big_kernel(FactorDev fa, FactorDev fb, FactorDev fc, const double* __restrict__ w, int64_t rows, double* __restrict__ out,
               int64_t rows_per_split) {
    constexpr int stA = 8 * MT + 4, stC = 8 * NT + 4;      // == 4 (mod 8) doubles: conflict-free fragment loads
    constexpr int BUF = RB_KC * (stA + stC + RB_PW);       // the weights live in a padding column of the fc tile
    __shared__ __align__(16) double sm[2 * BUF];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int p0 = blockIdx.x * RB_PW;
    const int p = p0 + warp;
    const int64_t k_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t k_end = min(rows, k_begin + rows_per_split);
    for (int i = tid; i < 2 * BUF; i += 256) sm[i] = 0.0;           // the padding columns stay zero

    // chunk elements owned by this thread: [32 x mA of fa | 32 x mC of fc | 32 x 8 of fb[p0 + q] | 32 of w].  The prefetch only
    // LOADS (raw values, nothing depends on them until the commit one chunk later); feature maps are applied at the commit.
    const int nA = RB_KC * fa.m, nC = RB_KC * fc.m, nB = RB_KC * RB_PW;
    const int total = nA + nC + nB + RB_KC;
    int desc[RB_MAXPRE];            // (kind << 16) | (k << 8) | column
    double pre[RB_MAXPRE];
#pragma unroll
    for (int q = 0; q < RB_MAXPRE; ++q) {
        const int e = tid + 256 * q;
        int d = -1;
        if (e < nA) { const int k = e / fa.m; d = (0 << 16) | (k << 8) | (e - k * fa.m); }
        else if (e < nA + nC) { const int e2 = e - nA; const int k = e2 / fc.m; d = (1 << 16) | (k << 8) | (e2 - k * fc.m); }
        else if (e < nA + nC + nB) { const int e2 = e - nA - nC; d = ((p0 + (e2 & 7) < fb.m ? 2 : 4) << 16) | ((e2 >> 3) << 8) | (e2 & 7); }
        else if (e < total) { d = (3 << 16) | ((e - nA - nC - nB) << 8); }
        desc[q] = d;
    }
    auto prefetch = [&](int64_t kb) {
#pragma unroll
        for (int q = 0; q < RB_MAXPRE; ++q) {
            const int d = desc[q];
            double v = 0.0;
            if (d >= 0) {
                const int kind = d >> 16, k = (d >> 8) & 255, c = d & 255;
                const int64_t row = kb + k;
                if (row < k_end && kind != 4) {
                    const double* src;
                    if (kind == 0) src = map_raw_ptr(fa.map_kind, fa.ptr + (fa.div == 1 ? row : row / fa.div) * fa.ld, c);
                    else if (kind == 1) src = map_raw_ptr(fc.map_kind, fc.ptr + (fc.div == 1 ? row : row / fc.div) * fc.ld, c);
                    else if (kind == 2) src = map_raw_ptr(fb.map_kind, fb.ptr + (fb.div == 1 ? row : row / fb.div) * fb.ld, p0 + c);
                    else src = w ? w + row : nullptr;
                    v = src ? *src : 1.0;
                } else if (kind != 3 && kind != 4 && row >= k_end) {
                    v = __longlong_as_double(0x7ff8000000000001ll);      // marks "row past the end": the commit stores 0
                }
            }
            pre[q] = v;
        }
    };
    auto commit = [&](int buf) {
        double* sA = sm + buf * BUF;
        double* sC = sA + RB_KC * stA;
        double* sB = sC + RB_KC * stC;
#pragma unroll
        for (int q = 0; q < RB_MAXPRE; ++q) {
            const int d = desc[q];
            if (d >= 0) {
                const int kind = d >> 16, k = (d >> 8) & 255, c = d & 255;
                const double raw = pre[q];
                const bool past = (__double_as_longlong(raw) == 0x7ff8000000000001ll);
                if (kind == 0) sA[k * stA + c] = past ? 0.0 : map_apply(fa.map_kind, raw, c);
                else if (kind == 1) sC[k * stC + c] = past ? 0.0 : map_apply(fc.map_kind, raw, c);
                else if (kind == 2) sB[k * RB_PW + c] = past ? 0.0 : map_apply(fb.map_kind, raw, p0 + c);
                else if (kind == 3) sC[k * stC + 8 * NT] = raw;      // 0 past the end (never loaded): the row contributes nothing
            }
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    __syncthreads();                 // zero fill done before the first commit
    if (k_begin < k_end) prefetch(k_begin);
    int buf = 0;
    for (int64_t kb = k_begin; kb < k_end; kb += RB_KC, buf ^= 1) {
        commit(buf);
        __syncthreads();
        if (kb + RB_KC < k_end) prefetch(kb + RB_KC);
        if (p < fb.m) {
            const double* sA = sm + buf * BUF;
            const double* sC = sA + RB_KC * stA;
            const double* sB = sC + RB_KC * stC;
#pragma unroll 2
            for (int k4 = 0; k4 < RB_KC; k4 += 4) {
                const double sc = sB[(k4 + fk) * RB_PW + warp] * sC[(k4 + fk) * stC + 8 * NT];
                double af[MT], bf[NT];
                const double* ap = sA + (k4 + fk) * stA + fr;
                const double* cp = sC + (k4 + fk) * stC + fr;
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = ap[i * 8] * sc;
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = cp[j * 8];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
    }
    if (p < fb.m) {
        double* o = out + (int64_t)blockIdx.y * ((int64_t)fa.m * fb.m * fc.m);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int a = i * 8 + fr;
            if (a >= fa.m) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int b = j * 8 + 2 * fk + e;
                    if (b < fc.m) o[((int64_t)a * fb.m + p) * fc.m + b] = acc[i][j][e];
                }
        }
    }
}

template <int MT>
static int launch_rhs_big_nt(int NT, dim3 grid, const FactorDev& a, const FactorDev& b, const FactorDev& c, const double* w, int64_t rows,
                             double* dst, int64_t rps, cudaStream_t st) {
    switch (NT) {
        case 1: rhs_big_kernel<MT, 1><<<grid, 256, 0, st>>>(a, b, c, w, rows, dst, rps); break;
        case 2: rhs_big_kernel<MT, 2><<<grid, 256, 0, st>>>(a, b, c, w, rows, dst, rps); break;
        case 3: rhs_big_kernel<MT, 3><<<grid, 256, 0, st>>>(a, b, c, w, rows, dst, rps); break;
        case 4: rhs_big_kernel<MT, 4><<<grid, 256, 0, st>>>(a, b, c, w, rows, dst, rps); break;
        default: rhs_big_kernel<MT, 5><<<grid, 256, 0, st>>>(a, b, c, w, rows, dst, rps); break;
    }
    return TN_OK;
}

__global__ void reduce_splits_kernel(const double* __restrict__ work, double* __restrict__ dst, int64_t n, int ksplit,
                                     int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = accumulate ? dst[i] : 0.0;
        for (int k = 0; k < ksplit; ++k) s += work[(int64_t)k * n + i];
        dst[i] = s;
    }
}

// Row splits of the Gram / right-hand-side GEMM.  Two CTAs are resident per SM and every CTA of a launch does the same work, so the
// launch takes ceil(tiles * ks / slots) rounds of rows / ks rows each: 600 CTAs on 296 slots (config 3's middle site with the former
// "4 CTAs per SM" rule) ran three rounds for two rounds' worth of work.  ks minimises rounds * (rows per split + a per-CTA overhead).
static int choose_ksplit(int64_t rows, int64_t nU, int64_t nV) {
    const int64_t tiles = ceil_div64(nU, GR_TU) * ceil_div64(nV, GR_TV);
    const int64_t slots = 2LL * sm_count();
    int64_t ks_max = ceil_div64(rows, 8 * GR_KC);
    const int64_t max_by_mem = (int64_t)(64LL << 20) / (nU * nV > 0 ? nU * nV : 1);  // <= 512 MB of partials
    if (ks_max > max_by_mem) ks_max = max_by_mem;
    if (ks_max > 1024) ks_max = 1024;
    if (ks_max < 1) ks_max = 1;
    if (tiles >= 4 * slots) return 1;
    const int64_t overhead = 2 * GR_KC;            // prologue + partial-tile store + its share of the reduction, in rows of work
    int64_t best = 1;
    double best_t = (double)ceil_div64(tiles, slots) * (double)(rows + overhead);
    for (int64_t ks = 2; ks <= ks_max && tiles * ks <= 8 * slots; ++ks) {
        const double t = (double)ceil_div64(tiles * ks, slots) * (double)(ceil_div64(rows, ks) + overhead);
        if (t < 0.98 * best_t) { best_t = t; best = ks; }
    }
    return (int)best;
}

template <int MODE>
static int launch_kr3(const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows,
                      double* dst, double* work, int ksplit, int accumulate, cudaStream_t st, const int* t1 = nullptr,
                      const int* t2 = nullptr, const int* t3 = nullptr, int P = 0) {
    constexpr bool PAIR = (MODE == 1);
    const FactorDev a = to_dev(fa), b = to_dev(fb), c = to_dev(fc);
    int nA = PAIR ? npairs(a.m) : a.m, nB = PAIR ? npairs(b.m) : b.m, nC = PAIR ? npairs(c.m) : c.m;
    if (MODE >= 2) { nA = P; nB = 1; nC = (MODE == 2) ? P : 1; }
    const int64_t nU = (int64_t)nA * nB;
    const int64_t n = nU * nC;
    if (ksplit < 1) ksplit = 1;
    int64_t rps = ceil_div64(ceil_div64(rows, ksplit), GR_KC) * GR_KC;
    if (rps < GR_KC) rps = GR_KC;
    const bool direct = (ksplit == 1 && !accumulate);
    TN_CHECK_ARG(direct || work != nullptr, "kr3: ksplit=%d / accumulate need a work buffer", ksplit);
    const size_t smem = (size_t)(GR_KC * ((a.m | 1) + (b.m | 1) + (c.m | 1)) + GR_KC + GR_KC * GR_SU + GR_KC * GR_SV) * sizeof(double) +
                        (size_t)(4 * GR_TU + 3 * GR_TV) * sizeof(short);
    TN_CHECK_ARG(smem <= 227 * 1024, "kr3: factor sizes %d,%d,%d need %zu B of shared memory", a.m, b.m, c.m, smem);
    TN_CHECK_ARG(a.m < 32768 && b.m < 32768 && c.m < 32768, "kr3: factor too large");
    const int64_t gx = ceil_div64(nU, GR_TU), gy = ceil_div64(nC, GR_TV);
    TN_CHECK_ARG(gy <= 65535 && ksplit <= 65535 && gx <= 0x7fffffff, "kr3: grid too large");
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)ksplit);
    // (A software-pipelined variant of the kernel below -- register prefetch of the raw factors, double-buffered 128-wide operand tiles,
    // two barriers per 16-row chunk -- gave the same bits and was 10-20 % SLOWER: profiles/r2_gram_f64_probe.jsonl; removed.)
    bool launched = false;
    if (MODE == 1 && !getenv("TN_GRAM_F64_UNFACTORED")) {
        // factored-operand kernel.  Tiling of the 128 U columns of a CTA: a 16 x 8 box of (qa, qb) (24 operand columns), or 128 consecutive
        // (qa, qb) indices (needs the second factor's pairs to fit: nB < 64; 127 / nB + 2 + nB operand columns); the first two factors may
        // be exchanged (the kernel writes row qb * nA + qa then).  The choice minimises the padded tile area; a site where no choice
        // comes within 25 % of the plain 128-column tiling, or whose operand columns are no fewer than U's, keeps the kernel above.
        const int64_t lin_pad = ceil_div64(nU, GR_TU) * GR_TU;
        int best_swap = 0, best_box = 0;
        int64_t best_pad = -1;
        int best_cols = 0;
        for (int sw = 0; sw < 2; ++sw) {
            const int xA = sw ? nB : nA, xB = sw ? nA : nB;
            for (int bx = 0; bx < 2; ++bx) {
                if (!bx && xB >= 64) continue;
                const int64_t pad = bx ? ceil_div64(xA, 16) * 16 * (ceil_div64(xB, 8) * 8) : lin_pad;
                const int cols = bx ? 24 : (GR_TU - 1) / xB + 2 + xB;
                if (cols > 64) continue;
                if (best_pad < 0 || pad < best_pad || (pad == best_pad && cols < best_cols)) {
                    best_pad = pad; best_cols = cols; best_swap = sw; best_box = bx;
                }
            }
        }
        if (best_pad >= 0 && 4 * best_pad <= 5 * lin_pad) {
            const FactorDev& xa = best_swap ? b : a;
            const FactorDev& xb = best_swap ? a : b;
            const int xA = best_swap ? nB : nA, xB = best_swap ? nA : nB;
            const bool box = best_box != 0;
            const int TAW = box ? 16 : (GR_TU - 1) / xB + 2, TBW = box ? 8 : xB;
            int stT = TAW + TBW + GR_TV;
            stT += (4 - (stT & 7) + 8) & 7;           // = 4 (mod 8): conflict-free DMMA fragment reads (row offset 8 banks per k)
            const int nTB = (xB + 7) / 8;
            const size_t fsmem = (size_t)(2 * GR_KC * (((a.m + 1) & ~1) + ((b.m + 1) & ~1) + ((c.m + 1) & ~1) + 1) + 2 * GR_KC * stT) * sizeof(double) +
                                 (size_t)(TAW + TBW + GR_TV) * sizeof(unsigned);
            const int64_t fgx = box ? (int64_t)((xA + 15) / 16) * nTB : gx;
            auto vec_ok = [](const FactorDev& f) {
                return f.map_kind == TN_MAP_IDENTITY && f.m >= 2 && (reinterpret_cast<uintptr_t>(f.ptr) & 15) == 0 && (f.ld & 1) == 0;
            };
            const int vecmask = getenv("TN_GRAM_F64_CP8") ? 0 : (vec_ok(xa) ? 1 : 0) | (vec_ok(xb) ? 2 : 0) | (vec_ok(c) ? 4 : 0);
            if (fsmem <= 227 * 1024 && fgx <= 0x7fffffff) {
                dim3 fgrid((unsigned)fgx, (unsigned)gy, (unsigned)ksplit);
                if (box) {
                    TN_SMEM(gram_f64_fact_kernel<true>, fsmem);
                    gram_f64_fact_kernel<true><<<fgrid, GR_THREADS, fsmem, st>>>(xa, xb, c, w, rows, direct ? dst : work, xA, xB, nC, rps, TAW, TBW, stT,
                                                                                 nTB, best_swap, vecmask);
                } else {
                    TN_SMEM(gram_f64_fact_kernel<false>, fsmem);
                    gram_f64_fact_kernel<false><<<fgrid, GR_THREADS, fsmem, st>>>(xa, xb, c, w, rows, direct ? dst : work, xA, xB, nC, rps, TAW, TBW, stT,
                                                                                  nTB, best_swap, vecmask);
                }
                launched = true;
            }
        }
    }
    if (!launched) {
        TN_SMEM(kr3_f64_kernel<MODE>, smem);
        kr3_f64_kernel<MODE><<<grid, GR_THREADS, smem, st>>>(a, b, c, w, rows, direct ? dst : work, nA, nB, nC, rps, t1, t2, t3);
    }
    TN_LAUNCH_CHECK();
    if (!direct) {
        int64_t blocks = ceil_div64(n, 256);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        reduce_splits_kernel<<<(unsigned)blocks, 256, 0, st>>>(work, dst, n, ksplit, accumulate);
        TN_LAUNCH_CHECK();
    }
    return TN_OK;
}

// ---------------------------------------------------------------------------------------------
// sigma, expansion, right-hand side preparation, node update

struct PosInfo {
    int m[3];     // size of parameter position t
    int role[3];  // role (0=a,1=b,2=c) of position t
    int n_role[3];  // pair counts per role
};

__device__ __forceinline__ int64_t m_index(const PosInfo& pi, const int* ii, const int* jj) {
    int q[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int lo = min(ii[t], jj[t]), hi = max(ii[t], jj[t]);
        q[pi.role[t]] = pair_index(lo, hi, pi.m[t]);
    }
    return ((int64_t)q[0] * pi.n_role[1] + q[1]) * pi.n_role[2] + q[2];
}

__global__ void sigma_kernel(const double* __restrict__ M, PosInfo pi, double* __restrict__ sigma) {
    __shared__ double red[32];
    const int64_t P = (int64_t)pi.m[0] * pi.m[1] * pi.m[2];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        int ii[3];
        ii[2] = (int)(i % pi.m[2]);
        const int64_t r = i / pi.m[2];
        ii[1] = (int)(r % pi.m[1]);
        ii[0] = (int)(r / pi.m[1]);
        s += fabs(M[m_index(pi, ii, ii)]);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) {
            double sc = v / (double)P;  // diag.abs().mean(), network.py:298
            if (sc == 0.0) sc = 1.0;    // network.py:299-300
            sigma[0] = sc;
        }
    }
}

__global__ void expand_kernel(const double* __restrict__ M, PosInfo pi, const double* __restrict__ sigma, double ridge,
                              double* __restrict__ A, int64_t lda) {
    const int64_t P = (int64_t)pi.m[0] * pi.m[1] * pi.m[2];
    const double sc = sigma[0];
    for (int64_t i = blockIdx.x; i < P; i += gridDim.x) {
        int ii[3];
        ii[2] = (int)(i % pi.m[2]);
        const int64_t r = i / pi.m[2];
        ii[1] = (int)(r % pi.m[1]);
        ii[0] = (int)(r / pi.m[1]);
        double* Ai = A + i * lda;
        for (int64_t j = threadIdx.x; j < P; j += blockDim.x) {
            int jj[3];
            jj[2] = (int)(j % pi.m[2]);
            const int64_t rj = j / pi.m[2];
            jj[1] = (int)(rj % pi.m[1]);
            jj[0] = (int)(rj / pi.m[1]);
            double v = M[m_index(pi, ii, jj)] / sc;  // A_f / scale, network.py:301
            if (i == j) v += ridge;                   // + 2 eps I, network.py:308,312
            Ai[j] = v;
        }
    }
}

__global__ void rhs_prepare_kernel(const double* __restrict__ b, const double* __restrict__ theta,
                                   const double* __restrict__ sigma, double ridge, double* __restrict__ rhs, int64_t P) {
    const double sc = sigma[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        double v = b[i] / sc;                       // network.py:302
        if (ridge != 0.0) v += ridge * theta[i];    // network.py:309,313
        rhs[i] = -v;                                // solve(A, -b), network.py:305,315
    }
}

__device__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    return t;
}

// One CTA: norms and the axpy of TensorNode.update_node (node.py:178-203).
__global__ void update_node_kernel(double* __restrict__ theta, const double* __restrict__ step, int64_t P, double lr,
                                   int adaptive, double max_norm) {
    __shared__ double red[32];
    double scale = 1.0;
    if (adaptive) {
        double s2 = 0.0, p2 = 0.0;
        for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
            s2 += step[i] * step[i];
            p2 += theta[i] * theta[i];
        }
        const double sn = sqrt(block_sum(s2, red)), pn = sqrt(block_sum(p2, red));
        if (sn > pn) scale = pn / sn;
    }
    double n2 = 0.0;
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        const double v = theta[i] + lr * (step[i] * scale);
        theta[i] = v;
        n2 += v * v;
    }
    if (max_norm > 0.0) {
        const double cn = sqrt(block_sum(n2, red));
        if (cn > max_norm) {
            const double f = max_norm / cn;
            for (int64_t i = threadIdx.x; i < P; i += blockDim.x) theta[i] *= f;
        }
    }
}

static int make_pos(const int* m_pos, const int* role_of_pos, PosInfo& pi) {
    bool seen[3] = {false, false, false};
    for (int t = 0; t < 3; ++t) {
        if (m_pos[t] < 1 || role_of_pos[t] < 0 || role_of_pos[t] > 2 || seen[role_of_pos[t]]) return -1;
        seen[role_of_pos[t]] = true;
        pi.m[t] = m_pos[t];
        pi.role[t] = role_of_pos[t];
        pi.n_role[role_of_pos[t]] = npairs(m_pos[t]);
    }
    return 0;
}

}  // namespace tn

extern "C" int tn_gram_ksplit(int64_t rows, int ma, int mb, int mc, int mode) {
    using namespace tn;
    if (mode != 0) return 1;
    return choose_ksplit(rows, (int64_t)npairs(ma) * npairs(mb), npairs(mc));
}

static int rhs_small_threads(int64_t P) {
    int threads = (int)((tn::ceil_div64(P, tn::RS_PER) + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > 1024) threads = 1024;
    return threads;
}

static int rhs_small_ctas(int64_t rows, int64_t P) {
    int64_t n = tn::ceil_div64(rows, 8 * tn::RS_ROWS);
    int per_sm = 1536 / rhs_small_threads(P);        // enough warps per SM to hide the shared-memory latency of the row loop
    if (per_sm < 2) per_sm = 2;
    if (per_sm > 6) per_sm = 6;
    const int64_t cap = (int64_t)per_sm * tn::sm_count();
    if (n > cap) n = cap;
    return (int)(n < 1 ? 1 : n);
}

extern "C" int tn_rhs_ksplit(int64_t rows, int ma, int mb, int mc) {
    if ((int64_t)ma * mb * mc <= tn::RS_MAXP) return rhs_small_ctas(rows, (int64_t)ma * mb * mc);   // small-core kernel: one partial per CTA
    return tn::choose_ksplit(rows, (int64_t)ma * mb, mc);
}

int tn_gram_kr3_tc(int mode, const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w,
                   int64_t rows, double* M, int accumulate, void* stream);

extern "C" int tn_gram_kr3(int mode, const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w,
                           int64_t rows, double* M, double* work, int ksplit, int accumulate, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(fa && fb && fc && M, "tn_gram_kr3: null argument");
    TN_CHECK_ARG(rows >= 0, "tn_gram_kr3: negative rows");
    TN_CHECK_ARG(fa->m >= 1 && fb->m >= 1 && fc->m >= 1, "tn_gram_kr3: empty factor");
    TN_CHECK_ARG(mode >= 0 && mode <= 3, "tn_gram_kr3: unknown mode %d", mode);
    if (rows == 0) {                   // an empty shard contributes nothing
        if (!accumulate)
            TN_CUDA(cudaMemsetAsync(M, 0, (size_t)npairs(fa->m) * npairs(fb->m) * npairs(fc->m) * sizeof(double), as_stream(stream)));
        return TN_OK;
    }
    if (mode == 0) return launch_kr3<1>(fa, fb, fc, w, rows, M, work, ksplit, accumulate, as_stream(stream));
    if (mode >= 1 && mode <= 3) return tn_gram_kr3_tc(mode, fa, fb, fc, w, rows, M, accumulate, stream);
    set_error("tn_gram_kr3: unknown mode %d", mode);
    return TN_EINVAL;
}

extern "C" int tn_rhs_kr3(const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows,
                          double* b, double* work, int ksplit, int accumulate, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(fa && fb && fc && b, "tn_rhs_kr3: null argument");
    TN_CHECK_ARG(rows >= 0, "tn_rhs_kr3: negative rows");
    const int64_t P = (int64_t)fa->m * fb->m * fc->m;
    if (rows == 0) {
        if (!accumulate) TN_CUDA(cudaMemsetAsync(b, 0, (size_t)P * sizeof(double), as_stream(stream)));
        return TN_OK;
    }
    if (P <= RS_MAXP) {
        const FactorDev a = to_dev(fa), bb = to_dev(fb), c = to_dev(fc);
        const int ctas = rhs_small_ctas(rows, P);
        TN_CHECK_ARG(ksplit == ctas && work != nullptr, "tn_rhs_kr3: small-core path needs work for %d partials (got ksplit=%d)", ctas, ksplit);
        const int64_t rpc = ceil_div64(ceil_div64(rows, ctas), RS_ROWS) * RS_ROWS;
        const int slots = a.m * bb.m * ((c.m + RS_PER - 1) / RS_PER);
        if (c.m >= 8 && slots <= 1024) {
            const int vthreads = ((slots + 31) / 32) * 32 < 128 ? 128 : ((slots + 31) / 32) * 32;
            const size_t vsmem = (size_t)RS_ROWS * ((a.m | 1) + (bb.m | 1) + ((c.m + RS_PER) | 1) + 1) * sizeof(double);
            TN_CHECK_ARG(vsmem <= 200 * 1024, "tn_rhs_kr3: factors too wide for the small-core path");
            TN_SMEM(rhs_small_vec_kernel, vsmem);
            rhs_small_vec_kernel<<<ctas, vthreads, vsmem, as_stream(stream)>>>(a, bb, c, w, rows, work, rpc);
        } else {
            const int threads = rhs_small_threads(P);
            const size_t smem = (size_t)RS_ROWS * ((a.m | 1) + (bb.m | 1) + (c.m | 1) + 1) * sizeof(double);
            TN_CHECK_ARG(smem <= 200 * 1024, "tn_rhs_kr3: factors too wide for the small-core path");
            TN_SMEM(rhs_small_kernel, smem);
            rhs_small_kernel<<<ctas, threads, smem, as_stream(stream)>>>(a, bb, c, w, rows, work, rpc);
        }
        TN_LAUNCH_CHECK();
        int64_t blocks = ceil_div64(P, 256);
        reduce_splits_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(work, b, P, ctas, accumulate);
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    if (fa->m <= 40 && fc->m <= 40 && fa->m * fc->m >= 64 && work != nullptr && ksplit >= 1 && !getenv("TN_RHS_NO_BIG")) {
        // register-blocked kernel: one wave of CTAs, rows split so that (p blocks) x (row splits) fills the SMs
        const FactorDev a = to_dev(fa), bb = to_dev(fb), c = to_dev(fc);
        const int pblocks = (bb.m + RB_PW - 1) / RB_PW;
        int64_t ks = sm_count() / pblocks;
        if (ks > ksplit) ks = ksplit;
        const int64_t max_by_rows = ceil_div64(rows, 4 * RB_KC);
        if (ks > max_by_rows) ks = max_by_rows;
        if (ks < 1) ks = 1;
        const int64_t rps = ceil_div64(ceil_div64(rows, ks), RB_KC) * RB_KC;
        ks = ceil_div64(rows, rps);
        const int MT = (a.m + 7) / 8, NT = (c.m + 7) / 8;
        dim3 grid((unsigned)pblocks, (unsigned)ks);
        cudaStream_t st = as_stream(stream);
        switch (MT) {
            case 1: launch_rhs_big_nt<1>(NT, grid, a, bb, c, w, rows, work, rps, st); break;
            case 2: launch_rhs_big_nt<2>(NT, grid, a, bb, c, w, rows, work, rps, st); break;
            case 3: launch_rhs_big_nt<3>(NT, grid, a, bb, c, w, rows, work, rps, st); break;
            case 4: launch_rhs_big_nt<4>(NT, grid, a, bb, c, w, rows, work, rps, st); break;
            default: launch_rhs_big_nt<5>(NT, grid, a, bb, c, w, rows, work, rps, st); break;
        }
        TN_LAUNCH_CHECK();
        int64_t blocks = ceil_div64(P, 256);
        reduce_splits_kernel<<<(unsigned)blocks, 256, 0, st>>>(work, b, P, (int)ks, accumulate);
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    return launch_kr3<0>(fa, fb, fc, w, rows, b, work, ksplit, accumulate, as_stream(stream));
}

extern "C" int tn_generic_ksplit(int64_t rows, int P, int rhs_only) {
    return tn::choose_ksplit(rows, P, rhs_only ? 1 : P);
}

extern "C" int tn_gram_generic(const tn_factor* f1, const tn_factor* f2, const tn_factor* f3, const int* t1, const int* t2,
                               const int* t3, int P, const double* w, int64_t rows, double* out, int rhs_only, double* work,
                               int ksplit, int accumulate, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(f1 && f2 && f3 && t1 && t2 && t3 && out && P >= 1 && rows >= 0, "tn_gram_generic: bad arguments");
    if (rhs_only) return launch_kr3<3>(f1, f2, f3, w, rows, out, work, ksplit, accumulate, as_stream(stream), t1, t2, t3, P);
    return launch_kr3<2>(f1, f2, f3, w, rows, out, work, ksplit, accumulate, as_stream(stream), t1, t2, t3, P);
}

extern "C" int tn_gram_sigma(const double* M, const int* m_pos, const int* role_of_pos, double* sigma_out, void* stream) {
    using namespace tn;
    PosInfo pi;
    TN_CHECK_ARG(M && sigma_out && m_pos && role_of_pos && make_pos(m_pos, role_of_pos, pi) == 0, "tn_gram_sigma: bad arguments");
    sigma_kernel<<<1, 1024, 0, as_stream(stream)>>>(M, pi, sigma_out);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_gram_expand(const double* M, const int* m_pos, const int* role_of_pos, const double* sigma, double ridge,
                              double* A, int64_t lda, void* stream) {
    using namespace tn;
    PosInfo pi;
    TN_CHECK_ARG(M && A && sigma && m_pos && role_of_pos && make_pos(m_pos, role_of_pos, pi) == 0, "tn_gram_expand: bad arguments");
    const int64_t P = (int64_t)pi.m[0] * pi.m[1] * pi.m[2];
    TN_CHECK_ARG(lda >= P, "tn_gram_expand: lda < P");
    int64_t blocks = P;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    expand_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(M, pi, sigma, ridge, A, lda);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_rhs_prepare(const double* b, const double* theta, const double* sigma, double ridge, double* rhs,
                              int64_t P, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(b && sigma && rhs && P >= 1 && (ridge == 0.0 || theta), "tn_rhs_prepare: bad arguments");
    int64_t blocks = ceil_div64(P, 256);
    if (blocks > 1024) blocks = 1024;
    rhs_prepare_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(b, theta, sigma, ridge, rhs, P);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_update_node(double* theta, const double* step, int64_t P, double lr, int adaptive_step, double max_norm,
                              double* scratch, void* stream) {
    using namespace tn;
    (void)scratch;
    TN_CHECK_ARG(theta && step && P >= 1, "tn_update_node: bad arguments");
    update_node_kernel<<<1, 1024, 0, as_stream(stream)>>>(theta, step, P, lr, adaptive_step, max_norm);
    TN_LAUNCH_CHECK();
    return TN_OK;
}
