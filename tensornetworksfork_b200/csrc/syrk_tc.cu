// Trailing update of the blocked Cholesky on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32).
//
//   C[i][q] -= sum_k L[i][k] * L[q][k]        C = A[c0:, c0:] (lower tiles only),  L = A[c0:, k0:k0+kb]  (fp64)
//
// This is the P^3/3 part of TensorNetwork.solve_system's factorisation (reference tensor/network.py:311-316) and the
// only part of it with GEMM shape.  FP64 on B200 peaks at 35 TF/s, the TF32 tensor pipe at 600+: the panel is copied once
// to fp32 (syrk_tc_convert_kernel), each CTA streams its 256 + 256 panel rows through a 3-stage cp.async ring straight
// into the UMMA canonical K-major layout, splits them in place into hi = tf32(x) and lo = x - hi, and one elected lane
// issues hi*hi + hi*lo + lo*hi into fp32 TMEM accumulators (two 128 x 256 tiles = all 512 TMEM columns).  The tile of C is
// read-modify-written in fp64 once per panel.  The factor that comes out is accurate to ~1e-5 |L||L^T| (fp32 accumulation
// over kb terms, truncating); tn_cholesky_solve_mixed uses it as the preconditioner of an fp64 conjugate-gradient
// refinement, so the solution that is returned satisfies the fp64 system to the requested residual.
#include "common.cuh"
#include "tc_common.cuh"

namespace tn {

constexpr int ST_M = 128;                                   // MMA M
constexpr int ST_BN = 256;                                  // MMA N
constexpr int ST_T = 2;                                     // M tiles per CTA
constexpr int ST_TILE = ST_M * ST_T;                        // 256 x 256 tile of C per CTA
constexpr int ST_KC = 16;                                   // k per pipeline stage (two MMA K steps of 8)
constexpr int ST_NS = 3;                                    // stages
constexpr int ST_PROD_WARPS = 8;
constexpr int ST_THREADS = 32 + 32 * ST_PROD_WARPS;
constexpr uint32_t ST_BLK = ST_TILE * ST_KC * 4;            // bytes of one 256-row x 16-k operand block: [piece of 4 k][256 rows][16 B]
constexpr uint32_t ST_LBO = ST_TILE * 16;                   // bytes between two 4-k pieces of a block
constexpr uint32_t ST_STAGE = 4 * ST_BLK;                   // [row block hi][column block hi][row block lo][column block lo]
constexpr size_t ST_SMEM = (size_t)ST_NS * ST_STAGE + (3 * ST_NS + 1) * 8 + 16;

struct SyrkTcParams {
    const float* X;     // n_pad x pitch, X[r][k] = (float)A[c0 + r][k0 + k], zero for r >= n
    int64_t pitch;      // floats per row of X (multiple of ST_KC)
    int64_t n;          // rows / columns of the trailing matrix
    int nchunks;        // pitch / ST_KC
    double* C;          // A + c0 * lda + c0
    int64_t lda;
    const int* info;
    int nt;             // tile rows of the trailing matrix
    int nct;            // tile columns to update (== nt: the whole lower triangle; smaller: only the leading column tiles)
};

// X = (float) of the panel, stored as the IMAGES the update kernel's stages are made of: for every block of 256 rows and every chunk
// of 16 k one contiguous 16 KB block [piece of 4 k][256 rows][4 floats] (the UMMA K-major core-matrix layout, LBO = 4096 B), so that a
// stage operand is ONE bulk copy.  (Round 1 filled the stages with 16-byte cp.async: its capture attributes 67 % of the kernel's
// shared-memory wavefronts to LDGSTS replays -- about one wavefront per 32-byte sector whatever the bank pattern.)
__global__ void __launch_bounds__(256)
syrk_tc_convert_kernel(const double* __restrict__ A, int64_t lda, int64_t c0, int64_t k0, int64_t n, int kb, float* __restrict__ X,
                       int64_t pitch, int64_t n_pad) {
    const int64_t q4 = pitch / 4;
    const int64_t total = n_pad * q4;
    const int64_t nchunks = pitch / ST_KC;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / q4;
        const int k = (int)(idx - r * q4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) {
            const double* src = A + (c0 + r) * lda + k0 + k;
            if (k < kb) v.x = (float)src[0];
            if (k + 1 < kb) v.y = (float)src[1];
            if (k + 2 < kb) v.z = (float)src[2];
            if (k + 3 < kb) v.w = (float)src[3];
        }
        const int64_t rb = r / ST_TILE, rr = r - rb * ST_TILE;
        const int64_t chunk = k / ST_KC, piece = (k % ST_KC) / 4;
        reinterpret_cast<float4*>(X)[((rb * nchunks + chunk) * 4 + piece) * ST_TILE + rr] = v;
    }
}

// SPLIT = true: 3xTF32 (hi/lo split, three MMAs per K step).  SPLIT = false: one TF32 MMA per K step on the fp32 panel as it
// lands (the tensor core ignores the low 13 mantissa bits) -- the producers only copy; for factors that merely precondition.
template <bool SPLIT>
__global__ void __launch_bounds__(ST_THREADS, 1)
syrk_tc_kernel(SyrkTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (*p.info != 0) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // lower tile (bi >= bj, bj < nct) of this CTA
    int bi, bj;
    if (p.nct == p.nt) {
        const int t = blockIdx.x;
        bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        while (bi * (bi + 1) / 2 > t) --bi;
        bj = t - bi * (bi + 1) / 2;
    } else {
        int t = blockIdx.x;
        bj = 0;
        while (t >= p.nt - bj) { t -= p.nt - bj; ++bj; }       // column bj holds nt - bj tiles; nct is small on this path
        bi = bj + t;
    }
    const int64_t r0 = (int64_t)bi * ST_TILE, q0 = (int64_t)bj * ST_TILE;

    uint8_t* stage_base = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ST_NS * ST_STAGE);
    uint64_t* full = bars;               // [NS]  producers -> MMA
    uint64_t* empty = bars + ST_NS;      // [NS]  tcgen05.commit -> producers
    uint64_t* raw_full = bars + 2 * ST_NS;   // [NS]  the two bulk copies of a stage have landed
    uint64_t* acc_full = bars + 3 * ST_NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * ST_NS + 1);

    if (tid == 0) {
        for (int s = 0; s < ST_NS; ++s) {
            mbar_init(&full[s], ST_PROD_WARPS);
            mbar_init(&empty[s], 1);
            mbar_init(&raw_full[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = p.nchunks;

    if (warp == 0) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(ST_M, ST_BN);
            constexpr uint32_t sbo = 128;
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)s * ST_STAGE);
                const uint32_t b_hi = sb + ST_BLK, b_lo = sb + 3 * ST_BLK;
#pragma unroll
                for (int t = 0; t < ST_T; ++t) {
                    const uint32_t a_hi = sb + (uint32_t)t * (ST_M * 16);          // rows t * 128 .. of the 256-row block
                    const uint32_t a_lo = a_hi + 2 * ST_BLK;
                    const uint32_t d = tmem_base + (uint32_t)(t * ST_BN);
#pragma unroll
                    for (int j = 0; j < ST_KC / 8; ++j) {
                        const uint32_t ko = (uint32_t)(2 * j) * ST_LBO;
                        umma_tf32(d, make_desc(a_hi + ko, ST_LBO, sbo), make_desc(b_hi + ko, ST_LBO, sbo), idesc, (c == 0 && j == 0) ? 0u : 1u);
                        if (SPLIT) {
                            umma_tf32(d, make_desc(a_hi + ko, ST_LBO, sbo), make_desc(b_lo + ko, ST_LBO, sbo), idesc, 1u);
                            umma_tf32(d, make_desc(a_lo + ko, ST_LBO, sbo), make_desc(b_hi + ko, ST_LBO, sbo), idesc, 1u);
                        }
                    }
                }
                umma_commit(&empty[s]);
                if (c + 1 == nchunks) umma_commit(acc_full);
                if (++s == ST_NS) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================== producers / epilogue ===============================
        // One thread keeps the ring full: two bulk copies per stage (row block and column block of the panel, 16 KB each) completing on
        // raw_full[s].  All producer threads then split THEIR 16-byte pieces of the hi images in place (hi = tf32(x), lo = x - hi into
        // the lo images; consecutive lanes = consecutive pieces: conflict free) and hand the stage to the MMA warp.
        const int pt = tid - 32;                       // 0..255
        const uint32_t stage_s = smem_u32(stage_base);
        const int64_t rb_a = r0 / ST_TILE, rb_b = q0 / ST_TILE;
        auto issue_chunk = [&](int chunk) {
            if (chunk < nchunks && pt == 0) {
                const int s = chunk % ST_NS;
                const uint32_t bar = smem_u32(&raw_full[s]);
                const uint32_t dst = stage_s + (uint32_t)s * ST_STAGE;
                const float* srca = p.X + (rb_a * nchunks + chunk) * (int64_t)(ST_BLK / 4);
                const float* srcb = p.X + (rb_b * nchunks + chunk) * (int64_t)(ST_BLK / 4);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * ST_BLK) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(srca), "r"(ST_BLK), "r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst + ST_BLK), "l"(srcb), "r"(ST_BLK), "r"(bar) : "memory");
            }
        };
        issue_chunk(0);
        issue_chunk(1);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % ST_NS;
            mbar_wait(&raw_full[s], (uint32_t)((c / ST_NS) & 1));       // this stage's two blocks have landed
            if (SPLIT) {
                const uint32_t sb = stage_s + (uint32_t)s * ST_STAGE + (uint32_t)pt * 16;
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = lds128(sb + (uint32_t)u * 4096);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t a = sb + (uint32_t)u * 4096;
                    const float4 h = make_float4(tf32_rn(v[u].x), tf32_rn(v[u].y), tf32_rn(v[u].z), tf32_rn(v[u].w));
                    sts128(a, h);
                    sts128(a + 2 * ST_BLK, make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w));
                }
                fence_proxy_async();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            // refill: chunk c+2 goes where chunk c-1 was; wait until its MMAs have read it
            const int m = c + 2;
            if (pt == 0 && m < nchunks) {
                if (m >= ST_NS) mbar_wait(&empty[m % ST_NS], (uint32_t)((m / ST_NS - 1) & 1));
                issue_chunk(m);
            }
        }

        // ---- epilogue: C tile -= accumulator (fp64 read-modify-write; the tile belongs to this CTA alone)
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 1) >> 2;             // the two warps of a quarter split the columns
        float* sc = reinterpret_cast<float*>(stage_base) + (size_t)(warp - 1) * (32 * 33);
        constexpr int groups = ST_T * ST_BN / 32;     // 16 column groups of 32
        for (int g = half * (groups / 2); g < (half + 1) * (groups / 2); ++g) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), r);
#pragma unroll
            for (int e = 0; e < 32; ++e) sc[lane * 33 + e] = __uint_as_float(r[e]);
            __syncwarp();
            const int col = g * 32 + lane;
            const int t = col / ST_BN;
            const int64_t gcol = q0 + (col - t * ST_BN);
            const int64_t grow0 = r0 + (int64_t)t * ST_M + q * 32;
            double* dst = p.C + grow0 * p.lda + gcol;
            if (gcol < p.n) {
#pragma unroll
                // 16 rows of 256 B in flight per warp.  (An L2 prefetch of the whole 512 KB tile at kernel start was measured and removed:
                // it evicts the panel the MMAs stream from L2 -- P = 41 876 mixed solve 337 -> 650-900 ms, profiles/r2_chol_prefetch.txt.)
                for (int rr0 = 0; rr0 < 32; rr0 += 16) {
                    double cv[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int64_t grow = grow0 + rr0 + e;
                        cv[e] = (grow < p.n && gcol <= grow) ? dst[(int64_t)(rr0 + e) * p.lda] : 0.0;
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int64_t grow = grow0 + rr0 + e;
                        if (grow < p.n && gcol <= grow) dst[(int64_t)(rr0 + e) * p.lda] = cv[e] - (double)sc[(rr0 + e) * 33 + lane];
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Workspace (floats) for one call with n trailing rows and a kb-wide panel.
int64_t syrk_tc_work_floats(int64_t n, int kb) {
    return ceil_div64(n, ST_TILE) * ST_TILE * (ceil_div64(kb, ST_KC) * ST_KC);
}

// C = A[c0:, c0:] -= A[c0:, k0:k0+kb] * A[c0:, k0:k0+kb]^T on the lower 256 x 256 tiles.  X: syrk_tc_work_floats(P - c0, kb).
// col_limit > 0 restricts the update to the first col_limit columns of C (a multiple of 256, or everything up to P).
static thread_local int g_syrk_passes = 3;
void syrk_tc_set_passes(int passes) { g_syrk_passes = (passes == 1) ? 1 : 3; }

int syrk_tc_update(double* A, int64_t lda, int64_t P, int64_t c0, int64_t k0, int kb, float* X, const int* info, cudaStream_t st,
                   int64_t col_limit) {
    const int64_t n = P - c0;
    if (n <= 0 || kb <= 0) return TN_OK;
    TN_SMEM(syrk_tc_kernel<true>, ST_SMEM);
    TN_SMEM(syrk_tc_kernel<false>, ST_SMEM);
    SyrkTcParams p;
    p.pitch = ceil_div64(kb, ST_KC) * ST_KC;
    const int64_t nt = ceil_div64(n, ST_TILE);
    const int64_t n_pad = nt * ST_TILE;
    p.X = X;
    p.n = n;
    p.nchunks = (int)(p.pitch / ST_KC);
    p.C = A + c0 * lda + c0;
    p.lda = lda;
    p.info = info;
    int64_t blocks = ceil_div64(n_pad * (p.pitch / 4), 256);
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    syrk_tc_convert_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, c0, k0, n, kb, X, p.pitch, n_pad);
    TN_LAUNCH_CHECK();
    int64_t nct = nt;
    if (col_limit > 0 && col_limit < n) {
        TN_CHECK_ARG(col_limit % ST_TILE == 0, "syrk_tc_update: column limit %lld is not a multiple of the tile", (long long)col_limit);
        nct = col_limit / ST_TILE;
    }
    p.nt = (int)nt;
    p.nct = (int)nct;
    const int64_t ntiles = nct * nt - nct * (nct - 1) / 2;
    TN_CHECK_ARG(ntiles <= 0x7fffffff, "syrk_tc_update: grid too large");
    if (g_syrk_passes == 1) syrk_tc_kernel<false><<<(unsigned)ntiles, ST_THREADS, ST_SMEM, st>>>(p);
    else syrk_tc_kernel<true><<<(unsigned)ntiles, ST_THREADS, ST_SMEM, st>>>(p);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

}  // namespace tn
