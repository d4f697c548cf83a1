// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), used for every fc1 / fc2
// product of the transfer network: forward, data-gradient and weight-gradient, in the step kernels
// and in the full-table transfer (updata).
//
// Precision scheme ("3xTF32"): every fp32 operand element x is split on the fly into
//     hi = tf32_rna(x),  lo = tf32_rna(x - hi)          (x - hi is exact in fp32)
// and   x*y ~= hi_x*hi_y + lo_x*hi_y + hi_x*lo_y   is accumulated in fp32 in tensor memory.  The
// dropped lo*lo term is <= 2^-22 relative per product, so results agree with an fp32 FFMA GEMM to
// ~1e-6 relative (tests/test_gpu_umma.py), inside the 1e-5 forward tolerance of the parity tests.
//
// Structure (one CTA = one 128 x BN output tile, 256 threads):
//   * operands are NOT moved by TMA: they need an elementwise transform (split, optional GELU) and,
//     for weight gradients, a transpose, so the eight warps load fp32 from global memory (coalesced
//     128-bit or 32-bit accesses), transform in registers and write 16-byte vectors into shared memory
//     in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices).  LBO = 144 B and
//     SBO = 1152 B (instead of the dense 128 / 1024) skew the core matrices so that both store patterns
//     are bank-conflict free;
//   * a 3-stage ring of {A_hi, A_lo, B_hi, B_lo} K-chunks (32 fp32 of K each) fed by 8 producer warps
//     from a register double buffer (two chunks of global loads in flight per thread); producers signal
//     full[s]; a ninth warp (one thread) waits on it, issues 12 tcgen05.mma.kind::tf32 (M=128, N=BN,
//     K=8) per chunk and commits them to empty[s], which the producers wait on before reusing the stage;
//   * the 128 x BN fp32 accumulator lives in TMEM (BN columns); the epilogue reads it back with
//     tcgen05.ld (one row per thread), stages it through shared memory and writes coalesced rows with
//     the fused epilogue (bias / multiply by GELU'(aux) / accumulate / transposed store).
#include <type_traits>

#include "sml_common.cuh"

namespace {

constexpr int UM_THREADS = 256;            // 8 producer/epilogue warps: two per scheduler, enough ILP/TLP for the operand transform
constexpr int UM_CTA_THREADS = UM_THREADS + 32;   // + 1 warp that only issues tcgen05.mma
constexpr int UM_BM = 128;
constexpr int UM_BK = 32;                  // fp32 elements of K per stage
constexpr int UM_STAGES = 3;
constexpr uint32_t UM_LBO = 144;           // bytes between K-adjacent core matrices
constexpr uint32_t UM_SBO = 8 * UM_LBO;    // bytes between 8-row groups (1152)
constexpr int UM_MAX_PROBS = 4;
constexpr int UM_NACC = 4;                 // independent TMEM accumulators the K chunks are dealt to

struct UmmaParams { SmlGemmProb p[UM_MAX_PROBS]; int transpose_out; int ksplit; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}

// round-to-nearest (ties away) to the 10-bit TF32 mantissa with full-rate integer ops (cvt.rna.tf32.f32 gives
// the same bits for finite values but issues on the quarter-rate conversion pipe); NaN / Inf pass through
__device__ __forceinline__ float tf32_rna(float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t r = (b + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float((b & 0x7F800000u) == 0x7F800000u ? b : r);
}

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_NONE (layout_type 0), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(UM_LBO >> 4) << 16) | ((uint64_t)(UM_SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_tf32_h(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t d_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(d_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// bytes of one operand tile (hi or lo) with ROWS rows of the M/N dimension and UM_BK of K
template <int ROWS> struct TileBytes { static constexpr uint32_t v = (ROWS / 8) * UM_SBO; };

// ---- operand staging ---------------------------------------------------------------------------
// K-contiguous source G[x][k] (pitch ld): 8 lanes cover one row's 128 B; each thread owns ROWS/32 (row, 16 B) pieces
template <int ROWS, bool GELU>
struct LoadKContig {
    static constexpr int RSTEP = UM_THREADS / 8;       // rows covered by one pass of the CTA
    static constexpr int NP = ROWS / RSTEP;
    float4 v[NP];
    __device__ __forceinline__ void load(const float *__restrict__ G, int ld, int x0, int X, int k0, int K) {
        const int kq = threadIdx.x & 7, xr0 = threadIdx.x >> 3;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int x = x0 + xr0 + RSTEP * j, k = k0 + 4 * kq;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x < X) {
                const float *src = G + (size_t)x * ld + k;
                if (k + 3 < K) t = __ldg(reinterpret_cast<const float4 *>(src));
                else { if (k < K) t.x = src[0]; if (k + 1 < K) t.y = src[1]; if (k + 2 < K) t.z = src[2]; }
            }
            v[j] = t;
        }
    }
    __device__ __forceinline__ void store(uint8_t *hi, uint8_t *lo) const {
        const int kq = threadIdx.x & 7, xr0 = threadIdx.x >> 3;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int xr = xr0 + RSTEP * j;
            float4 t = v[j];
            if (GELU) { t.x = sml_gelu(t.x); t.y = sml_gelu(t.y); t.z = sml_gelu(t.z); t.w = sml_gelu(t.w); }
            const float4 h = make_float4(tf32_rna(t.x), tf32_rna(t.y), tf32_rna(t.z), tf32_rna(t.w));
            const float4 l = make_float4(tf32_rna(t.x - h.x), tf32_rna(t.y - h.y), tf32_rna(t.z - h.z), tf32_rna(t.w - h.w));
            const uint32_t off = (xr >> 3) * UM_SBO + kq * UM_LBO + (xr & 7) * 16;
            *reinterpret_cast<float4 *>(hi + off) = h;
            *reinterpret_cast<float4 *>(lo + off) = l;
        }
    }
};

// X-contiguous source G[k][x] (pitch ld; weight-gradient operands and [K,N] weights): a thread owns one x
// (coalesced 4-byte loads across the warp) and UM_BK * ROWS / 128 consecutive k, i.e. whole 16-byte K-quads
template <int ROWS, bool GELU>
struct LoadXContig {
    static constexpr int KPT = UM_BK * ROWS / UM_THREADS;   // k per thread: 16 (ROWS=128) or 8 (ROWS=64)
    float v[KPT];
    __device__ __forceinline__ void load(const float *__restrict__ G, int ld, int x0, int X, int k0, int K) {
        const int xr = threadIdx.x % ROWS, kb = (threadIdx.x / ROWS) * KPT;
        const int x = x0 + xr;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int k = k0 + kb + i;
            v[i] = (x < X && k < K) ? __ldg(G + (size_t)k * ld + x) : 0.f;
        }
    }
    __device__ __forceinline__ void store(uint8_t *hi, uint8_t *lo) const {
        const int xr = threadIdx.x % ROWS, kb = (threadIdx.x / ROWS) * KPT;
#pragma unroll
        for (int q = 0; q < KPT / 4; ++q) {
            float t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = GELU ? sml_gelu(v[4 * q + i]) : v[4 * q + i];
            const float4 h = make_float4(tf32_rna(t[0]), tf32_rna(t[1]), tf32_rna(t[2]), tf32_rna(t[3]));
            const float4 l = make_float4(tf32_rna(t[0] - h.x), tf32_rna(t[1] - h.y), tf32_rna(t[2] - h.z), tf32_rna(t[3] - h.w));
            const uint32_t off = (xr >> 3) * UM_SBO + (kb / 4 + q) * UM_LBO + (xr & 7) * 16;
            *reinterpret_cast<float4 *>(hi + off) = h;
            *reinterpret_cast<float4 *>(lo + off) = l;
        }
    }
};

template <int ROWS, int MODE> struct Loader;   // MODE follows SML_A_* / SML_B_* (K-contig, K-contig+GELU, X-contig[, +GELU])

// smem: [stage][A_hi | A_lo | B_hi | B_lo], then the epilogue staging tile aliases stage memory
template <int BN> struct Smem {
    static constexpr uint32_t A_BYTES = TileBytes<UM_BM>::v, B_BYTES = TileBytes<BN>::v;
    static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr uint32_t EPI = UM_BM * (BN + 1) * 4;
    static constexpr uint32_t TOTAL = (UM_STAGES * STAGE > EPI ? UM_STAGES * STAGE : EPI) + 64;
};

// A_KCONTIG: A source is [m][k]; else [k][m].   B_KCONTIG: B source is [n][k]; else [k][n].
template <int BN, bool A_KCONTIG, bool A_GELU, bool B_KCONTIG, bool B_GELU, int EPI>
__global__ void __launch_bounds__(UM_CTA_THREADS, 1)
k_umma_gemm(UmmaParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    using S = Smem<BN>;
    // split-K: blockIdx.z = problem * ksplit + slice; a slice owns a 32-aligned K range and, when there is
    // more than one slice, adds its partial product into C with atomics (EPI_ACCUM only)
    SmlGemmProb p = P.p[blockIdx.z / P.ksplit];
    {
        const int per = ((p.K + UM_BK - 1) / UM_BK + P.ksplit - 1) / P.ksplit * UM_BK;
        const int kbeg = (int)(blockIdx.z % P.ksplit) * per;
        if (kbeg >= p.K) return;
        const int kend = kbeg + per < p.K ? kbeg + per : p.K;
        p.A += A_KCONTIG ? (size_t)kbeg : (size_t)kbeg * p.lda;
        p.B += B_KCONTIG ? (size_t)kbeg : (size_t)kbeg * p.ldb;
        p.K = kend - kbeg;
    }
    const bool atomic_acc = P.ksplit > 1;
    const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
    if (m0 >= p.M || n0 >= p.N) return;                       // whole CTA exits: nothing allocated yet
    // mbarriers: full[s] (256 producer arrivals), empty[s] (1 tcgen05.commit), done (1 commit)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::TOTAL - 64);
    uint64_t *empty = full + UM_STAGES;
    uint64_t *done = empty + UM_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::TOTAL - 8);
    const int warp = threadIdx.x >> 5;
    constexpr int MMA_WARP = UM_THREADS / 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < UM_STAGES; ++i) { mbar_init(&full[i], UM_THREADS); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN * UM_NACC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    sml_pdl_trigger();
    sml_pdl_wait();

    // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
    const int nchunks = (p.K + UM_BK - 1) / UM_BK;

    if (warp == MMA_WARP) {
        // ===== MMA issuer: the whole warp runs the loop (warp-uniform: descriptors stay in uniform registers), one elected
        // lane waits-for-full / issues the 12 MMAs of a stage / commits to empty[s] =====
        constexpr uint32_t DHI = (UM_SBO >> 4) | (1u << 14), KSTEP = (2 * UM_LBO) >> 4;
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        auto dlo = [](uint32_t a) { return ((a >> 4) & 0x3FFF) | ((UM_LBO >> 4) << 16); };
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % UM_STAGES;
            mbar_wait(&full[s], (c / UM_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + s * S::STAGE);
            const uint32_t a_hi = dlo(sa), a_lo = dlo(sa + S::A_BYTES), b_hi = dlo(sa + 2 * S::A_BYTES), b_lo = dlo(sa + 2 * S::A_BYTES + S::B_BYTES);
            // K chunks alternate over UM_NACC accumulators (see umma_packed.cu: the tensor core's fp32 adder
            // truncates, so short accumulation chains summed in the epilogue are more accurate)
            const uint32_t d = tmem + (uint32_t)(c % UM_NACC) * BN;
            if (leader) {
#pragma unroll
                for (int k = 0; k < UM_BK / 8; ++k) {                            // one MMA = K 8 = two core matrices
                    umma_tf32_h(d, a_lo + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, (c >= UM_NACC) || (k != 0));   // small terms first; first MMA overwrites
                    umma_tf32_h(d, a_hi + k * KSTEP, b_lo + k * KSTEP, DHI, IDESC, 1);
                    umma_tf32_h(d, a_hi + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, 1);
                }
                umma_commit(&empty[s]);                                          // stage reusable once these MMAs retire
            }
            __syncwarp();
        }
        if (leader) umma_commit(done);
        __syncwarp();
    } else {
        // ===== producers: global fp32 -> registers (two chunks in flight) -> hi/lo split -> canonical smem =====
        using LA = typename std::conditional<A_KCONTIG, LoadKContig<UM_BM, A_GELU>, LoadXContig<UM_BM, A_GELU>>::type;
        using LB = typename std::conditional<B_KCONTIG, LoadKContig<BN, B_GELU>, LoadXContig<BN, B_GELU>>::type;
        LA la0, la1;
        LB lb0, lb1;
        auto put = [&](int c, const LA &la, const LB &lb) {
            const int s = c % UM_STAGES;
            uint8_t *st = smem + s * S::STAGE;
            if (c >= UM_STAGES) mbar_wait(&empty[s], ((c / UM_STAGES) - 1) & 1); // the MMAs that read this stage are done
            la.store(st, st + S::A_BYTES);
            lb.store(st + 2 * S::A_BYTES, st + 2 * S::A_BYTES + S::B_BYTES);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> async proxy (UMMA)
            mbar_arrive(&full[s]);
        };
        la0.load(p.A, p.lda, m0, p.M, 0, p.K);
        lb0.load(p.B, p.ldb, n0, p.N, 0, p.K);
        if (nchunks > 1) {
            la1.load(p.A, p.lda, m0, p.M, UM_BK, p.K);
            lb1.load(p.B, p.ldb, n0, p.N, UM_BK, p.K);
        }
        for (int c = 0; c < nchunks; c += 2) {
            put(c, la0, lb0);
            if (c + 2 < nchunks) {
                la0.load(p.A, p.lda, m0, p.M, (c + 2) * UM_BK, p.K);
                lb0.load(p.B, p.ldb, n0, p.N, (c + 2) * UM_BK, p.K);
            }
            if (c + 1 < nchunks) {
                put(c + 1, la1, lb1);
                if (c + 3 < nchunks) {
                    la1.load(p.A, p.lda, m0, p.M, (c + 3) * UM_BK, p.K);
                    lb1.load(p.B, p.ldb, n0, p.N, (c + 3) * UM_BK, p.K);
                }
            }
        }
    }
    // ---- epilogue ----
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float *epi = reinterpret_cast<float *>(smem);                                // [128][BN+1], aliases the stages (all MMAs done)
    if (warp < MMA_WARP) {
    // TMEM lane == accumulator row; warp w may only touch lanes [32*(w%4), +32): warps w and w+4 share a
    // lane quarter and split the columns
    const int row = (warp & 3) * 32 + (threadIdx.x & 31);
#pragma unroll
    for (int c0 = (warp >> 2) * (BN / 2); c0 < (warp >> 2) * (BN / 2) + BN / 2; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
        for (int a = 1; a < (nchunks < UM_NACC ? nchunks : UM_NACC); ++a) {
            float w[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a * BN + c0, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) epi[row * (BN + 1) + c0 + i] = v[i];
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN * UM_NACC) : "memory");
    if (!P.transpose_out) {
        // coalesced rows: a warp writes 32 consecutive n of one m
        for (int idx = threadIdx.x; idx < UM_BM * BN; idx += UM_CTA_THREADS) {
            const int r = idx / BN, cc = idx % BN;
            const int m = m0 + r, n = n0 + cc;
            if (m < p.M && n < p.N) {
                float v = epi[r * (BN + 1) + cc];
                if (EPI == SML_EPI_BIAS) v += p.bias[n];
                if (EPI == SML_EPI_MUL_GELU_GRAD) v *= sml_gelu_grad(p.aux[(size_t)m * p.ldc + n]);
                float *dst = p.C + (size_t)m * p.ldc + n;
                if (EPI == SML_EPI_ACCUM && atomic_acc) { atomicAdd(dst, v); continue; }
                if (EPI == SML_EPI_ACCUM) v += *dst;
                *dst = v;
            }
        }
    } else {
        // C is stored [n][m] (pitch ldc): a warp writes 32 consecutive m of one n
        for (int idx = threadIdx.x; idx < UM_BM * BN; idx += UM_CTA_THREADS) {
            const int cc = idx / UM_BM, r = idx % UM_BM;
            const int m = m0 + r, n = n0 + cc;
            if (m < p.M && n < p.N) {
                float v = epi[r * (BN + 1) + cc];
                float *dst = p.C + (size_t)n * p.ldc + m;
                if (EPI == SML_EPI_ACCUM && atomic_acc) { atomicAdd(dst, v); continue; }
                if (EPI == SML_EPI_ACCUM) v += *dst;
                *dst = v;
            }
        }
    }
}

template <int BN, bool AK, bool AG, bool BK_, bool BG, int EPI>
int launch_one(const UmmaParams &P, dim3 grid, cudaStream_t st) {
    auto kern = k_umma_gemm<BN, AK, AG, BK_, BG, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        SML_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<BN>::TOTAL));
        attr_set = true;
    }
    SML_CUDA_OK(sml_launch(kern, grid, dim3(UM_CTA_THREADS), Smem<BN>::TOTAL, st, P));
    SML_LAUNCH_OK();
    return SML_OK;
}

template <int BN, bool AK, bool AG, bool BK_, bool BG>
int launch_epi(const UmmaParams &P, dim3 grid, int epi, cudaStream_t st) {
    switch (epi) {
        case SML_EPI_NONE: return launch_one<BN, AK, AG, BK_, BG, SML_EPI_NONE>(P, grid, st);
        case SML_EPI_BIAS: return launch_one<BN, AK, AG, BK_, BG, SML_EPI_BIAS>(P, grid, st);
        case SML_EPI_MUL_GELU_GRAD: return launch_one<BN, AK, AG, BK_, BG, SML_EPI_MUL_GELU_GRAD>(P, grid, st);
        case SML_EPI_ACCUM: return launch_one<BN, AK, AG, BK_, BG, SML_EPI_ACCUM>(P, grid, st);
    }
    sml_set_error("umma gemm: bad epilogue %d", epi);
    return SML_E_BADARG;
}

template <int BN>
int launch_bn(const UmmaParams &P, dim3 grid, int a_mode, int b_mode, int epi, cudaStream_t st) {
    if (a_mode == SML_A_MK && b_mode == SML_B_NK) return launch_epi<BN, true, false, true, false>(P, grid, epi, st);
    if (a_mode == SML_A_MK_GELU && b_mode == SML_B_NK) return launch_epi<BN, true, true, true, false>(P, grid, epi, st);
    if (a_mode == SML_A_MK && b_mode == SML_B_KN) return launch_epi<BN, true, false, false, false>(P, grid, epi, st);
    if (a_mode == SML_A_KM && b_mode == SML_B_KN) return launch_epi<BN, false, false, false, false>(P, grid, epi, st);
    if (a_mode == SML_A_KM_GELU && b_mode == SML_B_KN) return launch_epi<BN, false, true, false, false>(P, grid, epi, st);
    if (a_mode == SML_A_KM && b_mode == SML_B_KN_GELU) return launch_epi<BN, false, false, false, true>(P, grid, epi, st);
    sml_set_error("umma gemm: unsupported mode combination (%d, %d)", a_mode, b_mode);
    return SML_E_BADARG;
}

}  // namespace

// Same contract as sml_launch_sgemm plus transpose_out (C stored [n][m]); bn = 64 or 128.
int sml_launch_umma_gemm(const SmlGemmProb *probs, int n_probs, int a_mode, int b_mode, int epi, int transpose_out, int bn,
                         cudaStream_t st, int ksplit) {
    SML_REQUIRE(ksplit >= 1 && (ksplit == 1 || epi == SML_EPI_ACCUM), SML_E_BADARG, "umma gemm: split-K needs the accumulate epilogue");
    SML_REQUIRE(n_probs >= 1 && n_probs <= UM_MAX_PROBS, SML_E_BADARG, "umma gemm: bad problem count %d", n_probs);
    SML_REQUIRE(bn == 64 || bn == 128, SML_E_BADARG, "umma gemm: bn must be 64 or 128");
    SML_REQUIRE(!(transpose_out && (epi == SML_EPI_BIAS || epi == SML_EPI_MUL_GELU_GRAD)), SML_E_BADARG,
                "umma gemm: transposed store supports only the plain / accumulate epilogues");
    UmmaParams P;
    P.transpose_out = transpose_out;
    P.ksplit = ksplit;
    int maxM = 0, maxN = 0;
    for (int i = 0; i < n_probs; ++i) {
        P.p[i] = probs[i];
        if (probs[i].M > maxM) maxM = probs[i].M;
        if (probs[i].N > maxN) maxN = probs[i].N;
        SML_REQUIRE((probs[i].lda % 4) == 0 && (probs[i].ldb % 4) == 0, SML_E_BADARG, "umma gemm: lda/ldb must be multiples of 4");
        SML_REQUIRE(probs[i].K >= 1, SML_E_BADARG, "umma gemm: K must be positive");
    }
    if (maxM == 0 || maxN == 0) return SML_OK;
    dim3 grid((maxN + bn - 1) / bn, (maxM + UM_BM - 1) / UM_BM, n_probs * ksplit);
    return bn == 64 ? launch_bn<64>(P, grid, a_mode, b_mode, epi, st) : launch_bn<128>(P, grid, a_mode, b_mode, epi, st);
}
