// fp32-accurate GEMM on the tcgen05 tensor cores with PRE-PACKED operands (umma_pack.cuh): the four
// GEMMs on the critical chain of every step and of the full-table transfer
//     fc1:  Z1 = A W1^T + b1      (A from the conv prologue)         -> plain Z1 + packed GELU(Z1)
//     fc2:  Y  = GELU(Z1) W2^T + b2                                   -> plain Y
//     d2 :  dZ1 = (dY W2) * GELU'(Z1)                                 -> plain dZ1 + packed dZ1
//     d1 :  dA  = dZ1 W1                                              -> plain dA
// (model/conv_transfer.py:47-49 of the reference and their autograd).  The 3xTF32 split
// (x = hi + lo, hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM) is done ONCE per element by whoever
// produces the operand (conv kernel, loss kernel, the previous GEMM's epilogue, the theta packer), so
// this kernel moves operands with one cp.async.bulk per 36 KB block and spends no instructions on them.
//
// One CTA = one 128 x BN output tile, 10 warps:
//   warp 0 (1 thread)  bulk-copy producer: 3-stage ring, mbarrier expect_tx / complete_tx
//   warp 1 (1 thread)  MMA issuer: 12 tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per 32-wide K chunk,
//                      tcgen05.commit frees the stage; owns the TMEM allocation (BN fp32 columns)
//   warps 2-9          epilogue: tcgen05.ld (one accumulator row per thread) -> fused bias / GELU / GELU'
//                      -> 128-byte-contiguous packed stores for the next GEMM (+ plain row stores)
#include "sml_common.cuh"
#include <string.h>

#include "umma_pack.cuh"

namespace {

constexpr int PKG_THREADS = 320;
constexpr int PKG_EPI_THREADS = 256;
constexpr int PKG_STAGES = 3;
constexpr int PKG_MAX_PROBS = 4;
// The tensor core adds its products into the fp32 TMEM accumulator with a truncating adder: the error of one
// accumulator grows ~linearly with the number of K steps chained into it (measured 0.7e-6 relative at K = 64, 2.4e-6
// at 320, 4e-6 at 512).  K chunks are therefore dealt round-robin to PKG_NACC independent accumulators that the
// epilogue sums with round-to-nearest adds, which brings the GEMMs back to FFMA-class accuracy.
constexpr int PKG_NACC = 4;

struct PkParams { SmlPkProb p[PKG_MAX_PROBS]; int ksplit; SmlPkLoss L; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_NONE, version 1; LBO / SBO in 16-byte units
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(PK_LBO >> 4) << 16) | ((uint64_t)(PK_SBO >> 4) << 32) | (1ull << 46);
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

template <int BN, bool FUSE = false> struct PkSmem {
    static constexpr uint32_t A_BYTES = pk_block_bytes(128), B_BYTES = pk_block_bytes(BN);
    static constexpr uint32_t STAGE = A_BYTES + B_BYTES;
    static constexpr uint32_t W2_BYTES = FUSE ? 2 * pk_block_bytes(64) : 0;      // fused fc2: this CTA's 64-wide K slice of W2 (2 chunks)
    static constexpr uint32_t RED_BYTES = 384;                                    // fused loss: block reduction scratch
    static constexpr uint32_t TOTAL = PKG_STAGES * STAGE + W2_BYTES + RED_BYTES + 128;
};

// loss terms of one triple from its two scores (the arithmetic of k_loss, conv.cu; model/conv_transfer.py:120-134)
__device__ __forceinline__ void pk_loss_terms(int loss_kind, float sp, float sn, float invB, float &dsp, float &dsn, float &lpos, float &lneg) {
    if (loss_kind == SML_LOSS_BCE) {
        const float gp = sml_sigmoid(sp), gn = sml_sigmoid(sn);
        const float ap = gp + 1e-15f, an = (1.0f - gn) + 1e-15f;     // conv_transfer.py:124-125
        lpos = logf(ap); lneg = logf(an);
        dsp = -(gp * (1.0f - gp)) / ap * invB;
        dsn = (gn * (1.0f - gn)) / an * invB;
    } else {
        const float x = sp - sn;                                      // :128
        lpos = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));              // -logsigmoid(x) = softplus(-x)
        lneg = 0.f;
        dsp = -sml_sigmoid(-x);
        dsn = -dsp;
    }
}

// BROWS = rows per block of the packed B operand in memory.  BROWS == BN: one bulk copy per block.  BROWS = 128 with BN = 64
// (small batches: twice the CTAs, each pulling less through its SM's L2 port and running half the epilogue): the 64 rows
// of a tile are 8 consecutive 8-row groups inside the hi half and inside the lo half of a 128-row block, i.e. two copies.
// FUSE (fc1 only, BN = 64): the CTA goes on to multiply its 128 x 64 tile of GELU(Z1) -- kept in shared memory as the next
// A operand instead of being written to HBM in packed form -- with the matching 64-wide K slice of W2 and adds the partial
// Y[128 x 64] into the (pre-zeroed) fc2 output with fp32 atomics: fc1 and fc2 in one launch, an 8-way split-K fc2.
// LOSS (d2 only): the A operand (dY, K = 64: two chunks) is not read from memory but computed by the epilogue warps from Y before
// the main loop -- loss + dL/dY fused into the GEMM that consumes them (SmlPkLoss, sml_common.cuh).
template <int BN, int EPI, int BROWS = BN, bool FUSE = false, bool LOSS = false>
__global__ void __launch_bounds__(PKG_THREADS, 1)
k_umma_packed(PkParams P) {
    static_assert(BROWS == BN || (BROWS == 128 && BN == 64), "unsupported packed-B block shape");
    static_assert(!FUSE || (EPI == SML_PK_FC1 && BN == 64), "fc2 can only be fused into the 128 x 64 fc1 tiles");
    static_assert(!LOSS || (EPI == SML_PK_D2 && !FUSE), "the loss can only be fused into the d2 GEMM");
    extern __shared__ __align__(128) uint8_t smem[];
    using S = PkSmem<BN, FUSE>;
    // split-K: blockIdx.z = problem * ksplit + slice.  A single CTA streaming K = 512 pulls 0.6 MB through one SM's
    // L2 port (~7 us); slicing K spreads that over more SMs.  Slices add their partial tile with fp32 atomics into a
    // pre-zeroed C (plain-output epilogues only); slice 0 adds the bias.
    const SmlPkProb p = P.p[blockIdx.z / P.ksplit];
    const int ks = blockIdx.z % P.ksplit;
    const int tile_m = blockIdx.y, tile_n = blockIdx.x;
    if (tile_m >= p.m_tiles || tile_n * BN >= p.N) return;                    // nothing allocated yet
    const int c_per = (p.KC + P.ksplit - 1) / P.ksplit;
    const int c_beg = ks * c_per, c_end = (c_beg + c_per < p.KC) ? c_beg + c_per : p.KC;
    if (c_beg >= c_end) return;
    const bool split = P.ksplit > 1;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::TOTAL - 128);
    uint64_t *empty = full + PKG_STAGES;
    uint64_t *done = empty + PKG_STAGES;
    uint64_t *w2_full = done + 1, *g_full = done + 2, *y_done = done + 3;          // fused fc2 only
    uint64_t *a_full = done + 2;                                                   // fused loss only: the dY operand is in shared memory
    float *s_red = reinterpret_cast<float *>(smem + PKG_STAGES * S::STAGE + S::W2_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::TOTAL - 8);
    uint8_t *w2s = smem + PKG_STAGES * S::STAGE;                                    // W2 slice (fused fc2)
    constexpr int TMEM_COLS = FUSE ? 512 : BN * PKG_NACC;                           // + 64 columns for the partial Y
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < PKG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        if (FUSE) { mbar_init(w2_full, 1); mbar_init(g_full, PKG_EPI_THREADS); mbar_init(y_done, 1); }
        if (LOSS) mbar_init(a_full, PKG_EPI_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int KC = p.KC;
    sml_pdl_trigger();      // let the next kernel of the chain start its own prologue
    sml_pdl_wait();         // barriers and TMEM are set up; only now wait for the producer of our operands

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t *a = p.A + (size_t)(p.a_tile0 + tile_m) * KC * S::A_BYTES;
            constexpr uint32_t B_SRC = pk_block_bytes(BROWS);
            const uint8_t *b = p.B + (size_t)(tile_n * BN / BROWS) * KC * B_SRC + (size_t)((tile_n * BN) % BROWS / 8) * PK_SBO;
            if (FUSE) {     // W2[:, 64 tile_n .. +64) as two 64-row blocks of the packed P2 operand: needed only after the main loop
                mbar_expect_tx(w2_full, S::W2_BYTES);
                bulk_g2s(w2s, p.B2 + (size_t)(2 * tile_n) * pk_block_bytes(64), S::W2_BYTES, w2_full);
            }
            for (int c = c_beg; c < c_end; ++c) {
                const int i = c - c_beg, s = i % PKG_STAGES;
                if (i >= PKG_STAGES) mbar_wait(&empty[s], ((i / PKG_STAGES) - 1) & 1);
                uint8_t *st = smem + s * S::STAGE;
                mbar_expect_tx(&full[s], LOSS ? S::B_BYTES : S::STAGE);
                if (!LOSS) bulk_g2s(st, a + (size_t)c * S::A_BYTES, S::A_BYTES, &full[s]);
                if (BROWS == BN) {
                    bulk_g2s(st + S::A_BYTES, b + (size_t)c * B_SRC, S::B_BYTES, &full[s]);
                } else {
                    bulk_g2s(st + S::A_BYTES, b + (size_t)c * B_SRC, pk_half_bytes(BN), &full[s]);
                    bulk_g2s(st + S::A_BYTES + pk_half_bytes(BN), b + (size_t)c * B_SRC + pk_half_bytes(BROWS), pk_half_bytes(BN), &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp runs the loop (warp-uniform control flow keeps the descriptors in uniform registers), one elected lane
        // issues: ~3 instructions per tcgen05.mma instead of ~16
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t DHI = (PK_SBO >> 4) | (1u << 14), KSTEP = (2 * PK_LBO) >> 4;
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        if (LOSS) mbar_wait(a_full, 0);                                            // both dY chunks written by the epilogue warps
        for (int c = 0; c < c_end - c_beg; ++c) {
            const int s = c % PKG_STAGES;
            mbar_wait(&full[s], (c / PKG_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + s * S::STAGE);
            const uint32_t a_hi = ((sa >> 4) & 0x3FFF) | ((PK_LBO >> 4) << 16), a_lo = (((sa + pk_half_bytes(128)) >> 4) & 0x3FFF) | ((PK_LBO >> 4) << 16);
            const uint32_t b_hi = (((sa + S::A_BYTES) >> 4) & 0x3FFF) | ((PK_LBO >> 4) << 16),
                           b_lo = (((sa + S::A_BYTES + pk_half_bytes(BN)) >> 4) & 0x3FFF) | ((PK_LBO >> 4) << 16);
            const uint32_t d = tmem + (uint32_t)(c % PKG_NACC) * BN;               // accumulator of this chunk
            if (leader) {
#pragma unroll
                for (int k = 0; k < PK_BK / 8; ++k) {
                    umma_tf32_h(d, a_lo + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, (c >= PKG_NACC) || (k != 0));   // its first MMA overwrites
                    umma_tf32_h(d, a_hi + k * KSTEP, b_lo + k * KSTEP, DHI, IDESC, 1);
                    umma_tf32_h(d, a_hi + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, 1);
                }
                umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (leader) umma_commit(done);
        __syncwarp();
        if (FUSE) {
            // fc2 partial: Y[128 x 64] = GELU(Z1)[128 x 64 of this tile] W2[:, slice]^T, A operand written by the epilogue warps
            constexpr uint32_t ID64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            mbar_wait(w2_full, 0);
            mbar_wait(g_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ga = smem_u32(smem), wb = smem_u32(w2s);
            auto dlo = [](uint32_t a) { return ((a >> 4) & 0x3FFF) | ((PK_LBO >> 4) << 16); };
            if (leader) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t a_hi = dlo(ga + c * pk_block_bytes(128)), a_lo = dlo(ga + c * pk_block_bytes(128) + pk_half_bytes(128));
                    const uint32_t b_hi = dlo(wb + c * pk_block_bytes(64)), b_lo = dlo(wb + c * pk_block_bytes(64) + pk_half_bytes(64));
#pragma unroll
                    for (int k = 0; k < PK_BK / 8; ++k) {
                        umma_tf32_h(tmem + BN * PKG_NACC, a_lo + k * KSTEP, b_hi + k * KSTEP, DHI, ID64, (c | k) != 0);
                        umma_tf32_h(tmem + BN * PKG_NACC, a_hi + k * KSTEP, b_lo + k * KSTEP, DHI, ID64, 1);
                        umma_tf32_h(tmem + BN * PKG_NACC, a_hi + k * KSTEP, b_hi + k * KSTEP, DHI, ID64, 1);
                    }
                }
                umma_commit(y_done);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue warps: straight from TMEM registers, one accumulator row per thread =====
        // d2: the Z1 values of this thread's first 32 columns (the GELU' factor of the epilogue) do not depend on the MMAs: request them
        // now, one L2 round trip ahead of their use (the loss prologue and the MMAs run under it)
        float4 zpre[8];
        if (EPI == SML_PK_D2) {
            const int pr = (warp & 3) * 32 + lane, pc0 = ((warp - 2) >> 2) * (BN / 2);
            const bool pvalid = tile_m * 128 + pr < p.M;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                zpre[q] = pvalid ? __ldg(reinterpret_cast<const float4 *>(p.aux + (p.row0 + (int64_t)tile_m * 128 + pr) * p.ldc + tile_n * BN + pc0) + q)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (LOSS) {
            // ===== loss + dL/dY for the 128 rows of this tile, written as the packed A operand (two 32-column K chunks of dY) =====
            // Eight lanes per row (lane sub = lane & 7 owns columns 8 sub .. 8 sub + 7: every row is read as one coalesced 256 B
            // line), four rows per warp instruction, 16 rows per warp; all loads of a warp are issued before the first use, so the
            // prologue costs one L2 round trip plus ~4 x 150 instructions of sigmoid / log arithmetic per warp.
            const SmlPkLoss &L = P.L;
            const int lw = warp - 2, lg = lane >> 3, sub = lane & 7;
            const bool item_net = p.row0 >= L.row_pos;                  // row_pos = padded user rows > 0
            const bool lead = tile_n == 0;                              // this tile's CTA that emits what k_loss emitted
            const float invB = 1.0f / (float)L.B;
            float gb[8];                                                // fc2 bias gradient: column sums of dY (this lane's 8 columns)
#pragma unroll
            for (int i = 0; i < 8; ++i) gb[i] = 0.f;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;               // loss sums (sub == 0 lanes, user tiles)
            float4 u4[4][2], i4[4][2], j4[4][2];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int qrow = tile_m * 128 + lw * 16 + it * 4 + lg;  // row inside this net's rows
                const int kind = item_net ? (qrow < L.B ? 1 : 2) : 0;   // 0 user, 1 positive, 2 negative
                const int64_t b = qrow < p.M ? ((kind == 2) ? qrow - L.B : qrow) : 0;
                const float4 *yu = reinterpret_cast<const float4 *>(L.Y + b * SML_D) + 2 * sub;
                const float4 *yi = reinterpret_cast<const float4 *>(L.Y + (L.row_pos + b) * SML_D) + 2 * sub;
                const float4 *yj = reinterpret_cast<const float4 *>(L.Y + (L.row_neg + b) * SML_D) + 2 * sub;
                u4[it][0] = __ldcg(yu); u4[it][1] = __ldcg(yu + 1);
                i4[it][0] = __ldcg(yi); i4[it][1] = __ldcg(yi + 1);
                j4[it][0] = __ldcg(yj); j4[it][1] = __ldcg(yj + 1);
            }
            auto sum8 = [](float v) {                                   // over the 8 lanes of a row
                v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
                return v;
            };
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int lr = lw * 16 + it * 4 + lg;
                const int qrow = tile_m * 128 + lr;
                const bool ok = qrow < p.M;
                const int kind = item_net ? (qrow < L.B ? 1 : 2) : 0;
                const int64_t b = (kind == 2) ? qrow - L.B : qrow;
                float u[8] = {u4[it][0].x, u4[it][0].y, u4[it][0].z, u4[it][0].w, u4[it][1].x, u4[it][1].y, u4[it][1].z, u4[it][1].w};
                const float vi[8] = {i4[it][0].x, i4[it][0].y, i4[it][0].z, i4[it][0].w, i4[it][1].x, i4[it][1].y, i4[it][1].z, i4[it][1].w};
                const float vj[8] = {j4[it][0].x, j4[it][0].y, j4[it][0].z, j4[it][0].w, j4[it][1].x, j4[it][1].y, j4[it][1].z, j4[it][1].w};
                float inv_n = 1.0f;
                if (L.normalize_user) {   // ConvTransfer 'user': x / ||x||.detach()  (conv_transfer.py:62-63)
                    float q = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) q = fmaf(u[i], u[i], q);
                    const float nrm = sqrtf(sum8(q));
#pragma unroll
                    for (int i = 0; i < 8; ++i) u[i] = u[i] / nrm;
                    inv_n = 1.0f / nrm;
                }
                float sp = 0.f, sn = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) { sp = fmaf(u[i], vi[i], sp); sn = fmaf(u[i], vj[i], sn); }
                sp = sum8(sp); sn = sum8(sn);
                float dsp, dsn, lpos, lneg;
                pk_loss_terms(L.loss_kind, sp, sn, invB, dsp, dsn, lpos, lneg);
                float d[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float du = (dsp * vi[i] + dsn * vj[i]) * inv_n;
                    d[i] = !ok ? 0.f : (kind == 0 ? du : (kind == 1 ? dsp : dsn) * u[i]);
                }
                // columns 8 sub .. 8 sub + 7 = K chunk sub / 4, quads 2 (sub % 4) and 2 (sub % 4) + 1
                uint8_t *blk = smem + (size_t)(sub >> 2) * S::STAGE + (uint32_t)(lr >> 3) * PK_SBO + (uint32_t)(2 * (sub & 3)) * PK_LBO + (uint32_t)(lr & 7) * 16;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) pk_split(d[4 * q + i], h[i], l[i]);
                    *reinterpret_cast<float4 *>(blk + q * PK_LBO) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(blk + q * PK_LBO + pk_half_bytes(128)) = make_float4(l[0], l[1], l[2], l[3]);
                }
                if (lead && ok) {
                    if (L.dY) {
                        float4 *dst = reinterpret_cast<float4 *>(L.dY + (p.row0 + qrow) * SML_D) + 2 * sub;
                        dst[0] = make_float4(d[0], d[1], d[2], d[3]); dst[1] = make_float4(d[4], d[5], d[6], d[7]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) gb[i] += d[i];
                    if (kind == 0 && sub == 0) {
                        a0 += lpos; a1 += lneg;
                        if (L.rowsq) {
                            a2 += __ldcg(L.rowsq + b) + __ldcg(L.rowsq + L.row_pos + b) + __ldcg(L.rowsq + L.row_neg + b);
                            if (L.adaptive != 0.f) a3 += sqrtf(__ldcg(L.rowsq + b));       // transfer.py:490-499
                        }
                        if (L.scores) { L.scores[b] = sp; L.scores[L.B + b] = sn; }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(a_full)) : "memory");
            if (lead) {
                // s_red: [0, 32) loss sums of the 8 warps, [32, 96) the 64 column sums of dY
                a0 += __shfl_xor_sync(0xffffffffu, a0, 8); a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
                a1 += __shfl_xor_sync(0xffffffffu, a1, 8); a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
                a2 += __shfl_xor_sync(0xffffffffu, a2, 8); a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
                a3 += __shfl_xor_sync(0xffffffffu, a3, 8); a3 += __shfl_xor_sync(0xffffffffu, a3, 16);
                if (lane == 0) { s_red[4 * lw] = a0; s_red[4 * lw + 1] = a1; s_red[4 * lw + 2] = a2; s_red[4 * lw + 3] = a3; }
                if (L.gb_user && lw == 0) { s_red[32 + lane] = 0.f; s_red[64 + lane] = 0.f; }
                asm volatile("bar.sync 1, %0;" ::"n"(PKG_EPI_THREADS) : "memory");
                if (L.gb_user) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        gb[i] += __shfl_xor_sync(0xffffffffu, gb[i], 8); gb[i] += __shfl_xor_sync(0xffffffffu, gb[i], 16);
                    }
                    if (lane < 8) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) atomicAdd(&s_red[32 + 8 * sub + i], gb[i]);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(PKG_EPI_THREADS) : "memory");
                    if (lw < 2) atomicAdd((item_net ? L.gb_item : L.gb_user) + 32 * lw + lane, s_red[32 + 32 * lw + lane]);
                }
                if (!item_net && lw == 0 && lane == 0) {
                    // scalar loss: fixed-order sums (warp -> CTA -> partials[] -> the last user tile to arrive), like k_loss
                    float t4[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int ww = 0; ww < 8; ++ww) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) t4[k] += s_red[4 * ww + k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) L.partials[4 * tile_m + k] = t4[k];
                    __threadfence();
                    const unsigned n_user_tiles = (unsigned)P.p[0].m_tiles;
                    if (atomicAdd(L.ticket, 1u) == n_user_tiles - 1) {
                        __threadfence();
                        float ps = 0.f, ns = 0.f, qs = 0.f, ad = 0.f;
                        for (unsigned i = 0; i < n_user_tiles; ++i) {
                            ps += __ldcg(L.partials + 4 * i); ns += __ldcg(L.partials + 4 * i + 1);
                            qs += __ldcg(L.partials + 4 * i + 2); ad += __ldcg(L.partials + 4 * i + 3);
                        }
                        float loss = (L.loss_kind == SML_LOSS_BCE) ? (-(ps * invB)) + (-(ns * invB)) : ps;   // -mean - mean | -sum(logsigmoid)
                        loss = loss + L.l2 * (0.5f * qs);                          // transfer.py:486-488
                        if (L.adaptive != 0.f) loss = loss + L.adaptive * ad;       // :490-499
                        L.loss_out[0] = loss;
                        L.loss_out[1] += loss;
                        *L.ticket = 0;                                              // re-arm for the next launch (graph replay)
                    }
                }
            }
        }
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ew = warp - 2;                               // 0..7
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;                     // accumulator row of this thread
        const int chalf = ew >> 2;                             // the two warps of a quarter split the columns
        const int n0 = tile_n * BN;
        const int64_t m = p.row0 + (int64_t)tile_m * 128 + r;
        const bool valid = tile_m * 128 + r < p.M;
#pragma unroll
        for (int c0 = chalf * (BN / 2); c0 < chalf * (BN / 2) + BN / 2; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c0, v);
            {
                const int nacc = (c_end - c_beg) < PKG_NACC ? (c_end - c_beg) : PKG_NACC;   // accumulators actually written
                for (int a = 1; a < nacc; ++a) {
                    float w[32];
                    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + a * BN + c0, w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += w[i];
                }
            }
            const int nb = n0 + c0;
            if ((EPI == SML_PK_FC1 || EPI == SML_PK_FC2) && ks == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + nb) + q);
                    v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
                }
            }
            if (EPI == SML_PK_D2) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 == chalf * (BN / 2)) z = zpre[q];
                    else if (valid) z = __ldg(reinterpret_cast<const float4 *>(p.aux + m * p.ldc + nb) + q);
                    v[4 * q] *= sml_gelu_grad(z.x); v[4 * q + 1] *= sml_gelu_grad(z.y);
                    v[4 * q + 2] *= sml_gelu_grad(z.z); v[4 * q + 3] *= sml_gelu_grad(z.w);
                }
            }
            if (EPI == SML_PK_D2 && p.colsum) {
                // fc1 bias gradient = column sums of dZ1: transpose-reduce the warp's 32 rows x 32 columns with 31
                // shuffles (lane l ends with the sum of column l), one atomicAdd per lane
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = valid ? v[i] : 0.f;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const float send = up ? t[i] : t[i + off];
                        const float keep = up ? t[i + off] : t[i];
                        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                atomicAdd(p.colsum + nb + lane, t[0]);
            }
            if (p.C && valid) {
                // a thread writes 128 contiguous bytes of its row (the row pitch separates the lanes)
                if (split) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(p.C + m * p.ldc + nb + i, v[i]);
                } else {
                    float4 *dst = reinterpret_cast<float4 *>(p.C + m * p.ldc + nb);
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            if ((EPI == SML_PK_FC1 || EPI == SML_PK_D2) && (p.Cpk || FUSE)) {
                // column nb.. of this GEMM = K index of the next one: one 32-wide K chunk, 8 quads; lanes r%8
                // are 16 B apart => every store instruction writes whole 128 B core matrices
                // (fused fc2: the chunk goes to shared memory -- the pipeline stages are free once `done` has fired)
                uint8_t *blk = FUSE ? smem + (size_t)(c0 >> 5) * pk_block_bytes(128) + (uint32_t)(r >> 3) * PK_SBO + (uint32_t)(r & 7) * 16
                                    : p.Cpk + ((size_t)(p.c_tile0 + tile_m) * (p.N / PK_BK) + (nb >> 5)) * pk_block_bytes(128) +
                                          (uint32_t)(r >> 3) * PK_SBO + (uint32_t)(r & 7) * 16;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = (EPI == SML_PK_FC1) ? sml_gelu(v[4 * q + i]) : v[4 * q + i];
                        pk_split(x, h[i], l[i]);
                    }
                    *reinterpret_cast<float4 *>(blk + q * PK_LBO) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(blk + q * PK_LBO + pk_half_bytes(128)) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    }
    if (FUSE && warp >= 2) {
        // hand the GELU(Z1) tile to the tensor core, then add the partial fc2 product into Y
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(g_full)) : "memory");
        mbar_wait(y_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;
        const int64_t m = p.row0 + (int64_t)tile_m * 128 + r;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + BN * PKG_NACC + chalf * 32, v);
        if (tile_m * 128 + r < p.M) {
            float *y = p.Y + m * p.ldy + chalf * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(y + i, tile_n == 0 ? v[i] + __ldg(p.bias2 + chalf * 32 + i) : v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ---- theta packer ---------------------------------------------------------------------------------
// The four weight operands of one net, split and packed:  P1 = W1 as B[n=512][k=320] (128-row blocks),
// P2 = W2 as B[64][512] (64-row), P3 = W2^T as B[512][64] (128-row), P4 = W1^T as B[320][512] (64-row).
__global__ void __launch_bounds__(256) k_pack_theta(const float *__restrict__ theta, uint8_t *__restrict__ out, int n_nets,
                                                    int64_t *adam_state, double lr) {
    sml_pdl_wait();
    sml_pdl_trigger();
    if (adam_state && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        // the step's Adam tick (sml_adam_tick) rides along: step counter + step_size + sqrt(bias_correction2)
        sml_adam_tick_body(adam_state, lr, 0.9, 0.999);
    }
    const int net = blockIdx.y >> 2, which = blockIdx.y & 3;
    if (net >= n_nets) return;
    const float *th = theta + (size_t)net * SML_NET_STRIDE;
    uint8_t *dst = out + (size_t)net * SML_PK_THETA_BYTES;
    const float *W; int R, K, ld, rows; bool tr;
    if (which == 0) { W = th + SML_OFF_F1W; R = 512; K = 320; ld = 320; rows = 128; tr = false; dst += SML_PK_OFF_P1; }
    else if (which == 1) { W = th + SML_OFF_F2W; R = 64; K = 512; ld = 512; rows = 64; tr = false; dst += SML_PK_OFF_P2; }
    else if (which == 2) { W = th + SML_OFF_F2W; R = 512; K = 64; ld = 512; rows = 128; tr = true; dst += SML_PK_OFF_P3; }
    else { W = th + SML_OFF_F1W; R = 320; K = 512; ld = 320; rows = 64; tr = true; dst += SML_PK_OFF_P4; }
    const int KQ = K / 4, KC = K / 32;
    const int total = R * KQ;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int n, q;
        float x[4];
        if (!tr) {            // B[n][k] = W[n][k]: lanes along k (coalesced 16 B reads)
            q = idx % KQ; n = idx / KQ;
            const float4 t = __ldg(reinterpret_cast<const float4 *>(W + (size_t)n * ld + 4 * q));
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {              // B[n][k] = W[k][n]: lanes along n (coalesced 4 B reads, 128 B-contiguous packed stores)
            n = idx % R; q = idx / R;
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = __ldg(W + (size_t)(4 * q + i) * ld + n);
        }
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pk_split(x[i], h[i], l[i]);
        const int tile = n / rows, r = n % rows, k = 4 * q;
        uint8_t *blk = dst + ((size_t)tile * KC + (k >> 5)) * pk_block_bytes(rows) + pk_elem_off(r, k & 31);
        *reinterpret_cast<float4 *>(blk) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(blk + pk_half_bytes(rows)) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

template <int BN, int EPI, int BROWS = BN, bool FUSE = false, bool LOSS = false>
int launch_pk(const PkParams &P, dim3 grid, cudaStream_t st) {
    auto kern = k_umma_packed<BN, EPI, BROWS, FUSE, LOSS>;
    static bool attr_set = false;
    if (!attr_set) {
        SML_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PkSmem<BN, FUSE>::TOTAL));
        attr_set = true;
    }
    SML_CUDA_OK(sml_launch(kern, grid, dim3(PKG_THREADS), PkSmem<BN, FUSE>::TOTAL, st, P));
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // namespace

int sml_launch_umma_packed(const SmlPkProb *probs, int n_probs, int epi, cudaStream_t st, int ksplit, const SmlPkLoss *loss) {
    SML_REQUIRE(n_probs >= 1 && n_probs <= PKG_MAX_PROBS, SML_E_BADARG, "packed gemm: bad problem count %d", n_probs);
    SML_REQUIRE(ksplit >= 1 && (ksplit == 1 || epi == SML_PK_FC2 || epi == SML_PK_D1), SML_E_BADARG,
                "packed gemm: split-K only for the plain-output epilogues (fc2, d1)");
    SML_REQUIRE(!loss || (epi == SML_PK_D2 && n_probs == 2 && probs[0].KC == 2 && probs[0].row0 == 0 && loss->row_pos > 0), SML_E_BADARG,
                "packed gemm: the fused loss needs the d2 GEMM over the user net (first) and the item net");
    PkParams P;
    P.ksplit = ksplit;
    if (loss) P.L = *loss; else memset(&P.L, 0, sizeof(P.L));
    int max_mt = 0;
    const int N = probs[0].N;
    for (int i = 0; i < n_probs; ++i) {
        P.p[i] = probs[i];
        if (probs[i].m_tiles > max_mt) max_mt = probs[i].m_tiles;
        SML_REQUIRE(probs[i].N == N && probs[i].KC >= 1, SML_E_BADARG, "packed gemm: grouped problems must share N and have KC >= 1");
    }
    if (max_mt == 0) return SML_OK;
    // few row tiles (the B = 256 transfer step has 6): 128 x 64 tiles double the CTAs of the N = 512 GEMMs (24 -> 48), each
    // pulls 25 % less through its SM's L2 port and runs half the epilogue.  sml_debug_set_mask(512) keeps the wide tiles.
    int tot_mt = 0;
    for (int i = 0; i < n_probs; ++i) tot_mt += probs[i].m_tiles;
    const bool narrow = tot_mt <= 12 && !(sml_debug_mask() & 512);
    switch (epi) {
        case SML_PK_FC1_FC2:     // fc1 with fc2 fused into its epilogue (128 x 64 tiles); probs[i].B2 / bias2 / Y describe fc2
            SML_REQUIRE(N == 512 && ksplit == 1, SML_E_BADARG, "packed gemm: fused fc1 + fc2 needs N = 512 and no split-K");
            for (int i = 0; i < n_probs; ++i)
                SML_REQUIRE(probs[i].B2 && probs[i].bias2 && probs[i].Y, SML_E_BADARG, "packed gemm: fused fc1 + fc2 needs B2, bias2 and Y");
            return launch_pk<64, SML_PK_FC1, 128, true>(P, dim3(N / 64, max_mt, n_probs), st);
        case SML_PK_FC1:
            SML_REQUIRE(N % 128 == 0, SML_E_BADARG, "packed gemm: fc1 N");
            if (narrow) return launch_pk<64, SML_PK_FC1, 128>(P, dim3(N / 64, max_mt, n_probs), st);
            return launch_pk<128, SML_PK_FC1>(P, dim3(N / 128, max_mt, n_probs), st);
        case SML_PK_FC2: SML_REQUIRE(N % 64 == 0, SML_E_BADARG, "packed gemm: fc2 N"); return launch_pk<64, SML_PK_FC2>(P, dim3(N / 64, max_mt, n_probs * ksplit), st);
        case SML_PK_D2:
            SML_REQUIRE(N % 128 == 0, SML_E_BADARG, "packed gemm: d2 N");
            if (loss) {
                if (narrow) return launch_pk<64, SML_PK_D2, 128, false, true>(P, dim3(N / 64, max_mt, n_probs), st);
                return launch_pk<128, SML_PK_D2, 128, false, true>(P, dim3(N / 128, max_mt, n_probs), st);
            }
            if (narrow) return launch_pk<64, SML_PK_D2, 128>(P, dim3(N / 64, max_mt, n_probs), st);
            return launch_pk<128, SML_PK_D2>(P, dim3(N / 128, max_mt, n_probs), st);
        case SML_PK_D1: SML_REQUIRE(N % 64 == 0, SML_E_BADARG, "packed gemm: d1 N"); return launch_pk<64, SML_PK_D1>(P, dim3(N / 64, max_mt, n_probs * ksplit), st);
    }
    sml_set_error("packed gemm: bad epilogue %d", epi);
    return SML_E_BADARG;
}

int sml_launch_pack_theta(const float *theta, uint8_t *out, int n_nets, cudaStream_t st, int64_t *adam_state, double lr) {
    dim3 grid(40, 4 * n_nets);
    SML_CUDA_OK(sml_launch(k_pack_theta, grid, dim3(256), 0, st, theta, out, n_nets, adam_state, lr));
    SML_LAUNCH_OK();
    return SML_OK;
}
