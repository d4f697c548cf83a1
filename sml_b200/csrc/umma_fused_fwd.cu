// The transfer network forward as ONE persistent streaming kernel (BASELINE.json north_star item 2):
//     out[n] = fc2( GELU( fc1( GELU(conv2(GELU(conv1( stack(x_t[n], x_hat[n], x_t*x_hat/||x_t||) )))) ) ) )
// (one_transfer.forward + ConvTransfer_com.forward / ConvTransfer.forward, model/conv_transfer.py:37-50,57-69,92-110;
// meta_train.updata applies it to every row of both tables, model/transfer.py:884-902).  It reads the two 256 B source
// rows and writes the 256 B result row; nothing else touches HBM: the conv output (the fc1 operand), the fc1
// pre-activations and GELU(fc1) (the fc2 operand) live in shared memory / tensor memory only.  The unfused path
// (conv kernel -> fc1 GEMM -> fc2 GEMM, umma_packed.cu) moves ~12 KB per row through L2 / HBM for the same work.
//
// One CTA per SM, persistent over 128-row tiles.  Per tile:
//   fc1 phase   8 compute warps (two groups of 128 threads, thread = row) run the conv stage for 8 latent dims at a time and
//               write the resulting 40-wide K chunk of the fc1 operand -- already split into TF32 hi / lo halves, in the
//               canonical K-major core-matrix layout -- into a 2-slot shared-memory ring; a bulk-copy warp streams the
//               matching pre-packed W1 sub-blocks (128 outputs x 40 K, 40 KB) through a 3-slot ring; one thread issues
//               tcgen05.mma.kind::tf32 (3xTF32: lo*hi + hi*lo + hi*hi) accumulating Z1[128 x 512] in ALL 512 TMEM columns;
//   drain phase the same compute warps read Z1 back 32 columns at a time (tcgen05.ld, one row per thread), add the fc1 bias,
//               apply GELU, split, and write the 32-wide K chunk of the fc2 operand into the same ring; the MMA thread
//               accumulates Y[128 x 64] with W2 chunks from the weight ring.  Y lives in TMEM columns that have already been
//               drained, in FOUR accumulators (4 chunks each) that the final epilogue sums with round-to-nearest adds (the
//               tensor core's accumulate truncates: short chains keep fc2 at FFMA-class accuracy; fc1's single 120-MMA chain
//               costs ~2e-6 relative);
//   epilogue    + fc2 bias (optionally / ||row||, ConvTransfer 'user'), one 256 B row store per thread.
// The K index of fc1 is permuted (chunk c holds latent dims 8c..8c+7 of all five conv channels, k' = 8*m + d%8) so that a
// chunk needs only 8 dims of the source rows; W1 is packed in the same order by k_pack_fused.  Per tile the tensor core needs
// ~37 k cycles and the weights (1.57 MB per tile) ~37 k cycles of one SM's L2 port: the two bounds coincide.
#include "sml_common.cuh"
#include "umma_ptx.cuh"

namespace {

using namespace ptx;

constexpr int FF_THREADS = 320;
constexpr int FF_NB = 3;                           // weight ring slots
constexpr uint32_t FF_SLOT = 40960;                // bytes of every ring slot
constexpr uint32_t FF_LBO = 128;                   // K-adjacent core matrices are contiguous
constexpr uint32_t FF_SBO_A = 10 * 128;            // fc1 operands: 10 core matrices (40 K) per 8-row group
constexpr uint32_t FF_SBO_G = 8 * 128;             // fc2 operands: 8 core matrices (32 K) per 8-row group
constexpr uint32_t FF_HALF_A = 16 * FF_SBO_A;      // 128 rows: 20 480 B (hi), lo follows
constexpr uint32_t FF_HALF_G = 16 * FF_SBO_G;      // 16 384 B
constexpr uint32_t FF_HALF_W2 = 8 * FF_SBO_G;      // 64 rows: 8 192 B
constexpr uint32_t FF_W1_BYTES = 32 * FF_SLOT;     // 8 chunks x 4 sub-blocks
constexpr uint32_t FF_W2_CHUNK = 2 * FF_HALF_W2;   // 16 384 B
constexpr uint32_t FF_W2_BYTES = 16 * FF_W2_CHUNK;
constexpr size_t FF_PACKED_BYTES = (size_t)FF_W1_BYTES + FF_W2_BYTES;      // 1 572 864 B per net
constexpr uint32_t FF_OFF_B = 2 * FF_SLOT;
constexpr uint32_t FF_OFF_MISC = FF_OFF_B + FF_NB * FF_SLOT;
constexpr uint32_t FF_SMEM = FF_OFF_MISC + 512 * 4 + 64 * 4 + 512 + 256;

struct ConvW {
    float w1[10][3];
    float b1[10];
    float w2[5][10];
    float b2[5];
};

__device__ __forceinline__ uint32_t ff_desc_lo(uint32_t saddr) { return desc_lo(saddr, FF_LBO); }
__host__ __device__ constexpr uint32_t ff_desc_hi(uint32_t sbo) { return desc_hi(sbo); }
__device__ __forceinline__ void ff_mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t d_hi, uint32_t idesc, uint32_t acc) {
    umma_tf32_h(d, a_lo, b_lo, d_hi, idesc, acc);
}
__device__ __forceinline__ bool ff_elect_one() { return elect_one(); }

template <int R>
__device__ __forceinline__ void conv_point(const ConvW &w, float x0, float x1, float x2, float (&out)[5]) {
    float h1[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        float z = w.b1[c];
        z = fmaf(w.w1[c][0], x0, z);
        z = fmaf(w.w1[c][1], x1, z);
        if (R == 3) z = fmaf(w.w1[c][2], x2, z);
        h1[c] = sml_gelu(z);
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        float z = w.b2[m];
#pragma unroll
        for (int c = 0; c < 10; ++c) z = fmaf(w.w2[m][c], h1[c], z);
        out[m] = sml_gelu(z);
    }
}

struct FusedParams {
    const float *x_t, *x_hat;
    const int64_t *ids;
    int64_t n_rows, pitch;
    const float *theta;          // one net (conv weights, biases)
    const uint8_t *wpk;          // packed W1 / W2 of that net (k_pack_fused)
    float *out;
    int normalize_out;
};

template <int R>
__global__ void __launch_bounds__(FF_THREADS, 1)
k_transfer_fused(FusedParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *s_b1 = reinterpret_cast<float *>(smem + FF_OFF_MISC);
    float *s_b2 = s_b1 + 512;
    ConvW &sw = *reinterpret_cast<ConvW *>(s_b2 + 64);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + FF_OFF_MISC + 512 * 4 + 64 * 4 + 512);
    uint64_t *a_full = bars, *a_empty = bars + 2, *b_full = bars + 4, *b_empty = bars + 4 + FF_NB, *z_full = bars + 4 + 2 * FF_NB,
             *y_full = z_full + 1, *tmem_free = z_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(z_full + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (P.n_rows + 127) / 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < FF_NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(z_full, 1); mbar_init(y_full, 1); mbar_init(tmem_free, 128);
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_slot);
    sml_pdl_wait();
    // net parameters the compute warps need
    for (int i = threadIdx.x; i < 512; i += FF_THREADS) s_b1[i] = __ldg(P.theta + SML_OFF_F1B + i);
    for (int i = threadIdx.x; i < 64; i += FF_THREADS) s_b2[i] = __ldg(P.theta + SML_OFF_F2B + i);
    for (int i = threadIdx.x; i < 10 * R; i += FF_THREADS) sw.w1[i / R][i % R] = __ldg(P.theta + SML_OFF_C1W + i);
    for (int i = threadIdx.x; i < 10; i += FF_THREADS) sw.b1[i] = __ldg(P.theta + SML_OFF_C1B + i);
    for (int i = threadIdx.x; i < 50; i += FF_THREADS) sw.w2[i / 10][i % 10] = __ldg(P.theta + SML_OFF_C2W + i);
    for (int i = threadIdx.x; i < 5; i += FF_THREADS) sw.b2[i] = __ldg(P.theta + SML_OFF_C2B + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    sml_pdl_trigger();
    const uint32_t tmem = *tmem_slot;
    uint8_t *sA = smem, *sB = smem + FF_OFF_B;

    if (warp == 8) {
        // ===== weight producer: W1 sub-blocks and W2 chunks in consumption order, every tile the same 48 blocks =====
        if (lane == 0) {
            uint32_t item = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int i = 0; i < 48; ++i, ++item) {
                    const uint32_t s = item % FF_NB;
                    if (item >= FF_NB) mbar_wait(&b_empty[s], ((item / FF_NB) - 1) & 1);
                    const uint32_t bytes = i < 32 ? FF_SLOT : FF_W2_CHUNK;
                    const uint8_t *src = i < 32 ? P.wpk + (size_t)i * FF_SLOT : P.wpk + FF_W1_BYTES + (size_t)(i - 32) * FF_W2_CHUNK;
                    mbar_expect_tx(&b_full[s], bytes);
                    bulk_g2s(sB + s * FF_SLOT, src, bytes, &b_full[s]);
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer: the whole warp runs the loops (warp-uniform control flow keeps the descriptors in uniform registers),
        // one elected lane issues the tcgen05 instructions =====
        constexpr uint32_t ID128 = idesc_tf32(128), ID64 = idesc_tf32(64);
        constexpr uint32_t HI_A = ff_desc_hi(FF_SBO_A), HI_G = ff_desc_hi(FF_SBO_G);
        constexpr uint32_t KSTEP = (2 * FF_LBO) >> 4;                  // descriptor-lo increment per K step of 8
        const bool leader = ff_elect_one();
        uint32_t bitem = 0, acons0 = 0, acons1 = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            if (it > 0) { mbar_wait(tmem_free, (it - 1) & 1); tc_fence_after(); }         // the previous tile's Y has been read
            // ---- fc1: Z1[128 x 512] += A chunk (40 K) x W1 sub-block, 8 chunks x 4 sub-blocks ----
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                const int s = c & 1;
                if (s == 0) { mbar_wait(&a_full[0], acons0 & 1); ++acons0; } else { mbar_wait(&a_full[1], acons1 & 1); ++acons1; }
                tc_fence_after();
                const uint32_t a_hi = ff_desc_lo(smem_u32(sA + s * FF_SLOT)), a_lo = ff_desc_lo(smem_u32(sA + s * FF_SLOT) + FF_HALF_A);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb, ++bitem) {
                    const uint32_t bs = bitem % FF_NB;
                    mbar_wait(&b_full[bs], (bitem / FF_NB) & 1);
                    tc_fence_after();
                    const uint32_t b_hi = ff_desc_lo(smem_u32(sB + bs * FF_SLOT)), b_lo = ff_desc_lo(smem_u32(sB + bs * FF_SLOT) + FF_HALF_A);
                    const uint32_t d = tmem + nb * 128;
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            ff_mma(d, a_lo + k * KSTEP, b_hi + k * KSTEP, HI_A, ID128, (c | k) != 0);
                            ff_mma(d, a_hi + k * KSTEP, b_lo + k * KSTEP, HI_A, ID128, 1);
                            ff_mma(d, a_hi + k * KSTEP, b_hi + k * KSTEP, HI_A, ID128, 1);
                        }
                        umma_commit(&b_empty[bs]);
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(&a_empty[s]);
                __syncwarp();
            }
            if (leader) umma_commit(z_full);
            __syncwarp();
            // ---- fc2: Y_{c2/4}[128 x 64] += GELU(Z1) chunk (32 K) x W2 chunk; Y_a overlaps Z1 columns 0..63 ----
#pragma unroll 1
            for (int c2 = 0; c2 < 16; ++c2, ++bitem) {
                const int s = c2 & 1;
                if (s == 0) { mbar_wait(&a_full[0], acons0 & 1); ++acons0; } else { mbar_wait(&a_full[1], acons1 & 1); ++acons1; }
                if (c2 == 0) mbar_wait(&a_full[1], acons1 & 1);           // chunk 1 published => Z1 columns 32..63 have been read too
                const uint32_t bs = bitem % FF_NB;
                mbar_wait(&b_full[bs], (bitem / FF_NB) & 1);
                tc_fence_after();
                const uint32_t a_hi = ff_desc_lo(smem_u32(sA + s * FF_SLOT)), a_lo = ff_desc_lo(smem_u32(sA + s * FF_SLOT) + FF_HALF_G);
                const uint32_t b_hi = ff_desc_lo(smem_u32(sB + bs * FF_SLOT)), b_lo = ff_desc_lo(smem_u32(sB + bs * FF_SLOT) + FF_HALF_W2);
                const uint32_t d = tmem + (uint32_t)(c2 >> 2) * 64;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ff_mma(d, a_lo + k * KSTEP, b_hi + k * KSTEP, HI_G, ID64, ((c2 & 3) | k) != 0);
                        ff_mma(d, a_hi + k * KSTEP, b_lo + k * KSTEP, HI_G, ID64, 1);
                        ff_mma(d, a_hi + k * KSTEP, b_hi + k * KSTEP, HI_G, ID64, 1);
                    }
                    umma_commit(&b_empty[bs]);
                    umma_commit(&a_empty[s]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(y_full);
            __syncwarp();
        }
    } else {
        // ===== compute warps: conv stage (fc1 phase), GELU + re-split (drain phase), final epilogue =====
        const int g = warp >> 2, q = warp & 3;
        const int r = q * 32 + lane;                                   // row of the tile = TMEM lane
        uint8_t *slot = sA + g * FF_SLOT;                              // group g always fills ring slot g
        const uint32_t row_off = (uint32_t)(r >> 3) * FF_SBO_A + (uint32_t)(r & 7) * 16;
        const uint32_t row_off_g = (uint32_t)(r >> 3) * FF_SBO_G + (uint32_t)(r & 7) * 16;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t fills = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int64_t row = tile * 128 + r;
            const int64_t rr = row < P.n_rows ? row : P.n_rows - 1;     // rows past the end recompute the last row (never stored)
            const int64_t id = P.ids ? __ldg(P.ids + rr) : rr;
            const float4 *xt = reinterpret_cast<const float4 *>(P.x_t + id * P.pitch);
            const float4 *xh = reinterpret_cast<const float4 *>(P.x_hat + id * P.pitch);
            float inv_keep = 0.f;
            float nrm = 1.f;
            if (R == 3) {                                              // ||x_t|| (conv_transfer.py:94), same summation as a 64-term dot
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 t = __ldg(xt + j);
                    acc = fmaf(t.x, t.x, acc); acc = fmaf(t.y, t.y, acc); acc = fmaf(t.z, t.z, acc); acc = fmaf(t.w, t.w, acc);
                }
                nrm = sqrtf(acc);
            }
            (void)inv_keep;
            // ---- fc1 phase: chunks c = g, g + 2, g + 4, g + 6 ----
            for (int c = g; c < 8; c += 2, ++fills) {
                float4 t0 = __ldg(xt + 2 * c), t1 = __ldg(xt + 2 * c + 1), h0 = __ldg(xh + 2 * c), h1 = __ldg(xh + 2 * c + 1);
                if (fills > 0) mbar_wait(&a_empty[g], (fills - 1) & 1);
                const float x0[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                const float x1[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float hi[5][4], lo[5][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int dl = 4 * half + j;
                        const float x2 = (R == 3) ? (x0[dl] * x1[dl]) / nrm : 0.f;     // conv_transfer.py:93,98 (no eps: NaN on zero rows)
                        float v[5];
                        conv_point<R>(sw, x0[dl], x1[dl], x2, v);
#pragma unroll
                        for (int m = 0; m < 5; ++m) pk_split(v[m], hi[m][j], lo[m][j]);
                    }
#pragma unroll
                    for (int m = 0; m < 5; ++m) {                      // k' = 8 m + dl: core matrix 2 m + half
                        uint8_t *dst = slot + row_off + (uint32_t)(2 * m + half) * FF_LBO;
                        *reinterpret_cast<float4 *>(dst) = make_float4(hi[m][0], hi[m][1], hi[m][2], hi[m][3]);
                        *reinterpret_cast<float4 *>(dst + FF_HALF_A) = make_float4(lo[m][0], lo[m][1], lo[m][2], lo[m][3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&a_full[g]);
            }
            // ---- drain phase: Z1 chunks c2 = g, g + 2, ..., g + 14 -> GELU -> fc2 operand chunks ----
            mbar_wait(z_full, it & 1);
            tc_fence_after();
            for (int c2 = g; c2 < 16; c2 += 2, ++fills) {
                float v[32];
                tmem_ld32(taddr + 32 * c2, v);
                mbar_wait(&a_empty[g], (fills - 1) & 1);
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    float h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) pk_split(sml_gelu(v[4 * qq + j] + s_b1[32 * c2 + 4 * qq + j]), h[j], l[j]);
                    uint8_t *dst = slot + row_off_g + (uint32_t)qq * FF_LBO;
                    *reinterpret_cast<float4 *>(dst) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4 *>(dst + FF_HALF_G) = make_float4(l[0], l[1], l[2], l[3]);
                }
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&a_full[g]);
            }
            // ---- final epilogue (group 0): Y = Y_a + Y_b + Y_c + Y_d + b2 ----
            if (g == 0) {
                mbar_wait(y_full, it & 1);
                tc_fence_after();
                float y[64];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32], w[32];
                    tmem_ld32(taddr + 32 * hh, v);
#pragma unroll
                    for (int a = 1; a < 4; ++a) {
                        tmem_ld32(taddr + 64 * a + 32 * hh, w);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += w[i];
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[32 * hh + i] = v[i] + s_b2[32 * hh + i];
                }
                tc_fence_before();
                mbar_arrive(tmem_free);
                if (P.normalize_out) {                                  // ConvTransfer 'user': x / ||x||  (conv_transfer.py:62-63)
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) acc = fmaf(y[i], y[i], acc);
                    const float n2 = sqrtf(acc);
#pragma unroll
                    for (int i = 0; i < 64; ++i) y[i] = y[i] / n2;
                }
                if (row < P.n_rows) {
                    float4 *dst = reinterpret_cast<float4 *>(P.out + row * SML_D);
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<512>(tmem);
}

// W1 / W2 of one net -> the block order and K permutation k_transfer_fused streams them in (hi / lo TF32 halves).
__global__ void __launch_bounds__(256) k_pack_fused(const float *__restrict__ theta, uint8_t *__restrict__ out) {
    sml_pdl_wait();
    sml_pdl_trigger();
    const float *W1 = theta + SML_OFF_F1W, *W2 = theta + SML_OFF_F2W;
    const int total1 = 512 * 80, total2 = 64 * 128;                   // quads of 4 consecutive k'
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total1 + total2; idx += gridDim.x * blockDim.x) {
        float x[4];
        uint8_t *dst;
        uint32_t half;
        if (idx < total1) {
            // quad (n, c, kq): k' = 4 kq .. 4 kq + 3 of chunk c  <->  channel m = kq / 2, dims 8 c + 4 (kq & 1) + {0..3}
            const int n = idx / 80, rem = idx % 80, c = rem / 10, kq = rem % 10;
            const int m = kq >> 1, d0 = 8 * c + 4 * (kq & 1);
            const float4 t = __ldg(reinterpret_cast<const float4 *>(W1 + (size_t)n * 320 + m * 64 + d0));
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            const int nb = n >> 7, nl = n & 127;
            dst = out + (size_t)(c * 4 + nb) * FF_SLOT + (uint32_t)(nl >> 3) * FF_SBO_A + (uint32_t)kq * FF_LBO + (uint32_t)(nl & 7) * 16;
            half = FF_HALF_A;
        } else {
            const int j = idx - total1, n = j / 128, kq_all = j % 128, c2 = kq_all >> 3, kq = kq_all & 7;
            const float4 t = __ldg(reinterpret_cast<const float4 *>(W2 + (size_t)n * 512 + 4 * kq_all));
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            dst = out + FF_W1_BYTES + (size_t)c2 * FF_W2_CHUNK + (uint32_t)(n >> 3) * FF_SBO_G + (uint32_t)kq * FF_LBO + (uint32_t)(n & 7) * 16;
            half = FF_HALF_W2;
        }
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pk_split(x[i], h[i], l[i]);
        *reinterpret_cast<float4 *>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(dst + half) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

}  // namespace

size_t sml_fused_fwd_packed_bytes() { return FF_PACKED_BYTES; }

int sml_launch_pack_fused(const float *theta_net, uint8_t *out, cudaStream_t st) {
    SML_CUDA_OK(sml_launch(k_pack_fused, dim3(96), dim3(256), 0, st, theta_net, out));
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_launch_transfer_fused(const float *x_t, const float *x_hat, const int64_t *ids, int64_t n_rows, int64_t pitch, int variant,
                              const float *theta_net, const uint8_t *wpk, int normalize_out, float *out, cudaStream_t st) {
    if (n_rows <= 0) return SML_OK;
    static bool attr_set = false;
    if (!attr_set) {
        SML_CUDA_OK(cudaFuncSetAttribute(k_transfer_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        SML_CUDA_OK(cudaFuncSetAttribute(k_transfer_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        attr_set = true;
    }
    FusedParams P = {x_t, x_hat, ids, n_rows, pitch, theta_net, wpk, out, normalize_out};
    const int64_t tiles = (n_rows + 127) / 128;
    const int sms = sml_sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (variant == SML_VARIANT_COM) SML_CUDA_OK(sml_launch(k_transfer_fused<3>, dim3(grid), dim3(FF_THREADS), FF_SMEM, st, P));
    else SML_CUDA_OK(sml_launch(k_transfer_fused<2>, dim3(grid), dim3(FF_THREADS), FF_SMEM, st, P));
    SML_LAUNCH_OK();
    return SML_OK;
}
