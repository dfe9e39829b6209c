// Persistent attention-decoder loop: ALL U teacher-forced steps of
// AttnDecoder.__call__'s raw_rnn body (attn_decoder.py:76-166) in one
// cooperative launch per direction (forward / backward), replacing ~5 launches per
// step.  State-independent work is batched outside (ops.AttnDecoderFnV2); per step
// the kernel runs three grid-synchronous phases over all CTAs:
//
//   forward   G: gates = pre_g[t] + [ctx_{t-1} | h_{t-1}] . W_ch  (+ LSTM pointwise)
//                W_ch = [W_in_c.W_x ; W_h] folds InputProjection's ctx half into the
//                decoder-LSTM kernel, so xin is never materialised.  Tile = 16 rows x
//                8 units; the CTA's 32 gate columns of W_ch stay in shared memory.
//             Y: y = c_new . q_k + q_b            (query is the CELL state)
//             A: masked-softmax attention read-out, CTA per (row, half of D)
//   backward  A': attention backward per row -> ds, dy
//             P : dc_new += dy . q_k^T, LSTM pointwise backward -> dz_t
//             X : [dctx_{t-1} | dh_{t-1}] = dz_t . W_ch^T
// dHF / dEnc (sums over all steps) are produced after the loop by two parallel
// kernels from the stored ds_t, alpha_t, y_t, dctx_t -- no read-modify-write per step.
// The dense products use mma.sync m16n8k8 with error-compensated TF32 (3xTF32).
#include "../../include/e2e_asr_b200.h"
#include "common.cuh"

namespace e2e {

extern long long* g_rec_dbg;
int g_dec_p2p = 1;                 // 1 = point-to-point row-block counters, 0 = grid barriers (e2e_set_dec_sync test hook)

namespace {

constexpr int NTH = 256;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// D[16 x 8] += A[16 x K] (rows a_s, stride lda, K range [k_begin,k_end)) . B, B[k][n] = bt[n*ldb + k]
__device__ __forceinline__ void mma_block(float (&d)[4], const float* a_s, int lda, const float* bt, int ldb,
                                          int k_begin, int k_end, int g, int tq) {
    const float* a0 = a_s + g * lda;
    const float* a1 = a_s + (g + 8) * lda;
    const float* b0 = bt + g * ldb;
    // 4 independent accumulator chains: {even, odd k-step} x {hi*hi, cross terms}; a single chain
    // would serialise on the mma.sync result latency
    float dm[2][4], dx[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dm[i][j] = 0.f; dx[i][j] = 0.f; }
    int ph = 0;
#pragma unroll 4
    for (int k = k_begin + tq; k < k_end; k += 8, ph ^= 1) {
        uint32_t ah[4], al[4], bh[2], bl[2];
        split_tf32(a0[k], ah[0], al[0]);
        split_tf32(a1[k], ah[1], al[1]);
        split_tf32(a0[k + 4], ah[2], al[2]);
        split_tf32(a1[k + 4], ah[3], al[3]);
        split_tf32(b0[k], bh[0], bl[0]);
        split_tf32(b0[k + 4], bh[1], bl[1]);
        if (ph == 0) {
            mma_tf32(dx[0], al, bh);
            mma_tf32(dm[0], ah, bh);
            mma_tf32(dx[0], ah, bl);
        } else {
            mma_tf32(dx[1], al, bh);
            mma_tf32(dm[1], ah, bh);
            mma_tf32(dx[1], ah, bl);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] += (dx[0][j] + dx[1][j]) + (dm[0][j] + dm[1][j]);
}

__device__ __forceinline__ void mma_tf32_nv(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// D[nt][16 x 8] += A[16 x K-range] . B for NT n-tiles: the A fragment of a k-tile is loaded and split ONCE and
// reused by every n-tile (the warps of a CTA split K, not N).  B[k][n] = bt[n*ldb + k].
template <int NT>
__device__ __forceinline__ void mma_ksplit(float (&d)[NT][4], const float* a_s, int lda, const float* bt, int ldb,
                                           int k_begin, int k_end, int g, int tq) {
    const float* a0 = a_s + g * lda;
    const float* a1 = a_s + (g + 8) * lda;
    float dm[NT][4], dx[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dm[i][j] = 0.f; dx[i][j] = 0.f; }
#pragma unroll 2
    for (int k = k_begin + tq; k < k_end; k += 8) {
        uint32_t ah[4], al[4];
        split_tf32(a0[k], ah[0], al[0]);
        split_tf32(a1[k], ah[1], al[1]);
        split_tf32(a0[k + 4], ah[2], al[2]);
        split_tf32(a1[k + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float* b0 = bt + (8 * nt + g) * ldb;
            uint32_t bh[2], bl[2];
            split_tf32(b0[k], bh[0], bl[0]);
            split_tf32(b0[k + 4], bh[1], bl[1]);
            mma_tf32_nv(dx[nt], al, bh);
            mma_tf32_nv(dm[nt], ah, bh);
            mma_tf32_nv(dx[nt], ah, bl);
        }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = dm[i][j] + dx[i][j];
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch => co-resident)
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch, int* err) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic smem writes before later bulk copies
    __syncthreads();
    if (threadIdx.x == 0) {
        // red.release.gpu is cumulative over the CTA barrier above (the other threads' writes are ordered before it),
        // and the polling load is an acquire: no separate fence
        red_release_gpu_add(ctr, 1u);
        ++epoch;
        spin_wait_ge(ctr, epoch * gridDim.x, err);
    }
    __syncthreads();
}

// Point-to-point synchronisation per 16-row block instead of grid barriers (stride != 0): the batch rows of different
// row blocks never exchange data, and inside a block every phase only needs the PREVIOUS phase's tiles of that block.
// Each phase of each row block owns a monotonically increasing L2 counter; a CTA arrives once per tile it finished
// (release) and, before a tile, waits until the producers' counter reaches (tiles per round) x (rounds so far)
// (acquire).  16-32 arrivals on 12 different counters replace 128 arrivals on one, and nobody waits for another row
// block's stragglers.  Deadlock-free for any tile -> CTA mapping: every CTA walks (step, phase) in the same order and
// the grid is co-resident (cooperative launch).
struct P2P {
    unsigned* base;      // p.ctr + 1
    int stride, nrb;     // stride 0: disabled (grid barriers)
    __device__ __forceinline__ unsigned* at(int phase, int rb) const { return base + (size_t)stride * (phase * nrb + rb); }
};
__device__ __forceinline__ void tile_arrive(unsigned* c) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic smem writes before later bulk copies
    __syncthreads();
    if (threadIdx.x == 0) red_release_gpu_add(c, 1u);
}
__device__ __forceinline__ void tile_wait(const unsigned* c, unsigned target, int* err) {
    if (threadIdx.x == 0) spin_wait_ge(c, target, err);
    __syncthreads();
}

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
// ex2.approx + rcp.approx: absolute error ~2e-7, inside the 1e-4 parity budget
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_par(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > (1ll << 33)) __trap();      // never hang the GPU on a protocol bug
    }
}
// global -> own shared memory, completion (bytes) on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// profiling aid (e2e_set_rec_debug): clock64 stamps of CTA 0 / thread 0, 8 per step
__device__ long long* d_dec_dbg = nullptr;
#define DEC_STAMP(step, i) do { if (dbg) dbg[(step) * 32 + (i)] = clock64(); } while (0)

}  // namespace

// ======================================================================= forward
__global__ void __launch_bounds__(NTH, 1) dec_fwd_persist_kernel(e2e_dec_persist_args p, int attn_cap, int p2p_stride) {
    extern __shared__ __align__(128) float smem[];
    const int B = p.B, U = p.U, Hd = p.Hd, A = p.A, D = p.D, Tn = p.Tn, Tp = p.Tp;
    const int K = D + Hd, G4 = 4 * Hd, CAT = Hd + D;
    const int KS = K + 4;                      // Wt row stride (== 4 mod 32: conflict-free fragment loads)
    const int AS = K + 4;                      // A tile row stride
    const int QS_ = Hd + 4;                    // q_s row stride
    float* Wt = smem;                          // [32][K+8]  W_ch^T slice of this CTA's column block
    float* a_s = Wt + 32 * (K + 8);            // [16][AS]   A tile / scratch of the other phases
    float* red = a_s + 16 * AS;                // [8][32][4] cross-warp partials of phase Y
    float* gred = red + 8 * 32 * 4;            // [8][16][40] per-warp partial gate tiles of phase G
    float* qres = gred + 8 * 16 * 40;          // [8][Hd+4]  resident q_k^T slice of this CTA's phase-Y tile
    const int tid = threadIdx.x, w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;
    const int NCB = Hd / 8, nrb = (B + 15) / 16, gtiles = nrb * NCB;
    const bool resident = gtiles <= (int)gridDim.x;
    unsigned epoch = 0;
    const P2P pp{p.ctr + 1, p2p_stride, nrb};      // phases: 0 = G (gates), 1 = Y (query), 2 = A (attention)
    const bool p2p = p2p_stride != 0;
    // fast attention read-out: one (row, half) item per CTA and HF[b] / half of enc[b] fit the staging buffers
    // attn_cap: floats in attn_buf (0: not provisioned by the launcher)
    float* attn_buf = qres + 8 * (Hd + 4);
    const bool fastA = attn_cap > 0 && 2 * B <= (int)gridDim.x && Tn <= 128 && A % 4 == 0 && A <= 128 && D % 8 == 0 &&
                       D / 8 <= NTH && NTH % (D / 8) == 0 && Tn * A <= 16 * AS &&
                       ((Tn + 1) / 2) * (D / 2) <= 16 * AS && Tn * A <= attn_cap &&
                       ((Tn + 1) / 2) * (D / 2) <= attn_cap;
    __shared__ __align__(8) uint64_t abar[3];   // bulk-copy completion: [0] bufX (two uses per step), [1] bufY, [2] row tiles
    uint32_t aph0 = 0, aph1 = 0, lph = 0;
    if (tid == 0) {
        mbar_init(&abar[0], 1);
        mbar_init(&abar[1], 1);
        mbar_init(&abar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int ytiles = nrb * (A / 8);
    const bool yres = ytiles <= (int)gridDim.x;
    auto load_q = [&](int nt) {                  // qres[n][k] = q_k[k][8*nt + n]
        for (int i = tid; i < 8 * Hd; i += NTH) {
            int k = i / 8, n = i % 8;
            qres[n * QS_ + k] = p.q_k[(size_t)k * A + 8 * nt + n];
        }
    };
    if (yres && (int)blockIdx.x < ytiles) load_q(blockIdx.x % (A / 8));

    auto load_wt = [&](int cb) {
        // Wt[n][k] = W_ch[k][32*cb + n]
        for (int i = tid; i < 32 * K; i += NTH) {
            int k = i / 32, n = i % 32;
            Wt[n * KS + k] = p.W_ch[(size_t)k * G4 + 32 * cb + n];
        }
    };
    if (resident && (int)blockIdx.x < gtiles) load_wt(blockIdx.x % NCB);
    __syncthreads();

    long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? d_dec_dbg : nullptr;
    const int t_begin = p.t1 > 0 ? p.t0 : 0, t_end = p.t1 > 0 ? min(p.t1, U) : U;
    for (int t = t_begin; t < t_end; ++t) {
        const int rnd = t - t_begin;             // rounds of this launch (the row-block counters start at zero)
        DEC_STAMP(t, 0);
        // ------------------------------------------------------------ phase G
        for (int tile = blockIdx.x; tile < gtiles; tile += gridDim.x) {
            const int rb = tile / NCB, cb = tile % NCB;
            if (!resident) { __syncthreads(); load_wt(cb); }
            // A tile: [ctx_{t-1} (D) | h_{t-1} (Hd)] for rows rb*16 .. +16: one bulk copy per row segment
            const int nvalid = min(16, B - rb * 16);
            if (p2p && rnd > 0) tile_wait(pp.at(2, rb), (unsigned)(2 * nvalid * rnd), p.err);   // ctx_{t-1} of the block
            if (tid == 0) mbar_expect_tx(&abar[2], (uint32_t)(nvalid * (Hd + (t > 0 ? D : 0)) * 4));
            __syncwarp();
            if (lane * 8 + w < 32) {
                const int slot = lane * 8 + w, r = slot >> 1, part = slot & 1, b = rb * 16 + r;
                if (b < B) {
                    if (part == 0) {
                        if (t > 0) bulk_g2s(a_s + r * AS, p.cat + ((size_t)(t - 1) * B + b) * CAT + Hd, (uint32_t)(D * 4), &abar[2]);
                    } else {
                        bulk_g2s(a_s + r * AS + D, p.hprev + ((size_t)t * B + b) * Hd, (uint32_t)(Hd * 4), &abar[2]);
                    }
                }
            }
            if (t == 0 || nvalid < 16) {             // rows / segments no copy fills
                for (int i = tid; i < 16 * K; i += NTH) {
                    const int r = i / K, k = i % K;
                    if (rb * 16 + r >= B || (t == 0 && k < D)) a_s[r * AS + k] = 0.f;
                }
            }
            // epilogue operands of this thread's (row, unit) element, in flight during the product
            const int prow = tid >> 3, ul = tid & 7;
            const int eb = rb * 16 + prow, unit = cb * 8 + ul;
            const bool own = tid < 128 && eb < B;
            const size_t row = (size_t)t * B + eb;
            float4 pg = make_float4(0.f, 0.f, 0.f, 0.f);
            float cp = 0.f;
            if (own) {
                pg = __ldg(reinterpret_cast<const float4*>(p.pre_g + row * G4 + unit * 4));
                cp = __ldcg(p.cprev + row * Hd + unit);
            }
            DEC_STAMP(t, 16);
            mbar_wait_par(&abar[2], lph);
            lph ^= 1u;
            if (t == 0 || nvalid < 16) __syncthreads();
            DEC_STAMP(t, 17);
            {
                // the 8 warps split K; every warp produces a partial [16 x 32] gate tile
                float d[4][4];
                const int ksteps = K / 8, per = (ksteps + 7) / 8;
                const int k0 = min(ksteps, w * per) * 8, k1 = min(ksteps, (w + 1) * per) * 8;
                mma_ksplit<4>(d, a_s, AS, Wt, KS, k0, k1, g, tq);
                float* gw = gred + w * (16 * 40);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    *reinterpret_cast<float2*>(gw + g * 40 + 8 * nt + 2 * tq) = make_float2(d[nt][0], d[nt][1]);
                    *reinterpret_cast<float2*>(gw + (g + 8) * 40 + 8 * nt + 2 * tq) = make_float2(d[nt][2], d[nt][3]);
                }
            }
            DEC_STAMP(t, 18);
            __syncthreads();
            DEC_STAMP(t, 19);
            if (own) {
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) {
                    const float4 o = *reinterpret_cast<const float4*>(gred + ww * (16 * 40) + prow * 40 + 4 * ul);
                    z.x += o.x; z.y += o.y; z.z += o.z; z.w += o.w;
                }
                const float si = sigmoid_fast(z.x + pg.x);
                const float tj = tanh_fast(z.y + pg.y);
                const float sf = sigmoid_fast(z.z + pg.z + 1.0f);
                const float so = sigmoid_fast(z.w + pg.w);
                const float cn = cp * sf + si * tj;
                const float hn = tanh_fast(cn) * so;
                *reinterpret_cast<float4*>(p.acts + row * G4 + unit * 4) = make_float4(si, tj, sf, so);
                p.cat[row * CAT + unit] = cn;
                if (t + 1 < U) {      // committed state: finished rows keep (c, h)  (raw_rnn)
                    const bool live = t < p.lens[eb];
                    const float hp = a_s[prow * AS + D + unit];
                    p.cprev[(row + B) * Hd + unit] = live ? cn : cp;
                    p.hprev[(row + B) * Hd + unit] = live ? hn : hp;
                }
            }
            __syncthreads();
            if (p2p) tile_arrive(pp.at(0, rb));
        }
        DEC_STAMP(t, 1);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 2);
        // ------------------------------------------------------------ phase Y: y = c_new . q_k + q_b
        {
            float* c_s = a_s;                       // [16][Hd+4]
            const int CS_ = Hd + 4;
            for (int tile = blockIdx.x; tile < ytiles; tile += gridDim.x) {
                const int rb = tile / (A / 8), nt = tile % (A / 8);
                const int nvalid = min(16, B - rb * 16);
                __syncthreads();
                if (p2p) tile_wait(pp.at(0, rb), (unsigned)(NCB * (rnd + 1)), p.err);            // c_new of the block
                if (tid == 0) mbar_expect_tx(&abar[2], (uint32_t)(nvalid * Hd * 4));
                __syncwarp();
                if (lane * 8 + w < nvalid) {
                    const int r = lane * 8 + w;
                    bulk_g2s(c_s + r * CS_, p.cat + ((size_t)t * B + rb * 16 + r) * CAT, (uint32_t)(Hd * 4), &abar[2]);
                }
                if (nvalid < 16)
                    for (int i = tid; i < 16 * Hd; i += NTH)
                        if (i / Hd >= nvalid) c_s[(i / Hd) * CS_ + i % Hd] = 0.f;
                if (!yres) load_q(nt);
                mbar_wait_par(&abar[2], lph);
                lph ^= 1u;
                if (nvalid < 16 || !yres) __syncthreads();
                // 8 warps split K
                const int ksteps = Hd / 8, per = (ksteps + 7) / 8;
                const int k0 = min(ksteps, w * per) * 8, k1 = min(ksteps, (w + 1) * per) * 8;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                mma_block(d, c_s, CS_, qres, QS_, k0, k1, g, tq);
                float* part = red;                  // [8 warps][32 lanes][4]
                *reinterpret_cast<float4*>(part + (w * 32 + lane) * 4) = make_float4(d[0], d[1], d[2], d[3]);
                __syncthreads();
                if (w == 0) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int ww = 0; ww < 8; ++ww) {
                        float4 o = *reinterpret_cast<const float4*>(part + (ww * 32 + lane) * 4);
                        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                    }
                    const int col = 8 * nt + 2 * tq;
                    const float qb0 = p.q_b[col], qb1 = p.q_b[col + 1];
                    int b0 = rb * 16 + g, b1 = rb * 16 + g + 8;
                    if (b0 < B) *reinterpret_cast<float2*>(p.y + ((size_t)t * B + b0) * A + col) = make_float2(acc.x + qb0, acc.y + qb1);
                    if (b1 < B) *reinterpret_cast<float2*>(p.y + ((size_t)t * B + b1) * A + col) = make_float2(acc.z + qb0, acc.w + qb1);
                }
                if (p2p) tile_arrive(pp.at(1, rb));
            }
        }
        DEC_STAMP(t, 3);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 4);
        // ------------------------------------------------------------ phase A: attention read-out
        if (fastA) {
          if ((int)blockIdx.x < 2 * B) {
            // CTA = (row b, half of D), fixed for the whole sequence.  HF[b] and the CTA's half of enc[b] are
            // streamed L2 -> shared memory by bulk copies (mbarrier completion) in three transfers:
            // HF + y -> bufX | enc rows [0,TC) -> bufY | enc rows [TC,len) -> bufX once the scores are done.
            const int b = blockIdx.x / 2, half = blockIdx.x % 2;
            const int len = min(p.enc_len[b], Tn);
            const int dh = D / 2, TC = (Tn + 1) / 2;
            float* bufX = a_s;
            float* bufY = attn_buf;
            float* y_s = attn_buf + attn_cap;       // [A]
            float* v_s = y_s + A;                   // [A]
            float* s_s = v_s + A;                   // [Tn + pad]
            float* part_s = bufX;                   // [4][dh] partial read-outs (bufX is free again by then)
            const float* HFb = p.HF + (size_t)b * Tp * A;
            const float* encb = p.enc + (size_t)b * Tp * D + half * dh;
            const int c0 = min(TC, len), c1 = max(len - TC, 0);
            // HF[b] and the first enc rows do not depend on this step: their copies are in flight while the CTA waits
            // for the row block's query y_t
            if (tid == 0) {
                mbar_expect_tx(&abar[0], (uint32_t)((len + 1) * A * 4));
                bulk_g2s(bufX, HFb, (uint32_t)(len * A * 4), &abar[0]);
                mbar_expect_tx(&abar[1], (uint32_t)(c0 * dh * 4));
            }
            __syncwarp();
            // one bulk copy per enc row (dh floats), spread over the lanes of all warps
            // (a lane issues its copies serially, ~100 cycles each: slot = lane*8 + warp keeps that to a few per warp)
            for (int r = lane * 8 + w; r < c0; r += NTH) bulk_g2s(bufY + r * dh, encb + (size_t)r * D, (uint32_t)(dh * 4), &abar[1]);
            if (rnd == 0) for (int a = tid; a < A; a += NTH) v_s[a] = p.attn_v[a];
            if (p2p) tile_wait(pp.at(1, b / 16), (unsigned)((A / 8) * (rnd + 1)), p.err);        // y_t of the block
            if (tid == 0) bulk_g2s(y_s, p.y + ((size_t)t * B + b) * A, (uint32_t)(A * 4), &abar[0]);
            DEC_STAMP(t, 8);
            mbar_wait_par(&abar[0], aph0);
            aph0 ^= 1u;
            if (rnd == 0) __syncthreads();
            DEC_STAMP(t, 9);
            // scores: 8 lanes per tau (4 taus per warp pass), lane covers float4 columns sub, sub+8, ...
            {
                const int sub = lane & 7, tl = lane >> 3, nq = A / 4;
                float4 yy[4], vv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = sub + 8 * j;
                    yy[j] = q < nq ? *reinterpret_cast<const float4*>(y_s + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vv[j] = q < nq ? *reinterpret_cast<const float4*>(v_s + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                for (int tau0 = 4 * w; tau0 < len; tau0 += 32) {
                    const int tau = tau0 + tl;
                    float acc = 0.f;
                    if (tau < len) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int q = sub + 8 * j;
                            if (q < nq) {
                                const float4 hf = *reinterpret_cast<const float4*>(bufX + tau * A + q * 4);
                                acc += vv[j].x * tanh_fast(hf.x + yy[j].x) + vv[j].y * tanh_fast(hf.y + yy[j].y) +
                                       vv[j].z * tanh_fast(hf.z + yy[j].z) + vv[j].w * tanh_fast(hf.w + yy[j].w);
                            }
                        }
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    if (sub == 0 && tau < len) s_s[tau] = acc;
                }
            }
            __syncthreads();
            DEC_STAMP(t, 10);
            // bufX (HF) is free: the second half of the enc rows goes there while the softmax runs
            if (tid == 0) mbar_expect_tx(&abar[0], (uint32_t)(c1 * dh * 4));
            __syncwarp();
            if (c1 > 0)
                for (int r = lane * 8 + w; r < c1; r += NTH)
                    bulk_g2s(bufX + r * dh, encb + (size_t)(TC + r) * D, (uint32_t)(dh * 4), &abar[0]);
            // masked softmax over tau by warp 0 (Tn <= 128: at most 4 values per lane)
            if (w == 0) {
                float sv[4], mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int tau = lane + 32 * j;
                    sv[j] = tau < len ? s_s[tau] : -INFINITY;
                    mx = fmaxf(mx, sv[j]);
                }
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    sv[j] = lane + 32 * j < len ? __expf(sv[j] - mx) : 0.f;
                    sum += sv[j];
                }
                sum = warp_sum(sum);
                const float inv = 1.0f / sum;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int tau = lane + 32 * j;
                    if (tau < Tn) {
                        const float al = sv[j] * inv;
                        s_s[tau] = al;
                        if (half == 0) p.alpha[((size_t)t * B + b) * Tn + tau] = al;
                    }
                }
            }
            __syncthreads();
            DEC_STAMP(t, 11);
            // read-out: thread = (float4 column cq, tau group tg); 4 tau groups when dh <= 256
            const int ncq = dh / 4;                              // float4 columns of the half
            const int ntg = min(NTH / ncq, 4);                   // tau groups (ncq <= NTH guaranteed by fastA)
            const int cq = tid % ncq, tg = tid / ncq;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait_par(&abar[1], aph1);
            aph1 ^= 1u;
            DEC_STAMP(t, 12);
            if (tg < ntg) {
#pragma unroll 4
                for (int tau = tg; tau < c0; tau += ntg) {
                    const float al = s_s[tau];
                    const float4 e = *reinterpret_cast<const float4*>(bufY + tau * dh + cq * 4);
                    acc.x = fmaf(al, e.x, acc.x); acc.y = fmaf(al, e.y, acc.y);
                    acc.z = fmaf(al, e.z, acc.z); acc.w = fmaf(al, e.w, acc.w);
                }
            }
            DEC_STAMP(t, 13);
            mbar_wait_par(&abar[0], aph0);
            aph0 ^= 1u;
            DEC_STAMP(t, 14);
            if (tg < ntg) {
#pragma unroll 4
                for (int tau = tg; tau < c1; tau += ntg) {
                    const float al = s_s[TC + tau];
                    const float4 e = *reinterpret_cast<const float4*>(bufX + tau * dh + cq * 4);
                    acc.x = fmaf(al, e.x, acc.x); acc.y = fmaf(al, e.y, acc.y);
                    acc.z = fmaf(al, e.z, acc.z); acc.w = fmaf(al, e.w, acc.w);
                }
            }
            __syncthreads();                         // every thread is done with bufX
            if (tg < ntg) *reinterpret_cast<float4*>(part_s + tg * dh + cq * 4) = acc;
            __syncthreads();
            for (int dd = tid; dd < dh; dd += NTH) {
                float c = 0.f;
                for (int j = 0; j < ntg; ++j) c += part_s[j * dh + dd];
                p.cat[((size_t)t * B + b) * CAT + Hd + half * dh + dd] = c;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();                         // bufX = a_s is the next step's A tile
            if (p2p) tile_arrive(pp.at(2, b / 16));
          }
        } else
        {
            float* y_s = a_s;                       // [A]
            float* v_s = a_s + A;                   // [A]
            float* s_s = a_s + 2 * A;               // [Tn]
            float* redw = s_s + Tn;                 // [8]
            for (int item = blockIdx.x; item < 2 * B; item += gridDim.x) {
                const int b = item / 2, half = item % 2;
                const int len = min(p.enc_len[b], Tn);
                __syncthreads();
                if (p2p) tile_wait(pp.at(1, b / 16), (unsigned)((A / 8) * (rnd + 1)), p.err);
                for (int a = tid; a < A; a += NTH) { y_s[a] = __ldcg(p.y + ((size_t)t * B + b) * A + a); v_s[a] = p.attn_v[a]; }
                __syncthreads();
                const float* HFb = p.HF + (size_t)b * Tp * A;
#pragma unroll 2
                for (int tau = w; tau < len; tau += 8) {
                    float acc = 0.f;
                    for (int a = lane * 4; a < A; a += 128) {
                        float4 hf = __ldg(reinterpret_cast<const float4*>(HFb + (size_t)tau * A + a));
                        acc += v_s[a] * tanhf(hf.x + y_s[a]) + v_s[a + 1] * tanhf(hf.y + y_s[a + 1]) +
                               v_s[a + 2] * tanhf(hf.z + y_s[a + 2]) + v_s[a + 3] * tanhf(hf.w + y_s[a + 3]);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) s_s[tau] = acc;
                }
                __syncthreads();
                float mx = -INFINITY;
                for (int tau = tid; tau < len; tau += NTH) mx = fmaxf(mx, s_s[tau]);
                mx = warp_max(mx);
                if (lane == 0) redw[w] = mx;
                __syncthreads();
                mx = redw[0];
                for (int ww = 1; ww < 8; ++ww) mx = fmaxf(mx, redw[ww]);
                __syncthreads();
                float sum = 0.f;
                for (int tau = tid; tau < len; tau += NTH) {
                    float e = expf(s_s[tau] - mx);
                    s_s[tau] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                if (lane == 0) redw[w] = sum;
                __syncthreads();
                sum = 0.f;
                for (int ww = 0; ww < 8; ++ww) sum += redw[ww];
                const float inv = 1.0f / sum;
                __syncthreads();
                for (int tau = tid; tau < Tn; tau += NTH) {
                    float al = tau < len ? s_s[tau] * inv : 0.f;
                    if (tau < len) s_s[tau] = al;
                    if (half == 0) p.alpha[((size_t)t * B + b) * Tn + tau] = al;
                }
                __syncthreads();
                const float* encb = p.enc + (size_t)b * Tp * D;
                const int dh = D / 2;
                if (dh % 4 == 0 && dh <= 512) {
                    // warp w sums its taus (w, w+8, ...); lane owns float4 columns lane, lane+32, ...
                    float* part = a_s + ((2 * A + Tn + 16 + 3) & ~3);   // [8][dh], 16-byte aligned
                    const int nv = dh / 4;
                    float4 acc[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                    for (int tau = w; tau < len; tau += 8) {
                        const float al = s_s[tau];
                        const float4* er = reinterpret_cast<const float4*>(encb + (size_t)tau * D + half * dh);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int ci = lane + 32 * c;
                            if (ci < nv) {
                                float4 e = __ldg(er + ci);
                                acc[c].x = fmaf(al, e.x, acc[c].x); acc[c].y = fmaf(al, e.y, acc[c].y);
                                acc[c].z = fmaf(al, e.z, acc[c].z); acc[c].w = fmaf(al, e.w, acc[c].w);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ci = lane + 32 * c;
                        if (ci < nv) *reinterpret_cast<float4*>(part + w * dh + ci * 4) = acc[c];
                    }
                    __syncthreads();
                    for (int dd = tid; dd < dh; dd += NTH) {
                        float c = 0.f;
#pragma unroll
                        for (int ww = 0; ww < 8; ++ww) c += part[ww * dh + dd];
                        p.cat[((size_t)t * B + b) * CAT + Hd + half * dh + dd] = c;
                    }
                } else {
                    for (int dd = tid; dd < dh; dd += NTH) {
                        const int dcol = half * dh + dd;
                        float c = 0.f;
                        for (int tau = 0; tau < len; ++tau) c = fmaf(s_s[tau], encb[(size_t)tau * D + dcol], c);
                        p.cat[((size_t)t * B + b) * CAT + Hd + dcol] = c;
                    }
                }
                if (p2p) tile_arrive(pp.at(2, b / 16));
            }
        }
        DEC_STAMP(t, 5);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 6);
        if (dbg) dbg[t * 32 + 7] = fastA ? 1 : 0;
    }
}

// ======================================================================= backward
__global__ void __launch_bounds__(NTH, 1) dec_bwd_persist_kernel(e2e_dec_persist_args p, int attn_fast, int p2p_stride) {
    extern __shared__ __align__(16) float smem[];
    const int B = p.B, U = p.U, Hd = p.Hd, A = p.A, D = p.D, Tn = p.Tn, Tp = p.Tp;
    const int K = D + Hd, G4 = 4 * Hd, CAT = Hd + D;
    const int NX = 24;                          // output columns of [dctx|dh] per CTA tile (3 n-tiles)
    const int WS = G4 + 4;                      // Wt2 row stride (== 4 mod 32: conflict-free fragment loads)
    const int ZS = G4 + 4;                      // dz tile row stride
    const int QS_ = A + 4;                      // q_s row stride
    float* Wt2 = smem;                          // [NX][G4+8] rows of W_ch owned by this CTA (phase X)
    float* z_s = Wt2 + NX * (G4 + 8);           // [16][ZS]  dz tile / scratch of the other phases
    float* red = z_s + 16 * ZS;                 // [8][32][4]
    float* xred = red + 8 * 32 * 4;             // [8][16][40] per-warp partial tiles of phase X
    float* pres = xred + 8 * 16 * 40;           // [8][A+4]  resident q_k slice of this CTA's phase-P tile
    float* abuf = pres + 8 * (A + 4);           // fast attention backward: HF rows, y, v, dctx, alpha, ...
    __shared__ __align__(8) uint64_t lbar;      // bulk-copy completion of the row tiles
    __shared__ __align__(8) uint64_t ebar, hbar;  // ... of the enc rows + dctx / of the HF rows + y
    uint32_t lph = 0, eph = 0, hph = 0;
    // fast attention backward: CTA = (row b, half of the TIME axis); the two halves leave partial sums
    // S1_a = sum_tau alpha dalpha (1 - th^2), S2_a = sum_tau alpha (1 - th^2), dot = sum_tau alpha dalpha in an L2
    // scratch; phase P combines them:  dy_a = v_a (S1_a - dot S2_a)   (softmax backward folded in)
    const bool fastAp = attn_fast != 0 && 2 * B <= (int)gridDim.x;
    float* apart = reinterpret_cast<float*>(p.ctr) + 64;          // [B][2][2A+4]
    float* adots = apart + (size_t)B * 2 * (2 * A + 4);           // [U][B][2]
    const int tid = threadIdx.x, w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;
    const int nrb = (B + 15) / 16;
    const int NXB = (K + NX - 1) / NX, xtiles = nrb * NXB;
    const bool resident = xtiles <= (int)gridDim.x;
    unsigned epoch = 0;
    const P2P pp{p.ctr + 1, p2p_stride, nrb};      // phases: 0 = A' (attention backward), 1 = P (pointwise), 2 = X
    const bool p2p = p2p_stride != 0;
    const int a_arr = fastAp ? 2 : 1;              // A' arrivals per batch row and step

    auto load_wt2 = [&](int xb) {
        for (int i = tid; i < NX * (G4 / 4); i += NTH) {
            int n = i / (G4 / 4), c4 = (i % (G4 / 4)) * 4;
            int kr = xb * NX + n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kr < K) v = *reinterpret_cast<const float4*>(p.W_ch + (size_t)kr * G4 + c4);
            *reinterpret_cast<float4*>(Wt2 + n * WS + c4) = v;
        }
    };
    if (resident && (int)blockIdx.x < xtiles) load_wt2(blockIdx.x % NXB);
    const int NCBp = Hd / 8, ptiles = nrb * NCBp;
    const bool pres_ok = ptiles <= (int)gridDim.x;
    auto load_qp = [&](int cb) {                 // pres[n][k] = q_k[8*cb + n][k]
        for (int i = tid; i < 8 * A; i += NTH) {
            int n = i / A, k = i % A;
            pres[n * QS_ + k] = p.q_k[(size_t)(8 * cb + n) * A + k];
        }
    };
    if (pres_ok && (int)blockIdx.x < ptiles) load_qp(blockIdx.x % NCBp);
    if (tid == 0) {
        mbar_init(&lbar, 1);
        mbar_init(&ebar, 1);
        mbar_init(&hbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? d_dec_dbg : nullptr;
    for (int t = U - 1; t >= 0; --t) {
        DEC_STAMP(t, 0);
        // ------------------------------------------------------------ phase A': attention backward
        if (fastAp) {
          if ((int)blockIdx.x < 2 * B) {
            const int b = blockIdx.x / 2, half = blockIdx.x % 2;
            const int len = min(p.enc_len[b], Tn);
            const int TC = (Tn + 1) / 2, TCp = (TC + 3) & ~3, tau0 = half * TC;
            const int n = max(0, min(len - tau0, TC));
            const size_t row = (size_t)t * B + b;
            float* encbuf = z_s;                    // [TC][D] (spans z_s | red | xred)
            float* hfbuf = abuf;                    // [TC][A]
            float* y_s = hfbuf + TCp * A;           // [A]
            float* v_s = y_s + A;                   // [A]
            float* dctx_s = v_s + A;                // [D]
            float* al_s = dctx_s + D;               // [TCp] alpha
            float* w_s = al_s + TCp;                // [TCp] alpha * dalpha
            float* dal_s = w_s + TCp;               // [TCp] dalpha
            float* sred = dal_s + TCp;              // [4][2A] + [4]
            if (tid == 0) {
                mbar_expect_tx(&ebar, (uint32_t)((n + 1) * D * 4));
                if (n > 0) bulk_g2s(encbuf, p.enc + ((size_t)b * Tp + tau0) * D, (uint32_t)(n * D * 4), &ebar);
                if (!p2p) bulk_g2s(dctx_s, p.dcat + row * CAT + Hd, (uint32_t)(D * 4), &ebar);
            }
            if (tid == 32) {
                mbar_expect_tx(&hbar, (uint32_t)((n + 1) * A * 4));
                if (n > 0) bulk_g2s(hfbuf, p.HF + ((size_t)b * Tp + tau0) * A, (uint32_t)(n * A * 4), &hbar);
                bulk_g2s(y_s, p.y + row * A, (uint32_t)(A * 4), &hbar);
            }
            if (p2p) {
                // the enc / HF rows and y_t above do not depend on the previous phase: in flight during the wait for
                // d ctx_t (X of the previous round, all column tiles of the row block)
                if (t < U - 1) tile_wait(pp.at(2, b / 16), (unsigned)(NXB * (U - 1 - t)), p.err);
                if (tid == 0) bulk_g2s(dctx_s, p.dcat + row * CAT + Hd, (uint32_t)(D * 4), &ebar);
            }
            if (tid >= 64 && tid - 64 < n) al_s[tid - 64] = p.alpha[row * Tn + tau0 + tid - 64];
            if (t == U - 1) for (int a = tid; a < A; a += NTH) v_s[a] = p.attn_v[a];
            mbar_wait_par(&ebar, eph);
            eph ^= 1u;
            // dalpha_tau = dctx . enc[tau]: 8 lanes per tau, 4 taus per warp pass
            {
                const int sub = lane & 7, tl = lane >> 3, nq = D / 4;
                for (int tb = 4 * w; tb < n; tb += 32) {
                    const int tau = tb + tl;
                    float acc = 0.f;
                    if (tau < n) {
                        const float* er = encbuf + (size_t)tau * D;
#pragma unroll 4
                        for (int q = sub; q < nq; q += 8) {
                            const float4 e = *reinterpret_cast<const float4*>(er + q * 4);
                            const float4 dc = *reinterpret_cast<const float4*>(dctx_s + q * 4);
                            acc = fmaf(dc.x, e.x, acc); acc = fmaf(dc.y, e.y, acc);
                            acc = fmaf(dc.z, e.z, acc); acc = fmaf(dc.w, e.w, acc);
                        }
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    if (sub == 0 && tau < n) dal_s[tau] = acc;
                }
            }
            __syncthreads();
            if (w == 0) {
                float part = 0.f;
                for (int tl = lane; tl < n; tl += 32) {
                    const float wv = al_s[tl] * dal_s[tl];
                    w_s[tl] = wv;
                    part += wv;
                }
                part = warp_sum(part);
                if (lane == 0) sred[8 * A] = part;
            }
            mbar_wait_par(&hbar, hph);
            hph ^= 1u;
            __syncthreads();
            {
                const int nsub = min(NTH / A, 4);
                const int a = tid % A, sg = tid / A;
                if (sg < nsub) {
                    const float ya = y_s[a];
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
                    for (int tl = sg; tl < n; tl += nsub) {
                        const float th = tanh_fast(hfbuf[tl * A + a] + ya);
                        const float om = 1.f - th * th;
                        s1 = fmaf(w_s[tl], om, s1);
                        s2 = fmaf(al_s[tl], om, s2);
                    }
                    sred[sg * 2 * A + a] = s1;
                    sred[sg * 2 * A + A + a] = s2;
                }
                __syncthreads();
                float* pp = apart + ((size_t)b * 2 + half) * (2 * A + 4);
                for (int i = tid; i < 2 * A; i += NTH) {
                    float v = 0.f;
                    for (int j = 0; j < nsub; ++j) v += sred[j * 2 * A + i];
                    pp[i] = v;
                }
                if (tid == 0) {
                    pp[2 * A] = sred[8 * A];
                    adots[((size_t)t * B + b) * 2 + half] = sred[8 * A];
                }
                // raw alpha * dalpha; dec_dhf_kernel subtracts alpha * dot (the row total is not known here)
                const int hi = half == 0 ? TC : Tn - TC;
                for (int tl = tid; tl < hi; tl += NTH) p.ds[row * Tn + tau0 + tl] = tl < n ? w_s[tl] : 0.f;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (p2p) tile_arrive(pp.at(0, b / 16));
          }
        } else {
            float* y_s = z_s;                   // [A]
            float* v_s = y_s + A;               // [A]
            float* dctx_s = v_s + A;            // [D]
            float* ds_s = dctx_s + D;           // [Tn]
            float* acc_s = ds_s + Tn;           // [8][A]
            float* redw = acc_s + 8 * A;        // [8]
            for (int b = blockIdx.x; b < B; b += gridDim.x) {
                const int len = min(p.enc_len[b], Tn);
                const size_t row = (size_t)t * B + b;
                __syncthreads();
                if (p2p && t < U - 1) tile_wait(pp.at(2, b / 16), (unsigned)(NXB * (U - 1 - t)), p.err);
                for (int a = tid; a < A; a += NTH) { y_s[a] = p.y[row * A + a]; v_s[a] = p.attn_v[a]; }
                for (int dd = tid; dd < D; dd += NTH) dctx_s[dd] = __ldcg(p.dcat + row * CAT + Hd + dd);
                __syncthreads();
                const float* al = p.alpha + row * Tn;
                const float* encb = p.enc + (size_t)b * Tp * D;
                float part = 0.f;
#pragma unroll 2
                for (int tau = w; tau < len; tau += 8) {
                    float acc = 0.f;
                    for (int dd = lane * 4; dd < D; dd += 128) {
                        float4 e = __ldg(reinterpret_cast<const float4*>(encb + (size_t)tau * D + dd));
                        acc = fmaf(dctx_s[dd], e.x, acc); acc = fmaf(dctx_s[dd + 1], e.y, acc);
                        acc = fmaf(dctx_s[dd + 2], e.z, acc); acc = fmaf(dctx_s[dd + 3], e.w, acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) { ds_s[tau] = acc; part += al[tau] * acc; }
                }
                if (lane == 0) redw[w] = part;
                __syncthreads();
                float dot = 0.f;
                for (int ww = 0; ww < 8; ++ww) dot += redw[ww];
                for (int tau = tid; tau < Tn; tau += NTH) {
                    float dsv = tau < len ? al[tau] * (ds_s[tau] - dot) : 0.f;
                    if (tau < len) ds_s[tau] = dsv;
                    p.ds[row * Tn + tau] = dsv;
                }
                __syncthreads();
                const float* HFb = p.HF + (size_t)b * Tp * A;
                for (int a0 = 0; a0 < A; a0 += 32) {
                    const int a = a0 + lane;
                    float dya = 0.f;
                    if (a < A)
#pragma unroll 4
                        for (int tau = w; tau < len; tau += 8) {
                            float th = tanhf(__ldg(HFb + (size_t)tau * A + a) + y_s[a]);
                            dya += ds_s[tau] * v_s[a] * (1.f - th * th);
                        }
                    if (a < A) acc_s[w * A + a] = dya;
                }
                __syncthreads();
                for (int a = tid; a < A; a += NTH) {
                    float dya = 0.f;
                    for (int ww = 0; ww < 8; ++ww) dya += acc_s[ww * A + a];
                    p.dy[row * A + a] = dya;
                }
                if (p2p) tile_arrive(pp.at(0, b / 16));
            }
        }
        DEC_STAMP(t, 1);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 2);
        // ------------------------------------------------------------ phase P: dc_new += dy . q_k^T ; pointwise backward
        {
            float* dy_s = z_s;                       // [16][A+4]
            const int DS_ = A + 4;
            for (int tile = blockIdx.x; tile < ptiles; tile += gridDim.x) {
                const int rb = tile / NCBp, cb = tile % NCBp;
                const int nvalid = min(16, B - rb * 16);
                __syncthreads();
                if (p2p) tile_wait(pp.at(0, rb), (unsigned)(a_arr * nvalid * (U - t)), p.err);   // A' of the block
                if (fastAp) {
                    // dy_a = v_a (S1_a - dot S2_a) from the two time-halves' partial sums (L2 scratch); float4
                    // columns, every load of a thread in flight together
                    const int nq = A / 4;
#pragma unroll 2
                    for (int i = tid; i < 16 * nq; i += NTH) {
                        const int r = i / nq, q = i % nq, b = rb * 16 + r;
                        float4 dyv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (b < B) {
                            const float* p0 = apart + (size_t)b * 2 * (2 * A + 4);
                            const float* p1 = p0 + (2 * A + 4);
                            const float4 s10 = __ldcg(reinterpret_cast<const float4*>(p0) + q);
                            const float4 s11 = __ldcg(reinterpret_cast<const float4*>(p1) + q);
                            const float4 s20 = __ldcg(reinterpret_cast<const float4*>(p0 + A) + q);
                            const float4 s21 = __ldcg(reinterpret_cast<const float4*>(p1 + A) + q);
                            const float dot = __ldcg(p0 + 2 * A) + __ldcg(p1 + 2 * A);
                            const float4 vv = __ldg(reinterpret_cast<const float4*>(p.attn_v) + q);
                            dyv.x = vv.x * ((s10.x + s11.x) - dot * (s20.x + s21.x));
                            dyv.y = vv.y * ((s10.y + s11.y) - dot * (s20.y + s21.y));
                            dyv.z = vv.z * ((s10.z + s11.z) - dot * (s20.z + s21.z));
                            dyv.w = vv.w * ((s10.w + s11.w) - dot * (s20.w + s21.w));
                            if (cb == 0) *reinterpret_cast<float4*>(p.dy + ((size_t)t * B + b) * A + q * 4) = dyv;
                        }
                        *reinterpret_cast<float4*>(dy_s + r * DS_ + q * 4) = dyv;
                    }
                } else {
                    if (tid == 0) mbar_expect_tx(&lbar, (uint32_t)(nvalid * A * 4));
                    __syncwarp();
                    if (lane * 8 + w < nvalid) {
                        const int r = lane * 8 + w;
                        bulk_g2s(dy_s + r * DS_, p.dy + ((size_t)t * B + rb * 16 + r) * A, (uint32_t)(A * 4), &lbar);
                    }
                    if (nvalid < 16)
                        for (int i = tid; i < 16 * A; i += NTH)
                            if (i / A >= nvalid) dy_s[(i / A) * DS_ + i % A] = 0.f;
                }
                if (!pres_ok) load_qp(cb);
                // operands of this thread's (row, unit) element: in flight during the product
                const int prow = tid / 8, ul = tid % 8;
                const int eb = rb * 16 + prow, unit = cb * 8 + ul;
                const bool own = tid < 128 && eb < B;
                const size_t row = (size_t)t * B + eb;
                const bool live = own && t < p.lens[eb];
                float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
                float cnew = 0.f, dhn = 0.f, dcin = 0.f, carry = 0.f, cpv = 0.f;
                if (live) {
                    act = *reinterpret_cast<const float4*>(p.acts + row * G4 + unit * 4);
                    cnew = p.cat[row * CAT + unit];
                    dhn = (t + 1 < U) ? __ldcg(p.dch + (row + B) * K + D + unit) : 0.f;
                    dcin = __ldcg(p.dcat + row * CAT + unit);
                    carry = p.dc_carry[(size_t)eb * Hd + unit];
                    cpv = p.cprev[row * Hd + unit];
                }
                if (!fastAp) {
                    mbar_wait_par(&lbar, lph);
                    lph ^= 1u;
                }
                if (nvalid < 16 || !pres_ok || fastAp) __syncthreads();
                const int ksteps = A / 8, per = (ksteps + 7) / 8;
                const int k0 = min(ksteps, w * per) * 8, k1 = min(ksteps, (w + 1) * per) * 8;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                mma_block(d, dy_s, DS_, pres, QS_, k0, k1, g, tq);
                *reinterpret_cast<float4*>(red + (w * 32 + lane) * 4) = make_float4(d[0], d[1], d[2], d[3]);
                __syncthreads();
                // 16 rows x 8 units = 128 (row, unit) elements: thread tid < 128
                if (own) {
                    // fragment element (prow, ul): lane = (prow % 8) * 4 + ul / 2, reg = (prow / 8) * 2 + ul % 2
                    const int fl = (prow % 8) * 4 + ul / 2, fr = (prow / 8) * 2 + (ul % 2);
                    float dcq = 0.f;
#pragma unroll
                    for (int ww = 0; ww < 8; ++ww) dcq += red[(ww * 32 + fl) * 4 + fr];
                    float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
                    float dcc = 0.f;
                    if (live) {
                        const float si = act.x, tj = act.y, sf = act.z, so = act.w;
                        const float tc = tanh_fast(cnew);
                        const float dc = dcin + dcq + carry + dhn * so * (1.f - tc * tc);
                        dz.x = dc * tj * si * (1.f - si);
                        dz.y = dc * si * (1.f - tj * tj);
                        dz.z = dc * cpv * sf * (1.f - sf);
                        dz.w = dhn * tc * so * (1.f - so);
                        dcc = dc * sf;
                    }
                    p.dc_carry[(size_t)eb * Hd + unit] = dcc;
                    *reinterpret_cast<float4*>(p.dz + row * G4 + unit * 4) = dz;
                }
                if (p2p) tile_arrive(pp.at(1, rb));
            }
        }
        DEC_STAMP(t, 3);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 4);
        // ------------------------------------------------------------ phase X: [dctx_{t-1} | dh_{t-1}] = dz_t . W_ch^T
        for (int tile = blockIdx.x; tile < xtiles; tile += gridDim.x) {
            const int rb = tile / NXB, xb = tile % NXB;
            const int nvalid = min(16, B - rb * 16);
            __syncthreads();
            if (!resident) load_wt2(xb);
            if (p2p) tile_wait(pp.at(1, rb), (unsigned)(NCBp * (U - t)), p.err);                 // dz_t of the block
            if (tid == 0) mbar_expect_tx(&lbar, (uint32_t)(nvalid * G4 * 4));
            __syncwarp();
            if (lane * 8 + w < nvalid) {
                const int r = lane * 8 + w;
                bulk_g2s(z_s + r * ZS, p.dz + ((size_t)t * B + rb * 16 + r) * G4, (uint32_t)(G4 * 4), &lbar);
            }
            if (nvalid < 16)
                for (int i = tid; i < 16 * G4; i += NTH)
                    if (i / G4 >= nvalid) z_s[(i / G4) * ZS + i % G4] = 0.f;
            mbar_wait_par(&lbar, lph);
            lph ^= 1u;
            if (nvalid < 16 || !resident) __syncthreads();
            {
                // the 8 warps split K = 4 Hd; every warp produces a partial [16 x 24] tile (3 n-tiles)
                float d[3][4];
                const int ksteps = G4 / 8, per = (ksteps + 7) / 8;
                const int k0 = min(ksteps, w * per) * 8, k1 = min(ksteps, (w + 1) * per) * 8;
                mma_ksplit<3>(d, z_s, ZS, Wt2, WS, k0, k1, g, tq);
                float* xw = xred + w * (16 * 40);
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    *reinterpret_cast<float2*>(xw + g * 40 + 8 * nt + 2 * tq) = make_float2(d[nt][0], d[nt][1]);
                    *reinterpret_cast<float2*>(xw + (g + 8) * 40 + 8 * nt + 2 * tq) = make_float2(d[nt][2], d[nt][3]);
                }
            }
            __syncthreads();
            for (int idx = tid; idx < 16 * NX; idx += NTH) {
                const int r = idx / NX, c = idx % NX;
                const int b = rb * 16 + r, col = xb * NX + c;
                if (b < B && col < K) {
                    float v = 0.f;
#pragma unroll
                    for (int ww = 0; ww < 8; ++ww) v += xred[ww * (16 * 40) + r * 40 + c];
                    const size_t row = (size_t)t * B + b;
                    p.dch[row * K + col] = v;
                    // dctx_{t-1} joins the AttnProjection part already in dcat[t-1]
                    if (col < D && t > 0) p.dcat[(row - B) * CAT + Hd + col] += v;
                }
            }
            if (p2p) tile_arrive(pp.at(2, rb));
        }
        DEC_STAMP(t, 5);
        if (!p2p) grid_barrier(p.ctr, epoch, p.err);
        DEC_STAMP(t, 6);
    }
}

// ======================================================================= deferred sums over t
// denc[b,tau,:] += sum_t alpha_t[b,tau] * dctx_t[b,:]      (one CTA per (b, tau))
__global__ void __launch_bounds__(256) dec_denc_kernel(e2e_dec_persist_args p, float* __restrict__ denc) {
    const int b = blockIdx.x / p.Tn, tau = blockIdx.x % p.Tn;
    if (tau >= min(p.enc_len[b], p.Tn)) return;
    const int CAT = p.Hd + p.D;
    extern __shared__ float al_s[];          // [U]
    for (int t = threadIdx.x; t < p.U; t += blockDim.x) al_s[t] = p.alpha[((size_t)t * p.B + b) * p.Tn + tau];
    __syncthreads();
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < p.U; ++t) acc = fmaf(al_s[t], p.dcat[((size_t)t * p.B + b) * CAT + p.Hd + d], acc);
        denc[((size_t)b * p.Tp + tau) * p.D + d] += acc;
    }
}
// dHF[b,tau,a] = sum_t ds_t[b,tau] v_a (1 - th^2), th = tanh(HF[b,tau,a] + y_t[b,a]); dv_part[b*Tn+tau, a] = sum_t ds th
__global__ void __launch_bounds__(128) dec_dhf_kernel(e2e_dec_persist_args p, float* __restrict__ dHF,
                                                      float* __restrict__ dv_part, int attn_fast) {
    const int b = blockIdx.x / p.Tn, tau = blockIdx.x % p.Tn;
    const bool valid = tau < min(p.enc_len[b], p.Tn);
    extern __shared__ float ds_sm[];         // [U]
    // fast attention backward stored alpha*dalpha: ds = alpha (dalpha - dot), dot = sum of the two halves' partials
    const float* adots = reinterpret_cast<const float*>(p.ctr) + 64 + (size_t)p.B * 2 * (2 * p.A + 4);
    for (int t = threadIdx.x; t < p.U; t += blockDim.x) {
        const size_t row = (size_t)t * p.B + b;
        float v = valid ? p.ds[row * p.Tn + tau] : 0.f;
        if (valid && attn_fast) v -= p.alpha[row * p.Tn + tau] * (adots[row * 2] + adots[row * 2 + 1]);
        ds_sm[t] = v;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < p.A; a += blockDim.x) {
        float acc = 0.f, accv = 0.f;
        if (valid) {
            const float hf = p.HF[((size_t)b * p.Tp + tau) * p.A + a], va = p.attn_v[a];
            for (int t = 0; t < p.U; ++t) {
                float dsv = ds_sm[t];
                if (dsv != 0.f) {
                    float th = tanhf(hf + p.y[((size_t)t * p.B + b) * p.A + a]);
                    acc += dsv * va * (1.f - th * th);
                    accv = fmaf(dsv, th, accv);
                }
            }
            dHF[((size_t)b * p.Tp + tau) * p.A + a] = acc;
        }
        dv_part[(size_t)blockIdx.x * p.A + a] = accv;
    }
}

// floats before attn_buf: Wt | a_s | red | gred | qres
static size_t fwd_base_floats(const e2e_dec_persist_args& p) {
    int K = p.D + p.Hd;
    return (size_t)32 * (K + 8) + 16 * (K + 4) + 8 * 32 * 4 + 8 * 16 * 40 + 8 * (p.Hd + 4);
}
static size_t fwd_smem_bytes(const e2e_dec_persist_args& p) {
    int K = p.D + p.Hd;
    size_t aph = (size_t)32 * (K + 8) + 2 * p.A + p.Tn + 16;      // streaming attention fallback scratch (in a_s)
    return sizeof(float) * (max(fwd_base_floats(p), aph) + 64);
}
// staging buffer (floats) of the fast attention read-out, 0 when the shapes do not qualify
static int fwd_attn_cap(const e2e_dec_persist_args& p, size_t* smem) {
    int K = p.D + p.Hd;
    if (p.Tn > 128 || p.A > 128 || p.D / 8 > NTH || NTH % (p.D / 8) != 0) return 0;
    int cap = max(p.Tn * p.A, ((p.Tn + 1) / 2) * (p.D / 2));
    cap = (cap + 3) / 4 * 4;
    if (cap > 16 * (K + 4)) return 0;
    size_t total = sizeof(float) * (fwd_base_floats(p) + cap + 2 * p.A + p.Tn + 64);
    if (total > 227 * 1024) return 0;
    *smem = max(*smem, total);
    return cap;
}
static size_t bwd_smem_bytes(const e2e_dec_persist_args& p) {
    int G4 = 4 * p.Hd;
    size_t x = (size_t)24 * (G4 + 8) + 16 * (G4 + 4) + 8 * 32 * 4 + 8 * 16 * 40 + 8 * (p.A + 4);
    size_t a = (size_t)24 * (G4 + 8) + 2 * p.A + p.D + p.Tn + 8 * p.A + 16;
    size_t pp = (size_t)24 * (G4 + 8) + 16 * (p.A + 4) + 8 * (p.A + 8) + 16 * (G4 + 4) + 8 * 32 * 4;
    return sizeof(float) * (max(x, max(a, pp)) + 64);
}

// 1 when the time-split attention backward fits: shared-memory carve-up of dec_bwd_persist_kernel + L2 scratch
static int bwd_attn_fast(const e2e_dec_persist_args& p, size_t* smem) {
    const int G4 = 4 * p.Hd, TC = (p.Tn + 1) / 2, TCp = (TC + 3) & ~3;
    if (p.Tn > 128 || p.A > NTH || p.A % 4 != 0 || p.D % 4 != 0) return 0;
    const size_t enc_cap = (size_t)16 * (G4 + 4) + 8 * 32 * 4 + 8 * 16 * 40;     // z_s | red | xred
    if ((size_t)TC * p.D > enc_cap) return 0;
    const size_t base = (size_t)24 * (G4 + 8) + enc_cap + 8 * (p.A + 4);
    const size_t extra = (size_t)TCp * p.A + 2 * p.A + p.D + 3 * TCp + 8 * p.A + 8;
    const size_t total = sizeof(float) * (base + extra + 64);
    if (total > 227 * 1024) return 0;
    if ((size_t)p.B * 2 * (2 * p.A + 4) + (size_t)p.U * p.B * 2 + 64 > (size_t)(1 << 18)) return 0;   // ctr scratch >= 1 MB
    *smem = max(*smem, total);
    return 1;
}

// 1 when both persistent kernels fit the 227 KB of shared memory for these shapes (else the caller uses the
// per-step kernels of decoder.cu): e.g. cfg-5's D = 1024 needs 280 KB for the W_ch^T slice + [ctx|h] tile.
int dec_persist_fits(const e2e_dec_persist_args* a) {
    const e2e_dec_persist_args& p = *a;
    if (p.Hd % 8 != 0 || p.A % 8 != 0 || p.D % 8 != 0) return 0;
    if (!(2 * p.A + p.Tn + 32 + 4 * p.D <= 16 * (p.D + p.Hd + 4) && 10 * p.A + p.D + p.Tn + 16 <= 16 * (4 * p.Hd + 4)))
        return 0;
    return fwd_smem_bytes(p) <= 227 * 1024 && bwd_smem_bytes(p) <= 227 * 1024;
}

// The sums over the decoder steps of the attention backward -- denc += sum_t alpha_t dctx_t, dHF and dv_part [B*Tn, A]
// -- from stored per-step tensors; also serves the step-by-step loop (e2e_decoder_loop_bwd).  Reads B, U, Hd, A, D, Tn,
// Tp, enc_len, alpha, dcat, ds, y, HF, attn_v of `p`.
int dec_deferred_attn_grads(cudaStream_t st, const e2e_dec_persist_args& p, float* denc, float* dHF, float* dv_part) {
    if (p.B <= 0 || p.U <= 0 || p.Tn <= 0) return 0;
    dec_denc_kernel<<<p.B * p.Tn, 256, sizeof(float) * p.U, st>>>(p, denc);
    E2E_LAUNCH_CHECK();
    dec_dhf_kernel<<<p.B * p.Tn, 128, sizeof(float) * p.U, st>>>(p, dHF, dv_part, 0);
    E2E_LAUNCH_CHECK();
    return 0;
}

int dec_persist(cudaStream_t st, bool bwd, const e2e_dec_persist_args* a, float* denc, float* dHF, float* dv_part) {
    e2e_dec_persist_args p = *a;
    E2E_REQUIRE(p.Hd % 8 == 0 && p.A % 8 == 0 && p.D % 8 == 0, "decoder_persist: Hd, A, D must be multiples of 8");
    E2E_REQUIRE(2 * p.A + p.Tn + 32 + 4 * p.D <= 16 * (p.D + p.Hd + 4) &&
                    10 * p.A + p.D + p.Tn + 16 <= 16 * (4 * p.Hd + 4),
                "decoder_persist: attention length %d too long for the shared-memory scratch", p.Tn);
    if (p.B <= 0 || p.U <= 0) return 0;
    const void* fn = bwd ? (const void*)dec_bwd_persist_kernel : (const void*)dec_fwd_persist_kernel;
    size_t smem = bwd ? bwd_smem_bytes(p) : fwd_smem_bytes(p);
    E2E_REQUIRE(smem <= 227 * 1024, "decoder_persist: shapes need %zu B of shared memory", smem);
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nsm = sm_count();
    int nrb = (p.B + 15) / 16;
    int want = bwd ? max(max(p.B, nrb * (p.Hd / 8)), nrb * ((p.D + p.Hd + 23) / 24))
                   : max(max(2 * p.B, nrb * (p.Hd / 8)), nrb * (p.A / 8));
    int grid = min(nsm, want);
    E2E_CHECK_CUDA(cudaMemsetAsync(p.ctr, 0, 64 * sizeof(unsigned), st));   // barrier counter + row-block counters
    // point-to-point row-block counters: 3 phases x nrb counters in words 1 .. 63 of the scratch, `stride` words apart
    int p2p_stride = 0;
    if (g_dec_p2p && 3 * nrb <= 62) p2p_stride = max(1, min(8, 62 / (3 * nrb)));
    {
        static long long* last_dbg = nullptr;
        if (g_rec_dbg != last_dbg) {
            last_dbg = g_rec_dbg;
            E2E_CHECK_CUDA(cudaMemcpyToSymbolAsync(d_dec_dbg, &last_dbg, sizeof(last_dbg), 0, cudaMemcpyHostToDevice, st));
        }
    }
    int attn_cap = bwd ? bwd_attn_fast(p, &smem) : fwd_attn_cap(p, &smem);
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&p, &attn_cap, &p2p_stride};
    E2E_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NTH), args, smem, st));
    ++g_launches;
    if (bwd) {
        dec_denc_kernel<<<p.B * p.Tn, 256, sizeof(float) * p.U, st>>>(p, denc);
        E2E_LAUNCH_CHECK();
        dec_dhf_kernel<<<p.B * p.Tn, 128, sizeof(float) * p.U, st>>>(p, dHF, dv_part, attn_cap != 0 && 2 * p.B <= grid);
        E2E_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace e2e
