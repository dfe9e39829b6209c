// Register-resident LSTM recurrence for H in {128, 256} (forward and backward): the
// fast path of lstm_rec.cu / lstm_rec_cluster.cu.
//
// Same decomposition as lstm_rec_cluster.cu -- one 16-row batch slice of one direction
// is owned by ONE thread-block cluster of CS = H/16 CTAs, CTA j owns hidden units
// [16j, 16j+16) -- but designed from measurements on B200 (scratch/ubench.cu):
//   * mma.sync m16n8k8.tf32 issues at 8 cycles / SMSP; the old kernel spent 2.3x that per
//     step on shared-memory operand loads and on re-splitting W_hh.  Here the CTA's W_hh
//     slice lives in REGISTERS for the whole sequence as pre-split 3xTF32 B fragments
//     (64 hi + 64 lo registers per thread), so a step's k-loop is 16 LDS.128 + the
//     fp32->tf32 split of the h fragment + 96 MMAs per warp.  The K order inside the
//     contraction is permuted so one LDS.128 feeds two k-tiles (conflict-free, no padding).
//   * a remote-store + release/acquire exchange of h_t costs ~2700 cycles per step in a
//     16-CTA cluster; a 1 KB tile written to L2 and delivered to all 16 CTAs by ONE
//     multicast bulk copy (cp.async.bulk ... .multicast::cluster, data + mbarrier
//     transaction count in one operation) costs ~900.  Forward uses that (all-gather of
//     h_t); backward reduce-scatters the K-sliced partial d h_{t-1} with bulk
//     shared->distributed-shared copies (~1400).
//   * only 7 clusters of 16 single-CTA-per-SM blocks are co-resident on a B200.  With
//     NS = 2 a cluster interleaves TWO independent batch slices of the same direction
//     step by step, so the exchange of one slice is hidden behind the arithmetic of the
//     other and cfg-2 (8 slices) needs 4 clusters = 64 SMs.
// Layouts and length semantics are those documented in lstm_rec.cu.
#include "rec_frag.cuh"

namespace e2e {

namespace {

// ---------------------------------------------------------------- forward
// Warp w = (n-group ng = w % 4, k-half kh = w / 4): gate columns of the CTA's units [4ng, 4ng+4)
// (n-tile 0 = gates i,j, n-tile 1 = gates f,o), source tiles [kh*CS/2, (kh+1)*CS/2).  After the
// k-loop thread (g, tq) holds all four gate pre-activations of unit 4ng+tq for rows g and g+8; the
// two k-halves swap one float4 through shared memory (kh=0 finishes row g, kh=1 row g+8).
template <int CS, int NS>
__global__ void __launch_bounds__(NTH, 1) rec_fwd_mc_kernel(MParams p) {
    constexpr int H = CS * UPC;
    constexpr int KT = CS;                                       // k8-tiles per k-half (H / 2 / 8)
    extern __shared__ __align__(128) float smem[];
    float* h_s = smem;                                           // [NS][2][CS][TILE]
    float4* xchg = reinterpret_cast<float4*>(h_s + NS * 2 * CS * TILE);   // [2][8 warps][32 lanes]
    __shared__ __align__(8) uint64_t full[NS][2];
    __shared__ unsigned pubcnt[NS];                              // warps that have written their part of the tile

    const int ndir = p.ndir, T = p.T;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS;
    const int dir = cl % ndir;
    const int slice0 = (cl / ndir) * NS;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;
    const int ng = w % 4, kh = w / 4;

    if (tid == 0) {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_init(&full[sl][0], 1); mbar_init(&full[sl][1], 1); pubcnt[sl] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_expect_tx(&full[sl][0], CS * TILE * 4); mbar_expect_tx(&full[sl][1], CS * TILE * 4); }
    }
    // resident W_hh fragments: B[k][n], k = hidden unit of h_{t-1}, n = own gate column
    uint32_t bh[KT][2][2], bl[KT][2][2];
    {
        const float* Wg = p.Wh + (size_t)dir * H * H * 4;
        const int ncol_unit = rank * UPC + 4 * ng + (g >> 1);
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * (kh * KT + kt) + tq + 4 * e;
                    const float x = Wg[((size_t)k * H + ncol_unit) * 4 + 2 * nt + (g & 1)];
                    bh[kt][nt][e] = cvt_tf32(x);
                    bl[kt][nt][e] = cvt_tf32(x - __uint_as_float(bh[kt][nt][e]));
                }
        if (kMixed) {
            // bl[2q][nt] := bf16 W of pair q (b0 = tile 2q rows tq, tq+4; b1 = tile 2q+1), bl[2q+1][nt] := bf16 (W - tf32 W)
#pragma unroll
            for (int q = 0; q < KT / 2; ++q)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float x[2][2], r[2][2];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 8 * (kh * KT + 2 * q + j) + tq + 4 * e;
                            x[j][e] = Wg[((size_t)k * H + ncol_unit) * 4 + 2 * nt + (g & 1)];
                            r[j][e] = x[j][e] - __uint_as_float(bh[2 * q + j][nt][e]);
                        }
                    bl[2 * q][nt][0] = pack_bf16(x[0][0], x[0][1]);
                    bl[2 * q][nt][1] = pack_bf16(x[1][0], x[1][1]);
                    bl[2 * q + 1][nt][0] = pack_bf16(r[0][0], r[0][1]);
                    bl[2 * q + 1][nt][1] = pack_bf16(r[1][0], r[1][1]);
                }
        }
    }
    // this thread's pointwise element of every slice: row g + 8 kh, unit 4 ng + tq
    const int prow = g + 8 * kh;
    const int ul = 4 * ng + tq;
    const int unit = rank * UPC + ul;
    int pb[NS], plen[NS];
    float c_reg[NS], h_reg[NS];
    bool live[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        live[sl] = slice0 + sl < p.nslices;
        pb[sl] = (slice0 + sl) * R + prow;
        plen[sl] = (live[sl] && pb[sl] < p.B) ? p.lens[pb[sl]] : 0;
        c_reg[sl] = 0.f;
        h_reg[sl] = 0.f;
    }
    uint32_t phase = 0;                                          // bit (2 sl + buf)
    const bool rec = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    // x-projection pre-activations, prefetched one step ahead
    // element (b, t, dir, unit) of G (float4), Cst and Hout is idx = ((b sb + t st) ndir + dir) H + unit: kept as a
    // running index so the step loop carries no 64-bit address arithmetic
    float4 gxn[NS];
    long long idx[NS];
    const long long tstep = (dir == 0 ? 1 : -1) * p.st * ndir * H;
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        gxn[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int t0 = dir == 0 ? 0 : T - 1;
        idx[sl] = (((long long)pb[sl] * p.sb + (long long)t0 * p.st) * ndir + dir) * H + unit;
        if (t0 < plen[sl]) gxn[sl] = reinterpret_cast<const float4*>(p.G)[idx[sl]];
    }
    __syncthreads();
    cluster_sync_all();

    for (int s = 0; s < T; ++s) {
        const int buf = s & 1;
        const int t = dir == 0 ? s : T - 1 - s;
        const int tn = dir == 0 ? s + 1 : T - 2 - s;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            if (!live[sl]) continue;
            const bool active = t < plen[sl];
            const long long ix = idx[sl];
            idx[sl] = ix + tstep;
            const float4 gx = gxn[sl];
            if (rec) p.dbg[(s * NS + sl) * 8 + 0] = clock64();
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            if (s > 0) {
                float acc[2][4], accx[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { acc[nt][q] = 0.f; accx[nt][q] = 0.f; }
                mbar_wait(&full[sl][buf], (phase >> (2 * sl + buf)) & 1u);
                phase ^= 1u << (2 * sl + buf);
                if (tid == 0) mbar_expect_tx(&full[sl][buf], CS * TILE * 4);     // arm this buffer's next phase
                if (rec) p.dbg[(s * NS + sl) * 8 + 1] = clock64();
                const float* hb = h_s + (size_t)(sl * 2 + buf) * CS * TILE + kh * KT * 128 + lane * 4;
                if (kMixed) {
                    float4 a0 = *reinterpret_cast<const float4*>(hb), a1 = *reinterpret_cast<const float4*>(hb + 128);
#pragma unroll
                    for (int q = 0; q < KT / 2; ++q) {
                        const int qn = q + 1 < KT / 2 ? q + 1 : q;
                        const float4 n0 = *reinterpret_cast<const float4*>(hb + (2 * qn) * 128);
                        const float4 n1 = *reinterpret_cast<const float4*>(hb + (2 * qn + 1) * 128);
                        k16_mma<2>(acc, accx, a0, a1, bh[2 * q], bh[2 * q + 1], bl[2 * q], bl[2 * q + 1]);
                        a0 = n0; a1 = n1;
                    }
                } else {
                    float4 a_cur = *reinterpret_cast<const float4*>(hb);
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) {
                        const float4 a_nxt = *reinterpret_cast<const float4*>(hb + (kt + 1 < KT ? kt + 1 : kt) * 128);
                        ktile_mma<2>(acc, accx, a_cur, bh[kt], bl[kt]);
                        a_cur = a_nxt;
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[nt][q] += accx[nt][q];
                if (rec) p.dbg[(s * NS + sl) * 8 + 2] = clock64();
                // k-half combine: send the partials of the row the partner warp finishes
                const float4 send = kh == 0 ? make_float4(acc[0][2], acc[0][3], acc[1][2], acc[1][3])
                                            : make_float4(acc[0][0], acc[0][1], acc[1][0], acc[1][1]);
                // double-buffered by slice-step parity: the pair barrier of the step in between orders the reuse
                float4* xb = xchg + (((s * NS + sl) & 1) * 8) * 32;
                xb[w * 32 + lane] = send;
                pair_barrier(1 + ng);
                const float4 recv = xb[(w ^ 4) * 32 + lane];
                z[0] = (kh == 0 ? acc[0][0] : acc[0][2]) + recv.x;
                z[1] = (kh == 0 ? acc[0][1] : acc[0][3]) + recv.y;
                z[2] = (kh == 0 ? acc[1][0] : acc[1][2]) + recv.z;
                z[3] = (kh == 0 ? acc[1][1] : acc[1][3]) + recv.w;
            }
            if (rec) p.dbg[(s * NS + sl) * 8 + 3] = clock64();
            float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
            float cn = 0.f;
            if (active) {
                const float si = sigmoid_fast(z[0] + gx.x);
                const float tj = tanh_fast(z[1] + gx.y);
                const float sf = sigmoid_fast(z[2] + gx.z + 1.0f);
                const float so = sigmoid_fast(z[3] + gx.w);
                cn = c_reg[sl] * sf + si * tj;
                h_reg[sl] = tanh_fast(cn) * so;
                c_reg[sl] = cn;
                act = make_float4(si, tj, sf, so);
            }
            // publish h_t (state h: carried through for masked rows): tile -> L2 -> multicast to the cluster
            if (rec) p.dbg[(s * NS + sl) * 8 + 4] = clock64();
            if (s + 1 < T) {
                float* gt = p.xg + ((size_t)(sl * 2 + buf) * gridDim.x + blockIdx.x) * TILE;
                gt[(ng >> 1) * 128 + lane * 4 + 2 * (ng & 1) + kh] = h_reg[sl];      // fragment order
                asm volatile("fence.proxy.async.global;" ::: "memory");
                if (rec) p.dbg[(s * NS + sl) * 8 + 5] = clock64();
                __syncwarp();
                // no CTA barrier: the warp that completes the tile (8th arrival of this slice-step) multicasts it
                if (lane == 0) {
                    unsigned old;
                    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(s_u32(&pubcnt[sl])) : "memory");
                    if ((old & 7u) == 7u) {
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        bulk_multicast(s_u32(h_s + ((size_t)(sl * 2 + (buf ^ 1)) * CS + rank) * TILE), gt, TILE * 4,
                                       s_u32(&full[sl][buf ^ 1]), (uint16_t)((1u << CS) - 1u));
                    }
                }
            }
            if (rec) p.dbg[(s * NS + sl) * 8 + 6] = clock64();
            if (s + 1 < T && tn < plen[sl])
                gxn[sl] = reinterpret_cast<const float4*>(p.G)[ix + tstep];
            if (active) {
                reinterpret_cast<float4*>(p.G)[ix] = act;
                p.Cst[ix] = cn;
                p.Hout[ix] = h_reg[sl];
            }
            if (rec) p.dbg[(s * NS + sl) * 8 + 7] = clock64();
        }
    }
    cluster_sync_all();      // nobody exits while a peer's multicast may still target its shared memory
}

// ---------------------------------------------------------------- backward
// Per step and slice: partial[16 rows x H] = dz_t[16 x 64 own gate columns] . W_hh[:, own columns]^T
// (warp w: hidden units [2 CS w, 2 CS (w+1)) = CS/8 destination CTAs), staged as CS tiles of 1 KB and
// reduce-scattered with one bulk copy per destination, issued by the warp that produced the tile (no
// CTA barrier); the receiver sums the CS partial tiles.
template <int CS, int NS>
__global__ void __launch_bounds__(NTH, 1) rec_bwd_mc_kernel(MParams p) {
    constexpr int H = CS * UPC;
    constexpr int NTL = CS / 4;                                  // n-tiles (8 hidden units) per warp
    constexpr int ND = CS / 8;                                   // destination CTAs per warp
    extern __shared__ __align__(128) float smem[];
    float* recv = smem;                                          // [NS][2][CS][TILE] received partial tiles
    float* stage = recv + NS * 2 * CS * TILE;                    // [NS][2][CS][TILE] partial tiles to send
    float* dz_s = stage + NS * 2 * CS * TILE;                    // [NS][4 chunks][16 rows][16 gate columns]
    __shared__ __align__(8) uint64_t full[NS][2];

    const int ndir = p.ndir, T = p.T, Tp = p.Tp;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS;
    const int dir = cl % ndir;
    const int slice0 = (cl / ndir) * NS;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;

    if (tid == 0) {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_init(&full[sl][0], 1); mbar_init(&full[sl][1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_expect_tx(&full[sl][0], CS * TILE * 4); mbar_expect_tx(&full[sl][1], CS * TILE * 4); }
    }
    // resident fragments: B[k][n] = W_hh[hidden unit n][own gate column k]
    uint32_t bh[8][NTL][2], bl[8][NTL][2];
    {
        const float* Wg = p.Wh + (size_t)dir * H * H * 4;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    // C-fragment columns (2tq, 2tq+1) of n-tiles (2d, 2d+1) are 4 consecutive units
                    const int n_unit = 16 * (w * ND + nt / 2) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
                    const int kcol = 8 * kt + tq + 4 * e;
                    const float x = Wg[(size_t)n_unit * H * 4 + rank * 4 * UPC + kcol];
                    bh[kt][nt][e] = cvt_tf32(x);
                    bl[kt][nt][e] = cvt_tf32(x - __uint_as_float(bh[kt][nt][e]));
                }
        if (kMixed) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int n_unit = 16 * (w * ND + nt / 2) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
                    float x[2][2], r[2][2];
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kcol = 8 * (2 * q + j) + tq + 4 * e;
                            x[j][e] = Wg[(size_t)n_unit * H * 4 + rank * 4 * UPC + kcol];
                            r[j][e] = x[j][e] - __uint_as_float(bh[2 * q + j][nt][e]);
                        }
                    bl[2 * q][nt][0] = pack_bf16(x[0][0], x[0][1]);
                    bl[2 * q][nt][1] = pack_bf16(x[1][0], x[1][1]);
                    bl[2 * q + 1][nt][0] = pack_bf16(r[0][0], r[0][1]);
                    bl[2 * q + 1][nt][1] = pack_bf16(r[1][0], r[1][1]);
                }
        }
    }
    // pointwise element of this thread: row = tid / UPC, unit ul = tid % UPC
    const int prow = tid / UPC, ul = tid % UPC;
    const int unit = rank * UPC + ul;
    int pb[NS], plen[NS];
    float dc_reg[NS];
    bool live[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        live[sl] = slice0 + sl < p.nslices;
        pb[sl] = (slice0 + sl) * R + prow;
        plen[sl] = (live[sl] && pb[sl] < p.B) ? p.lens[pb[sl]] : 0;
        dc_reg[sl] = 0.f;
        if (live[sl] && pb[sl] < p.B)
            for (int t = T; t < Tp; ++t)
                reinterpret_cast<float4*>(p.G)[(((size_t)pb[sl] * p.sb + (size_t)t * p.st) * ndir + dir) * H + unit] =
                    make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t phase = 0;
    const bool rec = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    // saved activations / states / upstream gradient of a step, prefetched one step ahead
    float4 act_n[NS];
    float cst_n[NS], cprev_n[NS], dout_n[NS];
    // running element index of (b, t, dir, unit) in G (float4) / Cst / dOut; the walk goes against the forward one
    long long idx[NS];
    const long long tstep = (dir == 0 ? -1 : 1) * p.st * ndir * H;
    auto prefetch = [&](int sl, int t, long long ix) {
        act_n[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
        cst_n[sl] = 0.f; cprev_n[sl] = 0.f; dout_n[sl] = 0.f;
        if (t >= 0 && t < plen[sl]) {
            const int t_cprev = dir == 0 ? t - 1 : t + 1;
            act_n[sl] = reinterpret_cast<const float4*>(p.G)[ix];
            cst_n[sl] = p.Cst[ix];
            if (t_cprev >= 0 && t_cprev < plen[sl]) cprev_n[sl] = p.Cst[ix + tstep];
            dout_n[sl] = __ldg(p.dOut + ix);
        }
    };
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const int t0 = dir == 0 ? T - 1 : 0;
        idx[sl] = (((long long)pb[sl] * p.sb + (long long)t0 * p.st) * ndir + dir) * H + unit;
        prefetch(sl, t0, idx[sl]);
    }
    __syncthreads();
    cluster_sync_all();

    for (int s = 0; s < T; ++s) {
        const int buf = s & 1;
        const int t = dir == 0 ? T - 1 - s : s;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            if (!live[sl]) continue;
            const bool active = t < plen[sl];
            const long long ix = idx[sl];
            idx[sl] = ix + tstep;
            const float4 act = act_n[sl];
            const float cst = cst_n[sl], cprev = cprev_n[sl], dout = dout_n[sl];
            if (rec) p.dbg[(s * NS + sl) * 8 + 0] = clock64();
            float dh = 0.f;
            if (s > 0) {
                mbar_wait(&full[sl][buf], (phase >> (2 * sl + buf)) & 1u);
                phase ^= 1u << (2 * sl + buf);
                if (tid == 0) mbar_expect_tx(&full[sl][buf], CS * TILE * 4);
                if (rec) p.dbg[(s * NS + sl) * 8 + 1] = clock64();
                const float* rb = recv + (size_t)(sl * 2 + buf) * CS * TILE + tid;
                float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int src = 0; src < CS; ++src) part[src & 3] += rb[src * TILE];
                dh = (part[0] + part[1]) + (part[2] + part[3]);
            }
            if (rec) p.dbg[(s * NS + sl) * 8 + 2] = clock64();
            float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                dh += dout;
                const float si = act.x, tj = act.y, sf = act.z, so = act.w;
                const float tc = tanh_fast(cst);
                const float dct = dc_reg[sl] + dh * so * (1.f - tc * tc);
                dz.x = dct * tj * si * (1.f - si);
                dz.y = dct * si * (1.f - tj * tj);
                dz.z = dct * cprev * sf * (1.f - sf);
                dz.w = dh * tc * so * (1.f - so);
                dc_reg[sl] = dct * sf;
            }
            float* dzs = dz_s + sl * 4 * TILE;
            {
                // own dz tile in fragment order [k-tile = ul/2][16-byte chunk (g*4 + gate) ^ k-tile][2 (ul%2) + row/8]
                float* dq = dzs + (ul >> 1) * 128 + 2 * (ul & 1) + (prow >> 3);
                const int c0 = (prow & 7) * 4, kx = ul >> 1;
                dq[((c0 + 0) ^ kx) * 4] = dz.x;
                dq[((c0 + 1) ^ kx) * 4] = dz.y;
                dq[((c0 + 2) ^ kx) * 4] = dz.z;
                dq[((c0 + 3) ^ kx) * 4] = dz.w;
            }
            if (pb[sl] < p.B) reinterpret_cast<float4*>(p.G)[ix] = dz;
            if (s + 1 < T) prefetch(sl, dir == 0 ? t - 1 : t + 1, ix + tstep);
            if (rec) p.dbg[(s * NS + sl) * 8 + 3] = clock64();
            __syncthreads();
            if (rec) p.dbg[(s * NS + sl) * 8 + 4] = clock64();
            if (s + 1 < T) {
                float acc[NTL][4], accx[NTL][4];
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { acc[nt][q] = 0.f; accx[nt][q] = 0.f; }
                if (kMixed) {
                    float4 a0 = *reinterpret_cast<const float4*>(dzs + lane * 4);
                    float4 a1 = *reinterpret_cast<const float4*>(dzs + 128 + ((lane ^ 1) * 4));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k0 = q + 1 < 4 ? 2 * q + 2 : 2 * q, k1 = k0 + 1;
                        const float4 n0 = *reinterpret_cast<const float4*>(dzs + k0 * 128 + ((lane ^ k0) * 4));
                        const float4 n1 = *reinterpret_cast<const float4*>(dzs + k1 * 128 + ((lane ^ k1) * 4));
                        k16_mma<NTL>(acc, accx, a0, a1, bh[2 * q], bh[2 * q + 1], bl[2 * q], bl[2 * q + 1]);
                        a0 = n0; a1 = n1;
                    }
                } else {
                    float4 a_cur = *reinterpret_cast<const float4*>(dzs + lane * 4);
#pragma unroll
                    for (int kt = 0; kt < 8; ++kt) {
                        const int kn = kt + 1 < 8 ? kt + 1 : kt;
                        const float4 a_nxt = *reinterpret_cast<const float4*>(dzs + kn * 128 + ((lane ^ kn) * 4));
                        ktile_mma<NTL>(acc, accx, a_cur, bh[kt], bl[kt]);
                        a_cur = a_nxt;
                    }
                }
                if (rec) p.dbg[(s * NS + sl) * 8 + 5] = clock64();
                float* sg = stage + ((size_t)(sl * 2 + buf) * CS + w * ND) * TILE + g * UPC + 4 * tq;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    *reinterpret_cast<float4*>(sg + d * TILE) =
                        make_float4(acc[2 * d][0] + accx[2 * d][0], acc[2 * d][1] + accx[2 * d][1],
                                    acc[2 * d + 1][0] + accx[2 * d + 1][0], acc[2 * d + 1][1] + accx[2 * d + 1][1]);
                    *reinterpret_cast<float4*>(sg + d * TILE + 8 * UPC) =
                        make_float4(acc[2 * d][2] + accx[2 * d][2], acc[2 * d][3] + accx[2 * d][3],
                                    acc[2 * d + 1][2] + accx[2 * d + 1][2], acc[2 * d + 1][3] + accx[2 * d + 1][3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (rec) p.dbg[(s * NS + sl) * 8 + 6] = clock64();
                __syncwarp();
                // this warp wrote whole tiles: tile `dest` of my partials goes to slot `rank` of peer `dest`
                if (lane < ND) {
                    const int dest = w * ND + lane;
                    dsmem_bulk_copy(mapa(s_u32(recv + ((size_t)(sl * 2 + (buf ^ 1)) * CS + rank) * TILE), dest),
                                    mapa(s_u32(&full[sl][buf ^ 1]), dest),
                                    stage + ((size_t)(sl * 2 + buf) * CS + dest) * TILE, TILE * 4);
                }
            }
            if (rec) p.dbg[(s * NS + sl) * 8 + 7] = clock64();
        }
    }
    cluster_sync_all();
}

size_t fwd_smem(int CS, int NS) { return sizeof(float) * ((size_t)NS * 2 * CS * TILE) + 2 * 8 * 32 * 16; }
size_t bwd_smem(int CS, int NS) { return sizeof(float) * ((size_t)NS * 4 * CS * TILE + (size_t)NS * 4 * TILE); }

template <int CS, int NS>
int launch_mc(cudaStream_t st, bool bwd, const MParams& p, int nclusters, int* max_active) {
    auto kf = rec_fwd_mc_kernel<CS, NS>;
    auto kb = rec_bwd_mc_kernel<CS, NS>;
    const void* fn = bwd ? (const void*)kb : (const void*)kf;
    size_t smem = bwd ? bwd_smem(CS, NS) : fwd_smem(CS, NS);
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CS);
    cfg.blockDim = dim3(NTH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_active) {              // query only
        E2E_CHECK_CUDA(cudaOccupancyMaxActiveClusters(max_active, fn, &cfg));
        return 0;
    }
    MParams pc = p;
    if (bwd) E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kb, pc));
    else E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kf, pc));
    ++g_launches;
    return 0;
}

template <int CS>
int run_mc(cudaStream_t st, bool bwd, MParams& p, void* ws, size_t ws_bytes, int force_ns) {
    // NS = 1 (one slice per cluster) when every cluster is co-resident, else two interleaved slices
    static int max_active[2] = {-1, -1};
    if (max_active[bwd] < 0) {
        int rc = launch_mc<CS, 1>(st, bwd, p, 1, &max_active[bwd]);
        if (rc) return rc;
    }
    if (max_active[bwd] < 1) return -1;
    const int ngroups = p.ndir * p.nslices;
    int ns = (ngroups <= max_active[bwd] || p.nslices < 2) ? 1 : 2;
    if (force_ns == 1 || force_ns == 2) ns = force_ns;
    const int nclusters = p.ndir * cdiv(p.nslices, ns);
    if (!bwd) {
        const size_t need = sizeof(float) * (size_t)ns * 2 * nclusters * CS * TILE;
        if (need > ws_bytes) return -1;
        p.xg = (float*)ws;
    }
    return ns == 1 ? launch_mc<CS, 1>(st, bwd, p, nclusters, nullptr) : launch_mc<CS, 2>(st, bwd, p, nclusters, nullptr);
}

}  // namespace

extern long long* g_rec_dbg;
int g_rec_mc_ns = 0;       // 0 = automatic, 1 / 2 = force the number of interleaved slices (tests)

// returns 0 = launched, -1 = not eligible (caller uses the other kernels), >0 = error
int lstm_rec_mc(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int ndir, long long sb, long long stt,
                float* G, float* Hout, float* Cst, const float* Wh, const float* dOut, const int* lens, void* ws,
                size_t ws_bytes) {
    if (H != 128 && H != 256) return -1;
    if (B <= 0 || T <= 0) return 0;
    MParams p;
    p.G = G; p.Hout = Hout; p.Cst = Cst; p.Wh = Wh; p.dOut = dOut; p.lens = lens; p.xg = nullptr;
    p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.ndir = ndir; p.nslices = cdiv(B, R); p.sb = sb; p.st = stt;
    p.dbg = g_rec_dbg;
    p.carry_c = 0;
    if (H == 128) return run_mc<8>(st, bwd, p, ws, ws_bytes, g_rec_mc_ns);
    return run_mc<16>(st, bwd, p, ws, ws_bytes, g_rec_mc_ns);
}

}  // namespace e2e
