// Cluster-resident LSTM recurrence for H = 512 (BASELINE configs[4], the wide encoder): the warp-specialised scheme of
// lstm_rec_ws.cu with 32 hidden units (128 gate columns) per CTA, 16 CTAs per cluster, one 16-row batch slice per
// cluster.
//
// What changes against H = 256 is capacity.  A CTA's slice of W_hh is 512 x 128 values; split into fp16 hi and scaled
// lo' planes (rec_frag.cuh) that is 256 KB -- the whole register file.  So the hi plane stays in the MMA warps'
// registers as fragments (128 registers per thread) and the lo' plane lives in shared memory, also in fragment order
// (128 KB, one conflict-free LDS.128 per two n-tiles); the remaining ~95 KB of shared memory hold the exchange tiles.
// Per k16 block and n-tile the same three m16n8k16 f16 MMAs: hi.hi into the main accumulator, lo'.hi + hi.lo' into the
// cross accumulator (x 2^-11 at the end).
//   forward : h_t all-gather exactly as in lstm_rec_ws.cu -- every CTA publishes its [16 rows][32 units] tile pre-split
//             and packed in fragment order (2 KB) to L2 and ONE multicast bulk copy delivers it to the 16 CTAs.
//   backward: the products are K-sliced (a CTA owns the gate columns of its own units), so the partial d h tiles are
//             reduce-scattered.  There is no room for a staging copy in shared memory: an MMA warp writes the partial
//             tile of a destination CTA to L2 and delivers it with a single-destination multicast bulk copy into that
//             CTA's receive slot.  dz tiles are scaled per row by a power of two before the fp16 split (gradients span
//             many orders of magnitude across utterances) and the partial rows scaled back before they are sent.
// 8 MMA warps (setmaxnreg 200/208) + 8 epilogue warps (48/56 registers; one (row, unit) pair per thread and row half).
#include "rec_frag.cuh"

namespace e2e {

namespace {

constexpr int U5 = 32;                 // hidden units per CTA
constexpr int CS5 = 16;                // CTAs per cluster
constexpr int H5 = CS5 * U5;
constexpr int EW5 = 8;                 // epilogue warps
constexpr int NT5 = 256 + 32 * EW5;
constexpr int HT = R * U5;             // 32-bit words of an exchanged [16 rows][32 units] tile (2 KB): forward = fp16 hi | lo'
                                       // fragments of two k16 blocks, backward = fp32 partial sums
constexpr int ZST5 = 36;               // float4 slots per row of the z hand-off tile (32 used; 36 = conflict-free)
constexpr int WL_WORDS = 8 * 16 * 2 * 32 * 4;   // lo' plane of the CTA's W slice in fragment order (128 KB)

__device__ __forceinline__ void mbar_arrive5(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(NT5, 1) rec_fwd_h512_kernel(MParams p) {
    extern __shared__ __align__(128) float smem[];
    uint4* h_s = reinterpret_cast<uint4*>(smem);                          // [2][CS5][HT words]
    uint4* wl_s = h_s + 2 * CS5 * HT / 4;                                // [8 warps][16 k16][2 n-tile pairs][32 lanes]
    float4* zbuf = reinterpret_cast<float4*>(wl_s + WL_WORDS / 4);       // [2 k-halves][16 rows][ZST5]
    __shared__ __align__(8) uint64_t full[2];                            // h_{t-1} of all CTAs has arrived (tx bytes)
    __shared__ __align__(8) uint64_t zfull;                              // the 8 MMA warps have written their partial z
    __shared__ unsigned pubcnt;                                          // epilogue warps that have written their part of the tile

    const int ndir = p.ndir, T = p.T;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS5;
    const int dir = cl % ndir;
    const int slice = cl / ndir;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&zfull, 8);
        pubcnt = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&full[0], CS5 * HT * 4);
        mbar_expect_tx(&full[1], CS5 * HT * 4);
    }
    __syncthreads();
    cluster_sync_all();

    if (w < 8) {
        // =========================================================== MMA warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int g = lane / 4, tq = lane % 4, ng = w % 4, kh = w / 4;
        // warp (kh, ng): k16 blocks [16 kh, 16 kh + 16) (source CTAs [8 kh, 8 kh + 8)), units 8 ng .. 8 ng + 7 of the CTA =
        // n-tiles nt = 2 uq + gp: unit quad uq, gates (2 gp, 2 gp + 1); column g of a tile = unit 4 uq + g / 2, gate 2 gp + g % 2
        uint32_t bh[16][4][2];
        uint4* wl = wl_s + (size_t)w * 16 * 2 * 32 + lane;
        {
            const float* Wg = p.Wh + (size_t)dir * H5 * H5 * 4;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                uint32_t lo_frag[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int ncol_unit = rank * U5 + 8 * ng + 4 * (nt >> 1) + (g >> 1);
                    const int gate = 2 * (nt & 1) + (g & 1);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        unsigned short hi[2], lo[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 16 * (16 * kh + q) + 8 * j + tq + 4 * e;
                            split_f16(Wg[((size_t)k * H5 + ncol_unit) * 4 + gate], hi[e], lo[e]);
                        }
                        bh[q][nt][j] = pack_u16(hi[0], hi[1]);
                        lo_frag[nt][j] = pack_u16(lo[0], lo[1]);
                    }
                }
                wl[(q * 2 + 0) * 32] = make_uint4(lo_frag[0][0], lo_frag[0][1], lo_frag[1][0], lo_frag[1][1]);
                wl[(q * 2 + 1) * 32] = make_uint4(lo_frag[2][0], lo_frag[2][1], lo_frag[3][0], lo_frag[3][1]);
            }
        }
        __syncwarp();
        uint32_t phase = 0;                                      // bit buf
        for (int s = 1; s < T; ++s) {
            const int buf = s & 1;
            float acc[4][4], accx[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[nt][i] = 0.f; accx[nt][i] = 0.f; }
            mbar_wait(&full[buf], (phase >> buf) & 1u);
            phase ^= 1u << buf;
            if (tid == 0) mbar_expect_tx(&full[buf], CS5 * HT * 4);     // arm this buffer's next phase
            // source CTA 8 kh + q / 2 holds k16 blocks (q % 2): [hi fragments 512 B | lo' fragments 512 B] each
            const uint4* hb = h_s + (size_t)(buf * CS5 + 8 * kh) * (HT / 4) + lane;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const uint4 ah = hb[q * 64], al = hb[q * 64 + 32];
                const uint4 w0 = wl[(q * 2 + 0) * 32], w1 = wl[(q * 2 + 1) * 32];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_f16(acc[nt], ah, bh[q][nt][0], bh[q][nt][1]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_f16(accx[nt], al, bh[q][nt][0], bh[q][nt][1]);
                mma_f16(accx[0], ah, w0.x, w0.y);
                mma_f16(accx[1], ah, w0.z, w0.w);
                mma_f16(accx[2], ah, w1.x, w1.y);
                mma_f16(accx[3], ah, w1.z, w1.w);
            }
            // partial pre-activations of units 8 ng + 4 uq + tq, rows g and g + 8 (4 gates each) -> epilogue warps
            float4* zb = zbuf + (size_t)kh * 16 * ZST5 + 8 * ng + tq;
#pragma unroll
            for (int uq = 0; uq < 2; ++uq) {
                zb[g * ZST5 + 4 * uq] = make_float4(fmaf(accx[2 * uq][0], kF16LoInv, acc[2 * uq][0]), fmaf(accx[2 * uq][1], kF16LoInv, acc[2 * uq][1]),
                                                   fmaf(accx[2 * uq + 1][0], kF16LoInv, acc[2 * uq + 1][0]), fmaf(accx[2 * uq + 1][1], kF16LoInv, acc[2 * uq + 1][1]));
                zb[(g + 8) * ZST5 + 4 * uq] = make_float4(fmaf(accx[2 * uq][2], kF16LoInv, acc[2 * uq][2]), fmaf(accx[2 * uq][3], kF16LoInv, acc[2 * uq][3]),
                                                         fmaf(accx[2 * uq + 1][2], kF16LoInv, acc[2 * uq + 1][2]), fmaf(accx[2 * uq + 1][3], kF16LoInv, acc[2 * uq + 1][3]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive5(&zfull);
        }
    } else {
        // =========================================================== epilogue warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int e = tid - 256;
        const int ul = e % U5, r0 = e / U5;                      // rows r0 and r0 + 8
        const int unit = rank * U5 + ul;
        const long long tstep = (dir == 0 ? 1 : -1) * p.st * ndir * H5;
        int plen[2];
        float c_reg[2], h_reg[2];
        float4 gxn[2];
        long long idx[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pb = slice * R + r0 + 8 * j;
            plen[j] = pb < p.B ? p.lens[pb] : 0;
            c_reg[j] = 0.f;
            h_reg[j] = 0.f;
            const int t0 = dir == 0 ? 0 : T - 1;
            idx[j] = (((long long)pb * p.sb + (long long)t0 * p.st) * ndir + dir) * H5 + unit;
            gxn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t0 < plen[j]) gxn[j] = reinterpret_cast<const float4*>(p.G)[idx[j]];
        }
        uint32_t zph = 0;
        for (int s = 0; s < T; ++s) {
            const int buf = s & 1;
            const int t = dir == 0 ? s : T - 1 - s;
            const int tn = dir == 0 ? s + 1 : T - 2 - s;
            float4 z[2], gxc[2];
            long long ixc[2];
            // this step's x-projection (loaded a step ago) and the next step's prefetch, issued BEFORE the wait
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                gxc[j] = gxn[j];
                ixc[j] = idx[j];
                idx[j] = ixc[j] + tstep;
                if (s + 1 < T && tn < plen[j]) gxn[j] = reinterpret_cast<const float4*>(p.G)[ixc[j] + tstep];
            }
            if (s > 0) {
                mbar_wait(&zfull, zph);
                zph ^= 1u;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 za = zbuf[(size_t)(r0 + 8 * j) * ZST5 + ul];
                    const float4 zc = zbuf[(size_t)(16 + r0 + 8 * j) * ZST5 + ul];
                    z[j] = make_float4(za.x + zc.x, za.y + zc.y, za.z + zc.z, za.w + zc.w);
                }
            }
            float4 act[2];
            float cn[2];
            bool active[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                active[j] = t < plen[j];
                act[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                cn[j] = 0.f;
                if (active[j]) {
                    const float4 gx = gxc[j];
                    const float si = sigmoid_fast(z[j].x + gx.x);
                    const float tj = tanh_fast(z[j].y + gx.y);
                    const float sf = sigmoid_fast(z[j].z + gx.z + 1.0f);
                    const float so = sigmoid_fast(z[j].w + gx.w);
                    cn[j] = c_reg[j] * sf + si * tj;
                    h_reg[j] = tanh_fast(cn[j]) * so;
                    c_reg[j] = cn[j];
                    act[j] = make_float4(si, tj, sf, so);
                }
            }
            // publish h_t (state h: carried through for masked rows): tile -> L2 -> multicast to the cluster
            if (s + 1 < T) {
                float* gt = p.xg + ((size_t)buf * gridDim.x + blockIdx.x) * HT;
                // units 16 kb .. 16 kb + 15 of the CTA are one k16 block; 16-bit slot of (row, unit u) in a block's packed
                // fragments = lane (row % 8, u % 4) * 8 + (2 * (u / 8) + row / 8) * 2 + (u / 4) % 2; per block hi plane, then lo'
                unsigned short* gh = reinterpret_cast<unsigned short*>(gt) + (ul >> 4) * 512;
                const int u16 = ul & 15;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int row = r0 + 8 * j;
                    const int hp = ((row & 7) * 4 + (u16 & 3)) * 8 + (2 * (u16 >> 3) + (row >> 3)) * 2 + ((u16 >> 2) & 1);
                    unsigned short hi, lo;
                    split_f16(h_reg[j], hi, lo);
                    gh[hp] = hi;
                    gh[256 + hp] = lo;
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    unsigned old;
                    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(s_u32(&pubcnt)) : "memory");
                    if ((old & (EW5 - 1)) == EW5 - 1) {      // last epilogue warp of this step: the tile is complete
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        bulk_multicast(s_u32(h_s + (size_t)((buf ^ 1) * CS5 + rank) * (HT / 4)), gt, HT * 4, s_u32(&full[buf ^ 1]),
                                       (uint16_t)0xFFFFu);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (active[j]) {
                    reinterpret_cast<float4*>(p.G)[ixc[j]] = act[j];
                    p.Cst[ixc[j]] = cn[j];
                    p.Hout[ixc[j]] = h_reg[j];
                }
            }
        }
    }
    cluster_sync_all();      // nobody exits while a peer's multicast may still target its shared memory
}

// ---------------------------------------------------------------- backward
__global__ void __launch_bounds__(NT5, 1) rec_bwd_h512_kernel(MParams p) {
    extern __shared__ __align__(128) float smem[];
    float* recv = smem;                                                  // [2][CS5][HT] received partial d h tiles
    uint4* dz_s = reinterpret_cast<uint4*>(recv + 2 * CS5 * HT);         // [2][hi: 8 k16 x 32 lanes | lo': same] own dz fragments
    uint4* wl_s = dz_s + 2 * 512;                                        // [8 warps][8 k16][4 n-tile pairs][32 lanes]
    __shared__ __align__(8) uint64_t full[2];                            // the CS5 partial tiles of a step have arrived
    // the epilogue warps have written dz_t: two barriers alternating with the step, so that the barrier of step t cannot
    // complete the phase of t+2 before every MMA warp has tested the phase of t (see lstm_rec_ws.cu)
    __shared__ __align__(8) uint64_t dzready[2];
    __shared__ float rinv_s[2][R];                                       // 1 / (row scale) of the dz tile of a buffer

    const int ndir = p.ndir, T = p.T, Tp = p.Tp;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS5;
    const int dir = cl % ndir;
    const int slice = cl / ndir;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&dzready[0], EW5);
        mbar_init(&dzready[1], EW5);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&full[0], CS5 * HT * 4);
        mbar_expect_tx(&full[1], CS5 * HT * 4);
    }
    __syncthreads();
    cluster_sync_all();

    if (w < 8) {
        // =========================================================== MMA warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int g = lane / 4, tq = lane % 4;
        // resident fragments: B[k][n] = W_hh[hidden unit n][own gate column k]; warp w owns units [64 w, 64 w + 64) =
        // destination CTAs 2 w and 2 w + 1, n-tiles nt = 4 d + m: column g of tile m = unit 16 (m / 2) + 4 (g / 2) + 2 (m % 2) + g % 2
        // of the destination, so the C columns (2 tq, 2 tq + 1) of tiles (2 h, 2 h + 1) are its units 16 h + 4 tq .. + 3
        uint32_t bh[8][8][2];
        uint4* wl = wl_s + (size_t)w * 8 * 4 * 32 + lane;
        {
            const float* Wg = p.Wh + (size_t)dir * H5 * H5 * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t lo_frag[8][2];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int m = nt & 3;
                    const int n_unit = U5 * (2 * w + (nt >> 2)) + 16 * (m >> 1) + 4 * (g >> 1) + 2 * (m & 1) + (g & 1);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        unsigned short hi[2], lo[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kcol = 16 * q + 8 * j + tq + 4 * e;      // own gate column: unit 4 q + 2 j + e, gate tq
                            split_f16(Wg[(size_t)n_unit * H5 * 4 + rank * 4 * U5 + kcol], hi[e], lo[e]);
                        }
                        bh[q][nt][j] = pack_u16(hi[0], hi[1]);
                        lo_frag[nt][j] = pack_u16(lo[0], lo[1]);
                    }
                }
#pragma unroll
                for (int pr = 0; pr < 4; ++pr)
                    wl[(q * 4 + pr) * 32] = make_uint4(lo_frag[2 * pr][0], lo_frag[2 * pr][1], lo_frag[2 * pr + 1][0], lo_frag[2 * pr + 1][1]);
            }
        }
        __syncwarp();
        for (int s = 0; s + 1 < T; ++s) {
            const int buf = s & 1;
            mbar_wait(&dzready[buf], (uint32_t)((s >> 1) & 1));
            const uint4* dzb = dz_s + buf * 512 + lane;
            const float r0s = rinv_s[buf][g], r1s = rinv_s[buf][g + 8];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                float acc[4][4], accx[4][4];
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { acc[m][i] = 0.f; accx[m][i] = 0.f; }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 ah = dzb[q * 32], al = dzb[256 + q * 32];
                    const uint4 w0 = wl[(q * 4 + 2 * d) * 32], w1 = wl[(q * 4 + 2 * d + 1) * 32];
#pragma unroll
                    for (int m = 0; m < 4; ++m) mma_f16(acc[m], ah, bh[q][4 * d + m][0], bh[q][4 * d + m][1]);
#pragma unroll
                    for (int m = 0; m < 4; ++m) mma_f16(accx[m], al, bh[q][4 * d + m][0], bh[q][4 * d + m][1]);
                    mma_f16(accx[0], ah, w0.x, w0.y);
                    mma_f16(accx[1], ah, w0.z, w0.w);
                    mma_f16(accx[2], ah, w1.x, w1.y);
                    mma_f16(accx[3], ah, w1.z, w1.w);
                }
                // partial d h of the destination's 32 units, rows g and g + 8, back in true scale: tile [row][unit] in L2,
                // then one bulk copy into slot `rank` of the destination's receive buffer
                const int dest = 2 * w + d;
                float* gs = p.xg + (((size_t)buf * gridDim.x + blockIdx.x) * CS5 + dest) * HT;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    *reinterpret_cast<float4*>(gs + g * U5 + 16 * h + 4 * tq) =
                        make_float4(fmaf(accx[2 * h][0], kF16LoInv, acc[2 * h][0]) * r0s, fmaf(accx[2 * h][1], kF16LoInv, acc[2 * h][1]) * r0s,
                                    fmaf(accx[2 * h + 1][0], kF16LoInv, acc[2 * h + 1][0]) * r0s, fmaf(accx[2 * h + 1][1], kF16LoInv, acc[2 * h + 1][1]) * r0s);
                    *reinterpret_cast<float4*>(gs + (g + 8) * U5 + 16 * h + 4 * tq) =
                        make_float4(fmaf(accx[2 * h][2], kF16LoInv, acc[2 * h][2]) * r1s, fmaf(accx[2 * h][3], kF16LoInv, acc[2 * h][3]) * r1s,
                                    fmaf(accx[2 * h + 1][2], kF16LoInv, acc[2 * h + 1][2]) * r1s, fmaf(accx[2 * h + 1][3], kF16LoInv, acc[2 * h + 1][3]) * r1s);
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    bulk_multicast(s_u32(recv + (size_t)((buf ^ 1) * CS5 + rank) * HT), gs, HT * 4, s_u32(&full[buf ^ 1]),
                                   (uint16_t)(1u << dest));
                }
            }
        }
    } else {
        // =========================================================== epilogue warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int e = tid - 256;
        const int ul = e % U5, r0 = e / U5;                      // rows r0 and r0 + 8; a warp = the 32 units of one row pair
        const int unit = rank * U5 + ul;
        const long long tstep = (dir == 0 ? -1 : 1) * p.st * ndir * H5;
        int plen[2], pbv[2];
        float dc_reg[2];
        long long idx[2];
        float4 act_n[2];
        float cst_n[2], cprev_n[2], dout_n[2];
        auto prefetch = [&](int j, int t, long long ix) {
            act_n[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            cst_n[j] = 0.f; cprev_n[j] = 0.f; dout_n[j] = 0.f;
            if (t >= 0 && t < plen[j]) {
                const int t_cprev = dir == 0 ? t - 1 : t + 1;
                act_n[j] = reinterpret_cast<const float4*>(p.G)[ix];
                cst_n[j] = p.Cst[ix];
                if (t_cprev >= 0 && t_cprev < plen[j]) cprev_n[j] = p.Cst[ix + tstep];
                dout_n[j] = __ldg(p.dOut + ix);
            }
        };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pb = slice * R + r0 + 8 * j;
            pbv[j] = pb;
            plen[j] = pb < p.B ? p.lens[pb] : 0;
            dc_reg[j] = 0.f;
            const int t0 = dir == 0 ? T - 1 : 0;
            idx[j] = (((long long)pb * p.sb + (long long)t0 * p.st) * ndir + dir) * H5 + unit;
            if (pb < p.B)
                for (int t = T; t < Tp; ++t)
                    reinterpret_cast<float4*>(p.G)[(((size_t)pb * p.sb + (size_t)t * p.st) * ndir + dir) * H5 + unit] =
                        make_float4(0.f, 0.f, 0.f, 0.f);
            prefetch(j, t0, idx[j]);
        }
        uint32_t phase = 0;
        for (int s = 0; s < T; ++s) {
            const int buf = s & 1;
            const int t = dir == 0 ? T - 1 - s : s;
            float dh[2];
            float4 actc[2];
            float cstc[2], cprevc[2], doutc[2];
            long long ixc[2];
            // this step's operands (loaded a step ago) and the next step's prefetch, issued before the wait
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                dh[j] = 0.f;
                actc[j] = act_n[j]; cstc[j] = cst_n[j]; cprevc[j] = cprev_n[j]; doutc[j] = dout_n[j];
                ixc[j] = idx[j];
                idx[j] = ixc[j] + tstep;
                if (s + 1 < T) prefetch(j, dir == 0 ? t - 1 : t + 1, ixc[j] + tstep);
            }
            if (s > 0) {
                mbar_wait(&full[buf], (phase >> buf) & 1u);
                phase ^= 1u << buf;
                if (e == 0) mbar_expect_tx(&full[buf], CS5 * HT * 4);
                const float* rb = recv + (size_t)buf * CS5 * HT + r0 * U5 + ul;
                float pa[2][2];
#pragma unroll
                for (int j = 0; j < 2; ++j) { pa[j][0] = 0.f; pa[j][1] = 0.f; }
#pragma unroll
                for (int src = 0; src < CS5; ++src)
#pragma unroll
                    for (int j = 0; j < 2; ++j) pa[j][src & 1] += rb[src * HT + 8 * U5 * j];
#pragma unroll
                for (int j = 0; j < 2; ++j) dh[j] = pa[j][0] + pa[j][1];
            }
            float4 dz[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                dz[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < plen[j]) {
                    const float4 act = actc[j];
                    const float dhj = dh[j] + doutc[j];
                    const float si = act.x, tj = act.y, sf = act.z, so = act.w;
                    const float tc = tanh_fast(cstc[j]);
                    const float dct = dc_reg[j] + dhj * so * (1.f - tc * tc);
                    dz[j].x = dct * tj * si * (1.f - si);
                    dz[j].y = dct * si * (1.f - tj * tj);
                    dz[j].z = dct * cprevc[j] * sf * (1.f - sf);
                    dz[j].w = dhj * tc * so * (1.f - so);
                    dc_reg[j] = dct * sf;
                }
            }
            if (s + 1 < T) {
                // fp16 fragments of the CTA's dz tile, every row scaled by its own power of two (amax over the row's 128
                // gate columns = the warp): k16 block q = ul / 4 holds units 4 q .. 4 q + 3; 16-bit slot of (row, unit,
                // gate) = word [q][lane = (row % 8) * 4 + gate][reg = 2 * ((ul % 4) / 2) + row / 8], half ul % 2
                unsigned short* dqh = reinterpret_cast<unsigned short*>(dz_s + buf * 512);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int row = r0 + 8 * j;
                    float amax = fmaxf(fmaxf(fabsf(dz[j].x), fabsf(dz[j].y)), fmaxf(fabsf(dz[j].z), fabsf(dz[j].w)));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                    float sc = 1.0f;
                    if (amax >= 1.17549435e-38f) {
                        const int ex = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
                        sc = __uint_as_float((uint32_t)(min(max(13 - ex, -126), 126) + 127) << 23);
                    }
                    if (ul == 0) rinv_s[buf][row] = 1.0f / sc;
                    const float v[4] = {dz[j].x * sc, dz[j].y * sc, dz[j].z * sc, dz[j].w * sc};
                    const int wbase = (ul >> 2) * 128 + (row & 7) * 16 + 2 * ((ul & 3) >> 1) + (row >> 3);
#pragma unroll
                    for (int gt = 0; gt < 4; ++gt) {
                        unsigned short hi, lo;
                        split_f16(v[gt], hi, lo);
                        dqh[(wbase + gt * 4) * 2 + (ul & 1)] = hi;
                        dqh[(1024 + wbase + gt * 4) * 2 + (ul & 1)] = lo;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive5(&dzready[buf]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (pbv[j] < p.B) reinterpret_cast<float4*>(p.G)[ixc[j]] = dz[j];
            }
        }
    }
    cluster_sync_all();
}

size_t fwd_h512_smem() { return (size_t)2 * CS5 * HT * 4 + (size_t)WL_WORDS * 4 + (size_t)2 * 16 * ZST5 * 16; }
size_t bwd_h512_smem() { return (size_t)2 * CS5 * HT * 4 + (size_t)2 * 512 * 16 + (size_t)WL_WORDS * 4; }

}  // namespace

// workspace bytes the H = 512 kernels need for B rows and ndir directions (exchange tiles in L2)
size_t lstm_rec_h512_workspace(int B, int ndir, bool bwd) {
    const size_t grid = (size_t)ndir * cdiv(B, R) * CS5;
    return (size_t)2 * grid * HT * 4 * (bwd ? CS5 : 1);
}

// returns 0 = launched, -1 = not eligible (caller uses the other kernels), >0 = error
int lstm_rec_h512(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int ndir, long long sb, long long stt, float* G,
                  float* Hout, float* Cst, const float* Wh, const float* dOut, const int* lens, void* ws, size_t ws_bytes) {
    if (H != H5) return -1;
    if (B <= 0 || T <= 0) return 0;
    if (lstm_rec_h512_workspace(B, ndir, bwd) > ws_bytes) return -1;
    const void* fn = bwd ? (const void*)rec_bwd_h512_kernel : (const void*)rec_fwd_h512_kernel;
    const size_t smem = bwd ? bwd_h512_smem() : fwd_h512_smem();
    static int ready[2] = {0, 0};        // 1 = attributes set and a cluster fits, -1 = does not fit
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NT5);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS5;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ready[bwd] == 0) {
        E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3(CS5);
        int max_active = 0;
        E2E_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_active, fn, &cfg));
        ready[bwd] = max_active >= 1 ? 1 : -1;
    }
    if (ready[bwd] < 0) return -1;
    MParams p;
    p.carry_c = 0;
    p.G = G; p.Hout = Hout; p.Cst = Cst; p.Wh = Wh; p.dOut = dOut; p.lens = lens; p.xg = (float*)ws;
    p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.ndir = ndir; p.nslices = cdiv(B, R); p.sb = sb; p.st = stt;
    p.dbg = nullptr;
    cfg.gridDim = dim3(ndir * p.nslices * CS5);
    if (bwd) E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rec_bwd_h512_kernel, p));
    else E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rec_fwd_h512_kernel, p));
    ++g_launches;
    return 0;
}

}  // namespace e2e
