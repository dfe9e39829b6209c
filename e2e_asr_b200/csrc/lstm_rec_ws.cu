// Warp-specialised register-resident LSTM recurrence (H in {128, 256}).  One 16-row batch slice of one direction is
// owned by a cluster of CS = H / 16 CTAs; CTA j owns hidden units [16 j, 16 j + 16) = 64 gate columns of W_hh, held as
// MMA fragments in registers for the whole sequence; the h_t / d h tiles are exchanged in fragment order (multicast
// bulk copy forward, bulk DSMEM reduce-scatter backward).  The k-loop and the pointwise / publish work of a step run on
// DIFFERENT warps of the CTA, so they overlap instead of adding up.
//
// Measured on the non-specialised predecessor (cfg-2, two interleaved slices per cluster, cycles per slice-step):
// k-loop 1330, everything else (k-half combine, gates, publish, global stores) 1070 -- executed back to back by the same
// 8 warps, 4800 per step for the two slices, tensor pipe 42 % busy.  Here
//   warps 0..7  (2 warpgroups, setmaxnreg 216) hold W_hh as fragments and only run k-loops: for every slice
//               wait for h_{t-1} / dz_t, 64 MMAs per warp, hand the partial tile to the epilogue warps;
//   warps 8..11 (1 warpgroup, setmaxnreg 72) own the cell state: gates, h_t publish (multicast) resp. the
//               reduce-scatter receive + LSTM backward, global loads/stores.
// With two slices in flight the MMA warps always have the other slice's k-loop to run while a slice's
// epilogue + exchange completes: a step of both slices costs ~2 k-loops.
// Hand-offs are mbarriers inside the CTA (no CTA-wide barrier anywhere in the step loop).
#include "rec_frag.cuh"

namespace e2e {

namespace {

constexpr int NTW = 384;           // 8 MMA warps + 4 epilogue warps
constexpr int NEPI = 128;          // epilogue threads: thread e owns unit e % 16, rows e / 16 and e / 16 + 8
constexpr int ZST = 20;            // float4 slots per row of the z hand-off tile (16 used; 20 = conflict-free)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- forward
// F16: the fp16 split scheme of rec_frag.cuh (3 MMAs per k16 and n-tile, h_t exchanged pre-split) instead of the
// tf32 + bf16 scheme (4 MMAs, kept as test mode 7)
template <int CS, int NS, int EW, bool F16>
__global__ void __launch_bounds__(256 + 32 * EW, 1) rec_fwd_ws_kernel(MParams p) {
    constexpr int NR = 8 / EW;                                   // rows per epilogue thread (EW = 4: r0 and r0 + 8)
    constexpr int H = CS * UPC;
    constexpr int KT = CS;                                       // k8-tiles per k-half
    extern __shared__ __align__(128) float smem[];
    float* h_s = smem;                                           // [NS][2][CS][TILE]
    float4* zbuf = reinterpret_cast<float4*>(h_s + NS * 2 * CS * TILE);   // [NS][2 k-halves][16 rows][ZST]
    __shared__ __align__(8) uint64_t full[NS][2];                // h_{t-1} of all CTAs has arrived (tx bytes)
    __shared__ __align__(8) uint64_t zfull[NS];                  // the 8 MMA warps have written their partial z
    __shared__ unsigned pubcnt[NS];                              // epilogue warps that have written their part of the tile

    const int ndir = p.ndir, T = p.T;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS;
    const int dir = cl % ndir;
    const int slice0 = (cl / ndir) * NS;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32;

    if (tid == 0) {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            mbar_init(&full[sl][0], 1);
            mbar_init(&full[sl][1], 1);
            mbar_init(&zfull[sl], 8);
            pubcnt[sl] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_expect_tx(&full[sl][0], CS * TILE * 4); mbar_expect_tx(&full[sl][1], CS * TILE * 4); }
    }
    bool live[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) live[sl] = slice0 + sl < p.nslices;
    __syncthreads();
    cluster_sync_all();

    if (w < 8) {
        // =========================================================== MMA warps
        if (EW == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int g = lane / 4, tq = lane % 4, ng = w % 4, kh = w / 4;
        uint32_t bh[KT][2][2], bl[KT][2][2];
        if (!F16) {
            const float* Wg = p.Wh + (size_t)dir * H * H * 4;
            const int ncol_unit = rank * UPC + 4 * ng + (g >> 1);
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = 8 * (kh * KT + kt) + tq + 4 * e;
                        bh[kt][nt][e] = cvt_tf32(Wg[((size_t)k * H + ncol_unit) * 4 + 2 * nt + (g & 1)]);
                    }
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
        if (F16) {
            // the same registers hold the fp16 fragments: bh[q][nt] := packed fp16 W of pair q, bl[q][nt] := packed
            // fp16 of (W - fp16 W) 2^11 (only the first KT / 2 entries of each array are used)
            const float* Wg = p.Wh + (size_t)dir * H * H * 4;
            const int ncol_unit = rank * UPC + 4 * ng + (g >> 1);
#pragma unroll
            for (int q = 0; q < KT / 2; ++q)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        unsigned short hi[2], lo[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 8 * (kh * KT + 2 * q + j) + tq + 4 * e;
                            split_f16(Wg[((size_t)k * H + ncol_unit) * 4 + 2 * nt + (g & 1)], hi[e], lo[e]);
                        }
                        bh[q][nt][j] = pack_u16(hi[0], hi[1]);
                        bl[q][nt][j] = pack_u16(lo[0], lo[1]);
                    }
        }
        uint32_t phase = 0;                                      // bit (2 sl + buf)
        const bool rec = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
        for (int s = 1; s < T; ++s) {
            const int buf = s & 1;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                if (!live[sl]) continue;
                if (rec) p.dbg[(s * NS + sl) * 8 + 0] = clock64();
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
                if (F16) {
                    // source CTA kh*8 + q contributes one k16 pair: [hi fragments 512 B | lo' fragments 512 B]
                    uint4 a0 = *reinterpret_cast<const uint4*>(hb), a1 = *reinterpret_cast<const uint4*>(hb + 128);
#pragma unroll
                    for (int q = 0; q < KT / 2; ++q) {
                        const int qn = q + 1 < KT / 2 ? q + 1 : q;
                        const uint4 n0 = *reinterpret_cast<const uint4*>(hb + (2 * qn) * 128);
                        const uint4 n1 = *reinterpret_cast<const uint4*>(hb + (2 * qn + 1) * 128);
                        k16_mma_f16<2>(acc, accx, a0, a1, bh[q], bl[q]);
                        a0 = n0; a1 = n1;
                    }
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int q = 0; q < 4; ++q) accx[nt][q] *= kF16LoInv;
                } else {
                    float4 a0 = *reinterpret_cast<const float4*>(hb), a1 = *reinterpret_cast<const float4*>(hb + 128);
#pragma unroll
                    for (int q = 0; q < KT / 2; ++q) {
                        const int qn = q + 1 < KT / 2 ? q + 1 : q;
                        const float4 n0 = *reinterpret_cast<const float4*>(hb + (2 * qn) * 128);
                        const float4 n1 = *reinterpret_cast<const float4*>(hb + (2 * qn + 1) * 128);
                        k16_mma<2>(acc, accx, a0, a1, bh[2 * q], bh[2 * q + 1], bl[2 * q], bl[2 * q + 1]);
                        a0 = n0; a1 = n1;
                    }
                }
                // partial pre-activations of unit 4ng+tq, rows g and g+8 (4 gates each) -> epilogue warps
                float4* zb = zbuf + (size_t)(sl * 2 + kh) * 16 * ZST + 4 * ng + tq;
                zb[g * ZST] = make_float4(acc[0][0] + accx[0][0], acc[0][1] + accx[0][1], acc[1][0] + accx[1][0], acc[1][1] + accx[1][1]);
                zb[(g + 8) * ZST] = make_float4(acc[0][2] + accx[0][2], acc[0][3] + accx[0][3], acc[1][2] + accx[1][2], acc[1][3] + accx[1][3]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&zfull[sl]);
                if (rec) p.dbg[(s * NS + sl) * 8 + 2] = clock64();
            }
        }
    } else {
        // =========================================================== epilogue warps
        if (EW == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int e = tid - 8 * 32;
        const int ul = e % UPC, r0 = e / UPC;                    // rows r0 (+ 8 j)
        const int unit = rank * UPC + ul;
        const long long tstep = (dir == 0 ? 1 : -1) * p.st * ndir * H;
        // fragment-order position of (row r0 + 8 j, unit ul) in the exchanged tile: the two rows are adjacent floats
        const int fpos = (ul >> 3) * 128 + ((r0 & 7) * 4 + (ul & 3)) * 4 + 2 * ((ul >> 2) & 1) + (r0 >> 3);
        int plen[NS][NR];
        float c_reg[NS][NR], h_reg[NS][NR];
        float4 gxn[NS][NR];
        long long idx[NS][NR];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const int pb = (slice0 + sl) * R + r0 + 8 * j;
                plen[sl][j] = (live[sl] && pb < p.B) ? p.lens[pb] : 0;
                c_reg[sl][j] = 0.f;
                h_reg[sl][j] = 0.f;
                const int t0 = dir == 0 ? 0 : T - 1;
                idx[sl][j] = (((long long)pb * p.sb + (long long)t0 * p.st) * ndir + dir) * H + unit;
                gxn[sl][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t0 < plen[sl][j]) gxn[sl][j] = reinterpret_cast<const float4*>(p.G)[idx[sl][j]];
                // a continued sequence: c_{-1} from the row before the window (h_{-1} . Wh is already part of G_0)
                if (p.carry_c && plen[sl][j] > 0) c_reg[sl][j] = p.Cst[idx[sl][j] - tstep];
            }
        uint32_t zph = 0;
        const bool rec = p.dbg != nullptr && blockIdx.x == 0 && e == 0;
        for (int s = 0; s < T; ++s) {
            const int buf = s & 1;
            const int t = dir == 0 ? s : T - 1 - s;
            const int tn = dir == 0 ? s + 1 : T - 2 - s;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                if (!live[sl]) continue;
                float4 z[NR];
#pragma unroll
                for (int j = 0; j < NR; ++j) z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                // this step's x-projection (loaded a step ago) and the next step's prefetch, issued BEFORE the wait so
                // that the release fence of the publish below never waits for a load in flight
                float4 gxc[NR];
                long long ixc[NR];
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    gxc[j] = gxn[sl][j];
                    ixc[j] = idx[sl][j];
                    idx[sl][j] = ixc[j] + tstep;
                    if (s + 1 < T && tn < plen[sl][j]) gxn[sl][j] = reinterpret_cast<const float4*>(p.G)[ixc[j] + tstep];
                }
                if (rec) p.dbg[(s * NS + sl) * 8 + 3] = clock64();
                if (s > 0) {
                    mbar_wait(&zfull[sl], (zph >> sl) & 1u);
                    zph ^= 1u << sl;
                    if (rec) p.dbg[(s * NS + sl) * 8 + 4] = clock64();
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        const float4 za = zbuf[((size_t)(sl * 2 + 0) * 16 + r0 + 8 * j) * ZST + ul];
                        const float4 zc = zbuf[((size_t)(sl * 2 + 1) * 16 + r0 + 8 * j) * ZST + ul];
                        z[j] = make_float4(za.x + zc.x, za.y + zc.y, za.z + zc.z, za.w + zc.w);
                    }
                }
                float4 act[NR];
                float cn[NR];
                bool active[NR];
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    active[j] = t < plen[sl][j];
                    act[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    cn[j] = 0.f;
                    if (active[j]) {
                        const float4 gx = gxc[j];
                        const float si = sigmoid_fast(z[j].x + gx.x);
                        const float tj = tanh_fast(z[j].y + gx.y);
                        const float sf = sigmoid_fast(z[j].z + gx.z + 1.0f);
                        const float so = sigmoid_fast(z[j].w + gx.w);
                        cn[j] = c_reg[sl][j] * sf + si * tj;
                        h_reg[sl][j] = tanh_fast(cn[j]) * so;
                        c_reg[sl][j] = cn[j];
                        act[j] = make_float4(si, tj, sf, so);
                    }
                }
                if (rec) p.dbg[(s * NS + sl) * 8 + 5] = clock64();
                // publish h_t (state h: carried through for masked rows): tile -> L2 -> multicast to the cluster
                if (s + 1 < T) {
                    float* gt = p.xg + ((size_t)(sl * 2 + buf) * gridDim.x + blockIdx.x) * TILE;
                    if (F16) {
                        // the CTA's 16 units are one k16 pair: 16-bit slot of (row, unit) in the packed fragments =
                        // lane (row % 8, unit % 4) * 8 + (2 * (unit / 8) + row / 8) * 2 + (unit / 4) % 2; hi' plane, then lo'
                        unsigned short* gh = reinterpret_cast<unsigned short*>(gt);
#pragma unroll
                        for (int j = 0; j < NR; ++j) {
                            const int row = r0 + 8 * j;
                            const int hp = ((row & 7) * 4 + (ul & 3)) * 8 + (2 * (ul >> 3) + (row >> 3)) * 2 + ((ul >> 2) & 1);
                            unsigned short hi, lo;
                            split_f16(h_reg[sl][j], hi, lo);
                            gh[hp] = hi;
                            gh[256 + hp] = lo;
                        }
                    } else if (NR == 2) *reinterpret_cast<float2*>(gt + fpos) = make_float2(h_reg[sl][0], h_reg[sl][NR - 1]);
                    else gt[fpos] = h_reg[sl][0];
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        unsigned old;
                        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(s_u32(&pubcnt[sl])) : "memory");
                        if ((old & (EW - 1)) == EW - 1) {      // last epilogue warp of this slice-step: the tile is complete
                            asm volatile("fence.proxy.async.global;" ::: "memory");
                            bulk_multicast(s_u32(h_s + ((size_t)(sl * 2 + (buf ^ 1)) * CS + rank) * TILE), gt, TILE * 4,
                                           s_u32(&full[sl][buf ^ 1]), (uint16_t)((1u << CS) - 1u));
                        }
                    }
                }
                if (rec) p.dbg[(s * NS + sl) * 8 + 6] = clock64();
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    if (active[j]) {
                        reinterpret_cast<float4*>(p.G)[ixc[j]] = act[j];
                        p.Cst[ixc[j]] = cn[j];
                        p.Hout[ixc[j]] = h_reg[sl][j];
                    }
                }
                if (rec) p.dbg[(s * NS + sl) * 8 + 7] = clock64();
            }
        }
    }
    cluster_sync_all();      // nobody exits while a peer's multicast may still target its shared memory
}

// ---------------------------------------------------------------- backward
// F16: the fp16 split scheme for the backward products too.  dz spans many orders of magnitude ACROSS utterances (and
// fp16 has 5 exponent bits), so every row of the CTA's dz tile is scaled by its own power of two before the split
// (amax over the CTA's 64 gate columns of the row: a 16-lane shuffle reduction by the epilogue warps) and the MMA warps
// scale their partial d h rows back before the reduce-scatter -- the partial sums of the 16 CTAs stay in true scale.
template <int CS, int NS, int EW, bool F16>
__global__ void __launch_bounds__(256 + 32 * EW, 1) rec_bwd_ws_kernel(MParams p) {
    constexpr int NR = 8 / EW;                                   // rows per epilogue thread (EW = 4: r0 and r0 + 8)
    constexpr int H = CS * UPC;
    constexpr int NTL = CS / 4;                                  // n-tiles (8 hidden units) per MMA warp
    constexpr int ND = CS / 8;                                   // destination CTAs per MMA warp
    extern __shared__ __align__(128) float smem[];
    float* recv = smem;                                          // [NS][2][CS][TILE] received partial tiles
    float* stage = recv + NS * 2 * CS * TILE;                    // [NS][2][CS][TILE] partial tiles to send
    float* dz_s = stage + NS * 2 * CS * TILE;                    // [NS][2][8 k-tiles][32 chunks][4] own dz, fragment order
    __shared__ __align__(8) uint64_t full[NS][2];                // the CS partial tiles of a step have arrived
    // the epilogue warps have written dz_t: TWO barriers per slice, alternating with the step.  An epilogue warp's step
    // t+1 needs just ONE partial tile from every CTA (the MMA warp that owns this CTA as a destination), so with a single
    // barrier it could complete the phase of t+1 while a slow MMA warp of its own CTA (weight set-up at kernel start) has
    // not yet tested the phase of t -- the warp would then wait for a completion that depends on itself.  Step t+2 needs
    // every warp's step-t tiles, so a barrier that is signalled only every other step can never run a phase ahead.
    __shared__ __align__(8) uint64_t dzready[NS][2];
    __shared__ float rinv_s[NS][2][R];                           // F16: 1 / (row scale) of the dz tile of (slice, buffer)

    const int ndir = p.ndir, T = p.T, Tp = p.Tp;
    const uint32_t rank = cluster_rank();
    const int cl = blockIdx.x / CS;
    const int dir = cl % ndir;
    const int slice0 = (cl / ndir) * NS;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32;

    if (tid == 0) {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            mbar_init(&full[sl][0], 1);
            mbar_init(&full[sl][1], 1);
            mbar_init(&dzready[sl][0], EW);
            mbar_init(&dzready[sl][1], EW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) { mbar_expect_tx(&full[sl][0], CS * TILE * 4); mbar_expect_tx(&full[sl][1], CS * TILE * 4); }
    }
    bool live[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) live[sl] = slice0 + sl < p.nslices;
    __syncthreads();
    cluster_sync_all();

    if (w < 8) {
        // =========================================================== MMA warps
        if (EW == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int g = lane / 4, tq = lane % 4;
        // resident fragments: B[k][n] = W_hh[hidden unit n][own gate column k]
        uint32_t bh[8][NTL][2], bl[8][NTL][2];
        if (F16) {
            // bh[q][nt] := packed fp16 W of k16 pair q, bl[q][nt] := packed fp16 of (W - fp16 W) 2^11 (q < 4)
            const float* Wg = p.Wh + (size_t)dir * H * H * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int n_unit = 16 * (w * ND + nt / 2) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        unsigned short hi[2], lo[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kcol = 8 * (2 * q + j) + tq + 4 * e;
                            split_f16(Wg[(size_t)n_unit * H * 4 + rank * 4 * UPC + kcol], hi[e], lo[e]);
                        }
                        bh[q][nt][j] = pack_u16(hi[0], hi[1]);
                        bl[q][nt][j] = pack_u16(lo[0], lo[1]);
                    }
                }
        } else {
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
                        bh[kt][nt][e] = cvt_tf32(Wg[(size_t)n_unit * H * 4 + rank * 4 * UPC + kcol]);
                    }
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
        for (int s = 0; s + 1 < T; ++s) {
            const int buf = s & 1;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                if (!live[sl]) continue;
                float acc[NTL][4], accx[NTL][4];
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { acc[nt][q] = 0.f; accx[nt][q] = 0.f; }
                mbar_wait(&dzready[sl][buf], (uint32_t)((s >> 1) & 1));
                const float* dzs = dz_s + (size_t)(sl * 2 + buf) * 8 * 128;
                if (F16) {
                    // [hi fragments of the 4 k16 pairs: 4 x 512 B | lo' fragments], already in register order
                    uint4 a0 = *reinterpret_cast<const uint4*>(dzs + lane * 4);
                    uint4 a1 = *reinterpret_cast<const uint4*>(dzs + 512 + lane * 4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int qn = q + 1 < 4 ? q + 1 : q;
                        const uint4 n0 = *reinterpret_cast<const uint4*>(dzs + qn * 128 + lane * 4);
                        const uint4 n1 = *reinterpret_cast<const uint4*>(dzs + 512 + qn * 128 + lane * 4);
                        k16_mma_f16<NTL>(acc, accx, a0, a1, bh[q], bl[q]);
                        a0 = n0; a1 = n1;
                    }
                    const float r0s = rinv_s[sl][buf][g], r1s = rinv_s[sl][buf][g + 8];
#pragma unroll
                    for (int nt = 0; nt < NTL; ++nt) {
                        acc[nt][0] = fmaf(accx[nt][0], kF16LoInv, acc[nt][0]) * r0s;
                        acc[nt][1] = fmaf(accx[nt][1], kF16LoInv, acc[nt][1]) * r0s;
                        acc[nt][2] = fmaf(accx[nt][2], kF16LoInv, acc[nt][2]) * r1s;
                        acc[nt][3] = fmaf(accx[nt][3], kF16LoInv, acc[nt][3]) * r1s;
                        accx[nt][0] = accx[nt][1] = accx[nt][2] = accx[nt][3] = 0.f;
                    }
                } else {
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
                }
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
                __syncwarp();
                // this warp wrote whole tiles: tile `dest` of my partials goes to slot `rank` of peer `dest`
                if (lane < ND) {
                    const int dest = w * ND + lane;
                    dsmem_bulk_copy(mapa(s_u32(recv + ((size_t)(sl * 2 + (buf ^ 1)) * CS + rank) * TILE), dest),
                                    mapa(s_u32(&full[sl][buf ^ 1]), dest),
                                    stage + ((size_t)(sl * 2 + buf) * CS + dest) * TILE, TILE * 4);
                }
            }
        }
    } else {
        // =========================================================== epilogue warps
        if (EW == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int e = tid - 8 * 32;
        const int ul = e % UPC, r0 = e / UPC;                    // rows r0 and r0 + 8
        const int unit = rank * UPC + ul;
        const long long tstep = (dir == 0 ? -1 : 1) * p.st * ndir * H;
        int plen[NS][NR], pbv[NS][NR];
        float dc_reg[NS][NR];
        long long idx[NS][NR];
        float4 act_n[NS][NR];
        float cst_n[NS][NR], cprev_n[NS][NR], dout_n[NS][NR];
        auto prefetch = [&](int sl, int j, int t, long long ix) {
            act_n[sl][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            cst_n[sl][j] = 0.f; cprev_n[sl][j] = 0.f; dout_n[sl][j] = 0.f;
            if (t >= 0 && t < plen[sl][j]) {
                const int t_cprev = dir == 0 ? t - 1 : t + 1;
                act_n[sl][j] = reinterpret_cast<const float4*>(p.G)[ix];
                cst_n[sl][j] = p.Cst[ix];
                if (t_cprev >= 0 && t_cprev < plen[sl][j]) cprev_n[sl][j] = p.Cst[ix + tstep];
                dout_n[sl][j] = __ldg(p.dOut + ix);
            }
        };
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const int pb = (slice0 + sl) * R + r0 + 8 * j;
                pbv[sl][j] = pb;
                plen[sl][j] = (live[sl] && pb < p.B) ? p.lens[pb] : 0;
                dc_reg[sl][j] = 0.f;
                const int t0 = dir == 0 ? T - 1 : 0;
                idx[sl][j] = (((long long)pb * p.sb + (long long)t0 * p.st) * ndir + dir) * H + unit;
                if (live[sl] && pb < p.B)
                    for (int t = T; t < Tp; ++t)
                        reinterpret_cast<float4*>(p.G)[(((size_t)pb * p.sb + (size_t)t * p.st) * ndir + dir) * H + unit] =
                            make_float4(0.f, 0.f, 0.f, 0.f);
                prefetch(sl, j, t0, idx[sl][j]);
            }
        uint32_t phase = 0;
        for (int s = 0; s < T; ++s) {
            const int buf = s & 1;
            const int t = dir == 0 ? T - 1 - s : s;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) {
                if (!live[sl]) continue;
                float dh[NR];
                float4 actc[NR];
                float cstc[NR], cprevc[NR], doutc[NR];
                long long ixc[NR];
                // this step's operands (loaded a step ago) and the next step's prefetch, issued before the wait
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    dh[j] = 0.f;
                    actc[j] = act_n[sl][j]; cstc[j] = cst_n[sl][j]; cprevc[j] = cprev_n[sl][j]; doutc[j] = dout_n[sl][j];
                    ixc[j] = idx[sl][j];
                    idx[sl][j] = ixc[j] + tstep;
                    if (s + 1 < T) prefetch(sl, j, dir == 0 ? t - 1 : t + 1, ixc[j] + tstep);
                }
                if (s > 0) {
                    mbar_wait(&full[sl][buf], (phase >> (2 * sl + buf)) & 1u);
                    phase ^= 1u << (2 * sl + buf);
                    if (e == 0) mbar_expect_tx(&full[sl][buf], CS * TILE * 4);
                    const float* rb = recv + (size_t)(sl * 2 + buf) * CS * TILE + r0 * UPC + ul;
                    float pa[NR][2];
#pragma unroll
                    for (int j = 0; j < NR; ++j) { pa[j][0] = 0.f; pa[j][1] = 0.f; }
#pragma unroll
                    for (int src = 0; src < CS; ++src)
#pragma unroll
                        for (int j = 0; j < NR; ++j) pa[j][src & 1] += rb[src * TILE + 8 * UPC * j];
#pragma unroll
                    for (int j = 0; j < NR; ++j) dh[j] = pa[j][0] + pa[j][1];
                }
                float4 dz[NR];
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    dz[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t < plen[sl][j]) {
                        const float4 act = actc[j];
                        const float dhj = dh[j] + doutc[j];
                        const float si = act.x, tj = act.y, sf = act.z, so = act.w;
                        const float tc = tanh_fast(cstc[j]);
                        const float dct = dc_reg[sl][j] + dhj * so * (1.f - tc * tc);
                        dz[j].x = dct * tj * si * (1.f - si);
                        dz[j].y = dct * si * (1.f - tj * tj);
                        dz[j].z = dct * cprevc[j] * sf * (1.f - sf);
                        dz[j].w = dhj * tc * so * (1.f - so);
                        dc_reg[sl][j] = dct * sf;
                    }
                }
                if (F16 && s + 1 < T) {
                    // fp16 fragments of the CTA's dz tile, every row scaled by its own power of two: k16 pair q = ul / 4
                    // holds units 4q .. 4q+3; 16-bit slot of (row, unit, gate) = word [q][lane = (row % 8) * 4 + gate]
                    // [reg = 2 * ((ul % 4) / 2) + row / 8], half ul % 2; hi plane, then lo' plane (+512 words)
                    unsigned short* dqh = reinterpret_cast<unsigned short*>(dz_s + (size_t)(sl * 2 + buf) * 8 * 128);
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        const int row = r0 + 8 * j;
                        float amax = fmaxf(fmaxf(fabsf(dz[j].x), fabsf(dz[j].y)), fmaxf(fabsf(dz[j].z), fabsf(dz[j].w)));
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                        float sc = 1.0f;
                        if (amax >= 1.17549435e-38f) {
                            const int ex = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
                            sc = __uint_as_float((uint32_t)(min(max(13 - ex, -126), 126) + 127) << 23);
                        }
                        if (ul == 0) rinv_s[sl][buf][row] = 1.0f / sc;
                        const float v[4] = {dz[j].x * sc, dz[j].y * sc, dz[j].z * sc, dz[j].w * sc};
                        const int wbase = (ul >> 2) * 128 + (row & 7) * 16 + 2 * ((ul & 3) >> 1) + (row >> 3);
#pragma unroll
                        for (int gt = 0; gt < 4; ++gt) {
                            unsigned short hi, lo;
                            split_f16(v[gt], hi, lo);
                            dqh[(wbase + gt * 4) * 2 + (ul & 1)] = hi;
                            dqh[(512 + wbase + gt * 4) * 2 + (ul & 1)] = lo;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dzready[sl][buf]);
                } else if (s + 1 < T) {
                    // own dz tile in fragment order [k-tile = ul/2][16-byte chunk (r0*4 + gate) ^ k-tile][2 (ul%2) + row/8]:
                    // the thread's two rows are the adjacent floats of one slot pair
                    float* dq = dz_s + (size_t)(sl * 2 + buf) * 8 * 128 + (ul >> 1) * 128 + 2 * (ul & 1) + (r0 >> 3);
                    const int c0 = (r0 & 7) * 4, kx = ul >> 1;
                    if (NR == 2) {
                        *reinterpret_cast<float2*>(dq + ((c0 + 0) ^ kx) * 4) = make_float2(dz[0].x, dz[NR - 1].x);
                        *reinterpret_cast<float2*>(dq + ((c0 + 1) ^ kx) * 4) = make_float2(dz[0].y, dz[NR - 1].y);
                        *reinterpret_cast<float2*>(dq + ((c0 + 2) ^ kx) * 4) = make_float2(dz[0].z, dz[NR - 1].z);
                        *reinterpret_cast<float2*>(dq + ((c0 + 3) ^ kx) * 4) = make_float2(dz[0].w, dz[NR - 1].w);
                    } else {
                        dq[((c0 + 0) ^ kx) * 4] = dz[0].x;
                        dq[((c0 + 1) ^ kx) * 4] = dz[0].y;
                        dq[((c0 + 2) ^ kx) * 4] = dz[0].z;
                        dq[((c0 + 3) ^ kx) * 4] = dz[0].w;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dzready[sl][buf]);
                }
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    if (pbv[sl][j] < p.B) reinterpret_cast<float4*>(p.G)[ixc[j]] = dz[j];
                }
            }
        }
    }
    cluster_sync_all();
}

// epilogue warps: 8 (one (row, unit) element per thread) when two slices are interleaved -- the epilogue then
// competes with the other slice's k-loop for issue slots and must be short; 4 (two rows per thread) for NS = 1,
// where fewer warps mean fewer arrivals on the hand-off barriers (measured: 2844 vs 3037 cycles per step)
constexpr int ews(int NS) { return NS == 2 ? 8 : 4; }
size_t fwd_ws_smem(int CS, int NS) { return sizeof(float) * ((size_t)NS * 2 * CS * TILE) + (size_t)NS * 2 * 16 * ZST * 16; }
size_t bwd_ws_smem(int CS, int NS) { return sizeof(float) * ((size_t)NS * 4 * CS * TILE + (size_t)NS * 2 * 8 * 128); }

}  // namespace
int g_rec_fwd_f16 = 1;      // forward recurrence: 1 = fp16 split scheme, 0 = tf32 + bf16 scheme (test mode 7)
// backward recurrence: 1 = fp16 split scheme with per-row scaled dz tiles, 0 = tf32 + bf16 scheme (test mode 8).  Measured
// alone at cfg-2 (B = 64, T = 700): 1.95 vs 2.11 us per timestep.
int g_rec_bwd_f16 = 1;

namespace {

template <int CS, int NS>
int launch_ws(cudaStream_t st, bool bwd, const MParams& p, int nclusters, int* max_active) {
    auto kf = g_rec_fwd_f16 ? rec_fwd_ws_kernel<CS, NS, ews(NS), true> : rec_fwd_ws_kernel<CS, NS, ews(NS), false>;
    auto kb = g_rec_bwd_f16 ? rec_bwd_ws_kernel<CS, NS, ews(NS), true> : rec_bwd_ws_kernel<CS, NS, ews(NS), false>;
    const void* fn = bwd ? (const void*)kb : (const void*)kf;
    size_t smem = bwd ? bwd_ws_smem(CS, NS) : fwd_ws_smem(CS, NS);
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CS);
    cfg.blockDim = dim3(256 + 32 * ews(NS));
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
int run_ws(cudaStream_t st, bool bwd, MParams& p, void* ws, size_t ws_bytes, int force_ns) {
    // NS = 1 (one slice per cluster) when every cluster is co-resident, else two interleaved slices
    static int max_active[2] = {-1, -1};
    if (max_active[bwd] < 0) {
        int rc = launch_ws<CS, 1>(st, bwd, p, 1, &max_active[bwd]);
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
    return ns == 1 ? launch_ws<CS, 1>(st, bwd, p, nclusters, nullptr) : launch_ws<CS, 2>(st, bwd, p, nclusters, nullptr);
}

}  // namespace

extern long long* g_rec_dbg;

// returns 0 = launched, -1 = not eligible (caller uses the other kernels), >0 = error
int lstm_rec_ws(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int ndir, long long sb, long long stt,
                float* G, float* Hout, float* Cst, const float* Wh, const float* dOut, const int* lens, void* ws,
                size_t ws_bytes, int force_ns, int carry_c) {
    if (H != 128 && H != 256) return -1;
    if (B <= 0 || T <= 0) return 0;
    MParams p;
    p.carry_c = (!bwd && ndir == 1) ? carry_c : 0;
    p.G = G; p.Hout = Hout; p.Cst = Cst; p.Wh = Wh; p.dOut = dOut; p.lens = lens; p.xg = nullptr;
    p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.ndir = ndir; p.nslices = cdiv(B, R); p.sb = sb; p.st = stt;
    p.dbg = bwd ? nullptr : g_rec_dbg;
    if (H == 128) return run_ws<8>(st, bwd, p, ws, ws_bytes, force_ns);
    return run_ws<16>(st, bwd, p, ws, ws_bytes, force_ns);
}

}  // namespace e2e
