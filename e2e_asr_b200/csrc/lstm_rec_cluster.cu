// Thread-block-cluster LSTM recurrence (forward and backward): the fast path of
// lstm_rec.cu for H in {16,...,256} (H/16 CTAs per cluster, up to the 16-CTA
// non-portable cluster size).
//
// Same decomposition as lstm_rec.cu -- one GROUP per (direction, 16-row batch
// slice), CTA j of the group owns hidden units [16j, 16j+16) and keeps its W_hh
// slice resident in shared memory for the whole sequence -- but a group is now ONE
// thread-block cluster and nothing on the per-timestep critical path touches L2:
//   forward : every CTA all-gathers h_t (16 rows x 16 units = 1 KB per CTA) into
//             the 16 peers' shared memory with cp.async.bulk over DSMEM, completion
//             counted on the receiver's mbarrier (complete_tx) -- data and signal in
//             one transaction, no barrier, no flag polling in global memory.
//   backward: d h_{t-1} = dz_t . W_hh^T is computed as a K-sliced partial product
//             (each CTA contracts over ITS OWN 64 gate columns, so it reuses the
//             forward's W slice and never needs other CTAs' dz) followed by a
//             reduce-scatter of 1 KB partial tiles over DSMEM; 4x less traffic than
//             all-gathering dz.
// Layouts and length semantics are those documented in lstm_rec.cu.
#include "common.cuh"

namespace e2e {

struct RecParams;   // defined in lstm_rec.cu (same fields)

namespace {

struct CParams {
    float* G;
    float* Hout;
    float* Cst;
    const float* Wh;
    const float* dOut;
    const int* lens;
    int B, T, Tp, H, ndir;
    long long sb, st;
    long long* dbg;      // optional per-step clock64 stamps of cluster 0 / rank 0 (5 per step)
};

constexpr int R = 16, UPC = 16, NTH = 256;
constexpr int NC = 4 * UPC;                        // gate columns owned by one CTA (64)
constexpr int HSTR = 20;                           // row stride (floats) of an exchanged [R][UPC] tile: bank-conflict-free
constexpr int TILE_FLOATS = R * HSTR;              // one CTA's h / partial tile incl. padding
constexpr int TILE_BYTES = TILE_FLOATS * 4;        // 1280 B
constexpr int ZSTR = NC + 4;                       // row stride of the own-dz tile [R][NC]

// The recurrent product runs on the tensor cores with error-compensated TF32
// (3xTF32): x = hi + lo with hi = x & ~0x1fff (exactly a TF32 number) and
// lo = x - hi; D += lo*hi + hi*lo + hi*hi in the fp32 accumulator.  Relative error
// ~2^-21, inside the 1e-4 parity budget, at a fraction of the FFMA issue slots.
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
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
}

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 33)) __trap();      // never hang the GPU on a protocol bug
    }
}
// local smem -> (possibly remote) smem of CTA `rank`, completion on that CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_local_addr, uint32_t bar_local_addr, uint32_t rank,
                                                const void* src, uint32_t bytes) {
    uint32_t dst = mapa(dst_local_addr, rank), bar = mapa(bar_local_addr, rank);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(s_u32(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// direct DSMEM exchange: remote vector store + remote mbarrier arrive (a few hundred cycles end to end,
// against ~3000 for a cp.async.bulk round trip measured on B200)
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > (1ll << 33)) __trap();
    }
}
// all threads: push this CTA's [R][UPC] tile (local rows of HSTR floats at `src`) into slot `slot_addr`
// (local address of the same slot in every peer) of the CS peers, then one arrival per peer
template <int CS>
__device__ __forceinline__ void push_tile(const float* src, uint32_t slot_addr, uint32_t bar_addr, int tid) {
    const int chunk = tid % 64, row = chunk / 4, c4 = chunk % 4, dg = tid / 64;
    const float4 v = *reinterpret_cast<const float4*>(src + row * HSTR + c4 * 4);
    const uint32_t off = (uint32_t)(row * HSTR + c4 * 4) * 4u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int dest = dg * 4 + j;
        if (dest < CS) st_cluster_v4(mapa(slot_addr + off, dest), v);
    }
    // the CTA barrier orders every thread's remote stores before the release-arrives below
    __syncthreads();
    if (tid < CS) mbar_arrive_cluster(mapa(bar_addr, tid));
}

// ---------------------------------------------------------------- forward
// Per step: z[16 rows x 64 own gate columns] = h_{t-1}[16 x H] . W_hh[:, own columns] as
// mma.sync m16n8k8 (3xTF32); warp w owns gate columns [8w, 8w+8) = units 2w, 2w+1.
template <int CS>
__global__ void __launch_bounds__(NTH, 1) rec_fwd_cluster_kernel(CParams p) {
    constexpr int H = CS * UPC;
    constexpr int WKS = H + 8;                                            // Wt row stride: conflict-free fragments
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                                                     // [NC][WKS]  Wt[n][k] = W_hh[k][own col n]
    float* h_s = Wt + (size_t)NC * WKS;                                   // [2][CS][R][HSTR]
    float* stage = h_s + 2 * CS * TILE_FLOATS;                            // [2][R][HSTR]
    __shared__ __align__(8) uint64_t full[2];

    const int ndir = p.ndir, T = p.T;
    const uint32_t rank = cluster_rank();
    const int group = blockIdx.x / CS;
    const int dir = group % ndir;
    const int b0 = (group / ndir) * R;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;

    if (tid == 0) {
        mbar_init(&full[0], CS);
        mbar_init(&full[1], CS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float* Wg = p.Wh + (size_t)dir * H * H * 4;
        for (int i = tid; i < H * NC; i += NTH) {
            int k = i / NC, n = i % NC;
            Wt[n * WKS + k] = Wg[((size_t)k * H + rank * UPC) * 4 + n];
        }
        for (int i = tid; i < 2 * TILE_FLOATS; i += NTH) stage[i] = 0.f;
    }
    // this thread's pointwise element after the fragment exchange
    const int prow = g + 8 * (tq & 1);
    const int ul = 2 * w + (tq >> 1);
    const int unit = rank * UPC + ul;
    const int pb = b0 + prow;
    const int plen = pb < p.B ? p.lens[pb] : 0;
    float c_reg = 0.f, h_reg = 0.f;
    uint32_t ph[2] = {0u, 0u};
    __syncthreads();
    cluster_sync_all();

    const bool rec = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    for (int s = 0; s < T; ++s) {
        const int buf = s & 1;
        const int t = dir == 0 ? s : T - 1 - s;
        if (rec) p.dbg[s * 5 + 0] = clock64();
        float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t row = (size_t)pb * p.sb + (size_t)t * p.st;
        if (t < plen) gx = reinterpret_cast<const float4*>(p.G)[(row * ndir + dir) * H + unit];
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        if (s > 0) {
            mbar_wait_cluster(&full[buf], ph[buf]);
            ph[buf] ^= 1u;
            if (rec) p.dbg[s * 5 + 1] = clock64();
            const float* hb = h_s + (size_t)buf * CS * TILE_FLOATS;
            const float* wrow = Wt + (size_t)(8 * w + g) * WKS;
            // independent accumulator chains (the mma.sync result latency would otherwise serialise
            // the 96 MMAs): 2 k-phases x {hi*hi, cross terms}
            float dm[2][4], dx[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { dm[i][j] = 0.f; dx[i][j] = 0.f; }
#pragma unroll 4
            for (int kk = 0; kk < H / 8; kk += 2) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int k = 8 * (kk + i) + tq;
                    const int src = k / UPC, off = k % UPC;
                    const float* ha = hb + (size_t)(src * R + g) * HSTR + off;
                    uint32_t ah[4], al[4], bh[2], bl[2];
                    split_tf32(ha[0], ah[0], al[0]);
                    split_tf32(ha[8 * HSTR], ah[1], al[1]);
                    split_tf32(ha[4], ah[2], al[2]);
                    split_tf32(ha[8 * HSTR + 4], ah[3], al[3]);
                    split_tf32(wrow[k], bh[0], bl[0]);
                    split_tf32(wrow[k + 4], bh[1], bl[1]);
                    mma_tf32(dx[i], al, bh);
                    mma_tf32(dm[i], ah, bh);
                    mma_tf32(dx[i], ah, bl);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = (dx[0][j] + dx[1][j]) + (dm[0][j] + dm[1][j]);
        }
        if (rec) p.dbg[s * 5 + 2] = clock64();
        // fragment exchange: even tq keeps row g (gets f,o from its neighbour), odd tq keeps row g+8
        const bool even = (tq & 1) == 0;
        const float s0 = even ? d[2] : d[0], s1 = even ? d[3] : d[1];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        const float z0 = even ? d[0] : r0, z1 = even ? d[1] : r1, z2 = even ? r0 : d[2], z3 = even ? r1 : d[3];
        float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
        float cn = 0.f;
        const bool active = t < plen;
        if (active) {
            float si = sigmoidf_acc(z0 + gx.x);
            float tj = tanhf(z1 + gx.y);
            float sf = sigmoidf_acc(z2 + gx.z + 1.0f);
            float so = sigmoidf_acc(z3 + gx.w);
            cn = c_reg * sf + si * tj;
            h_reg = tanhf(cn) * so;
            c_reg = cn;
            act = make_float4(si, tj, sf, so);
        }
        // critical path first: publish h_t (state h: carried through for masked rows)
        stage[(size_t)buf * TILE_FLOATS + prow * HSTR + ul] = h_reg;
        if (rec) p.dbg[s * 5 + 3] = clock64();
        __syncthreads();
        if (rec) p.dbg[s * 5 + 4] = clock64();
        if (s + 1 < T)
            push_tile<CS>(stage + (size_t)buf * TILE_FLOATS,
                          s_u32(h_s + ((size_t)(buf ^ 1) * CS + rank) * TILE_FLOATS), s_u32(&full[buf ^ 1]), tid);
        if (active) {
            reinterpret_cast<float4*>(p.G)[(row * ndir + dir) * H + unit] = act;
            p.Cst[(row * ndir + dir) * H + unit] = cn;
            p.Hout[row * ndir * H + dir * H + unit] = h_reg;
        }
    }
    cluster_sync_all();      // nobody exits while peers may still write into its shared memory
}

// ---------------------------------------------------------------- backward
// Per step: partial[16 rows x H] = dz_t[16 x 64 own gate columns] . W_hh[:, own columns]^T as
// mma.sync (3xTF32), reduce-scattered over the cluster; then the pointwise LSTM backward.
template <int CS>
__global__ void __launch_bounds__(NTH, 1) rec_bwd_cluster_kernel(CParams p) {
    constexpr int H = CS * UPC;
    constexpr int WKS = H + 8;
    constexpr int NTILES = H / 8;                                          // 8-wide output tiles
    constexpr int MAXT = (NTILES + 7) / 8;                                 // per warp
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                                                      // [NC][WKS]
    float* red_s = Wt + (size_t)NC * WKS;                                  // [2][CS][R][HSTR] received partials
    float* pst = red_s + 2 * CS * TILE_FLOATS;                             // [2][CS][R][HSTR] partials to send
    float* dz_s = pst + 2 * CS * TILE_FLOATS;                              // [R][ZSTR] own dz of the previous step
    __shared__ __align__(8) uint64_t full[2];

    const int ndir = p.ndir, T = p.T, Tp = p.Tp;
    const uint32_t rank = cluster_rank();
    const int group = blockIdx.x / CS;
    const int dir = group % ndir;
    const int b0 = (group / ndir) * R;
    const int tid = threadIdx.x;
    const int w = tid / 32, lane = tid % 32, g = lane / 4, tq = lane % 4;

    if (tid == 0) {
        mbar_init(&full[0], CS);
        mbar_init(&full[1], CS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float* Wg = p.Wh + (size_t)dir * H * H * 4;
        for (int i = tid; i < H * NC; i += NTH) {
            int k = i / NC, n = i % NC;
            Wt[n * WKS + k] = Wg[((size_t)k * H + rank * UPC) * 4 + n];
        }
        for (int i = tid; i < 2 * CS * TILE_FLOATS; i += NTH) pst[i] = 0.f;
    }
    // pointwise element of this thread: row = tid / UPC, unit ul = tid % UPC
    const int prow = tid / UPC, ul = tid % UPC;
    const int unit = rank * UPC + ul;
    const int pb = b0 + prow;
    const int plen = pb < p.B ? p.lens[pb] : 0;
    float dc_reg = 0.f;
    if (pb < p.B)
        for (int t = T; t < Tp; ++t)
            reinterpret_cast<float4*>(p.G)[(((size_t)pb * p.sb + (size_t)t * p.st) * ndir + dir) * H + unit] =
                make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t ph[2] = {0u, 0u};
    __syncthreads();
    cluster_sync_all();

    for (int s = 0; s < T; ++s) {
        const int buf = s & 1;
        const int t = dir == 0 ? T - 1 - s : s;
        const int t_cprev = dir == 0 ? t - 1 : t + 1;
        const bool active = t < plen;
        const size_t row = (size_t)pb * p.sb + (size_t)t * p.st;
        float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
        float cst = 0.f, cprev = 0.f, dout = 0.f;
        if (active) {
            act = reinterpret_cast<const float4*>(p.G)[(row * ndir + dir) * H + unit];
            cst = p.Cst[(row * ndir + dir) * H + unit];
            if (t_cprev >= 0 && t_cprev < plen)
                cprev = p.Cst[(((size_t)pb * p.sb + (size_t)t_cprev * p.st) * ndir + dir) * H + unit];
            dout = __ldg(p.dOut + row * ndir * H + dir * H + unit);
        }
        float dh = 0.f;
        if (s > 0) {
            float* ps = pst + (size_t)buf * CS * TILE_FLOATS;
            float d[MAXT][4], dxx[MAXT][4];
#pragma unroll
            for (int i = 0; i < MAXT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { d[i][j] = 0.f; dxx[i][j] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < NC / 8; ++kk) {
                const int kc = 8 * kk + tq;
                uint32_t ah[4], al[4];
                split_tf32(dz_s[g * ZSTR + kc], ah[0], al[0]);
                split_tf32(dz_s[(g + 8) * ZSTR + kc], ah[1], al[1]);
                split_tf32(dz_s[g * ZSTR + kc + 4], ah[2], al[2]);
                split_tf32(dz_s[(g + 8) * ZSTR + kc + 4], ah[3], al[3]);
#pragma unroll
                for (int i = 0; i < MAXT; ++i) {
                    const int nt = w + 8 * i;
                    if (nt < NTILES) {
                        uint32_t bh[2], bl[2];
                        split_tf32(Wt[(size_t)kc * WKS + 8 * nt + g], bh[0], bl[0]);
                        split_tf32(Wt[(size_t)(kc + 4) * WKS + 8 * nt + g], bh[1], bl[1]);
                        mma_tf32(dxx[i], al, bh);
                        mma_tf32(d[i], ah, bh);
                        mma_tf32(dxx[i], ah, bl);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < MAXT; ++i) {
                const int nt = w + 8 * i;
                if (nt < NTILES) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[i][j] += dxx[i][j];
                    const int dest = nt / 2, du = (nt % 2) * 8 + 2 * tq;
                    *reinterpret_cast<float2*>(ps + (size_t)(dest * R + g) * HSTR + du) = make_float2(d[i][0], d[i][1]);
                    *reinterpret_cast<float2*>(ps + (size_t)(dest * R + g + 8) * HSTR + du) = make_float2(d[i][2], d[i][3]);
                }
            }
            __syncthreads();
            {
                // reduce-scatter: tile `dest` of my partials goes to slot `rank` of peer `dest`
                const int chunk = tid % 64, prow_c = chunk / 4, c4 = chunk % 4, dg = tid / 64;
                const uint32_t slot = s_u32(red_s + ((size_t)buf * CS + rank) * TILE_FLOATS) +
                                      (uint32_t)(prow_c * HSTR + c4 * 4) * 4u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int dest = dg * 4 + j;
                    if (dest < CS) {
                        const float4 v = *reinterpret_cast<const float4*>(ps + (size_t)dest * TILE_FLOATS +
                                                                           prow_c * HSTR + c4 * 4);
                        st_cluster_v4(mapa(slot, dest), v);
                    }
                }
                __syncthreads();
                if (tid < CS) mbar_arrive_cluster(mapa(s_u32(&full[buf]), tid));
            }
            mbar_wait_cluster(&full[buf], ph[buf]);
            ph[buf] ^= 1u;
            const float* rb = red_s + (size_t)buf * CS * TILE_FLOATS;
#pragma unroll
            for (int src = 0; src < CS; ++src) dh += rb[(size_t)(src * R + prow) * HSTR + ul];
        }
        float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            dh += dout;
            float si = act.x, tj = act.y, sf = act.z, so = act.w;
            float tc = tanhf(cst);
            float dct = dc_reg + dh * so * (1.f - tc * tc);
            dz.x = dct * tj * si * (1.f - si);
            dz.y = dct * si * (1.f - tj * tj);
            dz.z = dct * cprev * sf * (1.f - sf);
            dz.w = dh * tc * so * (1.f - so);
            dc_reg = dct * sf;
        }
        *reinterpret_cast<float4*>(dz_s + prow * ZSTR + ul * 4) = dz;
        __syncthreads();
        if (pb < p.B) reinterpret_cast<float4*>(p.G)[(row * ndir + dir) * H + unit] = dz;
    }
    cluster_sync_all();
}

size_t fwd_smem(int CS) { return sizeof(float) * ((size_t)NC * (CS * UPC + 8) + (2 * CS + 2) * TILE_FLOATS); }
size_t bwd_smem(int CS) { return sizeof(float) * ((size_t)NC * (CS * UPC + 8) + 4 * CS * TILE_FLOATS + R * ZSTR); }

template <int CS>
int launch_cluster(cudaStream_t st, bool bwd, const CParams& p, int ngroups) {
    auto kf = rec_fwd_cluster_kernel<CS>;
    auto kb = rec_bwd_cluster_kernel<CS>;
    const void* fn = bwd ? (const void*)kb : (const void*)kf;
    size_t smem = bwd ? bwd_smem(CS) : fwd_smem(CS);
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ngroups * CS);
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
    int nclusters = 0;
    E2E_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg));
    if (nclusters < 1) return -1;            // this cluster shape cannot be scheduled: caller falls back
    CParams pc = p;
    if (bwd) E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kb, pc));
    else E2E_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kf, pc));
    ++g_launches;
    return 0;
}

}  // namespace

// returns 0 = launched, -1 = not eligible (caller uses the L2/global-barrier kernel), >0 = error
long long* g_rec_dbg = nullptr;

int lstm_rec_cluster(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int ndir, long long sb, long long stt,
                     float* G, float* Hout, float* Cst, const float* Wh, const float* dOut, const int* lens) {
    if (H % UPC != 0) return -1;
    const int CS = H / UPC;
    if (CS != 1 && CS != 2 && CS != 4 && CS != 8 && CS != 16) return -1;
    if (B <= 0 || T <= 0) return 0;
    CParams p;
    p.G = G; p.Hout = Hout; p.Cst = Cst; p.Wh = Wh; p.dOut = dOut; p.lens = lens;
    p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.ndir = ndir; p.sb = sb; p.st = stt; p.dbg = bwd ? nullptr : g_rec_dbg;
    const int ngroups = ndir * cdiv(B, R);
    switch (CS) {
        case 1: return launch_cluster<1>(st, bwd, p, ngroups);
        case 2: return launch_cluster<2>(st, bwd, p, ngroups);
        case 4: return launch_cluster<4>(st, bwd, p, ngroups);
        case 8: return launch_cluster<8>(st, bwd, p, ngroups);
        default: return launch_cluster<16>(st, bwd, p, ngroups);
    }
}

}  // namespace e2e
