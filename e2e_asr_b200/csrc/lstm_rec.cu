// Persistent LSTM recurrence (forward and backward) for the pyramidal BiLSTM
// encoder layers and the decoder's LM-LSTM.
//
// Replaces the per-timestep TF while_loop of bidirectional_dynamic_rnn
// (reference encoder.py:77-81; BasicLSTMCell math pinned by basic_lstm.py:14-23):
// the all-timestep input projection x_t.W_x + b is a GEMM done beforehand; this
// kernel does h_{t-1}.W_hh + gates + state update for every t in ONE launch.
//
// Decomposition (hard part 1): batch rows are independent
// and so are the two directions, so the grid is cut into independent GROUPS
// (direction x batch-slice of R rows).  Inside a group, CTA j owns UPC hidden
// units: its 4*UPC gate columns of W_hh stay resident in shared memory for the
// whole sequence, c stays in registers, and only h_t (R x H floats) is
// exchanged per step through L2 (the layer-output buffer itself) behind a
// per-group release/acquire counter -- no grid-wide barrier.
//
// Layouts (all fp32, batch-major rows r = b*Tp + t, Tp >= max(len)+1):
//   G    [B][Tp][ndir][H][4]   in: x-projection+bias, gate-interleaved (i,j,f,o per
//                              unit);  fwd overwrites it with the activations
//                              (sig i, tanh j, sig(f+1), sig o); bwd overwrites
//                              it with d(pre-activation).
//   Hout [B][Tp][ndir*H]       layer output (zero where t >= len: must be zeroed
//                              by the caller before fwd)
//   Cst  [B][Tp][ndir][H]      c'_t for active steps
//   Wh   [ndir][H][H][4]       W_hh, [k][unit][gate]
// Length semantics: rows with t >= len emit 0 and keep state; the
// backward direction walks t downwards from T-1 under the same mask.
#include "common.cuh"

namespace e2e {

struct RecParams {
    float* G;
    float* Hout;
    float* Cst;
    const float* Wh;
    const float* dOut;   // bwd only: [B][Tp][ndir*H]
    const int* lens;
    unsigned* ctr;       // one counter per group, zero-initialised
    int* err;
    int B, T, Tp, H, ndir;
    long long sb, st;         // row index of (b, t) = b*sb + t*st (batch-major: Tp,1; time-major: 1,B)
    int b_begin, nb_slices;   // rows handled by this launch: [b_begin, b_begin + nb_slices*R)
};

// ---------------------------------------------------------------- forward
template <int RT, int UPC>
__global__ void __launch_bounds__(16 * UPC, 1) lstm_rec_fwd_kernel(RecParams p) {
    constexpr int R = 4 * RT;
    constexpr int NTH = 16 * UPC;
    constexpr int Q = (RT + 3) / 4;
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, ndir = p.ndir, T = p.T;
    const int HP = H + 4;
    float4* W_s = reinterpret_cast<float4*>(smem);               // [H][UPC] float4 (4 gates)
    float* h_s = smem + (size_t)H * UPC * 4;                     // [R][HP]

    const int nslices = H / UPC;
    const int slice = blockIdx.x % nslices;
    const int group = blockIdx.x / nslices;
    const int dir = group % ndir;
    const int b0 = p.b_begin + (group / ndir) * R;
    unsigned* ctr = p.ctr + group;

    const int tid = threadIdx.x;
    const int unit_lo = tid % 8, ks = (tid / 8) % 4, unit_hi = (tid / 32) % (UPC / 8);
    const int rg = tid / (32 * (UPC / 8));
    const int ul = unit_hi * 8 + unit_lo;            // unit within the CTA
    const int unit = slice * UPC + ul;               // unit within the layer

    // resident W_hh slice
    {
        const float4* Wg = reinterpret_cast<const float4*>(p.Wh) + (size_t)dir * H * H;
        for (int i = tid; i < H * UPC; i += NTH) {
            int k = i / UPC, u = i % UPC;
            W_s[i] = Wg[(size_t)k * H + slice * UPC + u];
        }
    }
    float c_reg[Q];
    int len_reg[Q];
    int b_reg[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c_reg[q] = 0.f;
        int r = ks + 4 * q;
        int b = b0 + rg * RT + r;
        b_reg[q] = b;
        len_reg[q] = (r < RT && b < p.B) ? p.lens[b] : 0;
    }
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        const int t_prev = dir == 0 ? t - 1 : t + 1;
        // prefetch this thread's x-projections (independent of the recurrence)
        float4 gx[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            gx[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < len_reg[q])
                gx[q] = reinterpret_cast<const float4*>(p.G)[
                    (((size_t)b_reg[q] * p.sb + (size_t)t * p.st) * ndir + dir) * H + unit];
        }
        float acc[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;

        if (s > 0) {
            if (tid == 0) spin_wait_ge(ctr, (unsigned)s * nslices, p.err);
            __syncthreads();
            // h_{t_prev} tile of this group's rows: Hout[b][t_prev][dir*H .. +H]
            const int chunks = H / 4;
            for (int i = tid; i < R * chunks; i += NTH) {
                int row = i / chunks, k4 = i % chunks;
                int b = b0 + row;
                float* dst = h_s + row * HP + k4 * 4;
                if (b < p.B)
                    cp_async16(dst, p.Hout + ((size_t)b * p.sb + (size_t)t_prev * p.st) * ndir * H + dir * H + k4 * 4);
                else
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            const float* hrow = h_s + (rg * RT) * HP;
            for (int k0 = 4 * ks; k0 < H; k0 += 16) {
                float4 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = W_s[(k0 + j) * UPC + ul];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    float4 hv = *reinterpret_cast<const float4*>(hrow + r * HP + k0);
                    const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[r][0] = fmaf(hk[j], w[j].x, acc[r][0]);
                        acc[r][1] = fmaf(hk[j], w[j].y, acc[r][1]);
                        acc[r][2] = fmaf(hk[j], w[j].z, acc[r][2]);
                        acc[r][3] = fmaf(hk[j], w[j].w, acc[r][3]);
                    }
                }
            }
            // all-reduce over the 4 k-split lanes (lane bits 3 and 4)
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v = acc[r][g];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    acc[r][g] = v;
                }
        }
        // pointwise: lane ks owns rows r = ks + 4q
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r == ks + 4 * q) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) z[g] = acc[r][g];
                }
            if (t < len_reg[q]) {
                float si = sigmoidf_acc(z[0] + gx[q].x);
                float tj = tanhf(z[1] + gx[q].y);
                float sf = sigmoidf_acc(z[2] + gx[q].z + 1.0f);
                float so = sigmoidf_acc(z[3] + gx[q].w);
                float cn = c_reg[q] * sf + si * tj;
                float hn = tanhf(cn) * so;
                c_reg[q] = cn;
                size_t row = (size_t)b_reg[q] * p.sb + (size_t)t * p.st;
                reinterpret_cast<float4*>(p.G)[(row * ndir + dir) * H + unit] = make_float4(si, tj, sf, so);
                p.Cst[(row * ndir + dir) * H + unit] = cn;
                p.Hout[row * ndir * H + dir * H + unit] = hn;
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            red_release_gpu_add(ctr, 1u);
        }
    }
}

// ---------------------------------------------------------------- backward
template <int RT, int UPC>
__global__ void __launch_bounds__(16 * UPC, 1) lstm_rec_bwd_kernel(RecParams p) {
    constexpr int R = 4 * RT;
    constexpr int NTH = 16 * UPC;
    constexpr int NOUT = 4 * RT;                 // outputs per (rg, unit-quad) thread group
    constexpr int Q = (NOUT + 15) / 16;
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, ndir = p.ndir, Tp = p.Tp, T = p.T;
    const int G4 = 4 * H, GP = 4 * H + 4;
    float* W_s = smem;                            // [UPC][GP]  rows of W_hh owned by this CTA
    float* dz_s = smem + (size_t)UPC * GP;        // [R][GP]

    const int nslices = H / UPC;
    const int slice = blockIdx.x % nslices;
    const int group = blockIdx.x / nslices;
    const int dir = group % ndir;
    const int b0 = p.b_begin + (group / ndir) * R;
    unsigned* ctr = p.ctr + group;

    const int tid = threadIdx.x;
    const int ks = tid % 16, uq = (tid / 16) % (UPC / 4), rg = tid / (16 * (UPC / 4));

    {
        const float* Wg = p.Wh + (size_t)dir * H * G4 + (size_t)slice * UPC * G4;
        for (int i = tid; i < UPC * (G4 / 4); i += NTH) {
            int u = i / (G4 / 4), c4 = i % (G4 / 4);
            *reinterpret_cast<float4*>(W_s + u * GP + c4 * 4) =
                *reinterpret_cast<const float4*>(Wg + (size_t)u * G4 + c4 * 4);
        }
    }
    // this thread's pointwise outputs: o = ks + 16q -> (row = o/4, unit quad member o%4)
    float dc_reg[Q];
    int len_reg[Q], b_reg[Q], unit_reg[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        dc_reg[q] = 0.f;
        int o = ks + 16 * q;
        int b = b0 + rg * RT + o / 4;
        b_reg[q] = b;
        unit_reg[q] = slice * UPC + uq * 4 + (o % 4);
        len_reg[q] = (o < NOUT && b < p.B) ? p.lens[b] : 0;
        // zero d(pre-activation) of the never-visited padded tail t in [T, Tp)
        if (o < NOUT && b < p.B)
            for (int t = T; t < Tp; ++t)
                reinterpret_cast<float4*>(p.G)[(((size_t)b * p.sb + (size_t)t * p.st) * ndir + dir) * H + unit_reg[q]] =
                    make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        // reverse of the forward order: fw walks t = T-1..0, bw walks t = 0..T-1
        const int t = dir == 0 ? T - 1 - s : s;
        const int t_done = dir == 0 ? t + 1 : t - 1;      // step processed just before (its dz feeds dh)
        const int t_cprev = dir == 0 ? t - 1 : t + 1;     // forward predecessor (c_{prev})
        float4 act[Q];
        float cst[Q], cprev[Q], dout[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            act[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            cst[q] = cprev[q] = dout[q] = 0.f;
            if (t < len_reg[q]) {
                size_t row = (size_t)b_reg[q] * p.sb + (size_t)t * p.st;
                act[q] = reinterpret_cast<const float4*>(p.G)[(row * ndir + dir) * H + unit_reg[q]];
                cst[q] = p.Cst[(row * ndir + dir) * H + unit_reg[q]];
                if (t_cprev >= 0 && t_cprev < len_reg[q])
                    cprev[q] = p.Cst[(((size_t)b_reg[q] * p.sb + (size_t)t_cprev * p.st) * ndir + dir) * H + unit_reg[q]];
                dout[q] = __ldg(p.dOut + row * ndir * H + dir * H + unit_reg[q]);
            }
        }
        float acc[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;

        if (s > 0) {
            if (tid == 0) spin_wait_ge(ctr, (unsigned)s * nslices, p.err);
            __syncthreads();
            const int chunks = G4 / 4;
            for (int i = tid; i < R * chunks; i += NTH) {
                int row = i / chunks, c4 = i % chunks;
                int b = b0 + row;
                float* dst = dz_s + row * GP + c4 * 4;
                if (b < p.B)
                    cp_async16(dst, p.G + (((size_t)b * p.sb + (size_t)t_done * p.st) * ndir + dir) * G4 + c4 * 4);
                else
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            const float* zrow = dz_s + (rg * RT) * GP;
            const float* wrow = W_s + (uq * 4) * GP;
            for (int c0 = 4 * ks; c0 < G4; c0 += 64) {
                float4 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(wrow + j * GP + c0);
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    float4 z = *reinterpret_cast<const float4*>(zrow + r * GP + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[r][j] = fmaf(z.x, w[j].x, acc[r][j]);
                        acc[r][j] = fmaf(z.y, w[j].y, acc[r][j]);
                        acc[r][j] = fmaf(z.z, w[j].z, acc[r][j]);
                        acc[r][j] = fmaf(z.w, w[j].w, acc[r][j]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[r][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    acc[r][j] = v;
                }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int o = ks + 16 * q;
            if (o >= NOUT || b_reg[q] >= p.B) continue;
            float dh = 0.f;
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r * 4 + j == o) dh = acc[r][j];
            size_t row = (size_t)b_reg[q] * p.sb + (size_t)t * p.st;
            float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < len_reg[q]) {
                dh += dout[q];
                float si = act[q].x, tj = act[q].y, sf = act[q].z, so = act[q].w;
                float tc = tanhf(cst[q]);
                float dct = dc_reg[q] + dh * so * (1.f - tc * tc);
                dz.x = dct * tj * si * (1.f - si);
                dz.y = dct * si * (1.f - tj * tj);
                dz.z = dct * cprev[q] * sf * (1.f - sf);
                dz.w = dh * tc * so * (1.f - so);
                dc_reg[q] = dct * sf;
            }
            reinterpret_cast<float4*>(p.G)[(row * ndir + dir) * H + unit_reg[q]] = dz;
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            red_release_gpu_add(ctr, 1u);
        }
    }
}

// ---------------------------------------------------------------- host side
template <int RT, int UPC>
static int launch_rec(cudaStream_t st, bool bwd, RecParams p, int ngroups, size_t smem) {
    auto kf = lstm_rec_fwd_kernel<RT, UPC>;
    auto kb = lstm_rec_bwd_kernel<RT, UPC>;
    const void* fn = bwd ? (const void*)kb : (const void*)kf;
    E2E_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = ngroups * (p.H / UPC);
    void* args[] = {&p};
    // cooperative launch: guarantees all CTAs of every group are co-resident
    E2E_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(16 * UPC), args, smem, st));
    ++g_launches;
    return 0;
}

static size_t rec_smem(bool bwd, int RT, int UPC, int H) {
    int R = 4 * RT;
    if (!bwd) return sizeof(float) * ((size_t)H * UPC * 4 + (size_t)R * (H + 4));
    return sizeof(float) * ((size_t)(UPC + R) * (4 * H + 4));
}

int lstm_rec_ws(cudaStream_t, bool, int, int, int, int, int, long long, long long, float*, float*, float*,
                const float*, const float*, const int*, void*, size_t, int, int);
int lstm_rec_h512(cudaStream_t, bool, int, int, int, int, int, long long, long long, float*, float*, float*,
                  const float*, const float*, const int*, void*, size_t);
// 0 = fastest eligible kernel (warp-specialised register-resident cluster kernel for H in {128, 256}, the H = 512
// kernel of lstm_rec_h512.cu when the workspace is large enough, else the L2-exchange kernel below); 1 = always the L2
// kernel; 5 / 6 = the warp-specialised kernel with 1 / 2 slices per cluster forced (tests); 7 / 8 = the warp-specialised
// kernel with the forward / backward pass on the tf32 + bf16 scheme instead of the fp16 split scheme.  (2 - 4 selected
// two earlier kernel generations, removed: they behave like 0.)
int g_rec_mode = 0;
long long* g_rec_dbg = nullptr;     // e2e_set_rec_debug: clock64 stamps of the forward cluster kernel / the decoder loop
extern int g_rec_fwd_f16, g_rec_bwd_f16;

// forward-only layer continuing a sequence (cell state carried in from Cst at time -1): the warp-specialised kernel only
int lstm_rec_fwd_carry(cudaStream_t st, int B, int T, int Tp, int H, long long sb, long long stt, float* G, float* Hout,
                       float* Cst, const float* Wh, const int* lens, void* ctr_ws, size_t ctr_ws_bytes) {
    E2E_REQUIRE(H == 128 || H == 256, "lstm_rec_fwd_carry: hidden size %d (the carried-state entry serves 128 / 256)", H);
    if (B <= 0 || T <= 0) return 0;
    g_rec_fwd_f16 = 1;
    int rc = lstm_rec_ws(st, false, B, T, Tp, H, 1, sb, stt, G, Hout, Cst, Wh, nullptr, lens, ctr_ws, ctr_ws_bytes, 0, 1);
    E2E_REQUIRE(rc >= 0, "lstm_rec_fwd_carry: the cluster kernel is not available for these shapes");
    return rc;
}

// workspace: ctr_ws must hold >= 4*ceil(B/4) unsigned + 1 int, zeroed by this function.
int lstm_rec(cudaStream_t st, bool bwd, int B, int T, int Tp, int H, int ndir, long long sb, long long stt,
             float* G, float* Hout, float* Cst, const float* Wh, const float* dOut, const int* lens,
             void* ctr_ws, size_t ctr_ws_bytes, int* err_flag) {
    E2E_REQUIRE(H % 8 == 0, "lstm_rec: hidden size %d must be a multiple of 8", H);
    E2E_REQUIRE(ndir == 1 || ndir == 2, "lstm_rec: ndir must be 1 or 2");
    E2E_REQUIRE(Tp >= T, "lstm_rec: Tp (%d) must be >= T (%d)", Tp, T);
    if (B <= 0 || T <= 0) return 0;
    if (g_rec_mode != 1) {
        g_rec_fwd_f16 = g_rec_mode != 7;
        g_rec_bwd_f16 = g_rec_mode != 8;
        int rc = lstm_rec_ws(st, bwd, B, T, Tp, H, ndir, sb, stt, G, Hout, Cst, Wh, dOut, lens, ctr_ws, ctr_ws_bytes,
                             (g_rec_mode == 5 || g_rec_mode == 6) ? g_rec_mode - 4 : 0, 0);
        if (rc >= 0) return rc;
    }
    if (g_rec_mode != 1) {
        int rc = lstm_rec_h512(st, bwd, B, T, Tp, H, ndir, sb, stt, G, Hout, Cst, Wh, dOut, lens, ctr_ws, ctr_ws_bytes);
        if (rc >= 0) return rc;
    }
    const int UPC = (H % 16 == 0) ? 16 : 8;
    const int nslices = H / UPC;
    const int nsm = sm_count();
    const size_t smem_cap = 220 * 1024;
    // smallest rows-per-CTA whose grid fits the SMs in one launch; otherwise the largest that fits smem
    int RT = 0;
    const int cands[4] = {1, 2, 4, 8};
    for (int i = 0; i < 4; ++i) {
        int rt = cands[i];
        if (rec_smem(bwd, rt, UPC, H) > smem_cap) break;
        RT = rt;
        int nb = cdiv(B, 4 * rt);
        if (ndir * nb * nslices <= nsm) break;
    }
    E2E_REQUIRE(RT > 0, "lstm_rec: hidden size %d does not fit shared memory", H);
    E2E_REQUIRE(ndir * nslices <= nsm, "lstm_rec: hidden size %d needs more than %d CTAs per batch slice", H, nsm);
    const int R = 4 * RT;
    const int max_slices_per_launch = nsm / (ndir * nslices);
    const int nb_total = cdiv(B, R);
    size_t smem = rec_smem(bwd, RT, UPC, H);
    for (int slice0 = 0; slice0 < nb_total; slice0 += max_slices_per_launch) {
        int nb = min(max_slices_per_launch, nb_total - slice0);
        int ngroups = nb * ndir;
        E2E_REQUIRE((size_t)ngroups * sizeof(unsigned) <= ctr_ws_bytes, "lstm_rec: counter workspace too small");
        E2E_CHECK_CUDA(cudaMemsetAsync(ctr_ws, 0, (size_t)ngroups * sizeof(unsigned), st));
        RecParams p;
        p.G = G; p.Hout = Hout; p.Cst = Cst; p.Wh = Wh; p.dOut = dOut; p.lens = lens;
        p.ctr = (unsigned*)ctr_ws; p.err = err_flag;
        p.B = B; p.T = T; p.Tp = Tp; p.H = H; p.ndir = ndir; p.sb = sb; p.st = stt;
        p.b_begin = slice0 * R; p.nb_slices = nb;
        int rc;
#define CASE(RT_, UPC_) if (RT == RT_ && UPC == UPC_) { rc = launch_rec<RT_, UPC_>(st, bwd, p, ngroups, smem); if (rc) return rc; continue; }
        CASE(1, 8) CASE(2, 8) CASE(4, 8) CASE(8, 8)
        CASE(1, 16) CASE(2, 16) CASE(4, 16) CASE(8, 16)
#undef CASE
        E2E_REQUIRE(false, "lstm_rec: no kernel for RT=%d UPC=%d", RT, UPC);
    }
    return 0;
}

// ------------------------------------------------------- weight (un)packing
// TF BasicLSTMCell kernel [(I+H), 4H], columns gate-blocked (i | j | f | o)
// (basic_lstm.py:17) -> Wx [I][ldwx] at column offset dir*4H, gate-interleaved
// [unit][4]; Wh [H][H][4]; bias likewise.
__global__ void lstm_pack_kernel(int I, int H, const float* __restrict__ kernel, const float* __restrict__ bias,
                                 float* __restrict__ Wx, int ldwx, int col0, float* __restrict__ Wh,
                                 float* __restrict__ bp) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = (size_t)(I + H + 1) * 4 * H;
    if (i >= total) return;
    int row = (int)(i / (4 * H)), c = (int)(i % (4 * H));
    int unit = c / 4, gate = c % 4;
    if (row < I) Wx[(size_t)row * ldwx + col0 + c] = kernel[(size_t)row * 4 * H + gate * H + unit];
    else if (row < I + H) Wh[(size_t)(row - I) * 4 * H + c] = kernel[(size_t)row * 4 * H + gate * H + unit];
    else bp[col0 + c] = bias[gate * H + unit];
}

__global__ void lstm_unpack_kernel(int I, int H, float* __restrict__ dkernel, float* __restrict__ dbias,
                                   const float* __restrict__ dWx, int ldwx, int col0,
                                   const float* __restrict__ dWh, const float* __restrict__ dbp, int accumulate) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = (size_t)(I + H + 1) * 4 * H;
    if (i >= total) return;
    int row = (int)(i / (4 * H)), c = (int)(i % (4 * H));   // c indexes the TF (gate-blocked) column
    int gate = c / H, unit = c % H;
    int pc = unit * 4 + gate;
    float v;
    float* dst;
    if (row < I) { v = dWx[(size_t)row * ldwx + col0 + pc]; dst = dkernel + (size_t)row * 4 * H + c; }
    else if (row < I + H) { v = dWh[(size_t)(row - I) * 4 * H + pc]; dst = dkernel + (size_t)row * 4 * H + c; }
    else { v = dbp[col0 + pc]; dst = dbias + c; }
    *dst = accumulate ? *dst + v : v;
}

int lstm_pack_weights(cudaStream_t st, int I, int H, const float* kernel, const float* bias, float* Wx,
                      int ldwx, int col0, float* Wh, float* bias_packed) {
    size_t total = (size_t)(I + H + 1) * 4 * H;
    lstm_pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(I, H, kernel, bias, Wx, ldwx, col0, Wh, bias_packed);
    E2E_LAUNCH_CHECK();
    return 0;
}

int lstm_unpack_grads(cudaStream_t st, int I, int H, float* dkernel, float* dbias, const float* dWx, int ldwx,
                      int col0, const float* dWh, const float* dbias_packed, int accumulate) {
    size_t total = (size_t)(I + H + 1) * 4 * H;
    lstm_unpack_kernel<<<cdiv(total, 256), 256, 0, st>>>(I, H, dkernel, dbias, dWx, ldwx, col0, dWh,
                                                        dbias_packed, accumulate);
    E2E_LAUNCH_CHECK();
    return 0;
}

}  // namespace e2e
