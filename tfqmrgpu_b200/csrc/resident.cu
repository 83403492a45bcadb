// The resident solver: a whole tfQMR solve of a SMALL system in ONE cooperative launch.
//
// A system like the reference's FD example (171 block rows of 8 x 8 blocks, one block column, 171 KB per vector) spends its time
// in latency, not in bandwidth: an iteration is 8 phases that each depend on a column-wide reduction or on the whole of the
// previous vector, and every phase of the per-kernel path is a chain of ~10 dependent L2 round trips (tile descriptor ->
// coefficients -> vector loads -> partial sums -> ticket -> last tile's sums -> scalars -> next launch): 64 us per iteration
// whether the eleven kernels are launched one by one, as a CUDA graph, or - a first version of this file - walked through by a
// persistent grid with a grid barrier in the place of every kernel boundary (2.64 / 2.63 / 2.67 ms per solve).
//
// What removes the round trips is ownership.  Here every CTA OWNS one vector tile for the whole solve:
//   * v1, v4, v5, v7, v8, v9 and the shadow vector v3 of the tile live in shared memory from the first to the last iteration;
//     only v6 - the vector the block-sparse product reads across tiles - is also written to global memory;
//   * the CTA computes the Y blocks of its own tile (one warp per Y block, A and X blocks staged in shared memory), so the
//     product's result never leaves the SM and E1 / E2 continue on it without a barrier;
//   * a column-wide sum is: tile partial -> global, ONE grid barrier, then every CTA of the column adds the column's partials in
//     the same fixed order and runs the scalar recurrence itself (dec35 / dec34 / decT of vec_body.cuh on shared-memory state),
//     so the per-column scalars and the solver's control block are replicated per CTA and never read from global memory;
//     the global iteration / probe decision (core.hxx:239-304) is taken redundantly by every CTA from the columns' monitors.
// An iteration is 6 grid barriers (K1 | P1 E1 | K2 | K3 | P2 E2 | K4) and no other cross-CTA wait.  The arithmetic statements
// are those of the per-kernel path (vec_body.cuh, core.hxx:189-233); sums are grouped differently, so results agree to rounding.
//
// Coherence: the L1 is not invalidated inside a launch, so everything another CTA wrote during the launch (v6 / v1 blocks, partial
// sums, monitors) is read with ld.global.cg; A, the index tables and B are constant and use the non-coherent path.  The file is
// also compiled with -Xptxas -dlcm=cg, so a plain load that slipped in would still be correct.
#include "vec_body.cuh"
#include <cstdlib>
#include <cstdio>
#include <vector>

namespace tfq {

namespace {

constexpr int kResThreads = 256;
constexpr int kResWarps = kResThreads/32;
constexpr int kResLoads = 16;               // 128-bit loads in flight per lane while a batch of entries is staged

template <typename real_t> struct ResidentArgs {
    real_t *v1, *v5, *v6;                     // global: X (written at the end and for probes), v5 = b after INIT, v6 (product input)
    float const *v3;
    real_t const *A, *B;
    uint32_t const *bpos; uint32_t nnzbB;
    real_t const *rho, *beta;                 // state after OP_INIT (solve_begin)
    double const *tau, *invBn2;
    int8_t *status, *snap;                    // out: per right-hand side (getRhsStatus)
    double *part, *colmon;
    Control *ctl;
    Tile const *tiles; uint32_t const *coltile; uint32_t nCols;
    uint32_t const *unit_e0, *ent_a, *ent_x, *unit_of_block;   // unit_of_block: (unit << 4) | column slot of the Y block in its unit
    uint32_t gmax;                            // block columns per unit (the entry table holds gmax X blocks per entry)
    unsigned *bar;                            // [0] arrivals, [1] generation
    uint32_t tile_elems;                      // reals per vector tile in shared memory (largest tile)
    uint32_t tile_blocks;                     // blocks of the largest tile
    int eb;                                   // entries staged per batch and warp
    int max_e;                                // entries per Y block whose indices (and, a_resident, A blocks) are kept in shared memory
    int warps_per_block;                      // warps that share one Y block in the product (a power of two)
    int a_resident;                           // the A blocks of the tile's rows stay in shared memory for the whole solve
    int ablate;                               // dev: 1 = skip the FMAs of the product, 2 = skip its global loads
    unsigned long long *trace;                // dev: cycles of CTA 0 in [barriers, products, column sums, total, ...], or nullptr
};

__device__ __forceinline__ unsigned long long res_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Grid-wide barrier; all CTAs of the cooperative launch are resident, so spinning is safe.  One monotonic counter: the k-th
// barrier is passed when nCta*k arrivals have been counted (red.release + ld.acquire polling: two L2 round trips, no reset, no
// generation word).  A barrier that is not released within 4 s (a lost CTA) traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nCta, unsigned &passed) {
    __syncthreads();
    ++passed;
    if (0 == threadIdx.x) {
        unsigned const target = nCta*passed;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
        unsigned seen;
        unsigned long long t0 = 0;
        for (unsigned spins = 1; ; ++spins) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
            if (int(seen - target) >= 0) break;
            if (0 == (spins & 0x3ffu)) {
                unsigned long long const now = res_now_ns();
                if (0 == t0) t0 = now; else if (now - t0 > 4000000000ull) __trap();
            }
        }
    }
    __syncthreads();
}

// columns of a Y block per lane in the product: the smallest divisor of LN that lets (LN / TJ) x LM lanes cover the block
constexpr int res_tj(int LM, int LN) {
    for (int tj = 1; tj <= LN; ++tj) if (0 == LN % tj && (LN/tj)*LM <= 32) return tj;
    return LN;
}

template <typename real_t, int LM, int LN>
__global__ void __launch_bounds__(kResThreads)
resident_solve_kernel(ResidentArgs<real_t> const a)
{
    constexpr int PL = LM*LN;                  // one plane (Re or Im) of a block
    constexpr int BE = 2*PL;                   // reals per X block
    constexpr int BA = 2*LM*LM;                // reals per A block
    constexpr int R = kResThreads/LN;          // thread rows per lane j
    constexpr int TJ = res_tj(LM, LN);
    constexpr int KS = (TJ <= 2) ? 4 : ((TJ <= 4) ? 2 : 1);       // accumulator sets in the product (see there)
    constexpr int aF4 = int(BA*sizeof(real_t)/16), xF4 = int(BE*sizeof(real_t)/16);
    static_assert((BA*sizeof(real_t)) % 16 == 0 && (BE*sizeof(real_t)) % 16 == 0, "blocks are whole float4s");
    static_assert(LN <= 64 && R >= 1 && 2*LN <= kResThreads, "lane count");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned const nCta = gridDim.x;
    Tile const t = a.tiles[blockIdx.x];
    uint32_t const c = t.col;
    int const nb = int(t.b1 - t.b0), nrows = nb*LM;
    uint32_t const t0 = a.coltile[c], t1 = a.coltile[c + 1];
    bool const first = (blockIdx.x == t0);     // publishes the column's monitor, status and snap

    // ---- shared memory: seven vector tiles, v3, the column's scalars, reduction scratch, control block, product staging ----
    size_t const TE = a.tile_elems;
    real_t *const s_v1 = reinterpret_cast<real_t*>(smem_raw);
    real_t *const s_v4 = s_v1 + TE, *const s_v5 = s_v4 + TE, *const s_v6 = s_v5 + TE, *const s_v7 = s_v6 + TE,
           *const s_v8 = s_v7 + TE, *const s_v9 = s_v8 + TE;
    float  *const s_v3 = reinterpret_cast<float*>(s_v9 + TE);
    double *const s_d = reinterpret_cast<double*>(smem_raw + ((7*TE*sizeof(real_t) + TE*sizeof(float) + 15)/16)*16);
    double *const s_tau = s_d, *const s_var = s_d + LN, *const s_inv = s_d + 2*LN;
    double *const s_red = s_d + 3*LN;          // [R][2][LN]
    double *const s_sum = s_red + 2*R*LN;      // [2][LN] column sums, then [2][LN] monitor scratch
    double *const s_mon = s_sum + 2*LN;
    real_t *const s_rho = reinterpret_cast<real_t*>(s_mon + 2*LN);      // complex scalars: [2][LN] each
    real_t *const s_alfa = s_rho + 2*LN, *const s_beta = s_alfa + 2*LN, *const s_c67 = s_beta + 2*LN, *const s_eta = s_c67 + 2*LN;
    int8_t *const s_status = reinterpret_cast<int8_t*>(s_eta + 2*LN), *const s_snap = s_status + LN;
    Control *const lc = reinterpret_cast<Control*>((reinterpret_cast<uintptr_t>(s_snap + LN) + 15) & ~uintptr_t(15));
    uint32_t *const s_e0 = reinterpret_cast<uint32_t*>(lc + 1);                       // [3][tile_blocks] first / end entry, column slot of every Y block
    uint32_t *const s_ea = s_e0 + 3*a.tile_blocks;                                  // [tile_blocks][max_e] A block of an entry
    uint32_t *const s_ex = s_ea + size_t(a.tile_blocks)*a.max_e;                      // [tile_blocks][max_e] X block (storage index)
    real_t *const s_A = reinterpret_cast<real_t*>((reinterpret_cast<uintptr_t>(s_ex + size_t(a.tile_blocks)*a.max_e) + 15) & ~uintptr_t(15));
    real_t *const s_yp = s_A + (a.a_resident ? size_t(a.tile_blocks)*a.max_e*BA : 0);             // [kResWarps][BE] partial Y blocks
    float4 *const s_stage = reinterpret_cast<float4*>(s_yp + size_t(kResWarps)*BE);

    // the recurrences of vec_body.cuh work on this CTA's copy of its column's scalars (column index 0)
    VecArgs<real_t> sv;
    sv.rho = s_rho; sv.alfa = s_alfa; sv.beta = s_beta; sv.c67 = s_c67; sv.eta = s_eta;
    sv.tau = s_tau; sv.var = s_var; sv.invBn2 = s_inv; sv.status = s_status; sv.snap = s_snap;
    sv.LM = LM; sv.LN = LN;

    // ---- load the tile: v5 = b and the state that OP_INIT left (core.hxx:114-131,153-165,189-192), everything else zero ----
    {
        real_t const *const g5 = a.v5 + size_t(t.b0)*BE;
        float const *const g3 = a.v3 + size_t(t.b0)*BE;
        for (int q = tid; q < nb*BE; q += kResThreads) {
            s_v5[q] = g5[q]; s_v3[q] = g3[q];
            s_v1[q] = 0; s_v4[q] = 0; s_v6[q] = 0; s_v7[q] = 0; s_v8[q] = 0; s_v9[q] = 0;
        }
        if (tid < LN) {
            size_t const s = size_t(c)*LN + tid, r = (size_t(c)*2 + 0)*LN + tid, m = (size_t(c)*2 + 1)*LN + tid;
            s_rho[tid] = a.rho[r]; s_rho[LN + tid] = a.rho[m];
            s_beta[tid] = a.beta[r]; s_beta[LN + tid] = a.beta[m];
            s_alfa[tid] = 0; s_alfa[LN + tid] = 0; s_c67[tid] = 0; s_c67[LN + tid] = 0; s_eta[tid] = 0; s_eta[LN + tid] = 0;
            s_tau[tid] = a.tau[s]; s_var[tid] = 0; s_inv[tid] = a.invBn2[s];
            s_status[tid] = a.status[s]; s_snap[tid] = a.snap[s];
        }
        if (0 == tid) *lc = *a.ctl;
    }
    __syncthreads();
    // entry lists of the tile's Y blocks (the entries of the block's unit, with the X block of the block's column), and the A blocks
    // themselves when they fit
    for (int yb = 0; yb < nb; ++yb) {
        uint32_t const ug = a.unit_of_block[t.b0 + yb];
        uint32_t const u = ug >> 4, g = ug & 15u;
        uint32_t const e0 = a.unit_e0[u], e1 = a.unit_e0[u + 1];
        int const ne = int(min(e1 - e0, uint32_t(a.max_e)));
        if (0 == tid) { s_e0[yb] = e0; s_e0[a.tile_blocks + yb] = e1; s_e0[2*a.tile_blocks + yb] = g; }
        for (int e = tid; e < ne; e += kResThreads) { s_ea[yb*a.max_e + e] = a.ent_a[e0 + e]; s_ex[yb*a.max_e + e] = a.ent_x[size_t(e0 + e)*a.gmax + g]; }
        if (a.a_resident) {
            float4 *const dst = reinterpret_cast<float4*>(s_A + size_t(yb)*a.max_e*BA);
            for (int q = tid; q < ne*aF4; q += kResThreads)
                dst[q] = __ldg(reinterpret_cast<float4 const*>(a.A) + size_t(a.ent_a[e0 + q/aF4])*aF4 + q % aF4);
        }
    }
    __syncthreads();

    int const j = tid % LN, rt = tid / LN;
    bool const act = rt < R;
    unsigned redphase = 0;                     // partial sums are double-buffered by the parity of the reduction
    unsigned passed = 0;                       // grid barriers passed so far
    bool const tracing = (nullptr != a.trace) && 0 == blockIdx.x && 0 == tid;     // dev: cycle counts of CTA 0, thread 0, kept in registers
    long long tr_bar = 0, tr_prod = 0, tr_sum = 0, tr_stage = 0, tr_fma = 0;
    long long const nRHS = (long long)(a.nCols)*LN;

    // element (row q of the tile, lane j): Re at idx, Im at idx + PL
    auto idx = [&](int q) { return (q / LM)*BE + (q % LM)*LN + j; };

    // tile sums acc[0..D) per lane -> global partials, grid barrier, column sums in s_sum[d*LN + j].  The column's partials are
    // added by kResThreads / (D*LN) slices of threads (slice s takes the tiles t0 + s, t0 + s + S, ...: independent loads in
    // flight instead of one dependent chain), then the slices in order: the same grouping in every CTA of the column.
    auto reduce = [&](double acc0, double acc1, int D) {
        long long const tr0 = tracing ? clock64() : 0;
        if (act) { s_red[(rt*2 + 0)*LN + j] = acc0; s_red[(rt*2 + 1)*LN + j] = acc1; }
        __syncthreads();
        unsigned const par = (redphase++) & 1u;
        if (tid < D*LN) {
            int const d = tid / LN, jj = tid - d*LN;
            double s = 0;
            for (int r = 0; r < R; ++r) s += s_red[(r*2 + d)*LN + jj];
            a.part[(size_t(blockIdx.x)*kPartD + par*2 + d)*LN + jj] = s;
        }
        long long const tb0 = tracing ? clock64() : 0;
        grid_barrier(a.bar, nCta, passed);
        if (tracing) tr_bar += clock64() - tb0;
        int const nq = D*LN, S = kResThreads/nq;
        int const q = tid % nq, sl = tid / nq;
        if (sl < S) {
            int const d = q / LN, jj = q - d*LN;
            double s = 0;
            #pragma unroll 4
            for (uint32_t tt = t0 + sl; tt < t1; tt += S) s += __ldcg(&a.part[(size_t(tt)*kPartD + par*2 + d)*LN + jj]);
            s_red[sl*nq + q] = s;
        }
        __syncthreads();
        if (tid < nq) {
            double s = s_red[tid];
            for (int s2 = 1; s2 < S; ++s2) s += s_red[s2*nq + tid];
            s_sum[tid] = s;                      // [d*LN + j]
        }
        __syncthreads();
        if (tracing) tr_sum += clock64() - tr0;
    };

    // the tile's v6 (or v1) to global memory for the other tiles' products
    auto publish = [&](real_t *g, real_t const *s) {
        real_t *const dst = g + size_t(t.b0)*BE;
        for (int q = tid; q < nb*BE; q += kResThreads) dst[q] = s[q];
    };

    // Y blocks of this tile: y := A * x  (blocksparse.hxx:71-199, blockmult.hxx:28-82).
    // One warp per Y block.  The entry indices are in shared memory (s_ea / s_ex, loaded once), the A blocks too when they fit
    // (a_resident: A is read from global memory ONCE per solve), and the X blocks of a batch of entries are fetched with all
    // their 128-bit loads in flight before the first one is stored to the staging area.
    auto product = [&](real_t *sy, real_t const *gx) {
        long long const tp0 = tracing ? clock64() : 0;
        int const i = lane % LM, jg = lane / LM;
        bool const on = jg < LN/TJ;
        int const perE = a.a_resident ? xF4 : (aF4 + xF4);         // float4s staged per entry
        float4 *const st4 = s_stage + size_t(warp)*a.eb*perE;
        float4 const *const A4 = reinterpret_cast<float4 const*>(a.A);
        float4 const *const X4 = reinterpret_cast<float4 const*>(gx);
        // W warps share a Y block: warp `ws` of the group takes the block's entries ws, ws + W, ... and the partial sums are
        // added in the order of the warps (a fixed grouping: W follows from the tile size alone)
        int const W = a.warps_per_block, groups = kResWarps/W;
        int const grp = warp / W, ws = warp - grp*W;
        for (int yb0 = 0; yb0 < nb; yb0 += groups) {
            int const yb = yb0 + grp;
            if (yb < nb) {
                uint32_t const e0 = s_e0[yb], e1 = s_e0[a.tile_blocks + yb], gslot = s_e0[2*a.tile_blocks + yb];
                int const mine = (int(e1 - e0) > ws) ? (int(e1 - e0) - ws + W - 1)/W : 0;     // entries of this warp
                // KS independent accumulators per output (k = ks mod KS): one warp per scheduler cannot hide the latency of a
                // chain of dependent fp64 FMAs (measured ~40 cycles each), so the chains are made short and many
                real_t acr[KS][TJ], aci[KS][TJ];
                #pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    #pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) { acr[ks][jj] = 0; aci[ks][jj] = 0; }
                }
                for (int m0 = 0; m0 < mine; m0 += a.eb) {
                    int const ne = min(a.eb, mine - m0);
                    long long const tl0 = tracing ? clock64() : 0;
                    __syncwarp();
                    // per entry the block(s) are contiguous: lane l fetches float4 l, l + 32, ... of every block; all loads of a
                    // chunk of entries are in flight before the first is stored (few instructions: one warp per scheduler pays
                    // ~4 cycles for each)
                    constexpr int LPE = (aF4 + xF4 + 31)/32;                  // loads per entry and lane (upper bound)
                    constexpr int CHE = (kResLoads/LPE > 0) ? kResLoads/LPE : 1;   // entries per chunk
                    for (int c0 = 0; c0 < ne; c0 += CHE) {
                        float4 reg[CHE][LPE];
                        #pragma unroll
                        for (int ce = 0; ce < CHE; ++ce) {
                            int const e = c0 + ce;
                            int const le = ws + W*(m0 + e);                     // entry number within the Y block
                            bool const live = (e < ne) && !(a.ablate & 2);
                            uint32_t ia = 0, ix = kNoBlock;
                            if (live) {
                                if (!a.a_resident) ia = (le < a.max_e) ? s_ea[yb*a.max_e + le] : a.ent_a[e0 + le];
                                ix = (le < a.max_e) ? s_ex[yb*a.max_e + le] : a.ent_x[size_t(e0 + le)*a.gmax + gslot];
                            }
                            #pragma unroll
                            for (int l = 0; l < LPE; ++l) {
                                int const r = l*32 + lane;                      // float4 within the staged entry
                                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (live && r < perE) {
                                    if (r < perE - xF4) val = __ldg(A4 + size_t(ia)*aF4 + r);
                                    else if (kNoBlock != ix) val = __ldcg(X4 + size_t(ix)*xF4 + (r - (perE - xF4)));
                                }
                                reg[ce][l] = val;
                            }
                        }
                        #pragma unroll
                        for (int ce = 0; ce < CHE; ++ce) {
                            #pragma unroll
                            for (int l = 0; l < LPE; ++l) {
                                int const r = l*32 + lane;
                                if (c0 + ce < ne && r < perE) st4[(c0 + ce)*perE + r] = reg[ce][l];
                            }
                        }
                    }
                    __syncwarp();
                    if (tracing) tr_stage += clock64() - tl0;
                    if (on && !(a.ablate & 1)) {
                        for (int e = 0; e < ne; ++e) {
                            int const le = ws + W*(m0 + e);
                            real_t const *const Xs = reinterpret_cast<real_t const*>(st4 + size_t(e)*perE + (perE - xF4));
                            real_t const *const As = a.a_resident ? (s_A + (size_t(yb)*a.max_e + le)*BA)
                                                                  : reinterpret_cast<real_t const*>(st4 + size_t(e)*perE);
                            #pragma unroll
                            for (int k = 0; k < LM; ++k) {
                                real_t const ar = As[k*LM + i], ai = As[LM*LM + k*LM + i];
                                #pragma unroll
                                for (int jj = 0; jj < TJ; ++jj) {
                                    real_t const xr = Xs[k*LN + jg*TJ + jj], xi = Xs[PL + k*LN + jg*TJ + jj];
                                    acr[k % KS][jj] = fma( ar, xr, acr[k % KS][jj]);         // complex multiply-accumulate (blockmult.hxx:76-77)
                                    acr[k % KS][jj] = fma(-ai, xi, acr[k % KS][jj]);
                                    aci[k % KS][jj] = fma( ar, xi, aci[k % KS][jj]);
                                    aci[k % KS][jj] = fma( ai, xr, aci[k % KS][jj]);
                                }
                            }
                        }
                    }
                }
                if (tracing) tr_fma += clock64() - tp0;
                if (on) {
                    real_t *const dst = (1 == W) ? (sy + yb*BE) : (s_yp + size_t(warp)*BE);
                    #pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) {
                        real_t sr = acr[0][jj], si = aci[0][jj];
                        #pragma unroll
                        for (int ks = 1; ks < KS; ++ks) { sr += acr[ks][jj]; si += aci[ks][jj]; }
                        dst[i*LN + jg*TJ + jj] = sr;
                        dst[PL + i*LN + jg*TJ + jj] = si;
                    }
                }
            }
            if (W > 1) {
                __syncthreads();
                for (int q = tid; q < groups*BE; q += kResThreads) {
                    int const g = q / BE, el = q - g*BE;
                    if (yb0 + g < nb) {
                        real_t s = s_yp[size_t(g*W)*BE + el];
                        for (int w2 = 1; w2 < W; ++w2) s += s_yp[size_t(g*W + w2)*BE + el];
                        sy[(yb0 + g)*BE + el] = s;
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (tracing) tr_prod += clock64() - tp0;
    };

    // the columns' monitors -> (max, sum, sum) over all columns, identical in every CTA
    auto all_columns = [&](double &m0, double &m1, double &m2) {
        m0 = 0; m1 = 0; m2 = 0;
        for (uint32_t cc = 0; cc < a.nCols; ++cc) {
            double const x0 = __ldcg(&a.colmon[size_t(cc)*4 + 0]);
            m0 = (m0 < x0) ? x0 : m0;
            m1 += __ldcg(&a.colmon[size_t(cc)*4 + 1]);
            m2 += __ldcg(&a.colmon[size_t(cc)*4 + 2]);
        }
    };

    long long const tt0 = tracing ? clock64() : 0;
    while (true) {
        int st = lc->state;                    // replicated: every CTA takes the same decisions from the same numbers
        if (STATE_DONE == st) break;
        if (STATE_RUN == st) {
            // ---- K1: v6 := v5 + beta v6 (core.hxx:194) -----------------------------------------------------------------
            if (act) {
                real_t const pr = s_beta[j], pi = s_beta[LN + j];
                for (int q = rt; q < nrows; q += R) {
                    int const o = idx(q);
                    real_t const xr = s_v5[o], xi = s_v5[o + PL], yr = s_v6[o], yi = s_v6[o + PL];
                    s_v6[o] = xr + pr*yr - pi*yi; s_v6[o + PL] = xi + pi*yr + pr*yi;
                }
            }
            __syncthreads();
            publish(a.v6, s_v6);
            grid_barrier(a.bar, nCta, passed);
            // ---- v9 := A v6 (core.hxx:198); E1: v4 := v9 + beta (v8 + beta v4), z34 = v3.v4 -> alfa, c67 (core.hxx:196-205) ----
            product(s_v9, a.v6);
            {
                double z0 = 0, z1 = 0;
                if (act) {
                    real_t const pr = s_beta[j], pi = s_beta[LN + j];
                    for (int q = rt; q < nrows; q += R) {
                        int const o = idx(q);
                        real_t const ar = s_v8[o], ai = s_v8[o + PL], br = s_v9[o], bi = s_v9[o + PL], yr = s_v4[o], yi = s_v4[o + PL];
                        float const wr = s_v3[o], wi = s_v3[o + PL];
                        real_t const tr = ar + pr*yr - pi*yi, ti = ai + pi*yr + pr*yi;
                        real_t const nr = br + pr*tr - pi*ti, ni = bi + pi*tr + pr*ti;
                        s_v4[o] = nr; s_v4[o + PL] = ni;
                        real_t const dr = nr*wr - ni*wi, di = nr*wi + ni*wr;      // linalg.hxx:506-507
                        z0 += double(dr); z1 += double(di);
                    }
                }
                reduce(z0, z1, 2);
                if (tid < LN) dec34(sv, 0, tid, s_sum[tid], s_sum[LN + tid]);
                __syncthreads();
            }
            // ---- K2: v7 := v6 + c67 v7; v5 += alfa v9; |v5|^2 -> decT (core.hxx:207-214) ---------------------------------
            {
                double d0 = 0;
                if (act) {
                    real_t const pr = s_c67[j], pi = s_c67[LN + j], qr = s_alfa[j], qi = s_alfa[LN + j];
                    for (int q = rt; q < nrows; q += R) {
                        int const o = idx(q);
                        real_t const ar = s_v6[o], ai = s_v6[o + PL], yr = s_v7[o], yi = s_v7[o + PL];
                        real_t const br = s_v9[o], bi = s_v9[o + PL], zr = s_v5[o], zi = s_v5[o + PL];
                        s_v7[o] = ar + pr*yr - pi*yi; s_v7[o + PL] = ai + pi*yr + pr*yi;
                        real_t const mr = qr*br - qi*bi + zr, mi = qi*br + qr*bi + zi;      // linalg.hxx:656-657
                        s_v5[o] = mr; s_v5[o + PL] = mi;
                        d0 += double(mr)*double(mr) + double(mi)*double(mi);
                    }
                }
                reduce(d0, 0., 1);
                if (tid < LN) decT(sv, 0, tid, s_sum[tid], true);
                __syncthreads();
            }
            // ---- K3: v1 += eta v7; v6 += alfa v4; v7 := v6 + c67 v7 (core.hxx:216-220) ----------------------------------
            if (act) {
                real_t const pr = s_eta[j], pi = s_eta[LN + j], qr = s_alfa[j], qi = s_alfa[LN + j], sr = s_c67[j], si = s_c67[LN + j];
                for (int q = rt; q < nrows; q += R) {
                    int const o = idx(q);
                    real_t const yr = s_v7[o], yi = s_v7[o + PL], br = s_v4[o], bi = s_v4[o + PL], zr = s_v6[o], zi = s_v6[o + PL];
                    s_v1[o] = pr*yr - pi*yi + s_v1[o]; s_v1[o + PL] = pi*yr + pr*yi + s_v1[o + PL];
                    real_t const mr = qr*br - qi*bi + zr, mi = qi*br + qr*bi + zi;
                    s_v6[o] = mr; s_v6[o + PL] = mi;
                    s_v7[o] = mr + sr*yr - si*yi; s_v7[o + PL] = mi + si*yr + sr*yi;
                }
            }
            __syncthreads();
            publish(a.v6, s_v6);
            grid_barrier(a.bar, nCta, passed);
            // ---- v8 := A v6 (core.hxx:224); E2: v5 += alfa v8; |v5|^2 -> decT (core.hxx:226-231) -------------------------
            product(s_v8, a.v6);
            {
                double d0 = 0;
                if (act) {
                    real_t const pr = s_alfa[j], pi = s_alfa[LN + j];
                    for (int q = rt; q < nrows; q += R) {
                        int const o = idx(q);
                        real_t const br = s_v8[o], bi = s_v8[o + PL], zr = s_v5[o], zi = s_v5[o + PL];
                        real_t const mr = pr*br - pi*bi + zr, mi = pi*br + pr*bi + zi;
                        s_v5[o] = mr; s_v5[o + PL] = mi;
                        d0 += double(mr)*double(mr) + double(mi)*double(mi);
                    }
                }
                reduce(d0, 0., 1);
                if (tid < LN) decT(sv, 0, tid, s_sum[tid], false);
                __syncthreads();
            }
            // ---- K4: v1 += eta v7; next z35 = v3.v5 -> beta, rho; convergence monitor and the iteration decision
            //      (core.hxx:233,189-192,235-260).  The monitor only needs the state after E2, so the column's first tile publishes
            //      it BEFORE the barrier of this reduction and no extra barrier is needed. -------------------------------------
            {
                double z0 = 0, z1 = 0;
                if (act) {
                    real_t const pr = s_eta[j], pi = s_eta[LN + j];
                    for (int q = rt; q < nrows; q += R) {
                        int const o = idx(q);
                        real_t const yr = s_v7[o], yi = s_v7[o + PL], zr = s_v5[o], zi = s_v5[o + PL];
                        float const wr = s_v3[o], wi = s_v3[o + PL];
                        s_v1[o] = pr*yr - pi*yi + s_v1[o]; s_v1[o + PL] = pi*yr + pr*yi + s_v1[o + PL];
                        real_t const dr = zr*wr - zi*wi, di = zr*wi + zi*wr;
                        z0 += double(dr); z1 += double(di);
                    }
                }
                if (first && 0 == tid) {
                    double mx = 0, b4 = 0, b5 = 0;
                    for (int q = 0; q < LN; ++q) {
                        double const res2 = s_tau[q]*s_inv[q];
                        mx = (mx < res2) ? res2 : mx;            // std::max semantics (NaN never wins)
                        b4 += (-2 == s_status[q]); b5 += (-1 == s_status[q]);
                    }
                    a.colmon[size_t(c)*4 + 0] = mx; a.colmon[size_t(c)*4 + 1] = b4; a.colmon[size_t(c)*4 + 2] = b5;
                }
                reduce(z0, z1, 2);
                if (tid < LN) { s_snap[tid] = s_status[tid]; dec35(sv, 0, tid, s_sum[tid], s_sum[LN + tid]); }
                if (0 == tid) {
                    double m0, m1, m2;
                    all_columns(m0, m1, m2);
                    decide_iteration(*lc, m0, m1, m2, nRHS);
                }
                __syncthreads();
            }
            st = lc->state;
        }
        if (STATE_PROBE == st) {
            // ---- residual probe: v9 := A v1 - b, |v9|^2 per right-hand side (core.hxx:263-304) ---------------------------
            publish(a.v1, s_v1);
            grid_barrier(a.bar, nCta, passed);
            product(s_v9, a.v1);
            for (uint32_t b = 0; b < a.nnzbB; ++b) {
                uint32_t const pos = a.bpos[b];
                if (pos >= t.b0 && pos < t.b1) {
                    for (int q = tid; q < BE; q += kResThreads) s_v9[(pos - t.b0)*BE + q] -= a.B[size_t(b)*BE + q];
                }
            }
            __syncthreads();
            double d0 = 0;
            if (act) {
                for (int q = rt; q < nrows; q += R) {
                    int const o = idx(q);
                    real_t const xr = s_v9[o], xi = s_v9[o + PL];
                    d0 += double(xr)*double(xr) + double(xi)*double(xi);
                }
            }
            reduce(d0, 0., 1);
            if (tid < LN) {
                double const res2 = s_sum[tid]*s_inv[tid];
                double notdone = 0;
                if (res2 > lc->tol2) { if (0 == s_snap[tid]) notdone = 1; }
                else if (res2 <= 0) {           // core.hxx:282-285: the component counts as converged
                    if (s_status[tid] == s_snap[tid]) s_status[tid] = 1;
                    s_snap[tid] = 1;
                } else if (lc->freeze && 0 == s_snap[tid] && 0 == s_status[tid]) {
                    s_status[tid] = kFrozen; s_snap[tid] = kFrozen;     // early-freeze extension (vec_body.cuh)
                }
                s_mon[tid] = res2; s_mon[LN + tid] = notdone;
            }
            __syncthreads();
            if (first && 0 == tid) {
                double mx = 0, nd = 0;
                for (int q = 0; q < LN; ++q) { double const res2 = s_mon[q]; mx = (mx < res2) ? res2 : mx; nd += s_mon[LN + q]; }
                a.colmon[size_t(c)*4 + 0] = mx; a.colmon[size_t(c)*4 + 1] = nd; a.colmon[size_t(c)*4 + 2] = 0;
            }
            grid_barrier(a.bar, nCta, passed);
            if (0 == tid) {
                double m0, m1, m2;
                all_columns(m0, m1, m2);
                decide_probe(*lc, m0, m1);
            }
            grid_barrier(a.bar, nCta, passed);           // everyone has read the monitors before an iteration's K4 rewrites them
        }
    }

    if (tracing) {
        a.trace[0] = tr_bar; a.trace[1] = tr_prod; a.trace[2] = tr_sum; a.trace[3] = clock64() - tt0; a.trace[4] = tr_stage; a.trace[5] = tr_fma;
        a.trace[6] = (unsigned long long)(lc->iteration); a.trace[7] = (unsigned long long)(lc->probes);
    }
    // ---- results: X, the per-right-hand-side status, the control block ---------------------------------------------------
    publish(a.v1, s_v1);
    if (first && tid < LN) { a.status[size_t(c)*LN + tid] = s_status[tid]; a.snap[size_t(c)*LN + tid] = s_snap[tid]; }
    if (0 == blockIdx.x && 0 == tid) { lc->cols_done = 0; *a.ctl = *lc; }
}

__global__ void invert_units_kernel(uint32_t *unit_of_block, uint32_t const *unit_y, uint32_t nUnits, uint32_t gmax) {
    uint32_t const q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q < nUnits*gmax) { uint32_t const iy = unit_y[q]; if (kNoBlock != iy) unit_of_block[iy] = ((q / gmax) << 4) | (q % gmax); }
}

// shared memory of one CTA: tiles of tile_blocks blocks, max_e entries per Y block kept (indices; A blocks if a_res), eb staged
// entries per warp
template <typename real_t, int LM, int LN>
size_t resident_smem(size_t tile_blocks, int max_e, bool a_res, int eb) {
    constexpr int R = kResThreads/LN;
    size_t const tile_elems = tile_blocks*2*LM*LN;
    size_t b = ((7*tile_elems*sizeof(real_t) + tile_elems*sizeof(float) + 15)/16)*16;
    b += (3*LN + 2*R*LN + 4*LN)*sizeof(double) + 10*LN*sizeof(real_t) + 2*LN + 16 + sizeof(Control);
    b += (3*tile_blocks + 2*tile_blocks*max_e)*sizeof(uint32_t) + 16;
    if (a_res) b += tile_blocks*max_e*2*LM*LM*sizeof(real_t);
    b += size_t(kResWarps)*2*LM*LN*sizeof(real_t);
    b += size_t(kResWarps)*eb*((a_res ? 0 : 2*LM*LM) + 2*LM*LN)*sizeof(real_t);
    return b + 128;
}

// The resident solver cuts the vectors into its OWN tiles (not the plan's streaming tiles): one CTA per tile, so about one tile
// per SM (TFQMRGPU_RESIDENT_TILES_PER_SM of them), each a range of blocks of one block column; tiles of 1 / 2 blocks put 4 / 2
// warps on every Y block of the product.
struct ResidentTiling { std::vector<Tile> tiles; std::vector<uint32_t> coltile; uint32_t tile_blocks = 0; };
ResidentTiling resident_tiling(Plan const &p, int nsm) {
    ResidentTiling rt;
    char const *const e = std::getenv("TFQMRGPU_RESIDENT_TILES_PER_SM");
    int const per_sm = e ? std::max(1, std::atoi(e)) : 1;
    size_t const target = size_t(nsm)*per_sm;
    // the smallest tile size that needs at most `target` tiles
    uint32_t tb = std::max<uint32_t>(1, uint32_t((size_t(p.nnzbX) + target - 1)/target));
    for (;; ++tb) {
        size_t n = 0;
        for (uint32_t c = 0; c < p.nCols; ++c) n += std::max<size_t>(1, (size_t(p.h_colstart[c + 1] - p.h_colstart[c]) + tb - 1)/tb);
        if (n <= target || tb >= uint32_t(std::max(p.nnzbX, 1))) break;
    }
    rt.coltile.assign(size_t(p.nCols) + 1, 0);
    for (uint32_t c = 0; c < p.nCols; ++c) {
        uint32_t const b0 = p.h_colstart[c], n = p.h_colstart[c + 1] - b0;
        uint32_t const nt = std::max<uint32_t>(1, (n + tb - 1)/tb);
        rt.coltile[c] = uint32_t(rt.tiles.size());
        for (uint32_t t = 0; t < nt; ++t) {
            Tile tile; tile.col = c; tile.pad = 0;
            tile.b0 = b0 + uint32_t((uint64_t(n)*t)/nt);
            tile.b1 = b0 + uint32_t((uint64_t(n)*(t + 1))/nt);
            rt.tile_blocks = std::max(rt.tile_blocks, tile.b1 - tile.b0);
            rt.tiles.push_back(tile);
        }
    }
    rt.coltile[p.nCols] = uint32_t(rt.tiles.size());
    rt.tile_blocks = std::max<uint32_t>(rt.tile_blocks, 1);
    return rt;
}

template <typename real_t, int LM, int LN>
tfqmrgpuStatus_t launch_resident(Plan &p, cudaStream_t stream, bool dry)
{
    auto kernel = resident_solve_kernel<real_t, LM, LN>;
    int dev = 0, nsm = 148, coop = 0;
    TFQ_CUDA(cudaGetDevice(&dev));
    TFQ_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    TFQ_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop || p.gmax < 1 || p.gmax > 16 || p.nUnits >= (1u << 28) || p.nnzbX < 1 || p.h_rpA.empty() || p.h_colstart.size() != size_t(p.nCols) + 1)
        return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    ResidentTiling const rt = resident_tiling(p, nsm);
    uint32_t const nTiles = uint32_t(rt.tiles.size());
    size_t const tile_blocks = rt.tile_blocks;
    // Only really small systems: with more than 4 blocks per tile a CTA's four warps take several rounds over its Y blocks and
    // the per-kernel path, which spreads the product over all SMs' threads, wins (measured on the block-size sweep, 1728 block
    // rows: 4x5 fp32 with 20736 X blocks 10.6 vs 1.3 ms per solve, 4x32 with 3456 blocks 2.25 vs 1.79 ms; the FD example,
    // 171 blocks: 1.06 vs 2.63 ms).  TFQMRGPU_RESIDENT_MAX_TILE overrides (dev).
    {
        char const *const e_mt = std::getenv("TFQMRGPU_RESIDENT_MAX_TILE");
        size_t const max_tile = e_mt ? size_t(std::max(1, std::atoi(e_mt))) : 4;
        if (tile_blocks > max_tile) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
    int max_row = 1;                              // entries of a Y block <= blocks in its row of A
    for (int r = 0; r < p.mb; ++r) max_row = std::max(max_row, int(p.h_rpA[r + 1] - p.h_rpA[r]));
    int const max_e = std::min(max_row, 64);
    int W = 1;                                    // warps per Y block: as many as the CTA's warps allow for the largest tile
    while (2*W <= kResWarps && size_t(2*W)*tile_blocks <= size_t(kResWarps)) W *= 2;
    // all tiles co-resident: k CTAs per SM, nsm*k >= nTiles; within that budget keep A in shared memory if it fits with at
    // least 4 staged entries per warp, and stage as many entries per batch as fit (at most a warp's share of a row)
    int const k = int((nTiles + nsm - 1)/nsm);
    size_t const budget = (size_t(220)*1024)/size_t(k);
    int const share = (max_row + W - 1)/W;
    bool a_res = (max_row <= max_e) && resident_smem<real_t, LM, LN>(tile_blocks, max_e, true, std::min(4, share)) <= budget;
    char const *const e_ares = std::getenv("TFQMRGPU_RESIDENT_A");           // dev switch: 0 = stage A with X
    if (e_ares && '0' == e_ares[0]) a_res = false;
    int eb = std::min(share, 32);
    while (eb > 1 && resident_smem<real_t, LM, LN>(tile_blocks, max_e, a_res, eb) > budget) --eb;
    size_t const smem = resident_smem<real_t, LM, LN>(tile_blocks, max_e, a_res, eb);
    if (smem > budget) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    static size_t configured[kMaxDevices] = {0};
    TFQ_CUDA(ensure_dynamic_smem(kernel, smem, configured));
    int per_sm = 0;
    TFQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kResThreads, smem));
    if (size_t(per_sm)*nsm < nTiles) return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    if (dry) return TFQMRGPU_STATUS_SUCCESS;

    if (nullptr == p.d_resident_bar) TFQ_CUDA(cudaMalloc((void**)&p.d_resident_bar, 2*sizeof(unsigned)));
    TFQ_CUDA(cudaMemsetAsync(p.d_resident_bar, 0, 2*sizeof(unsigned), stream));
    if (nullptr == p.d_unit_of_block) {          // once per configured plan: tiles, partial-sum scratch, block -> unit
        TFQ_CUDA(cudaMalloc((void**)&p.d_res_tiles, rt.tiles.size()*sizeof(Tile)));
        TFQ_CUDA(cudaMalloc((void**)&p.d_res_coltile, rt.coltile.size()*sizeof(uint32_t)));
        TFQ_CUDA(cudaMalloc((void**)&p.d_res_part, size_t(nTiles)*kPartD*LN*sizeof(double)));
        TFQ_CUDA(cudaMemcpyAsync(p.d_res_tiles, rt.tiles.data(), rt.tiles.size()*sizeof(Tile), cudaMemcpyHostToDevice, stream));
        TFQ_CUDA(cudaMemcpyAsync(p.d_res_coltile, rt.coltile.data(), rt.coltile.size()*sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        TFQ_CUDA(cudaStreamSynchronize(stream));                 // (the host vectors are locals)
        TFQ_CUDA(cudaMalloc((void**)&p.d_unit_of_block, std::max<size_t>(p.nnzbX, 1)*sizeof(uint32_t)));
        invert_units_kernel<<<(p.nUnits*p.gmax + 255)/256, 256, 0, stream>>>(p.d_unit_of_block, p.d_unit_y, p.nUnits, p.gmax);
        TFQ_CUDA(cudaGetLastError());
    }

    ResidentArgs<real_t> a;
    a.v1 = ws<real_t>(p, p.off_v[1]); a.v5 = ws<real_t>(p, p.off_v[5]); a.v6 = ws<real_t>(p, p.off_v[6]);
    a.v3 = ws<float const>(p, p.off_v[3]);
    a.A = ws<real_t const>(p, p.off_A); a.B = ws<real_t const>(p, p.off_B);
    a.bpos = p.d_bpos; a.nnzbB = uint32_t(std::max(p.nnzbB, 0));
    a.rho = ws<real_t const>(p, p.off_rho); a.beta = ws<real_t const>(p, p.off_beta);
    a.tau = ws<double const>(p, p.off_tau); a.invBn2 = ws<double const>(p, p.off_invBn2);
    a.status = ws<int8_t>(p, p.off_status); a.snap = ws<int8_t>(p, p.off_snap);
    a.part = p.d_res_part; a.colmon = ws<double>(p, p.off_colmon);
    a.ctl = ws<Control>(p, p.off_ctl);
    a.tiles = p.d_res_tiles; a.coltile = p.d_res_coltile; a.nCols = p.nCols;
    a.unit_e0 = p.d_unit_e0; a.ent_a = p.d_ent_a; a.ent_x = p.d_ent_x; a.unit_of_block = p.d_unit_of_block; a.gmax = p.gmax;
    a.bar = p.d_resident_bar;
    a.tile_elems = uint32_t(tile_blocks*2*size_t(LM)*LN); a.tile_blocks = uint32_t(tile_blocks); a.eb = eb;
    a.max_e = max_e; a.a_resident = a_res ? 1 : 0; a.warps_per_block = W;
    a.trace = nullptr;
    { char const *const e_abl = std::getenv("TFQMRGPU_RESIDENT_ABLATE"); a.ablate = e_abl ? std::atoi(e_abl) : 0; }
    char const *const e_trace = std::getenv("TFQMRGPU_RESIDENT_TRACE");       // dev: where CTA 0 spends its time
    if (e_trace && '0' != e_trace[0]) {
        if (nullptr == p.d_resident_trace) TFQ_CUDA(cudaMalloc((void**)&p.d_resident_trace, 8*sizeof(unsigned long long)));
        TFQ_CUDA(cudaMemsetAsync(p.d_resident_trace, 0, 8*sizeof(unsigned long long), stream));
        a.trace = p.d_resident_trace;
    }
    void *args[] = { &a };
    TFQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(kernel), dim3(nTiles), dim3(kResThreads), args, smem, stream));
    if (a.trace) {
        unsigned long long h[8];
        TFQ_CUDA(cudaMemcpyAsync(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost, stream));
        TFQ_CUDA(cudaStreamSynchronize(stream));
        std::printf("# resident: %u CTAs of %zu block(s), %d warp(s) per Y block, %zu B shared, eb %d, A resident %d; %llu iterations, %llu probes; CTA 0: total %.1f us, "
                    "barriers %.1f, products %.1f (staging %.1f, until the sums are done %.1f), column sums incl. their barrier %.1f\n",
                    nTiles, tile_blocks, W, smem, eb, int(a_res), h[6], h[7], h[3]*1e-3, h[0]*1e-3, h[1]*1e-3, h[4]*1e-3, h[5]*1e-3, h[2]*1e-3);
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t resident_dispatch(Plan &p, cudaStream_t stream, bool dry)
{
    bool const z = ('z' == p.precision);
    switch (p.LM*1000 + p.LN) {
#define TFQ_CASE(LM, LN) case LM*1000 + LN: return z ? launch_resident<double, LM, LN>(p, stream, dry) : launch_resident<float, LM, LN>(p, stream, dry);
        TFQ_CASE(4, 4) TFQ_CASE(4, 5) TFQ_CASE(4, 8) TFQ_CASE(4, 32)
        TFQ_CASE(8, 8) TFQ_CASE(8, 9) TFQ_CASE(8, 10) TFQ_CASE(8, 32) TFQ_CASE(8, 64)
#undef TFQ_CASE
        default: return TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
}

} // namespace

// Which plans run resident: blocks with LM <= 8 (the SIMT product's sizes), at most 4 X blocks per SM (see launch_resident), all tiles
// co-resident with their seven vectors in shared memory, no user-defined operator, no shard exchange, no per-product profiling.
// TFQMRGPU_RESIDENT=0 switches it off.
bool resident_supported(Plan &p)
{
    // (read per solve: a getenv call is nothing next to a launch, and tests and callers can switch between solves)
    char const *const e_on = std::getenv("TFQMRGPU_RESIDENT");
    if (e_on && '0' == e_on[0]) return false;
    if (p.use_tc16 || p.use_dmma || p.user_op || p.exch.slots || p.profile || p.multi) return false;
    if ('z' != p.precision && 'c' != p.precision) return false;
    if (p.LM > 8) return false;
    // the structural part of the answer (tiling, shared memory, occupancy) is the same for every solve of a configured plan
    if (p.resident_fits < 0 || std::getenv("TFQMRGPU_RESIDENT_MAX_TILE") || std::getenv("TFQMRGPU_RESIDENT_TILES_PER_SM"))
        p.resident_fits = (TFQMRGPU_STATUS_SUCCESS == resident_dispatch(p, nullptr, true)) ? 1 : 0;
    return 1 == p.resident_fits;
}

tfqmrgpuStatus_t launch_resident_solve(Plan &p, cudaStream_t stream, int /*maxIterations*/)
{
    return resident_dispatch(p, stream, false);
}

} // namespace tfq
