// Device code shared by the per-kernel launches (vecops.cu) and the resident solver (resident.cu): the arguments of the fused
// tfQMR vector kernels, the scalar recurrences, the reference's host decisions restated for the device, and the body of one
// vector-kernel tile.  See vecops.cu for the overview.
#pragma once
#include "tfq_internal.hpp"
#include <cuda_fp16.h>
#include <type_traits>
#include <algorithm>

namespace tfq {

namespace {

constexpr double kEpsilon = 2.5e-308; // linalg.hxx:31
// Early-freeze extension (not in the reference): status of a right-hand side whose true residual passed a probe;
// its X is kept from then on (eta = 0 like after a breakdown) while the other right-hand sides continue.
constexpr int8_t kFrozen = 2;

template <typename real_t> struct VecArgs {
    real_t *v1, *v4, *v5, *v6, *v7, *v8, *v9;
    float const *v3;
    real_t *rho, *alfa, *beta, *c67, *eta;
    double *tau, *var, *invBn2;
    int8_t *status, *snap;
    double *part, *colmon;
    unsigned *ticket;
    Control *ctl;
    Tile const *tiles;
    uint32_t const *coltile;
    uint32_t nCols;
    int LM, LN, lmShift;
    // fp16-pair operand of the tensor-core product (xop.cu, spmm_tc16.cu): per-column maxima of |v4|, |v5|, |v6| kept by the kernels
    // that write those vectors, tile maxima scratch, and the operand itself with its column scales (written by K1 / K3)
    float *mx4, *mx5, *mx6, *partmax, *xs, *xsinv;
    uint4 *xop;
    double *slot_out;         // column-sharded runs: where K4 / N3 export (max, count, count) instead of deciding
};

template <typename T, int N> struct alignas(sizeof(T)*N) Vec { T v[N]; };

template <typename T, int N> __device__ __forceinline__ Vec<T, N> ldv(T const *p) { return *reinterpret_cast<Vec<T, N> const*>(p); }
template <typename T, int N> __device__ __forceinline__ void stv(T *p, Vec<T, N> const &x) { *reinterpret_cast<Vec<T, N>*>(p) = x; }

// ---- scalar recurrences, one (block column c, lane j) each ---------------------------------------
// tfQMRdec35: linalg.hxx:50-75
template <typename real_t>
__device__ __forceinline__ void dec35(VecArgs<real_t> const &a, uint32_t c, int j, double z_Re, double z_Im) {
    size_t const r = (size_t(c)*2 + 0)*a.LN + j, m = (size_t(c)*2 + 1)*a.LN + j;
    double const rho_Re = double(a.rho[r]), rho_Im = double(a.rho[m]);
    double const abs2rho = rho_Re*rho_Re + rho_Im*rho_Im;
    double const abs2z = z_Re*z_Re + z_Im*z_Im;
    if ((abs2z < kEpsilon) || (abs2rho < kEpsilon)) {
        if (kFrozen != a.status[size_t(c)*a.LN + j]) a.status[size_t(c)*a.LN + j] = -1;
        a.beta[r] = 0; a.beta[m] = 0; a.rho[r] = 0; a.rho[m] = 0;
    } else {
        double const den = 1./abs2rho;
        a.beta[r] = real_t((z_Re*rho_Re + z_Im*rho_Im)*den);
        a.beta[m] = real_t((z_Im*rho_Re - z_Re*rho_Im)*den);
        a.rho[r] = real_t(z_Re); a.rho[m] = real_t(z_Im);
    }
}
// tfQMRdec34: linalg.hxx:116-151
template <typename real_t>
__device__ __forceinline__ void dec34(VecArgs<real_t> const &a, uint32_t c, int j, double z_Re, double z_Im) {
    size_t const r = (size_t(c)*2 + 0)*a.LN + j, m = (size_t(c)*2 + 1)*a.LN + j;
    double const rho_Re = double(a.rho[r]), rho_Im = double(a.rho[m]);
    double const abs2rho = rho_Re*rho_Re + rho_Im*rho_Im;
    double const abs2z = z_Re*z_Re + z_Im*z_Im;
    if ((abs2z < kEpsilon) || (abs2rho < kEpsilon)) {
        if (kFrozen != a.status[size_t(c)*a.LN + j]) a.status[size_t(c)*a.LN + j] = -2;
        a.alfa[r] = 0; a.alfa[m] = 0; a.c67[r] = 0; a.c67[m] = 0;
    } else {
        double const eta_Re = double(a.eta[r]), eta_Im = double(a.eta[m]);
        double const zden = -1./abs2z;
        a.alfa[r] = real_t((rho_Re*z_Re + rho_Im*z_Im)*zden);
        a.alfa[m] = real_t((rho_Im*z_Re - rho_Re*z_Im)*zden);
        double const vden = a.var[size_t(c)*a.LN + j]/abs2rho;
        double const t_Re = (eta_Re*rho_Re + eta_Im*rho_Im)*vden;
        double const t_Im = (eta_Im*rho_Re - eta_Re*rho_Im)*vden;
        a.c67[r] = real_t(z_Re*t_Re - z_Im*t_Im);
        a.c67[m] = real_t(z_Im*t_Re + z_Re*t_Im);
    }
}
// tfQMRdecT: linalg.hxx:195-229
template <typename real_t>
__device__ __forceinline__ void decT(VecArgs<real_t> const &a, uint32_t c, int j, double D55, bool with_c67) {
    size_t const r = (size_t(c)*2 + 0)*a.LN + j, m = (size_t(c)*2 + 1)*a.LN + j, s = size_t(c)*a.LN + j;
    double cosi = 0;
    real_t r67 = 1;
    double const Tau = a.tau[s];
    if (fabs(Tau) > kEpsilon) {
        double const Var = D55/Tau;
        cosi = 1./(1. + Var);
        a.var[s] = Var;
        a.tau[s] = D55*cosi;
        r67 = real_t(Var*cosi);
    } else {
        if (kFrozen != a.status[s]) a.status[s] = -3;
        a.var[s] = 0; a.tau[s] = 0;
    }
    if (a.status[s] < 0 || kFrozen == a.status[s]) { a.eta[r] = 0; a.eta[m] = 0; }   // (frozen: only with the early-freeze extension)
    else { a.eta[r] = real_t(-cosi*double(a.alfa[r])); a.eta[m] = real_t(-cosi*double(a.alfa[m])); }
    if (with_c67) { a.c67[r] = r67; a.c67[m] = 0; }
}

// ---- the reference's host logic after an iteration / after a residual probe, evaluated on the device -------------
// m0 = max over right-hand sides of tau/|b|^2, m1 / m2 = right-hand sides with status -2 / -1 (core.hxx:239-260)
__device__ __forceinline__ void decide_iteration(Control &ctl, double m0, double m1, double m2, long long nRHS) {
    int const it = ctl.iteration + 1;
    double const max_bound2 = m0*(2*it + 1);                        // core.hxx:252
    bool const probe = (max_bound2 <= ctl.target_bound2) || (it >= ctl.max_iterations); // core.hxx:254
    ctl.iteration = it;
    ctl.max_bound2 = max_bound2;
    if (nRHS == (long long)(m1 + m2)) {                             // core.hxx:255-260
        ctl.result = TFQMRGPU_STATUS_BREAKDOWN;
        ctl.state = STATE_DONE;
    } else {
        ctl.state = probe ? STATE_PROBE : STATE_RUN;
    }
}
// m0 = max over right-hand sides of the true relative residual^2, m1 = right-hand sides that are not done (core.hxx:274-298)
__device__ __forceinline__ void decide_probe(Control &ctl, double m0, double m1) {
    double max_res2 = 1.4e-76;                                      // core.hxx:274
    max_res2 = (max_res2 < m0) ? m0 : max_res2;
    ctl.residual2_reached = max_res2;                               // core.hxx:287
    ctl.target_bound2 = (ctl.max_bound2/max_res2)*ctl.tol2;        // core.hxx:290
    ctl.probes += 1;
    if (0 == m1) {                                                  // isDone, core.hxx:294-298
        ctl.iterations_needed = ctl.iteration;
        ctl.result = TFQMRGPU_STATUS_SUCCESS;
        ctl.state = STATE_DONE;
    } else {
        ctl.state = (ctl.iteration < ctl.max_iterations) ? STATE_RUN : STATE_DONE;
    }
}

template <int OP> struct OpTraits;
template <> struct OpTraits<OP_INIT> { static constexpr int D = 3, expect = -1; };
template <> struct OpTraits<OP_K1>   { static constexpr int D = 0, expect = STATE_RUN; };
template <> struct OpTraits<OP_E1>   { static constexpr int D = 2, expect = STATE_RUN; };
template <> struct OpTraits<OP_K2>   { static constexpr int D = 1, expect = STATE_RUN; };
template <> struct OpTraits<OP_K3>   { static constexpr int D = 0, expect = STATE_RUN; };
template <> struct OpTraits<OP_E2>   { static constexpr int D = 1, expect = STATE_RUN; };
template <> struct OpTraits<OP_K4>   { static constexpr int D = 2, expect = STATE_RUN; };
template <> struct OpTraits<OP_N3>   { static constexpr int D = 1, expect = STATE_PROBE; };

// XM: also keep the per-column maximum magnitude of the vector this kernel writes (INIT, K2, E2: v5; E1: v4) - the
// operand-emitting K1 / K3 (vec_xop_kernel) bound the magnitude of their result with it
// One tile of one vector kernel.  tileIndex: the CTA index of the per-kernel launches (vecops.cu); the resident solver (resident.cu)
// walks its CTAs over the tiles instead.  red: dynamic shared memory, sized by vec_smem_bytes().
template <typename real_t, int VEC, int OP, bool XM>
__device__ __forceinline__ void vec_tile(VecArgs<real_t> const &a, uint32_t const tileIndex, double *const red)
{
    constexpr int D = OpTraits<OP>::D;
    constexpr int DD = (D > 0) ? D : 1;
    __shared__ int s_flag;

    Tile const t = a.tiles[tileIndex];
    uint32_t const c = t.col;
    int const LM = a.LM, LN = a.LN;
    int const tid = threadIdx.x;
    int const LNV = LN/VEC;
    int const jv = tid % LNV, r0 = tid / LNV, rstep = int(blockDim.x)/LNV;
    int const j0 = jv*VEC;
    size_t const plane = size_t(LM)*LN;

    // CTA-uniform coefficients of this block column: a[c][Re|Im][j0 .. j0+VEC)
    auto coef = [&](real_t const *s, real_t (&re)[VEC], real_t (&im)[VEC]) {
        #pragma unroll
        for (int v = 0; v < VEC; ++v) { re[v] = s[(size_t(c)*2 + 0)*LN + j0 + v]; im[v] = s[(size_t(c)*2 + 1)*LN + j0 + v]; }
    };
    real_t pr[VEC], pi[VEC], qr[VEC], qi[VEC], sr[VEC], si[VEC];
    (void)pr; (void)pi; (void)qr; (void)qi; (void)sr; (void)si;
    if (OP == OP_K1 || OP == OP_E1) coef(a.beta, pr, pi);
    if (OP == OP_K2) { coef(a.c67, pr, pi); coef(a.alfa, qr, qi); }
    if (OP == OP_K3) { coef(a.eta, pr, pi); coef(a.alfa, qr, qi); coef(a.c67, sr, si); }
    if (OP == OP_E2) coef(a.alfa, pr, pi);
    if (OP == OP_K4) coef(a.eta, pr, pi);

    double acc[DD][VEC];
    float vmax[VEC];
    #pragma unroll
    for (int d = 0; d < DD; ++d) {
        #pragma unroll
        for (int v = 0; v < VEC; ++v) acc[d][v] = 0;
    }
    #pragma unroll
    for (int v = 0; v < VEC; ++v) vmax[v] = 0.f;
    (void)vmax;

    int const nrows = int(t.b1 - t.b0) << a.lmShift;
    #pragma unroll 2
    for (int rr = r0; rr < nrows; rr += rstep) {
        uint32_t const blk = t.b0 + uint32_t(rr >> a.lmShift);
        int const i = rr & (LM - 1);
        size_t const ore = (size_t(blk)*2*LM + i)*LN + j0, oim = ore + plane;
        using V = Vec<real_t, VEC>;
        using F = Vec<float, VEC>;

        if (OP == OP_INIT) {
            // tau = |b|^2 (core.hxx:154-155) and the first z35 = v3.v5 (core.hxx:189)
            V const xr = ldv<real_t, VEC>(a.v5 + ore), xi = ldv<real_t, VEC>(a.v5 + oim);
            F const wr = ldv<float, VEC>(a.v3 + ore), wi = ldv<float, VEC>(a.v3 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc[0][v] += double(xr.v[v])*double(xr.v[v]) + double(xi.v[v])*double(xi.v[v]);
                if (XM) vmax[v] = fmaxf(vmax[v], fmaxf(fabsf(float(xr.v[v])), fabsf(float(xi.v[v]))));
                real_t const tr = xr.v[v]*wr.v[v] - xi.v[v]*wi.v[v];
                real_t const ti = xr.v[v]*wi.v[v] + xi.v[v]*wr.v[v];
                acc[1 % DD][v] += double(tr); acc[2 % DD][v] += double(ti);
            }
        }
        if (OP == OP_K1) { // v6 := v5 + beta*v6   (core.hxx:194, linalg.hxx:660-661)
            V const xr = ldv<real_t, VEC>(a.v5 + ore), xi = ldv<real_t, VEC>(a.v5 + oim);
            V yr = ldv<real_t, VEC>(a.v6 + ore), yi = ldv<real_t, VEC>(a.v6 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                real_t const nr = xr.v[v] + pr[v]*yr.v[v] - pi[v]*yi.v[v];
                real_t const ni = xi.v[v] + pi[v]*yr.v[v] + pr[v]*yi.v[v];
                yr.v[v] = nr; yi.v[v] = ni;
            }
            stv<real_t, VEC>(a.v6 + ore, yr); stv<real_t, VEC>(a.v6 + oim, yi);
        }
        if (OP == OP_E1) { // v4 := v8 + beta*v4 ; v4 := v9 + beta*v4 ; z34 += v4.v3  (core.hxx:196,200,202)
            V const ar = ldv<real_t, VEC>(a.v8 + ore), ai = ldv<real_t, VEC>(a.v8 + oim);
            V const br = ldv<real_t, VEC>(a.v9 + ore), bi = ldv<real_t, VEC>(a.v9 + oim);
            V yr = ldv<real_t, VEC>(a.v4 + ore), yi = ldv<real_t, VEC>(a.v4 + oim);
            F const wr = ldv<float, VEC>(a.v3 + ore), wi = ldv<float, VEC>(a.v3 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                real_t const tr = ar.v[v] + pr[v]*yr.v[v] - pi[v]*yi.v[v];
                real_t const ti = ai.v[v] + pi[v]*yr.v[v] + pr[v]*yi.v[v];
                real_t const nr = br.v[v] + pr[v]*tr - pi[v]*ti;
                real_t const ni = bi.v[v] + pi[v]*tr + pr[v]*ti;
                yr.v[v] = nr; yi.v[v] = ni;
                if (XM) vmax[v] = fmaxf(vmax[v], fmaxf(fabsf(float(nr)), fabsf(float(ni))));
                real_t const dr = nr*wr.v[v] - ni*wi.v[v]; // linalg.hxx:506-507, products in real_t x float
                real_t const di = nr*wi.v[v] + ni*wr.v[v];
                acc[0][v] += double(dr); acc[1 % DD][v] += double(di);
            }
            stv<real_t, VEC>(a.v4 + ore, yr); stv<real_t, VEC>(a.v4 + oim, yi);
        }
        if (OP == OP_K2) { // v7 := v6 + c67*v7 ; v5 := alfa*v9 + v5 ; d55 += |v5|^2  (core.hxx:207-211)
            V const ar = ldv<real_t, VEC>(a.v6 + ore), ai = ldv<real_t, VEC>(a.v6 + oim);
            V yr = ldv<real_t, VEC>(a.v7 + ore), yi = ldv<real_t, VEC>(a.v7 + oim);
            V const br = ldv<real_t, VEC>(a.v9 + ore), bi = ldv<real_t, VEC>(a.v9 + oim);
            V zr = ldv<real_t, VEC>(a.v5 + ore), zi = ldv<real_t, VEC>(a.v5 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                real_t const nr = ar.v[v] + pr[v]*yr.v[v] - pi[v]*yi.v[v];
                real_t const ni = ai.v[v] + pi[v]*yr.v[v] + pr[v]*yi.v[v];
                yr.v[v] = nr; yi.v[v] = ni;
                real_t const mr = qr[v]*br.v[v] - qi[v]*bi.v[v] + zr.v[v]; // linalg.hxx:656-657
                real_t const mi = qi[v]*br.v[v] + qr[v]*bi.v[v] + zi.v[v];
                zr.v[v] = mr; zi.v[v] = mi;
                if (XM) vmax[v] = fmaxf(vmax[v], fmaxf(fabsf(float(mr)), fabsf(float(mi))));
                acc[0][v] += double(mr)*double(mr) + double(mi)*double(mi);
            }
            stv<real_t, VEC>(a.v7 + ore, yr); stv<real_t, VEC>(a.v7 + oim, yi);
            stv<real_t, VEC>(a.v5 + ore, zr); stv<real_t, VEC>(a.v5 + oim, zi);
        }
        if (OP == OP_K3) { // v1 += eta*v7 ; v6 += alfa*v4 ; v7 := v6 + c67*v7  (core.hxx:216-220)
            V yr = ldv<real_t, VEC>(a.v7 + ore), yi = ldv<real_t, VEC>(a.v7 + oim);
            V xr = ldv<real_t, VEC>(a.v1 + ore), xi = ldv<real_t, VEC>(a.v1 + oim);
            V const br = ldv<real_t, VEC>(a.v4 + ore), bi = ldv<real_t, VEC>(a.v4 + oim);
            V zr = ldv<real_t, VEC>(a.v6 + ore), zi = ldv<real_t, VEC>(a.v6 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                xr.v[v] = pr[v]*yr.v[v] - pi[v]*yi.v[v] + xr.v[v];
                xi.v[v] = pi[v]*yr.v[v] + pr[v]*yi.v[v] + xi.v[v];
                real_t const mr = qr[v]*br.v[v] - qi[v]*bi.v[v] + zr.v[v];
                real_t const mi = qi[v]*br.v[v] + qr[v]*bi.v[v] + zi.v[v];
                zr.v[v] = mr; zi.v[v] = mi;
                real_t const nr = mr + sr[v]*yr.v[v] - si[v]*yi.v[v];
                real_t const ni = mi + si[v]*yr.v[v] + sr[v]*yi.v[v];
                yr.v[v] = nr; yi.v[v] = ni;
            }
            stv<real_t, VEC>(a.v1 + ore, xr); stv<real_t, VEC>(a.v1 + oim, xi);
            stv<real_t, VEC>(a.v6 + ore, zr); stv<real_t, VEC>(a.v6 + oim, zi);
            stv<real_t, VEC>(a.v7 + ore, yr); stv<real_t, VEC>(a.v7 + oim, yi);
        }
        if (OP == OP_E2) { // v5 := alfa*v8 + v5 ; d55 += |v5|^2  (core.hxx:226-228)
            V const br = ldv<real_t, VEC>(a.v8 + ore), bi = ldv<real_t, VEC>(a.v8 + oim);
            V zr = ldv<real_t, VEC>(a.v5 + ore), zi = ldv<real_t, VEC>(a.v5 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                real_t const mr = pr[v]*br.v[v] - pi[v]*bi.v[v] + zr.v[v];
                real_t const mi = pi[v]*br.v[v] + pr[v]*bi.v[v] + zi.v[v];
                zr.v[v] = mr; zi.v[v] = mi;
                if (XM) vmax[v] = fmaxf(vmax[v], fmaxf(fabsf(float(mr)), fabsf(float(mi))));
                acc[0][v] += double(mr)*double(mr) + double(mi)*double(mi);
            }
            stv<real_t, VEC>(a.v5 + ore, zr); stv<real_t, VEC>(a.v5 + oim, zi);
        }
        if (OP == OP_K4) { // v1 += eta*v7 (core.hxx:233) ; next z35 += v5.v3 (core.hxx:189)
            V const yr = ldv<real_t, VEC>(a.v7 + ore), yi = ldv<real_t, VEC>(a.v7 + oim);
            V xr = ldv<real_t, VEC>(a.v1 + ore), xi = ldv<real_t, VEC>(a.v1 + oim);
            V const zr = ldv<real_t, VEC>(a.v5 + ore), zi = ldv<real_t, VEC>(a.v5 + oim);
            F const wr = ldv<float, VEC>(a.v3 + ore), wi = ldv<float, VEC>(a.v3 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) {
                xr.v[v] = pr[v]*yr.v[v] - pi[v]*yi.v[v] + xr.v[v];
                xi.v[v] = pi[v]*yr.v[v] + pr[v]*yi.v[v] + xi.v[v];
                real_t const dr = zr.v[v]*wr.v[v] - zi.v[v]*wi.v[v];
                real_t const di = zr.v[v]*wi.v[v] + zi.v[v]*wr.v[v];
                acc[0][v] += double(dr); acc[1 % DD][v] += double(di);
            }
            stv<real_t, VEC>(a.v1 + ore, xr); stv<real_t, VEC>(a.v1 + oim, xi);
        }
        if (OP == OP_N3) { // |v9|^2 with v9 = A*v1 - b  (core.hxx:265-269)
            V const xr = ldv<real_t, VEC>(a.v9 + ore), xi = ldv<real_t, VEC>(a.v9 + oim);
            #pragma unroll
            for (int v = 0; v < VEC; ++v) acc[0][v] += double(xr.v[v])*double(xr.v[v]) + double(xi.v[v])*double(xi.v[v]);
        }
    }

    if (D == 0) return;

    // ---- tile-level reduction over the threads that share a lane j (fixed order) -------------------
    #pragma unroll
    for (int d = 0; d < DD; ++d) {
        #pragma unroll
        for (int v = 0; v < VEC; ++v) red[(size_t(r0)*DD + d)*LN + j0 + v] = acc[d][v];
    }
    __syncthreads();
    int half = 1; while (half < rstep) half <<= 1;
    for (half >>= 1; half > 0; half >>= 1) {
        if (r0 < half && r0 + half < rstep) {
            #pragma unroll
            for (int d = 0; d < DD; ++d) {
                #pragma unroll
                for (int v = 0; v < VEC; ++v)
                    red[(size_t(r0)*DD + d)*LN + j0 + v] += red[(size_t(r0 + half)*DD + d)*LN + j0 + v];
            }
        }
        __syncthreads();
    }
    if (0 == r0 && tid < LNV) {
        #pragma unroll
        for (int d = 0; d < DD; ++d) {
            #pragma unroll
            for (int v = 0; v < VEC; ++v) a.part[(size_t(tileIndex)*kPartD + d)*LN + j0 + v] = red[size_t(d)*LN + j0 + v];
        }
    }

    if (XM) {   // tile maximum per lane j (the order of a maximum does not matter)
        float *const fr = reinterpret_cast<float*>(red);
        __syncthreads();
        #pragma unroll
        for (int v = 0; v < VEC; ++v) fr[r0*LN + j0 + v] = vmax[v];
        __syncthreads();
        int hm = 1; while (hm < rstep) hm <<= 1;
        for (hm >>= 1; hm > 0; hm >>= 1) {
            if (r0 < hm && r0 + hm < rstep) {
                #pragma unroll
                for (int v = 0; v < VEC; ++v) fr[r0*LN + j0 + v] = fmaxf(fr[r0*LN + j0 + v], fr[(r0 + hm)*LN + j0 + v]);
            }
            __syncthreads();
        }
        if (0 == r0 && tid < LNV) {
            #pragma unroll
            for (int v = 0; v < VEC; ++v) a.partmax[size_t(tileIndex)*64 + j0 + v] = fr[j0 + v];
        }
    }

    // ---- the last tile of this block column finishes the column ------------------------------------
    uint32_t const t0 = a.coltile[c], t1 = a.coltile[c + 1];
    __threadfence();
    __syncthreads();
    if (0 == tid) s_flag = (atomicAdd(&a.ticket[c], 1u) == (t1 - t0) - 1u);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (0 == tid) a.ticket[c] = 0;

    if (XM && tid < LN) {
        float m = 0.f;
        for (uint32_t tt = t0; tt < t1; ++tt) m = fmaxf(m, __ldcg(&a.partmax[size_t(tt)*64 + tid]));
        float *const mx = (OP == OP_E1) ? a.mx4 : a.mx5;
        mx[size_t(c)*LN + tid] = m;
    }
    int const nq = DD*LN;
    int const nsl = int(blockDim.x)/nq;
    {
        int const q = tid % nq, sl = tid / nq;
        double s = 0;
        if (sl < nsl) {
            int const d = q / LN, j = q - d*LN;
            for (uint32_t tt = t0 + sl; tt < t1; tt += nsl) s += __ldcg(&a.part[(size_t(tt)*kPartD + d)*LN + j]);
        }
        __syncthreads();
        if (sl < nsl) red[sl*nq + q] = s;
        __syncthreads();
        if (0 == sl) {
            for (int s2 = 1; s2 < nsl; ++s2) s += red[s2*nq + q];
        }
        __syncthreads();
        if (0 == sl) red[q] = s; // red[d*LN + j] = column sum
        __syncthreads();
    }

    int const j = tid;
    double *const mon = red + nq; // scratch behind the sums: [LN] values + flags
    if (OP == OP_INIT) {
        if (j < LN) {
            size_t const s = size_t(c)*LN + j, r = (size_t(c)*2 + 0)*LN + j, m = (size_t(c)*2 + 1)*LN + j;
            double const d = red[j];
            a.tau[s] = d;                 // core.hxx:155
            a.invBn2[s] = 1./d;           // core.hxx:165
            a.var[s] = 0; a.status[s] = 0; a.snap[s] = 0;      // core.hxx:123,127
            a.eta[r] = 0; a.eta[m] = 0;   // core.hxx:121
            a.rho[r] = 1; a.rho[m] = 0;   // core.hxx:122
            dec35(a, c, j, red[LN + j], red[2*LN + j]);
        }
        return;
    }
    if (OP == OP_E1) { if (j < LN) dec34(a, c, j, red[j], red[LN + j]); return; }
    if (OP == OP_K2) { if (j < LN) decT(a, c, j, red[j], true);  return; }
    if (OP == OP_E2) { if (j < LN) decT(a, c, j, red[j], false); return; }

    if (OP == OP_K4) {
        // convergence monitor on the state after this iteration (core.hxx:235-247), BEFORE dec35 touches status
        if (j < LN) {
            size_t const s = size_t(c)*LN + j;
            int8_t const st = a.status[s];
            a.snap[s] = st;
            mon[j] = a.tau[s]*a.invBn2[s];
            mon[LN + j] = double(st);
        }
        __syncthreads();
        if (0 == tid) {
            double mx = 0, b4 = 0, b5 = 0;
            for (int q = 0; q < LN; ++q) {
                double const res2 = mon[q];
                mx = (mx < res2) ? res2 : mx;            // std::max semantics (NaN never wins)
                b4 += (-2. == mon[LN + q]); b5 += (-1. == mon[LN + q]);
            }
            a.colmon[size_t(c)*4 + 0] = mx; a.colmon[size_t(c)*4 + 1] = b4; a.colmon[size_t(c)*4 + 2] = b5;
        }
        if (j < LN) dec35(a, c, j, red[j], red[LN + j]); // next iteration's beta, rho (core.hxx:192)
    }
    if (OP == OP_N3) {
        // probe: relative residual per right-hand side (core.hxx:276-286)
        double const tol2 = a.ctl->tol2;
        if (j < LN) {
            size_t const s = size_t(c)*LN + j;
            double const res2 = red[j]*a.invBn2[s];
            double notdone = 0;
            if (res2 > tol2) { if (0 == a.snap[s]) notdone = 1; }
            else if (res2 <= 0) {           // core.hxx:282-285: the host marks the component as converged
                // the fused dec35 of the NEXT iteration has already run; keep its verdict if it changed the status
                if (a.status[s] == a.snap[s]) a.status[s] = 1;
                a.snap[s] = 1;              // snap = the reference's status_h (what getRhsStatus reports)
            } else if (a.ctl->freeze && 0 == a.snap[s] && 0 == a.status[s]) {
                a.status[s] = kFrozen; a.snap[s] = kFrozen;     // early-freeze extension: this right-hand side is done
            }
            mon[j] = res2; mon[LN + j] = notdone;
        }
        __syncthreads();
        if (0 == tid) {
            double mx = 0, nd = 0;
            for (int q = 0; q < LN; ++q) { double const res2 = mon[q]; mx = (mx < res2) ? res2 : mx; nd += mon[LN + q]; }
            a.colmon[size_t(c)*4 + 0] = mx; a.colmon[size_t(c)*4 + 1] = nd;
        }
    }

    // ---- the last block column evaluates the reference's host logic on the device --------------------
    __threadfence();
    __syncthreads();
    if (0 == tid) s_flag = (atomicAdd(&a.ctl->cols_done, 1u) == a.nCols - 1u);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double m0 = 0, m1 = 0, m2 = 0;
    for (uint32_t cc = tid; cc < a.nCols; cc += blockDim.x) {
        double const x0 = __ldcg(&a.colmon[size_t(cc)*4 + 0]);
        m0 = (m0 < x0) ? x0 : m0;
        m1 += __ldcg(&a.colmon[size_t(cc)*4 + 1]);
        m2 += __ldcg(&a.colmon[size_t(cc)*4 + 2]);
    }
    // block reduction (max, sum, sum) in a fixed order: shared memory, then the first warp (always complete: blockDim >= 224)
    // strides over the entries and finishes with shuffles.  (A serial loop of one thread over blockDim entries cost ~10 us.)
    __syncthreads();
    red[tid] = m0; red[blockDim.x + tid] = m1; red[2*blockDim.x + tid] = m2;
    __syncthreads();
    if (tid >= 32) return;
    m0 = 0; m1 = 0; m2 = 0;
    for (unsigned q = tid; q < blockDim.x; q += 32) {
        double const x0 = red[q];
        m0 = (m0 < x0) ? x0 : m0; m1 += red[blockDim.x + q]; m2 += red[2*blockDim.x + q];
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double const x0 = __shfl_down_sync(0xffffffffu, m0, off);
        m0 = (m0 < x0) ? x0 : m0;
        m1 += __shfl_down_sync(0xffffffffu, m1, off);
        m2 += __shfl_down_sync(0xffffffffu, m2, off);
    }
    if (0 == tid) {
        Control &ctl = *a.ctl;
        ctl.cols_done = 0;
        if (a.slot_out) {            // column-sharded run: export this shard's part, decide_kernel follows
            a.slot_out[0] = m0; a.slot_out[1] = m1; a.slot_out[2] = m2; a.slot_out[3] = 0;
            __threadfence_system();
        } else {
            if (OP == OP_K4) decide_iteration(ctl, m0, m1, m2, (long long)(a.nCols)*LN);
            if (OP == OP_N3) decide_probe(ctl, m0, m1);
        }
    }
}

template <typename real_t>
VecArgs<real_t> make_args(Plan const &p) {
    VecArgs<real_t> a;
    a.v1 = ws<real_t>(p, p.off_v[1]); a.v4 = ws<real_t>(p, p.off_v[4]); a.v5 = ws<real_t>(p, p.off_v[5]);
    a.v6 = ws<real_t>(p, p.off_v[6]); a.v7 = ws<real_t>(p, p.off_v[7]); a.v8 = ws<real_t>(p, p.off_v[8]);
    a.v9 = ws<real_t>(p, p.off_v[9]); a.v3 = ws<float const>(p, p.off_v[3]);
    a.rho = ws<real_t>(p, p.off_rho); a.alfa = ws<real_t>(p, p.off_alfa); a.beta = ws<real_t>(p, p.off_beta);
    a.c67 = ws<real_t>(p, p.off_c67); a.eta = ws<real_t>(p, p.off_eta);
    a.tau = ws<double>(p, p.off_tau); a.var = ws<double>(p, p.off_var); a.invBn2 = ws<double>(p, p.off_invBn2);
    a.status = ws<int8_t>(p, p.off_status); a.snap = ws<int8_t>(p, p.off_snap);
    a.part = ws<double>(p, p.off_part); a.colmon = ws<double>(p, p.off_colmon);
    a.ticket = ws<unsigned>(p, p.off_ticket); a.ctl = ws<Control>(p, p.off_ctl);
    a.tiles = p.d_tiles; a.coltile = p.d_coltile; a.nCols = p.nCols;
    a.LM = p.LM; a.LN = p.LN;
    int sh = 0; while ((1 << sh) < p.LM) ++sh;
    a.lmShift = sh;
    a.mx4 = a.mx5 = a.mx6 = a.partmax = a.xs = a.xsinv = nullptr; a.xop = nullptr;
    a.slot_out = nullptr;
    if (p.use_tc16) {
        size_t const n = size_t(p.nCols)*p.LN;
        a.mx4 = ws<float>(p, p.off_mx); a.mx5 = a.mx4 + n; a.mx6 = a.mx5 + n;
        a.partmax = ws<float>(p, p.off_xpart); a.xs = ws<float>(p, p.off_xs); a.xsinv = ws<float>(p, p.off_xsinv);
        a.xop = ws<uint4>(p, p.off_xop);
    }
    return a;
}

// dynamic shared memory of a vector kernel with `threads` threads: tile reduction rstep*D*LN doubles; the column finish needs
// >= threads + (D + 2)*LN; the monitor 3*threads
template <int OP>
inline size_t vec_smem_bytes(int threads, int LNV, int LN) {
    constexpr int D = OpTraits<OP>::D;
    if (D <= 0) return 0;
    size_t const rstep = size_t(threads/LNV);
    size_t smem = std::max<size_t>(rstep*D*LN, size_t(threads) + size_t(D)*LN + 2*size_t(LN));
    smem = std::max<size_t>(smem, 3*size_t(threads));
    return smem*sizeof(double);
}

} // namespace
} // namespace tfq
