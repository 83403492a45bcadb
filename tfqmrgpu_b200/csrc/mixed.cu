// Precision 'm': "start with float and converge double" (tfqmrgpu.h:72).
//
// The reference names this precision in its header and accepts it in bufferSize (tfqmrgpu.cu:386, "ToDo: test"), but its solver
// dispatch has the case commented out (tfqmrgpu.cu:42) and solve returns PRECISION_MISSMATCH.  Here it is an fp64-accurate solve
// whose iterations run in fp32 on the tensor cores: iterative refinement around the ordinary complex-fp32 tfQMR solver.
//
//   X = 0 (or the caller's X with tfqmrgpux_bsrsv_setInitialGuess)
//   repeat:   R = B - A*X              fp64 product (DMMA / SIMT kernel of the fp64 plan), fp64 column norms
//             stop when max_rhs |R|/|B| <= threshold
//             solve A*D = R            fp32 plan ("inner": same A and X patterns, right-hand-side pattern = X pattern since R is
//                                      X-shaped), every right-hand-side column scaled to unit size by a power of two,
//                                      inner threshold 1e-3 (less when less is missing), converged right-hand sides frozen
//             X += D                   fp64
//
// The caller hands over double data ('z' layouts) and gets doubles back; the operator is kept twice (fp64 for the residuals and as the
// fp32 operand of the inner plan, converted on the device after the upload), and the caller's single workspace holds both plans.
// Per pass the fp64 side costs one product and three streaming kernels; everything else is the fp32 solver at its own speed.
#include "tfq_internal.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>

namespace tfq {

struct MixedPlan {
    Plan  *inner = nullptr;            // complex fp32 plan
    size_t off_inner = 0;              // byte offset of its workspace inside the caller's buffer
    size_t off_part = 0, off_rn2 = 0, off_scale = 0, off_unscale = 0;   // fp64 scratch behind it
    size_t totalBytes = 0;
    int    slices = 1;                 // partial sums per block column
    double *h_rn2 = nullptr;           // pinned: column norms read back per pass
    uint32_t *d_colstart = nullptr;    // [nCols+1] first storage block of every block column
    std::vector<double> bn2;           // |b|^2 per right-hand side
    bool   a_ready = false;
    int    passes = 0;
};

namespace {

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// dst = float(src), two elements per thread
__global__ void to_float_kernel(float2 *__restrict__ dst, double2 const *__restrict__ src, size_t n2) {
    for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n2; i += size_t(gridDim.x)*blockDim.x) {
        double2 const v = src[i];
        dst[i] = make_float2(float(v.x), float(v.y));
    }
}

// Partial sums of squares per right-hand side: CTA (block column c, slice s) runs over its share of the column's blocks (they are
// contiguous in the column-sorted storage); thread (r, j) adds rows r, r + R, ... of every block, then the R row groups are added in
// a fixed order.  part[(c*S + s)*LN + j].
__global__ void colnorm_part_kernel(double *__restrict__ part, double const *__restrict__ v, uint32_t const *__restrict__ colstart,
                                    int S, int LN, int rows /* 2*LM */, int R) {
    extern __shared__ double sh[];
    uint32_t const c = blockIdx.x/S, s = blockIdx.x % S;
    uint32_t const b0 = colstart[c], n = colstart[c + 1] - b0;
    uint32_t const lo = b0 + uint32_t((uint64_t(n)*s)/S), hi = b0 + uint32_t((uint64_t(n)*(s + 1))/S);
    int const j = threadIdx.x % LN, r = threadIdx.x/LN;
    double acc = 0;
    size_t const blockElems = size_t(rows)*LN;
    for (uint32_t b = lo; b < hi; ++b) {
        double const *const blk = v + size_t(b)*blockElems;
#pragma unroll 4
        for (int i = r; i < rows; i += R) { double const x = blk[size_t(i)*LN + j]; acc += x*x; }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (0 == r) {
        double sum = 0;
        for (int q = 0; q < R; ++q) sum += sh[q*LN + j];
        part[(size_t(c)*S + s)*LN + j] = sum;
    }
}

// rn2[c][j] = sum of the slices in order; scale = the power of two that brings the column to unit size (0 for a zero column)
__global__ void colnorm_finish_kernel(double *__restrict__ rn2, double *__restrict__ scale, double *__restrict__ unscale,
                                      double const *__restrict__ part, int S, int LN, uint32_t nCols) {
    size_t const i = size_t(blockIdx.x)*blockDim.x + threadIdx.x;
    if (i >= size_t(nCols)*LN) return;
    uint32_t const c = uint32_t(i/LN); int const j = int(i % LN);
    double sum = 0;
    for (int s = 0; s < S; ++s) sum += part[(size_t(c)*S + s)*LN + j];
    rn2[i] = sum;
    double sc = 0, un = 0;
    if (sum > 0 && isfinite(sum)) {
        int e = 0;
        frexp(sqrt(sum), &e);            // sqrt(sum) = m * 2^e, m in [0.5, 1)
        sc = ldexp(1.0, -e); un = ldexp(1.0, e);
    }
    scale[i] = sc; unscale[i] = un;
}

// right-hand sides of the inner plan (its B has the pattern of X, blocks in the caller's order): B[b] = float(-scale * v[bpos[b]])
// where v = A*X - B of the outer plan in storage order
__global__ void to_inner_rhs_kernel(float *__restrict__ Bin, double const *__restrict__ v, double const *__restrict__ scale,
                                    uint32_t const *__restrict__ bpos, uint32_t const *__restrict__ blockcol, int blockElems, int LN) {
    uint32_t const sidx = bpos[blockIdx.x];
    double const *const sc = scale + size_t(blockcol[sidx])*LN;
    double const *const src = v + size_t(sidx)*blockElems;
    float *const dst = Bin + size_t(blockIdx.x)*blockElems;
    for (int q = threadIdx.x; q < blockElems; q += blockDim.x) dst[q] = float(-src[q]*sc[q % LN]);
}

// X += unscale * D   (both in the same column-sorted storage order); a correction that is not finite is dropped
__global__ void add_correction_kernel(double *__restrict__ X, float const *__restrict__ D, double const *__restrict__ unscale,
                                      uint32_t const *__restrict__ blockcol, int blockElems, int LN) {
    double const *const un = unscale + size_t(blockcol[blockIdx.x])*LN;
    size_t const base = size_t(blockIdx.x)*blockElems;
    for (int q = threadIdx.x; q < blockElems; q += blockDim.x) {
        double const d = double(D[base + q])*un[q % LN];
        if (isfinite(d)) X[base + q] += d;
    }
}

double env_double(char const *name, double dflt) {
    char const *e = std::getenv(name);
    if (nullptr == e) return dflt;
    double const v = std::atof(e);
    return (v > 0) ? v : dflt;
}

} // namespace

void mixed_destroy(Plan &p)
{
    if (nullptr == p.mixed) return;
    MixedPlan *m = p.mixed;
    if (m->inner) { plan_release(*m->inner); delete m->inner; }
    if (m->h_rn2) cudaFreeHost(m->h_rn2);
    if (m->d_colstart) cudaFree(m->d_colstart);
    delete m;
    p.mixed = nullptr;
    p.lean_vectors = false;
}

Plan* mixed_inner(Plan const &p) { return p.mixed ? p.mixed->inner : nullptr; }
int   mixed_passes(Plan const &p) { return p.mixed ? p.mixed->passes : 0; }

// bufferSize(..., 'm'): the fp64 plan, the fp32 plan and the scratch of the refinement in ONE caller-owned workspace
tfqmrgpuStatus_t mixed_buffer_size(Plan &p, cudaStream_t stream, int LM, int LN, size_t *bytes)
{
    p.lean_vectors = true;                 // X, Y and a scratch vector: the iteration's v4..v7 live in the fp32 plan only
    tfqmrgpuStatus_t st = plan_configure(p, stream, LM, LN, 'z');
    if (TFQMRGPU_STATUS_SUCCESS != st) { p.lean_vectors = (nullptr != p.mixed); return st; }
    if (nullptr == p.mixed) {
        p.mixed = new (std::nothrow) MixedPlan();
        if (nullptr == p.mixed) { p.lean_vectors = false; return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED); }
    }
    MixedPlan &m = *p.mixed;
    m.a_ready = false;
    if (nullptr == m.inner) {
        m.inner = new (std::nothrow) Plan();
        if (nullptr == m.inner) { mixed_destroy(p); return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED); }
        Plan &q = *m.inner;
        // the residual of a refinement pass is X-shaped: the inner plan's right-hand sides have the pattern of X
        q.mb = p.mb; q.nnzbA = p.nnzbA; q.nnzbX = p.nnzbX; q.nnzbB = p.nnzbX; q.indexOffset = 0;
        st = plan_analyse(q, stream, p.h_rpA.data(), p.h_ciA.data(), p.h_rowptrX.data(), p.h_ciX.data(),
                          p.h_rowptrX.data(), p.h_ciX.data(), 0);
        if (TFQMRGPU_STATUS_SUCCESS != st) { mixed_destroy(p); return st; }
    }
    Plan &q = *m.inner;
    q.pBuffer = nullptr; q.v3_ready = false;
    st = plan_configure(q, stream, LM, LN, 'c');
    if (TFQMRGPU_STATUS_SUCCESS != st) { mixed_destroy(p); return st; }
    q.configured = true;
    if (q.nCols != p.nCols || q.h_colstart != p.h_colstart) { mixed_destroy(p); return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR); }

    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    m.slices = int(std::max<size_t>(1, std::min<size_t>(64, (size_t(4)*nsm + p.nCols - 1)/p.nCols)));
    size_t off = align256(p.bufferBytes);
    auto take = [&](size_t b) { size_t const at = off; off = align256(off + b); return at; };
    m.off_inner = take(q.bufferBytes);
    size_t const nrhs = size_t(p.nCols)*LN;
    m.off_part = take(nrhs*m.slices*8);
    m.off_rn2 = take(nrhs*8); m.off_scale = take(nrhs*8); m.off_unscale = take(nrhs*8);
    m.totalBytes = off + 256;
    if (m.h_rn2) { cudaFreeHost(m.h_rn2); m.h_rn2 = nullptr; }
    TFQ_CUDA(cudaMallocHost((void**)&m.h_rn2, nrhs*8));
    if (m.d_colstart) { cudaFree(m.d_colstart); m.d_colstart = nullptr; }
    TFQ_CUDA(cudaMalloc((void**)&m.d_colstart, (size_t(p.nCols) + 1)*4));
    TFQ_CUDA(cudaMemcpy(m.d_colstart, p.h_colstart.data(), (size_t(p.nCols) + 1)*4, cudaMemcpyHostToDevice));
    m.bn2.assign(nrhs, 0.0);
    *bytes = m.totalBytes;
    return TFQMRGPU_STATUS_SUCCESS;
}

// setBuffer: the inner plan's window of the caller's workspace (the outer plan has been attached by the caller of this function)
tfqmrgpuStatus_t mixed_set_buffer(Plan &p, cudaStream_t stream)
{
    MixedPlan &m = *p.mixed;
    Plan &q = *m.inner;
    q.pBuffer = p.pBuffer + m.off_inner;
    plan_drop_graph(q);
    TFQ_CUDA(cudaMemsetAsync(q.pBuffer + q.off_zero, 0, q.bufferBytes - 256 - q.off_zero, stream));
    tfqmrgpuStatus_t const st = fill_v3(q, stream);
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    q.v3_ready = true;
    m.a_ready = false;
    return TFQMRGPU_STATUS_SUCCESS;
}

// after setMatrix('A') of the outer plan: the fp32 copy of the operator in the inner plan's form
tfqmrgpuStatus_t mixed_after_set_a(Plan &p, cudaStream_t stream)
{
    MixedPlan &m = *p.mixed;
    Plan &q = *m.inner;
    if (nullptr == q.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    size_t const n2 = size_t(p.nnzbA)*p.LM*p.LM;           // pairs of elements
    if (n2 > 0) {
        int const grid = int(std::min<size_t>((n2 + 255)/256, size_t(148)*16));
        to_float_kernel<<<grid, 256, 0, stream>>>(ws<float2>(q, q.off_A), ws<double2 const>(p, p.off_A), n2);
        TFQ_CUDA(cudaGetLastError());
        if (q.use_tc16) {
            tfqmrgpuStatus_t st = launch_aop_blockmax(q, 0, uint32_t(p.nnzbA), stream);
            if (TFQMRGPU_STATUS_SUCCESS == st) st = launch_aop_convert(q, stream);
            if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        }
    }
    m.a_ready = true;
    return TFQMRGPU_STATUS_SUCCESS;
}

namespace {

// column norms of an X-shaped fp64 vector of the outer plan -> m.h_rn2 (blocking) and scale / unscale on the device
tfqmrgpuStatus_t column_norms(Plan &p, cudaStream_t stream, double const *v)
{
    MixedPlan &m = *p.mixed;
    uint32_t const *const d_colstart = m.d_colstart;
    int const LN = p.LN, rows = 2*p.LM;
    int const R = std::max(1, std::min(rows, 256/LN));
    int const threads = R*LN;
    size_t const nrhs = size_t(p.nCols)*LN;
    colnorm_part_kernel<<<p.nCols*m.slices, threads, size_t(threads)*8, stream>>>(ws<double>(p, m.off_part), v, d_colstart, m.slices, LN, rows, R);
    colnorm_finish_kernel<<<unsigned((nrhs + 127)/128), 128, 0, stream>>>(ws<double>(p, m.off_rn2), ws<double>(p, m.off_scale),
                                                                         ws<double>(p, m.off_unscale), ws<double const>(p, m.off_part),
                                                                         m.slices, LN, p.nCols);
    TFQ_CUDA(cudaGetLastError());
    TFQ_CUDA(cudaMemcpyAsync(m.h_rn2, p.pBuffer + m.off_rn2, nrhs*8, cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace

tfqmrgpuStatus_t mixed_solve(Plan &p, cudaStream_t stream, double tolerance, int maxIterations)
{
    if (nullptr == p.mixed || nullptr == p.pBuffer || !p.configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    MixedPlan &m = *p.mixed;
    Plan &q = *m.inner;
    if (nullptr == q.pBuffer || !q.configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (!m.a_ready && p.nnzbA > 0) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);     // setMatrix('A') has not been called
    if (p.user_op || p.precond || p.exch.slots) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);

    double *const X = ws<double>(p, p.off_v[1]), *const Y = ws<double>(p, p.off_v[9]);
    int const blockElems = 2*p.LM*p.LN;
    int const threads = std::min(256, ((blockElems + 31)/32)*32);
    size_t const nrhs = size_t(p.nCols)*p.LN;
    double const N = double(p.nnzbX)*p.LM*p.LN, M = double(p.nPairs)*8.*p.LM*p.LM*p.LN;
    double const inner_floor = env_double("TFQMRGPU_MIXED_INNER_TOL", 1e-3);
    int const pass_cap = [&] { char const *e = std::getenv("TFQMRGPU_MIXED_INNER_ITER"); int const v = e ? std::atoi(e) : 0;
                               return (v > 0) ? v : std::max(40, maxIterations/4); }();

    tfqmrgpuStatus_t st;
#define TFQ_DO(call) do { st = (call); if (TFQMRGPU_STATUS_SUCCESS != st) return st; } while (0)
    p.flops_performed = 0; p.iterations_needed = maxIterations; p.residuum_reached = 1e150; m.passes = 0;
    p.xop_of_x = false;
    // The fp32 passes freeze a right-hand side as soon as its true residual passes a probe (setEarlyFreeze): the reference's rule - all
    // right-hand sides below the threshold at the SAME probe - lets converged fp32 columns drift while it waits for the last one, and a
    // pass then runs into its iteration cap (config-4 shard, 128 right-hand sides: 75 instead of 42 fp32 iterations, 1282 vs 736 ms);
    // the fp64 residual of the next pass checks every column anyway.  TFQMRGPU_MIXED_FREEZE=0 switches it off.
    { char const *e = std::getenv("TFQMRGPU_MIXED_FREEZE"); q.early_freeze = (e && '0' == e[0]) ? 0 : 1; }
    if (!p.initial_guess) TFQ_CUDA(cudaMemsetAsync(X, 0, p.vecBytes, stream));       // like the reference (core.hxx:125)

    // |b|^2 per right-hand side
    TFQ_CUDA(cudaMemsetAsync(Y, 0, p.vecBytes, stream));
    TFQ_DO(launch_add_rhs(p, Y, 1.0, -1, stream));
    TFQ_DO(column_norms(p, stream, Y));
    for (size_t i = 0; i < nrhs; ++i) m.bn2[i] = m.h_rn2[i];

    int total_it = 0, stagnant = 0;
    double prev = 1e300, probes = 0, launches = 0;
    tfqmrgpuStatus_t result = TFQMRGPU_STATUS_MAX_ITERATIONS;
    for (int pass = 0; ; ++pass) {
        // ---- R = B - A*X in fp64 (kept with the opposite sign in Y) and its column norms --------------------------------------
        if (0 == pass && !p.initial_guess) {
            TFQ_CUDA(cudaMemsetAsync(Y, 0, p.vecBytes, stream));
        } else {
            TFQ_DO(launch_spmm(p, Y, X, -1, stream));
            p.flops_performed += M;
        }
        TFQ_DO(launch_add_rhs(p, Y, -1.0, -1, stream));
        p.flops_performed += 4.*N;
        TFQ_DO(column_norms(p, stream, Y));
        launches += 5;
        double rel2 = 0;
        for (size_t i = 0; i < nrhs; ++i) {
            double const r = m.h_rn2[i], b = m.bn2[i];
            double const q2 = (b > 0) ? r/b : ((r > 0) ? 1e300 : 0.0);
            if (!(q2 <= rel2)) rel2 = q2;               // (a NaN ends up here as well)
        }
        if (!(rel2 == rel2)) rel2 = 1e300;
        p.residuum_reached = std::sqrt(rel2);
        if (verbosity() > 1) std::printf("# tfQMRgpu(B200) mixed: pass %d, %d inner iterations so far, residual %.3e\n", pass, total_it, p.residuum_reached);
        if (rel2 <= tolerance*tolerance) { result = TFQMRGPU_STATUS_SUCCESS; p.iterations_needed = total_it; break; }
        if (total_it >= maxIterations) break;
        if (pass > 0) {
            stagnant = (rel2 > 0.25*prev) ? stagnant + 1 : 0;     // a pass that gains less than a factor of two
            if (stagnant >= 2) break;
        }
        prev = rel2;

        // ---- A*D = R in fp32 -------------------------------------------------------------------------------------------------
        to_inner_rhs_kernel<<<q.nnzbB, threads, 0, stream>>>(ws<float>(q, q.off_B), Y, ws<double const>(p, m.off_scale),
                                                            q.d_bpos, p.d_blockcol, blockElems, p.LN);
        TFQ_CUDA(cudaGetLastError());
        double const missing = tolerance/std::sqrt(rel2);         // what is still to gain, < 1
        double const tol_in = std::min(0.1, std::max(inner_floor, 0.25*missing));
        int const it_in = std::min(pass_cap, maxIterations - total_it);
        tfqmrgpuStatus_t const ist = solve(q, stream, tol_in, it_in);
        if (TFQMRGPU_STATUS_SUCCESS != ist && TFQMRGPU_STATUS_MAX_ITERATIONS != ist && TFQMRGPU_STATUS_BREAKDOWN != ist) return ist;
        total_it += q.iterations_run;
        p.flops_performed += q.flops_performed;
        probes += q.stat_probes; launches += q.stat_launches + 2;
        add_correction_kernel<<<p.nnzbX, threads, 0, stream>>>(X, ws<float const>(q, q.off_v[1]), ws<double const>(p, m.off_unscale),
                                                              p.d_blockcol, blockElems, p.LN);
        TFQ_CUDA(cudaGetLastError());
        p.flops_performed += 2.*N;
        m.passes = pass + 1;
        if (0 == q.iterations_run) { if (++stagnant >= 2) break; }
    }
#undef TFQ_DO
    TFQ_CUDA(cudaStreamSynchronize(stream));
    p.flops_performed_all += p.flops_performed;
    p.solved = true;
    p.iterations_run = total_it;
    if (TFQMRGPU_STATUS_SUCCESS != result) p.iterations_needed = maxIterations;
    p.stat_probes = probes; p.stat_launches = launches; p.stat_bodies = total_it;
    return result;
}

} // namespace tfq
