// The tfQMR iteration driver.
//
// Role of tfqmrgpu::solve (tfqmrgpu_core.hxx:20-335).  The reference decides on the HOST after every
// iteration (two blocking device->host copies of tau and status, core.hxx:235-236) whether to probe
// the true residual.  Here the solver state lives on the device (tfq::Control): the last CTA of the
// last kernel of an iteration evaluates the reference's rule and sets state = RUN | PROBE | DONE, and
// every kernel starts by checking that state.  The host therefore only ENQUEUES iteration bodies
// (iteration kernels followed by the probe kernels, which are no-ops unless state == PROBE) and looks
// at an asynchronous read-back of the control block a few bodies behind, so host and device never
// serialise inside the loop while the iteration/probe sequence stays exactly the reference's.
#include "tfq_internal.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <dlfcn.h>

namespace tfq {

// The reference brackets its solver with the NVTX ranges "tfQMR preparation" and "tfQMR iterations" (tfqmrgpu.hxx:6-27,
// core.hxx:29,177).  NVTX is resolved at run time (libnvToolsExt.so.1, only if TFQMRGPU_NVTX=1 asks for it), so the library
// has no link dependency on it and the ranges cost nothing when switched off.
namespace {
struct NvtxApi { int (*push)(char const*) = nullptr; int (*pop)() = nullptr; bool tried = false; };
NvtxApi &nvtx() {
    static NvtxApi api;
    if (!api.tried) {
        api.tried = true;
        char const *e = std::getenv("TFQMRGPU_NVTX");
        if (e && '0' != e[0]) {
            void *lib = dlopen("libnvToolsExt.so.1", RTLD_NOW | RTLD_GLOBAL);
            if (nullptr == lib) lib = dlopen("libnvToolsExt.so", RTLD_NOW | RTLD_GLOBAL);
            if (lib) {
                api.push = reinterpret_cast<int (*)(char const*)>(dlsym(lib, "nvtxRangePushA"));
                api.pop  = reinterpret_cast<int (*)()>(dlsym(lib, "nvtxRangePop"));
            }
        }
    }
    return api;
}
struct TfqRange {
    mutable bool open = false;
    explicit TfqRange(char const *name) { NvtxApi &n = nvtx(); if (n.push && n.pop) { n.push(name); open = true; } }
    void end() const { if (open) { nvtx().pop(); open = false; } }
    ~TfqRange() { end(); }
};
} // namespace

namespace {
__global__ void set_control_kernel(Control *ctl, Control const init) { *ctl = init; }
// r := r - y  (initial guess: v5 = b - A*x0)
template <typename real_t>
__global__ void subtract_kernel(real_t *__restrict__ r, real_t const *__restrict__ y, size_t n) {
    for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x)*blockDim.x) r[i] -= y[i];
}
}

namespace {
constexpr int kAhead = 3;      // iteration bodies the host may run ahead of the last control read-back
constexpr int kRing = 6;       // read-back slots; slot 6 = final state, slot 7 = upload staging
}

namespace {
constexpr int kBodyKernels = 11;
}

// Right preconditioner (the slot the reference left commented out: core.hxx:37,57 `has_preconditioner`, `vP`): the solver iterates on
// A*P, every product becomes y = A*(P*x) through the plan's extra vector, and X = P*v1 once the iterations have ended.
static tfqmrgpuStatus_t apply_precond(Plan &p, void const *x, int expect, cudaStream_t stream)
{
    if (nullptr == p.d_precond_tmp) TFQ_CUDA(cudaMalloc((void**)&p.d_precond_tmp, p.vecBytes));
    int32_t const st = p.precond(p.precond_ctx, p.d_precond_tmp, x, reinterpret_cast<int32_t const*>(&ws<Control const>(p, p.off_ctl)->state), expect, stream);
    return st ? tfqmrgpuStatus_t(st) : TFQMRGPU_STATUS_SUCCESS;
}

// one tfQMR iteration (core.hxx:189-233); every kernel checks the device-resident state first and runs only in state RUN
// (events: nullptr, or four events recorded around the two A*v6 products when profiling)
tfqmrgpuStatus_t enqueue_iteration(Plan &p, cudaStream_t stream, cudaEvent_t const *events)
{
    void *const v6 = p.pBuffer + p.off_v[6];
    void *const v8 = p.pBuffer + p.off_v[8], *const v9 = p.pBuffer + p.off_v[9];
    tfqmrgpuStatus_t st;
#define TFQ_DO(call) do { st = (call); if (TFQMRGPU_STATUS_SUCCESS != st) return st; } while (0)
    // plans on the fp16-pair tensor-core product: K1 and K3 write v6 AND its tensor-core operand (TFQMRGPU_XOP_FUSED=0: separate pass)
    static bool const fuse_env = [] { char const *e = std::getenv("TFQMRGPU_XOP_FUSED"); return !(e && '0' == e[0]); }();
    bool const fused = p.use_tc16 && fuse_env && nullptr == p.user_op && nullptr == p.precond;
    void const *xin = v6;                                              // what the products multiply: v6, or P*v6
    TFQ_DO(fused ? launch_vecop_xop(p, OP_K1, stream) : launch_vecop(p, OP_K1, stream));
    if (p.precond) { TFQ_DO(apply_precond(p, v6, STATE_RUN, stream)); xin = p.d_precond_tmp; }
    if (events) TFQ_CUDA(cudaEventRecord(events[0], stream));
    TFQ_DO(fused ? launch_spmm_operand_ready(p, v9, v6, STATE_RUN, stream)
                 : launch_spmm(p, v9, xin, STATE_RUN, stream));        // v9 := A*v6     (core.hxx:198)
    if (events) TFQ_CUDA(cudaEventRecord(events[1], stream));
    TFQ_DO(launch_vecop(p, OP_E1, stream));
    TFQ_DO(launch_vecop(p, OP_K2, stream));
    TFQ_DO(fused ? launch_vecop_xop(p, OP_K3, stream) : launch_vecop(p, OP_K3, stream));
    if (p.precond) TFQ_DO(apply_precond(p, v6, STATE_RUN, stream));
    if (events) TFQ_CUDA(cudaEventRecord(events[2], stream));
    TFQ_DO(fused ? launch_spmm_operand_ready(p, v8, v6, STATE_RUN, stream)
                 : launch_spmm(p, v8, xin, STATE_RUN, stream));        // v8 := A*v6     (core.hxx:224)
    if (events) TFQ_CUDA(cudaEventRecord(events[3], stream));
    TFQ_DO(launch_vecop(p, OP_E2, stream));
    TFQ_DO(launch_vecop(p, OP_K4, stream));
#undef TFQ_DO
    return TFQMRGPU_STATUS_SUCCESS;
}

// the residual probe (core.hxx:263-304); its kernels run only in state PROBE
tfqmrgpuStatus_t enqueue_probe(Plan &p, cudaStream_t stream)
{
    void *const v1 = p.pBuffer + p.off_v[1], *const v9 = p.pBuffer + p.off_v[9];
    tfqmrgpuStatus_t st;
#define TFQ_DO(call) do { st = (call); if (TFQMRGPU_STATUS_SUCCESS != st) return st; } while (0)
    void const *xin = v1;
    if (p.precond) { TFQ_DO(apply_precond(p, v1, STATE_PROBE, stream)); xin = p.d_precond_tmp; }      // the true residual is A*(P*v1) - b
    TFQ_DO(launch_spmm(p, v9, xin, STATE_PROBE, stream));              // v9 := A*v1     (core.hxx:265)
    TFQ_DO(launch_add_rhs(p, v9, -1.0, STATE_PROBE, stream));          // v9 -= b        (core.hxx:267)
    TFQ_DO(launch_vecop(p, OP_N3, stream));
#undef TFQ_DO
    return TFQMRGPU_STATUS_SUCCESS;
}

namespace {

// all shards' convergence monitors of (kind, parity) to every shard, then the common decision (multi-process runs: the hook
// all-gathers on the solver's stream; one process with several devices exchanges through multi.cu instead)
tfqmrgpuStatus_t exchange_and_decide(Plan &p, int kind, cudaStream_t stream)
{
    if (p.exch.hook) {
        int32_t const hst = p.exch.hook(p.exch.hook_ctx, exchange_slots(p, kind), 4*p.exch.nshards, stream);
        if (hst) return tfqmrgpuStatus_t(hst);
    }
    return launch_decide(p, kind, stream);
}

// iteration followed by the probe; with an exchange registered, the shards' monitors are combined after K4 and after N3
tfqmrgpuStatus_t enqueue_body(Plan &p, cudaStream_t stream, cudaEvent_t const *events)
{
    tfqmrgpuStatus_t st = enqueue_iteration(p, stream, events);
    if (TFQMRGPU_STATUS_SUCCESS == st && p.exch.slots) st = exchange_and_decide(p, 0, stream);
    if (TFQMRGPU_STATUS_SUCCESS == st) st = enqueue_probe(p, stream);
    if (TFQMRGPU_STATUS_SUCCESS == st && p.exch.slots) st = exchange_and_decide(p, 1, stream);
    return st;
}

// capture the body once into a CUDA graph: one launch per iteration instead of eleven (small systems such as the
// reference's FD example are bound by launch latency).  The legacy default stream cannot be captured, so the capture
// runs on a private stream and the instantiated graph is launched into the caller's stream.
tfqmrgpuStatus_t build_body_graph(Plan &p)
{
    if (nullptr == p.capture_stream) TFQ_CUDA(cudaStreamCreateWithFlags(&p.capture_stream, cudaStreamNonBlocking));
    TFQ_CUDA(cudaStreamBeginCapture(p.capture_stream, cudaStreamCaptureModeThreadLocal));
    tfqmrgpuStatus_t const st = enqueue_body(p, p.capture_stream, nullptr);
    cudaGraph_t graph = nullptr;
    cudaError_t const e = cudaStreamEndCapture(p.capture_stream, &graph);
    if (TFQMRGPU_STATUS_SUCCESS != st || cudaSuccess != e || nullptr == graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return (TFQMRGPU_STATUS_SUCCESS != st) ? st : TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED);
    }
    cudaError_t const e2 = cudaGraphInstantiate(&p.body_exec, graph, 0);
    cudaGraphDestroy(graph);
    TFQ_CUDA(e2);
    return TFQMRGPU_STATUS_SUCCESS;
}

bool graphs_enabled() {
    static int on = -1;
    if (on < 0) { char const *e = std::getenv("TFQMRGPU_GRAPH"); on = (e && '0' == e[0]) ? 0 : 1; }
    return 1 == on;
}
} // namespace

tfqmrgpuStatus_t solve_begin(Plan &p, cudaStream_t stream, double tolerance, int maxIterations)
{
    // block size / precision dispatch of the reference (tfqmrgpu.cu:40-72)
    if (!block_size_allowed(p.LM, p.LN))
        return TFQMRGPU_BLOCKSIZE_MISSING + TFQMRGPU_CODE_CHAR*p.LM + TFQMRGPU_CODE_LINE*p.LN;
    if ('z' != p.precision && 'c' != p.precision) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, p.precision);
    if (nullptr == p.pBuffer || !p.configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (p.precond && p.initial_guess) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);   // (v1 would mix x0 with the unpreconditioned iterate)

    // lazily created host-side resources; every step is retried by the next solve if it fails here
    for (auto &e : p.ev) if (nullptr == e) TFQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (nullptr == p.h_ctl) TFQ_CUDA(cudaMallocHost((void**)&p.h_ctl, 8*sizeof(Control)));
    Control *const d_ctl = ws<Control>(p, p.off_ctl);

    // ---- initial state (core.hxx:114-131,170-174) ----------------------------------------------------
    Control &c0 = p.h_ctl[7];
    std::memset(&c0, 0, sizeof(Control));
    c0.state = (maxIterations > 0) ? STATE_RUN : STATE_DONE;
    c0.max_iterations = maxIterations;
    c0.result = TFQMRGPU_STATUS_MAX_ITERATIONS;
    c0.iterations_needed = maxIterations;
    c0.tol2 = tolerance*tolerance;
    c0.target_bound2 = c0.tol2*100*100;
    c0.residual2_reached = 1e300;
    c0.freeze = p.early_freeze;
    // by a kernel, not by a host->device copy: a copy would queue on the H2D copy engine behind whatever uploads the caller has in
    // flight on OTHER streams (the next system's 7 GB operator in a double-buffered caller) and the solve would wait for them
    set_control_kernel<<<1, 1, 0, stream>>>(d_ctl, c0);
    TFQ_CUDA(cudaGetLastError());
    TFQ_CUDA(cudaMemsetAsync(p.pBuffer + p.off_ticket, 0, (size_t(p.nCols) + 8)*4, stream));
    // v1 and v4..v9 are contiguous: the initial guess is discarded like in the reference (core.hxx:125) unless the caller asked
    // for it (tfqmrgpux_bsrsv_setInitialGuess: v1 keeps what setMatrix('X') uploaded / the previous solve left)
    size_t const first = p.initial_guess ? p.off_v[4] : p.off_v[1];
    TFQ_CUDA(cudaMemsetAsync(p.pBuffer + first, 0, (p.off_v[9] + p.vecBytes) - first, stream));
    if (p.use_tc16) TFQ_CUDA(cudaMemsetAsync(p.pBuffer + p.off_mx, 0, 3*size_t(p.nCols)*p.LN*sizeof(float), stream));   // max|v6| = 0
    p.exch.parity = 0;
    p.xop_of_x = false;             // the iterations overwrite the tensor-core operand (and X itself)

    tfqmrgpuStatus_t st;
    st = launch_add_rhs(p, p.pBuffer + p.off_v[5], 1.0, -1, stream);       // v5 := b        (core.hxx:153)
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    st = launch_vecop(p, OP_INIT, stream);                                 // tau, 1/|b|^2, first dec35
    p.guess_flops = 0;
    if (TFQMRGPU_STATUS_SUCCESS != st || !p.initial_guess) return st;

    // ---- initial guess x0 = v1 (not in the reference, which zeroes X: core.hxx:125): tfQMR on the residual r0 = b - A*x0.  The
    //      iteration is unchanged - v1 accumulates the corrections on top of x0, the probe evaluates A*v1 - b - only its start
    //      differs: v5 = r0, tau = |r0|^2 and the first rho = v3.r0, while the convergence bound stays relative to |b|
    //      (1/|b|^2 from the first INIT is kept).
    size_t const nrhs = size_t(p.nCols)*p.LN;
    if (nullptr == p.d_guess_scratch) TFQ_CUDA(cudaMalloc((void**)&p.d_guess_scratch, nrhs*sizeof(double)));
    TFQ_CUDA(cudaMemcpyAsync(p.d_guess_scratch, p.pBuffer + p.off_invBn2, nrhs*sizeof(double), cudaMemcpyDeviceToDevice, stream));
    st = launch_spmm(p, p.pBuffer + p.off_v[9], p.pBuffer + p.off_v[1], -1, stream);
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    size_t const n = size_t(p.nnzbX)*2*p.LM*p.LN;
    int const grid = int(std::min<size_t>((n + 255)/256, size_t(148)*16));
    if ('z' == p.precision) subtract_kernel<double><<<grid, 256, 0, stream>>>(ws<double>(p, p.off_v[5]), ws<double const>(p, p.off_v[9]), n);
    else                    subtract_kernel<float ><<<grid, 256, 0, stream>>>(ws<float >(p, p.off_v[5]), ws<float const >(p, p.off_v[9]), n);
    TFQ_CUDA(cudaGetLastError());
    TFQ_CUDA(cudaMemsetAsync(p.pBuffer + p.off_v[9], 0, p.vecBytes, stream));
    st = launch_vecop(p, OP_INIT, stream);                                 // tau = |r0|^2, first dec35 on r0
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    TFQ_CUDA(cudaMemcpyAsync(p.pBuffer + p.off_invBn2, p.d_guess_scratch, nrhs*sizeof(double), cudaMemcpyDeviceToDevice, stream));
    p.guess_flops = double(p.nPairs)*8.*p.LM*p.LM*p.LN + 2.*double(n);
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t solve_finish(Plan &p, Control const &fin, int bodies, double launches)
{
    // ---- bookkeeping (core.hxx:133-138,324-325; flop formula of SURVEY.md a14) ------------------------
    double const N = double(p.nnzbX)*p.LM*p.LN;
    double const M = double(p.nPairs)*8.*p.LM*p.LM*p.LN;
    p.flops_performed = fin.iteration*(104.*N + 2.*M) + 4.*N + fin.probes*(M + 4.*N) + p.guess_flops;
    p.flops_performed_all += p.flops_performed; // the reference never accumulates this (defect, fixed)
    p.residuum_reached = std::sqrt(fin.residual2_reached);
    p.iterations_needed = fin.iterations_needed;
    p.iterations_run = fin.iteration;
    p.solved = true;
    p.stat_probes = fin.probes; p.stat_launches = launches; p.stat_bodies = bodies;
    p.stat_bound2 = fin.max_bound2; p.stat_target2 = fin.target_bound2;
    return fin.result;
}

tfqmrgpuStatus_t solve(Plan &p, cudaStream_t stream, double tolerance, int maxIterations)
{
    auto const t_start = std::chrono::steady_clock::now();
    TfqRange const range_prep("tfQMR preparation");     // the reference's NVTX ranges (core.hxx:29,177), behind TFQMRGPU_NVTX=1
    tfqmrgpuStatus_t st = solve_begin(p, stream, tolerance, maxIterations);
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    Control *const d_ctl = ws<Control>(p, p.off_ctl);
    double launches = 2;
    // profiling: CUDA events on the solver's own stream around the solve and around every A*v6 product
    constexpr int kProfBodies = 256;
    if (p.profile && p.prof_ev.empty()) p.prof_ev.resize(2 + 4*kProfBodies, nullptr);
    if (p.profile) {
        for (auto &e : p.prof_ev) if (nullptr == e) TFQ_CUDA(cudaEventCreate(&e));
    }
    auto mark = [&](int which) { return p.profile ? cudaEventRecord(p.prof_ev[which], stream) : cudaSuccess; };
    TFQ_CUDA(mark(0));
    // (a callback cannot be captured blindly; the exchange hook of a sharded run is a host call per iteration)
    bool const use_graph = graphs_enabled() && !p.profile && maxIterations > 0 && nullptr == p.user_op && nullptr == p.precond && nullptr == p.exch.slots
                           && !(resident_supported(p) && !p.initial_guess);
    if (use_graph && nullptr == p.body_exec) {
        tfqmrgpuStatus_t const gst = build_body_graph(p);
        if (TFQMRGPU_STATUS_SUCCESS != gst) return gst;
    }

    if (verbosity() > 0) { // the reference prints this line unconditionally (core.hxx:167)
        std::vector<double> inv(size_t(p.nCols)*p.LN);
        TFQ_CUDA(cudaMemcpyAsync(inv.data(), p.pBuffer + p.off_invBn2, inv.size()*8, cudaMemcpyDeviceToHost, stream));
        TFQ_CUDA(cudaStreamSynchronize(stream));
        double mn = 9e99, mx = -1;
        for (double v : inv) { double const n2 = 1./v; mn = std::min(mn, n2); mx = std::max(mx, n2); }
        std::printf("# norms of B within [%g, %g]\n", std::sqrt(mn), std::sqrt(mx));
    }
    range_prep.end();
    TfqRange const range_iter("tfQMR iterations");

    int bodies = 0;
    // small systems: the whole solve in one cooperative launch (resident.cu); the loop below is skipped
    bool const resident = maxIterations > 0 && !p.initial_guess && nullptr == p.precond && resident_supported(p);   // (the resident solver starts from X = 0)
    if (resident) {
        st = launch_resident_solve(p, stream, maxIterations);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        launches += 1;
    }
    for (int i = 0; i < maxIterations && !resident; ++i) {
        if (i >= kAhead) {
            int const slot = (i - kAhead) % kRing;
            TFQ_CUDA(cudaEventSynchronize(p.ev[slot]));
            if (STATE_DONE == p.h_ctl[slot].state) break;
        }
        if (use_graph) {
            TFQ_CUDA(cudaGraphLaunch(p.body_exec, stream));
            launches += kBodyKernels;
        } else {
            bool const prof = p.profile && (i < kProfBodies);
            p.exch.parity = i & 1;
            st = enqueue_body(p, stream, prof ? &p.prof_ev[2 + 4*i] : nullptr);
            if (TFQMRGPU_STATUS_SUCCESS != st) return st;
            launches += kBodyKernels;
        }
        int const slot = i % kRing;
        TFQ_CUDA(cudaMemcpyAsync(&p.h_ctl[slot], d_ctl, sizeof(Control), cudaMemcpyDeviceToHost, stream));
        TFQ_CUDA(cudaEventRecord(p.ev[slot], stream));
        ++bodies;
    }
    if (p.precond && maxIterations > 0) {      // X = P*v1 (the speculative bodies behind DONE did not touch v1)
        st = apply_precond(p, p.pBuffer + p.off_v[1], -1, stream);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        TFQ_CUDA(cudaMemcpyAsync(p.pBuffer + p.off_v[1], p.d_precond_tmp, p.vecBytes, cudaMemcpyDeviceToDevice, stream));
        launches += 1;
    }
    Control &fin = p.h_ctl[6];
    TFQ_CUDA(cudaMemcpyAsync(&fin, d_ctl, sizeof(Control), cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(mark(1));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    TFQ_CUDA(cudaGetLastError());
    if (p.profile) {
        float ms = 0;
        TFQ_CUDA(cudaEventElapsedTime(&ms, p.prof_ev[0], p.prof_ev[1]));
        p.prof_solve_ms = ms; p.prof_spmm_ms = 0; p.prof_spmm_launches = 0; p.prof_iterations = fin.iteration;
        // only the bodies that really iterated: later ones were no-ops (state != RUN)
        for (int i = 0; i < fin.iteration && i < kProfBodies && i < bodies; ++i) {
            for (int h = 0; h < 2; ++h) {
                TFQ_CUDA(cudaEventElapsedTime(&ms, p.prof_ev[2 + 4*i + 2*h], p.prof_ev[3 + 4*i + 2*h]));
                p.prof_spmm_ms += ms; p.prof_spmm_launches += 1;
            }
        }
    }
    if (resident) bodies = fin.iteration;
    tfqmrgpuStatus_t const result = solve_finish(p, fin, bodies, launches);
    p.stat_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    if (verbosity() > 1)
        std::printf("# tfQMRgpu(B200): %d iterations, %d probes, residual %.3e, status %d, %.3f ms\n",
                    fin.iteration, fin.probes, p.residuum_reached, fin.result, p.stat_ms);
    return result;
}

} // namespace tfq
