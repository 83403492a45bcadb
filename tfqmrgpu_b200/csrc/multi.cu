// Several GPUs in one process behind the ordinary C-ABI (tfqmrgpux_bsrsv_setDevices / TFQMRGPU_NUM_GPUS).
//
// The reference has no multi-GPU code.  What shards: every tfQMR scalar is per right-hand-side column and A*X never mixes
// block columns (tfqmrgpu_core.hxx:189-233, tfqmrgpu_blocksparse.hxx:71-199), so the block columns of X/B are cut into
// contiguous ranges (balanced by X blocks), A is replicated, and every device runs the ordinary single-GPU kernels on a
// sub-plan of its own.  What is shared: the reference's iteration counter and probe schedule are GLOBAL - one maximum over
// all right-hand sides decides (core.hxx:239-299) - so after K4 and after N3 every shard exports (max, count, count) into
// pinned host memory that all devices map, the devices' streams wait for each other's events, and decide_kernel (vecops.cu)
// takes the same decision everywhere: an N-GPU run has the single-GPU iteration count, and - because a shard tiles its
// vectors like the unsharded plan and the product's accumulation chains do not depend on the schedule - the single-GPU bits.
//
// Data movement: A is uploaded in N row ranges, one per device over its own PCIe link, converted there (layout, fp16 operand
// pairs) and pushed device-to-device (NVLink) to the other N-1 windows; B and an optional initial X are scattered by the host;
// X is gathered device-to-device into the caller's workspace when getMatrix asks for it.  The caller's single workspace lives
// on the "home" device (current at createPlan): [shard 0's workspace | gathered X | conversion scratch]; the other devices'
// workspaces are allocated here.  One host thread enqueues for all devices (its order gives the cross-device event order);
// host threads are only used to drive the N uploads of A concurrently.
#include "tfq_internal.hpp"
#include <curand.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

namespace tfq {

namespace {

constexpr int kAhead = 3, kRing = 6;

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct DeviceGuard {         // restores the caller's current device
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
};

__global__ void gather_blocks_f32(float *__restrict__ dst, float const *__restrict__ src, uint32_t const *__restrict__ sel, uint32_t blockElems) {
    size_t const so = size_t(sel[blockIdx.x])*blockElems, dof = size_t(blockIdx.x)*blockElems;
    for (uint32_t q = threadIdx.x; q < blockElems; q += blockDim.x) dst[dof + q] = src[so + q];
}

} // namespace

struct Shard {
    int dev = 0;
    cudaStream_t stream = nullptr;
    Plan *plan = nullptr;                 // analysis of this shard's block columns, on `dev`
    void *ws = nullptr; bool own_ws = false;
    std::vector<uint32_t> selX, selB;     // the shard's blocks: indices into the caller's X and B block arrays
    uint32_t *d_selX = nullptr;
    uint32_t c0 = 0, c1 = 0;              // dense block-column range
    uint32_t g0 = 0, g1 = 0;              // the same range in the column-sorted storage order of the unsharded X
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // after K4, after N3, after the pushes of A
    int a_row0 = 0, a_row1 = 0;           // block rows of A this device uploads and converts
};

struct MultiPlan {
    std::vector<Shard> shards;
    double *slots = nullptr;              // pinned, portable, mapped: [2 kinds][2 parities][n][4]
    int home = 0;
    size_t off_gx = 0, off_scratch = 0;   // gathered X and conversion scratch in the caller's workspace
    Control *h_fin = nullptr;             // pinned: final control block of every shard
};

static void multi_release_shards(MultiPlan &m)
{
    DeviceGuard guard;
    for (Shard &s : m.shards) {
        cudaSetDevice(s.dev);
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.plan) { plan_release(*s.plan); delete s.plan; }
        if (s.own_ws && s.ws) cudaFree(s.ws);
        if (s.d_selX) cudaFree(s.d_selX);
        for (auto &e : s.ev) if (e) cudaEventDestroy(e);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    m.shards.clear();
    if (m.slots) { cudaFreeHost(m.slots); m.slots = nullptr; }
    if (m.h_fin) { cudaFreeHost(m.h_fin); m.h_fin = nullptr; }
}

void multi_destroy(Plan &p)
{
    if (nullptr == p.multi) return;
    multi_release_shards(*p.multi);
    delete p.multi;
    p.multi = nullptr;
}

// ---- setDevices: partition the block columns, analyse every shard on its device ------------------------------------------
tfqmrgpuStatus_t multi_set_devices(Plan &p, int nDevices, int const *devices)
{
    multi_destroy(p);
    p.configured = false; p.pBuffer = nullptr; p.bufferBytes = 0; p.v3_ready = false;
    if (nDevices <= 1) return TFQMRGPU_STATUS_SUCCESS;
    int count = 0;
    TFQ_CUDA(cudaGetDeviceCount(&count));
    int n = std::min<int>(nDevices, int(p.nCols));          // at least one block column per device
    std::vector<int> devs(n);
    for (int s = 0; s < n; ++s) {
        devs[s] = devices ? devices[s] : s;
        if (devs[s] < 0 || devs[s] >= count) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    }
    if (n <= 1) return TFQMRGPU_STATUS_SUCCESS;
    if (p.h_ciX.size() != size_t(p.nnzbX)) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);   // createPlan keeps the index arrays

    DeviceGuard guard;
    MultiPlan *m = new (std::nothrow) MultiPlan();
    if (nullptr == m) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
    p.multi = m;
    m->home = guard.saved;
    // the analysis of the unsharded problem lives on the home device: dense column ids and the B -> X map
    std::vector<uint16_t> colindx(p.nnzbX);
    std::vector<uint32_t> subset(std::max(p.nnzbB, 1));
    TFQ_CUDA(cudaMemcpy(colindx.data(), p.d_colindx, size_t(p.nnzbX)*sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (p.nnzbB > 0) TFQ_CUDA(cudaMemcpy(subset.data(), p.d_subset, size_t(p.nnzbB)*sizeof(uint32_t), cudaMemcpyDeviceToHost));

    // contiguous column ranges balanced by the number of X blocks, at least one column each
    std::vector<uint32_t> bounds(n + 1, 0);
    {
        uint32_t const nc = p.nCols;
        double const total = double(p.nnzbX);
        for (int r = 1; r < n; ++r) {
            double const target = total*r/n;
            uint32_t c = uint32_t(std::lower_bound(p.h_colstart.begin(), p.h_colstart.end(), uint32_t(std::ceil(target))) - p.h_colstart.begin());
            c = std::max(c, bounds[r - 1] + 1);
            c = std::min(c, nc - uint32_t(n - r));
            bounds[r] = c;
        }
        bounds[n] = nc;
    }

    m->shards.resize(n);
    int const mb = p.mb;
    for (int s = 0; s < n; ++s) {
        Shard &sh = m->shards[s];
        sh.dev = devs[s];
        sh.c0 = bounds[s]; sh.c1 = bounds[s + 1];
        sh.g0 = p.h_colstart[sh.c0]; sh.g1 = p.h_colstart[sh.c1];
        // the shard's sub-patterns of X and B (zero-based row pointers, the caller's column values)
        std::vector<int32_t> rpX(mb + 1, 0), ciX, rpB(mb + 1, 0), ciB;
        for (int r = 0; r < mb; ++r) {
            for (int i = p.h_rowptrX[r]; i < p.h_rowptrX[r + 1]; ++i)
                if (colindx[i] >= sh.c0 && colindx[i] < sh.c1) { ciX.push_back(p.h_ciX[i]); sh.selX.push_back(uint32_t(i)); }
            rpX[r + 1] = int32_t(ciX.size());
            for (int i = p.h_rpB[r]; i < p.h_rpB[r + 1]; ++i) {
                uint16_t const c = colindx[subset[i]];
                if (c >= sh.c0 && c < sh.c1) { ciB.push_back(p.h_ciB[i]); sh.selB.push_back(uint32_t(i)); }
            }
            rpB[r + 1] = int32_t(ciB.size());
        }
        TFQ_CUDA(cudaSetDevice(sh.dev));
        if (sh.dev != m->home) {              // peer access both ways (X gather, pushes of A); "already enabled" is fine
            cudaError_t e = cudaDeviceEnablePeerAccess(m->home, 0);
            if (cudaSuccess != e && cudaErrorPeerAccessAlreadyEnabled != e) { cudaGetLastError(); }
            cudaGetLastError();
        }
        for (int t = 0; t < n; ++t) if (devs[t] != sh.dev) { cudaDeviceEnablePeerAccess(devs[t], 0); cudaGetLastError(); }
        TFQ_CUDA(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking));
        for (auto &e : sh.ev) TFQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sh.plan = new (std::nothrow) Plan();
        if (nullptr == sh.plan) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
        Plan &sp = *sh.plan;
        sp.mb = mb; sp.nnzbA = p.nnzbA; sp.nnzbX = int(ciX.size()); sp.nnzbB = int(ciB.size()); sp.indexOffset = 0;
        tfqmrgpuStatus_t const st = plan_analyse(sp, sh.stream, p.h_rpA.data(), p.h_ciA.data(), rpX.data(), ciX.data(), rpB.data(), ciB.data(), 0);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        if (sp.nCols != sh.c1 - sh.c0) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
        TFQ_CUDA(cudaMalloc((void**)&sh.d_selX, std::max<size_t>(sh.selX.size(), 1)*sizeof(uint32_t)));
        TFQ_CUDA(cudaMemcpy(sh.d_selX, sh.selX.data(), sh.selX.size()*sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    // row ranges of A for the N concurrent uploads, balanced by blocks
    for (int s = 0; s < n; ++s) {
        auto row_at = [&](double frac) {
            int32_t const target = int32_t(std::llround(frac*p.nnzbA));
            return int(std::lower_bound(p.h_rpA.begin(), p.h_rpA.end(), target) - p.h_rpA.begin());
        };
        m->shards[s].a_row0 = (0 == s) ? 0 : std::min(mb, row_at(double(s)/n));
        m->shards[s].a_row1 = (n - 1 == s) ? mb : std::min(mb, row_at(double(s + 1)/n));
    }
    for (int s = 1; s < n; ++s) m->shards[s].a_row0 = m->shards[s - 1].a_row1;
    TFQ_CUDA(cudaHostAlloc((void**)&m->slots, size_t(2*2*n*4)*sizeof(double), cudaHostAllocPortable | cudaHostAllocMapped));
    std::memset(m->slots, 0, size_t(2*2*n*4)*sizeof(double));
    TFQ_CUDA(cudaHostAlloc((void**)&m->h_fin, size_t(n)*sizeof(Control), cudaHostAllocPortable));
    return TFQMRGPU_STATUS_SUCCESS;
}

int multi_get_devices(Plan const &p, int *devices, int arrayLength)
{
    if (nullptr == p.multi) return 1;
    int const n = int(p.multi->shards.size());
    for (int s = 0; s < n && s < arrayLength && devices; ++s) devices[s] = p.multi->shards[s].dev;
    return n;
}

// ---- bufferSize -------------------------------------------------------------------------------------------------------
tfqmrgpuStatus_t multi_buffer_size(Plan &p, int LM, int LN, char prec, size_t *bytes)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    bool const is_double = ('z' == prec);
    size_t const blockBytes = 2*size_t(LM)*LN*(is_double ? 8 : 4);
    p.LM = LM; p.LN = LN; p.precision = prec;
    p.vecBytes = size_t(p.nnzbX)*blockBytes;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, m.home);
    int const n = int(m.shards.size());
    for (int s = 0; s < n; ++s) {
        Shard &sh = m.shards[s];
        TFQ_CUDA(cudaSetDevice(sh.dev));
        if (sh.own_ws && sh.ws) { cudaFree(sh.ws); sh.ws = nullptr; sh.own_ws = false; }
        Plan &sp = *sh.plan;
        sp.pBuffer = nullptr; sp.bufferBytes = 0; sp.v3_ready = false; sp.configured = false;
        plan_drop_graph(sp);
        sp.tile_blocks_hint = p.tile_blocks_hint;   // (tiles are cut per column: a shard tiles its columns like the unsharded plan anyway)
        sp.max_cols_hint = p.maxColsPerRow;
        tfqmrgpuStatus_t const st = plan_configure(sp, sh.stream, LM, LN, prec);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        sp.configured = true;
        sp.exch.nshards = n; sp.exch.shard = s; sp.exch.nrhs_global = (long long)(p.nCols)*LN; sp.exch.slots = m.slots;
    }
    m.off_gx = align256(m.shards[0].plan->bufferBytes);
    m.off_scratch = m.off_gx + align256(p.vecBytes);
    p.bufferBytes = m.off_scratch + align256(p.vecBytes) + 256;
    *bytes = p.bufferBytes;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- setBuffer: workspaces, zeroed scalars, the shards' slices of the single-GPU shadow vector v3 -----------------------------
tfqmrgpuStatus_t multi_set_buffer(Plan &p, void *pBuffer)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    size_t const nGlobal = size_t(p.nnzbX)*2*p.LM*p.LN;            // floats of the unsharded v3 (always float, core.hxx:60)
    uint32_t const blockElems = 2u*p.LM*p.LN;
    for (size_t s = 0; s < m.shards.size(); ++s) {
        Shard &sh = m.shards[s];
        Plan &sp = *sh.plan;
        TFQ_CUDA(cudaSetDevice(sh.dev));
        if (sh.own_ws && sh.ws) { cudaFree(sh.ws); sh.ws = nullptr; sh.own_ws = false; }
        if (0 == s) sh.ws = pBuffer;
        else { TFQ_CUDA(cudaMalloc(&sh.ws, sp.bufferBytes)); sh.own_ws = true; }
        sp.pBuffer = static_cast<char*>(sh.ws);
        plan_drop_graph(sp);
        TFQ_CUDA(cudaMemsetAsync(sp.pBuffer + sp.off_zero, 0, sp.bufferBytes - 256 - sp.off_zero, sh.stream));
        // the reference's stream (cuRAND XORWOW, seed 1234, linalg.hxx:784-797) for the UNSHARDED block order; keep this shard's blocks
        float *full = nullptr; bool own_full = false;
        if (0 == s && p.vecBytes >= nGlobal*sizeof(float)) full = reinterpret_cast<float*>(static_cast<char*>(pBuffer) + m.off_scratch);
        else { TFQ_CUDA(cudaMalloc((void**)&full, nGlobal*sizeof(float))); own_full = true; }
        curandGenerator_t gen;
        tfqmrgpuStatus_t st = TFQMRGPU_STATUS_SUCCESS;
        if (CURAND_STATUS_SUCCESS != curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
        else {
            if (CURAND_STATUS_SUCCESS != curandSetStream(gen, sh.stream)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
            else if (CURAND_STATUS_SUCCESS != curandSetPseudoRandomGeneratorSeed(gen, 1234ull)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
            else if (CURAND_STATUS_SUCCESS != curandGenerateUniform(gen, full, nGlobal)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
            if (TFQMRGPU_STATUS_SUCCESS == st && sp.nnzbX > 0) {
                float *const scratch = ws<float>(sp, sp.off_v[9]);
                gather_blocks_f32<<<sp.nnzbX, std::min(256u, ((blockElems + 31)/32)*32), 0, sh.stream>>>(scratch, full, sh.d_selX, blockElems);
                st = permute_v3(sp, ws<float>(sp, sp.off_v[3]), scratch, true, sh.stream);
            }
            cudaStreamSynchronize(sh.stream);      // the generator must outlive its asynchronous work
            curandDestroyGenerator(gen);
        }
        if (own_full) cudaFree(full);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
        TFQ_CUDA(cudaGetLastError());
        sp.v3_ready = true;
    }
    p.pBuffer = static_cast<char*>(pBuffer);
    p.v3_ready = true;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- setMatrix --------------------------------------------------------------------------------------------------------
tfqmrgpuStatus_t multi_set_matrix(Plan &p, char v, void const *val, char precision, char transposition, tfqmrgpuDataLayout_t layout,
                                  bool trans, double scal_imag)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    bool const is_double = ('z' == p.precision);
    size_t const s_ = is_double ? 8 : 4;
    int const n = int(m.shards.size());
    if ('a' == v) {
        size_t const blockBytes = 2*size_t(p.LM)*p.LM*s_;
        // every device uploads its row range into its own A window over its own PCIe link and converts it there ...
        std::vector<tfqmrgpuStatus_t> status(n, TFQMRGPU_STATUS_SUCCESS);
        auto upload = [&](int s) {
            Shard &sh = m.shards[s];
            Plan &sp = *sh.plan;
            if (cudaSuccess != cudaSetDevice(sh.dev)) { status[s] = TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED); return; }
            uint32_t const b0 = uint32_t(p.h_rpA[sh.a_row0]), b1 = uint32_t(p.h_rpA[sh.a_row1]);
            if (b1 <= b0) return;
            char *const dst = sp.pBuffer + sp.off_A + size_t(b0)*blockBytes;
            if (cudaSuccess != cudaMemcpyAsync(dst, static_cast<char const*>(val) + size_t(b0)*blockBytes, size_t(b1 - b0)*blockBytes,
                                               cudaMemcpyHostToDevice, sh.stream)) { status[s] = TFQ_ERR(TFQMRGPU_STATUS_LAUNCH_FAILED); return; }
            tfqmrgpuStatus_t st = convert_inplace(sp, dst, b1 - b0, p.LM, p.LM, is_double, layout, trans, scal_imag, sh.stream);
            if (TFQMRGPU_STATUS_SUCCESS == st && sp.use_tc16) st = launch_aop_blockmax(sp, b0, b1 - b0, sh.stream);
            if (TFQMRGPU_STATUS_SUCCESS == st && sp.use_tc16) st = launch_aop_convert_rows(sp, sh.a_row0, sh.a_row1, sh.stream);
            status[s] = st;
        };
        {
            std::vector<std::thread> threads;
            for (int s = 1; s < n; ++s) threads.emplace_back(upload, s);
            upload(0);
            for (auto &t : threads) t.join();
        }
        for (int s = 0; s < n; ++s) if (TFQMRGPU_STATUS_SUCCESS != status[s]) return status[s];
        // ... and pushes the converted range (and its row scales) to the other windows, device to device
        for (int s = 0; s < n; ++s) {
            Shard &sh = m.shards[s];
            Plan &sp = *sh.plan;
            TFQ_CUDA(cudaSetDevice(sh.dev));
            uint32_t const b0 = uint32_t(p.h_rpA[sh.a_row0]), b1 = uint32_t(p.h_rpA[sh.a_row1]);
            for (int t = 0; t < n && b1 > b0; ++t) {
                if (t == s) continue;
                Plan &tp = *m.shards[t].plan;
                TFQ_CUDA(cudaMemcpyPeerAsync(tp.pBuffer + tp.off_A + size_t(b0)*blockBytes, m.shards[t].dev,
                                             sp.pBuffer + sp.off_A + size_t(b0)*blockBytes, sh.dev, size_t(b1 - b0)*blockBytes, sh.stream));
                if (sp.use_tc16 && sh.a_row1 > sh.a_row0)
                    TFQ_CUDA(cudaMemcpyPeerAsync(tp.pBuffer + tp.off_ainv + size_t(sh.a_row0)*4, m.shards[t].dev,
                                                 sp.pBuffer + sp.off_ainv + size_t(sh.a_row0)*4, sh.dev, size_t(sh.a_row1 - sh.a_row0)*4, sh.stream));
            }
            TFQ_CUDA(cudaEventRecord(sh.ev[2], sh.stream));
        }
        for (int s = 0; s < n; ++s) {          // a window is complete when every other device's push has arrived
            TFQ_CUDA(cudaSetDevice(m.shards[s].dev));
            for (int t = 0; t < n; ++t) if (t != s) TFQ_CUDA(cudaStreamWaitEvent(m.shards[s].stream, m.shards[t].ev[2], 0));
        }
        return TFQMRGPU_STATUS_SUCCESS;
    }
    // 'b' / 'x': the host scatters the shard's blocks (caller order) and every shard converts its own
    size_t const blockBytes = 2*size_t(p.LM)*p.LN*s_;
    for (int s = 0; s < n; ++s) {
        Shard &sh = m.shards[s];
        Plan &sp = *sh.plan;
        std::vector<uint32_t> const &sel = ('b' == v) ? sh.selB : sh.selX;
        if (sel.empty()) continue;
        std::vector<char> tmp(sel.size()*blockBytes);
        for (size_t i = 0; i < sel.size(); ++i)
            std::memcpy(tmp.data() + i*blockBytes, static_cast<char const*>(val) + size_t(sel[i])*blockBytes, blockBytes);
        TFQ_CUDA(cudaSetDevice(sh.dev));
        tfqmrgpuStatus_t st;
        if ('b' == v) {
            char *const dst = sp.pBuffer + sp.off_B;
            TFQ_CUDA(cudaMemcpyAsync(dst, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, sh.stream));
            st = convert_inplace(sp, dst, uint32_t(sel.size()), p.LM, p.LN, is_double, layout, trans, scal_imag, sh.stream);
        } else {
            char *const scratch = sp.pBuffer + sp.off_v[9];
            TFQ_CUDA(cudaMemcpyAsync(scratch, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, sh.stream));
            st = convert_permuted(sp, sp.pBuffer + sp.off_v[1], scratch, uint32_t(sel.size()), p.LM, p.LN, is_double, layout, trans, scal_imag, true, sh.stream);
        }
        TFQ_CUDA(cudaStreamSynchronize(sh.stream));        // tmp goes out of scope
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    }
    (void)precision; (void)transposition;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- solve: all shards in lockstep, the reference's global rule ---------------------------------------------------------------
tfqmrgpuStatus_t multi_solve(Plan &p, double tolerance, int maxIterations)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    int const n = int(m.shards.size());
    tfqmrgpuStatus_t st;
    for (Shard &sh : m.shards) {
        TFQ_CUDA(cudaSetDevice(sh.dev));
        st = solve_begin(*sh.plan, sh.stream, tolerance, maxIterations);
        if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    }
    Plan &p0 = *m.shards[0].plan;
    int bodies = 0;
    for (int i = 0; i < maxIterations; ++i) {
        if (i >= kAhead) {      // every shard takes the same decisions: shard 0's control block tells when to stop
            int const slot = (i - kAhead) % kRing;
            TFQ_CUDA(cudaEventSynchronize(p0.ev[slot]));
            if (STATE_DONE == p0.h_ctl[slot].state) break;
        }
        for (int kind = 0; kind < 2; ++kind) {
            for (Shard &sh : m.shards) {
                TFQ_CUDA(cudaSetDevice(sh.dev));
                sh.plan->exch.parity = i & 1;
                st = (0 == kind) ? enqueue_iteration(*sh.plan, sh.stream, nullptr) : enqueue_probe(*sh.plan, sh.stream);
                if (TFQMRGPU_STATUS_SUCCESS != st) return st;
                TFQ_CUDA(cudaEventRecord(sh.ev[kind], sh.stream));
            }
            for (int s = 0; s < n; ++s) {     // every shard's monitors are in the mapped host slots once all K4 (N3) have finished
                Shard &sh = m.shards[s];
                TFQ_CUDA(cudaSetDevice(sh.dev));
                for (int t = 0; t < n; ++t) if (t != s) TFQ_CUDA(cudaStreamWaitEvent(sh.stream, m.shards[t].ev[kind], 0));
                st = launch_decide(*sh.plan, kind, sh.stream);
                if (TFQMRGPU_STATUS_SUCCESS != st) return st;
            }
        }
        TFQ_CUDA(cudaSetDevice(m.shards[0].dev));
        int const slot = i % kRing;
        TFQ_CUDA(cudaMemcpyAsync(&p0.h_ctl[slot], ws<Control>(p0, p0.off_ctl), sizeof(Control), cudaMemcpyDeviceToHost, m.shards[0].stream));
        TFQ_CUDA(cudaEventRecord(p0.ev[slot], m.shards[0].stream));
        ++bodies;
    }
    for (int s = 0; s < n; ++s) {
        Shard &sh = m.shards[s];
        TFQ_CUDA(cudaSetDevice(sh.dev));
        TFQ_CUDA(cudaMemcpyAsync(&m.h_fin[s], ws<Control>(*sh.plan, sh.plan->off_ctl), sizeof(Control), cudaMemcpyDeviceToHost, sh.stream));
    }
    tfqmrgpuStatus_t result = TFQMRGPU_STATUS_SUCCESS;
    double flops = 0;
    for (int s = 0; s < n; ++s) {
        Shard &sh = m.shards[s];
        TFQ_CUDA(cudaSetDevice(sh.dev));
        TFQ_CUDA(cudaStreamSynchronize(sh.stream));
        TFQ_CUDA(cudaGetLastError());
        tfqmrgpuStatus_t const r = solve_finish(*sh.plan, m.h_fin[s], bodies, double(2 + 13*bodies));
        if (0 == s) result = r;
        flops += sh.plan->flops_performed;
    }
    // getInfo of the unsharded problem (the flop formula is linear in the number of right-hand sides)
    p.flops_performed = flops;
    p.flops_performed_all += flops;
    p.residuum_reached = p0.residuum_reached;
    p.iterations_needed = p0.iterations_needed;
    p.solved = true;
    p.stat_probes = p0.stat_probes; p.stat_bodies = bodies; p.stat_launches = double(n)*(2 + 13*bodies);
    p.stat_bound2 = p0.stat_bound2; p.stat_target2 = p0.stat_target2;
    return result;
}

// ---- getMatrix('X'): gather the shards' solutions device to device, then the ordinary conversion on the home device ------------
tfqmrgpuStatus_t multi_gather_x(Plan &p, cudaStream_t homeStream)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    size_t const blockBytes = 2*size_t(p.LM)*p.LN*(('z' == p.precision) ? 8 : 4);
    char *const gx = p.pBuffer + m.off_gx;
    for (Shard &sh : m.shards) {
        Plan &sp = *sh.plan;
        TFQ_CUDA(cudaSetDevice(sh.dev));
        // a shard's column-sorted storage IS the range [g0, g1) of the unsharded column-sorted storage
        if (sh.g1 > sh.g0)
            TFQ_CUDA(cudaMemcpyPeerAsync(gx + size_t(sh.g0)*blockBytes, m.home, sp.pBuffer + sp.off_v[1], sh.dev, size_t(sh.g1 - sh.g0)*blockBytes, sh.stream));
        TFQ_CUDA(cudaEventRecord(sh.ev[2], sh.stream));
    }
    TFQ_CUDA(cudaSetDevice(m.home));
    for (Shard &sh : m.shards) TFQ_CUDA(cudaStreamWaitEvent(homeStream, sh.ev[2], 0));
    return TFQMRGPU_STATUS_SUCCESS;
}

size_t multi_off_gx(Plan const &p) { return p.multi->off_gx; }
size_t multi_off_scratch(Plan const &p) { return p.multi->off_scratch; }

// per right-hand-side status in the unsharded (block column, lane) order
void multi_set_early_freeze(Plan &p)
{
    if (nullptr == p.multi) return;
    for (auto &sh : p.multi->shards) if (sh.plan) sh.plan->early_freeze = p.early_freeze;
}

tfqmrgpuStatus_t multi_rhs_status(Plan &p, int8_t *statusHost)
{
    MultiPlan &m = *p.multi;
    DeviceGuard guard;
    for (Shard &sh : m.shards) {
        Plan &sp = *sh.plan;
        TFQ_CUDA(cudaSetDevice(sh.dev));
        TFQ_CUDA(cudaMemcpyAsync(statusHost + size_t(sh.c0)*p.LN, sp.pBuffer + sp.off_snap, size_t(sp.nCols)*p.LN, cudaMemcpyDeviceToHost, sh.stream));
        TFQ_CUDA(cudaStreamSynchronize(sh.stream));
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace tfq
