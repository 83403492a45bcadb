// The C-ABI of libtfQMRgpu.so: the 21 entry points of tfqmrgpu.h plus the tfqmrgpux_ extensions.
// Behavioural reference: tfQMRgpu/source/tfqmrgpu.cu (cited per function).
#include "tfq_internal.hpp"
#include <curand.h>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <atomic>

namespace tfq {

const int kAllowedBlockSizes[15][2] = { // allowed_block_sizes.h:4-18, same order
    {4, 4}, {4, 5}, {4, 8}, {4, 32}, {8, 8}, {8, 9}, {8, 10}, {8, 32}, {8, 64},
    {16, 16}, {16, 32}, {16, 64}, {32, 32}, {32, 64}, {64, 64}};

bool block_size_allowed(int lm, int ln) {
    for (auto const &bs : kAllowedBlockSizes) if (bs[0] == lm && bs[1] == ln) return true;
    return false;
}

// (read from the upload threads of multi-device plans as well: an atomic, initialised once)
static std::atomic<int> g_verbosity{-1};
int verbosity() {
    int v = g_verbosity.load(std::memory_order_relaxed);
    if (v < 0) {
        char const *e = std::getenv("TFQMRGPU_VERBOSE");
        v = e ? std::max(0, std::atoi(e)) : 0;
        g_verbosity.store(v, std::memory_order_relaxed);
    }
    return v;
}
void set_verbosity(int level) { g_verbosity.store(level < 0 ? 0 : level, std::memory_order_relaxed); }

static inline Plan* P(tfqmrgpuBsrsvPlan_t plan) { return reinterpret_cast<Plan*>(plan); }
static inline char lower(char c) { return char(c | 32); } // the reference's "| IgnoreCase" (util.hxx:12)
// the precision argument of set/getMatrix: a mixed-precision plan ('m') exchanges doubles, under the name 'z' or 'm'
static inline bool data_is_double(Plan const &p, char precision) {
    char const c = lower(precision);
    return ('z' == c) || (p.mixed && 'm' == c);
}

// random shadow vector: cuRAND XORWOW, seed 1234, generated in the caller's block order like the
// reference (linalg.hxx:777-797), then moved into column-sorted storage order
tfqmrgpuStatus_t fill_v3(Plan &p, cudaStream_t stream) {
    size_t const n = size_t(p.nnzbX)*2*p.LM*p.LN;
    float *const scratch = ws<float>(p, p.off_v[9]);
    curandGenerator_t gen;
    if (CURAND_STATUS_SUCCESS != curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT)) return TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    tfqmrgpuStatus_t st = TFQMRGPU_STATUS_SUCCESS;
    if (CURAND_STATUS_SUCCESS != curandSetStream(gen, stream)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    else if (CURAND_STATUS_SUCCESS != curandSetPseudoRandomGeneratorSeed(gen, 1234ull)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    else if (CURAND_STATUS_SUCCESS != curandGenerateUniform(gen, scratch, n)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    if (TFQMRGPU_STATUS_SUCCESS == st) st = permute_v3(p, ws<float>(p, p.off_v[3]), scratch, true, stream);
    cudaStreamSynchronize(stream); // the generator must outlive its asynchronous work
    curandDestroyGenerator(gen);
    return st;
}

// shared by setMatrix/getMatrix/getVector: parse layout and transposition (tfqmrgpu.cu:480-501)
static tfqmrgpuStatus_t parse_layout_trans(tfqmrgpuDataLayout_t layout, char transposition, bool &trans, double &scal_imag) {
    switch (layout) {
        case TFQMRGPU_LAYOUT_RRRRIIII: case TFQMRGPU_LAYOUT_RIRIRIRI: case TFQMRGPU_LAYOUT_RRIIRRII: break;
        default: return TFQMRGPU_DATALAYOUT_UNKNOWN + TFQMRGPU_CODE_LINE*layout;
    }
    scal_imag = 1;
    switch (lower(transposition)) {
        case 'h': case 'c': scal_imag = -1; trans = true; break;
        case '*':           scal_imag = -1; trans = false; break;
        case 't': trans = true; break;
        case 'n': trans = false; break;
        default: return TFQ_ERRC(TFQMRGPU_TANSPOSITION_UNKNOWN, lower(transposition));
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

static tfqmrgpuStatus_t download_vector(Handle *h, Plan &p, size_t off_vec, void *val, char precision, char transposition,
                                        tfqmrgpuDataLayout_t layout) {
    bool trans = false; double scal_imag = 1;
    tfqmrgpuStatus_t st = parse_layout_trans(layout, transposition, trans, scal_imag);
    if (st) return st;
    if (p.nnzbX < 1) return TFQMRGPU_STATUS_SUCCESS;
    if (nullptr == p.pBuffer || nullptr == val) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    bool const is_double = ('z' == p.precision);
    if (data_is_double(p, precision) != is_double) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, precision);
    if (p.multi) {      // several devices: gather the shards' X device to device, convert on the home device
        if (off_vec != p.off_v[1]) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
        st = multi_gather_x(p, h->stream);
        if (st) return st;
        char *const gx = p.pBuffer + multi_off_gx(p), *const scratch = p.pBuffer + multi_off_scratch(p);
        st = convert_permuted(p, scratch, gx, p.nnzbX, p.LM, p.LN, is_double, layout, trans, scal_imag, false, h->stream);
        if (st) return st;
        TFQ_CUDA(cudaMemcpyAsync(val, scratch, p.vecBytes, cudaMemcpyDeviceToHost, h->stream));
        TFQ_CUDA(cudaStreamSynchronize(h->stream));
        return TFQMRGPU_STATUS_SUCCESS;
    }
    // out of place through a scratch vector, so the device copy keeps its solver layout
    // (the reference transposes X in place and leaves it in host layout, tfqmrgpu.cu:552-600)
    size_t const off_scratch = (off_vec == p.off_v[9]) ? p.off_v[8] : p.off_v[9];
    st = convert_permuted(p, p.pBuffer + off_scratch, p.pBuffer + off_vec, p.nnzbX, p.LM, p.LN, is_double, layout, trans, scal_imag, false, h->stream);
    if (st) return st;
    TFQ_CUDA(cudaMemcpyAsync(val, p.pBuffer + off_scratch, p.vecBytes, cudaMemcpyDeviceToHost, h->stream));
    TFQ_CUDA(cudaStreamSynchronize(h->stream));
    return TFQMRGPU_STATUS_SUCCESS;
}

} // namespace tfq

using namespace tfq;

extern "C" {

// ---- handle / stream: tfqmrgpu.cu:110-134 -----------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpuCreateHandle(tfqmrgpuHandle_t *handle) {
    if (nullptr == handle) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nullptr != *handle) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    Handle *h = new (std::nothrow) Handle();
    if (nullptr == h) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
    *handle = h;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpuDestroyHandle(tfqmrgpuHandle_t handle) {
    if (nullptr == handle) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    delete static_cast<Handle*>(handle);
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpuSetStream(tfqmrgpuHandle_t handle, cudaStream_t const streamId) {
    if (nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    static_cast<Handle*>(handle)->stream = streamId;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpuGetStream(tfqmrgpuHandle_t handle, cudaStream_t *streamId) {
    if (nullptr == handle || nullptr == streamId) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    *streamId = static_cast<Handle*>(handle)->stream;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- workspace: tfqmrgpu.cu:682-698 -----------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpuCreateWorkspace(void* *pBuffer, size_t const pBufferSizeInBytes, char const memType) {
    if (nullptr == pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaError_t const err = ('m' == lower(memType)) ? cudaMallocManaged(pBuffer, pBufferSizeInBytes)
                                                    : cudaMalloc(pBuffer, pBufferSizeInBytes);
    if (cudaSuccess != err) { cudaGetLastError(); return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED); }
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpuDestroyWorkspace(void* pBuffer) {
    return (cudaSuccess == cudaFree(pBuffer)) ? TFQMRGPU_STATUS_SUCCESS : TFQ_ERR(TFQMRGPU_POINTER_INVALID);
}

// ---- block sizes: tfqmrgpu.cu:75-106 ----------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_allowedBlockSizes(int32_t *number, int32_t *blockSizes, int const arrayLength) {
    if (nullptr == number) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nullptr == blockSizes) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (0 != *number) for (int i = 0; i < arrayLength; ++i) blockSizes[i] = 0; // tfqmrgpu.cu:83
    int n = 0, written = 0;
    for (auto const &bs : kAllowedBlockSizes) {
        ++n;
        if (2*n < arrayLength) { blockSizes[2*written] = bs[0]; blockSizes[2*written + 1] = bs[1]; ++written; } // tfqmrgpu.cu:88
    }
    *number = n;
    return (n == written) ? TFQMRGPU_STATUS_SUCCESS : TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
}
tfqmrgpuStatus_t tfqmrgpu_bsrsv_blockSizeMissing(int const ldA, int const ldB) {
    if (block_size_allowed(ldA, ldB)) return TFQMRGPU_STATUS_SUCCESS;
    return TFQMRGPU_BLOCKSIZE_MISSING + TFQMRGPU_CODE_CHAR*ldA + TFQMRGPU_CODE_LINE*ldB;
}

// ---- plan: tfqmrgpu.cu:136-361 ----------------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_createPlan(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t *plan, int const mb,
    int32_t const *bsrRowPtrA, int const nnzbA, int32_t const *bsrColIndA,
    int32_t const *bsrRowPtrX, int const nnzbX, int32_t const *bsrColIndX,
    int32_t const *bsrRowPtrB, int const nnzbB, int32_t const *bsrColIndB,
    int const indexOffset, int const echo)
{
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (nullptr != *plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);             // tfqmrgpu.cu:161
    if (mb < 1) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);                    // tfqmrgpu.cu:166-172
    if (nnzbX < 1) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nnzbB > nnzbX) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nnzbA < 0 || nnzbB < 0) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if ((long long)nnzbA > (long long)mb*mb) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (!bsrRowPtrA || !bsrRowPtrX || !bsrRowPtrB) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if ((nnzbA && !bsrColIndA) || !bsrColIndX || (nnzbB && !bsrColIndB)) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (nnzbA != bsrRowPtrA[mb] - bsrRowPtrA[0]) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nnzbX != bsrRowPtrX[mb] - bsrRowPtrX[0]) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nnzbB != bsrRowPtrB[mb] - bsrRowPtrB[0]) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);

    Plan *p = new (std::nothrow) Plan();
    if (nullptr == p) return TFQ_ERR(TFQMRGPU_STATUS_ALLOCATION_FAILED);
    p->mb = mb; p->nnzbA = nnzbA; p->nnzbX = nnzbX; p->nnzbB = nnzbB; p->indexOffset = indexOffset;
    cudaStream_t const stream = handle ? static_cast<Handle*>(handle)->stream : nullptr;
    tfqmrgpuStatus_t const st = plan_analyse(*p, stream, bsrRowPtrA, bsrColIndA, bsrRowPtrX, bsrColIndX, bsrRowPtrB, bsrColIndB, echo);
    if (TFQMRGPU_STATUS_SUCCESS != st) { plan_release(*p); delete p; return st; }
    // TFQMRGPU_NUM_GPUS=N: shard the right-hand-side block columns over devices 0..N-1 (same as tfqmrgpux_bsrsv_setDevices)
    if (char const *e = std::getenv("TFQMRGPU_NUM_GPUS")) {
        int const ngpu = std::atoi(e);
        if (ngpu > 1) {
            tfqmrgpuStatus_t const mst = multi_set_devices(*p, ngpu, nullptr);
            if (TFQMRGPU_STATUS_SUCCESS != mst) { plan_release(*p); delete p; return mst; }
        }
    }
    *plan = reinterpret_cast<tfqmrgpuBsrsvPlan_t>(p);
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpu_bsrsv_destroyPlan(tfqmrgpuHandle_t, tfqmrgpuBsrsvPlan_t plan) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    plan_release(*P(plan));
    delete P(plan);
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- bufferSize: tfqmrgpu.cu:364-412 ----------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_bufferSize(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan,
    int const ldA, int const blockDim, int const ldB, int const RhsBlockDim, char const precision, size_t *pBufferSizeInBytes)
{
    int const LM = ldA, LN = ldB;
    if (LM != blockDim) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (LM > LN) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (LN != RhsBlockDim) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    char prec;
    switch (lower(precision)) { // tfqmrgpu.cu:383-390
        case 'f': case 'c': prec = 'c'; break;
        case 'm': prec = 'm'; break;
        case 'd': case 'z': prec = 'z'; break;
        default: prec = 'z';
    }
    // Whatever happens below, a previously registered workspace was laid out for the OLD configuration: forget it first, so
    // that no later setMatrix/solve can run with a stale layout after a failed re-configuration.
    p.pBuffer = nullptr; p.bufferBytes = 0; p.v3_ready = false; p.configured = false;
    plan_drop_graph(p);
    if (nullptr == pBufferSizeInBytes) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    *pBufferSizeInBytes = 0;
    if (!block_size_allowed(LM, LN)) return TFQMRGPU_BLOCKSIZE_MISSING + TFQMRGPU_CODE_CHAR*LM + TFQMRGPU_CODE_LINE*LN; // tfqmrgpu.cu:70
    if ('m' != prec) mixed_destroy(p);
    if ('m' == prec) {
        // The reference accepts 'm' here (tfqmrgpu.cu:386) and fails in solve (tfqmrgpu.cu:42-44: the case is commented out).
        // This library implements it (mixed.cu); TFQMRGPU_MIXED=0 restores the reference's refusal.
        char const *e = std::getenv("TFQMRGPU_MIXED");
        if ((e && '0' == e[0]) || p.multi) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, prec);                      // tfqmrgpu.cu:44
        tfqmrgpuStatus_t const xst = mixed_buffer_size(p, static_cast<Handle*>(handle)->stream, LM, LN, pBufferSizeInBytes);
        if (TFQMRGPU_STATUS_SUCCESS != xst) { p.bufferBytes = 0; *pBufferSizeInBytes = 0; return xst; }
        p.configured = true;
        return TFQMRGPU_STATUS_SUCCESS;
    }
    if (p.multi) {
        tfqmrgpuStatus_t const mst = multi_buffer_size(p, LM, LN, prec, pBufferSizeInBytes);
        if (TFQMRGPU_STATUS_SUCCESS != mst) { p.bufferBytes = 0; *pBufferSizeInBytes = 0; return mst; }
        p.configured = true;
        return TFQMRGPU_STATUS_SUCCESS;
    }
    tfqmrgpuStatus_t const st = plan_configure(p, static_cast<Handle*>(handle)->stream, LM, LN, prec);
    if (TFQMRGPU_STATUS_SUCCESS != st) { p.bufferBytes = 0; return st; }
    p.configured = true;
    *pBufferSizeInBytes = p.bufferBytes;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- setBuffer / getBuffer: tfqmrgpu.cu:415-462 -----------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_setBuffer(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, void* const pBuffer) {
    if (nullptr == pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (!p.configured || 0 == p.bufferBytes) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR); // bufferSize has not been called (successfully)
    if (size_t(pBuffer) & 255) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);  // 2^TFQMRGPU_MEMORY_ALIGNMENT
    if (p.multi) return multi_set_buffer(p, pBuffer);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    p.pBuffer = static_cast<char*>(pBuffer);
    plan_drop_graph(p);               // the captured iteration body holds pointers into the old workspace
    TFQ_CUDA(cudaMemsetAsync(p.pBuffer + p.off_zero, 0, p.bufferBytes - 256 - p.off_zero, stream)); // zero block, scalars, tickets, control
    tfqmrgpuStatus_t const st = fill_v3(p, stream);
    if (TFQMRGPU_STATUS_SUCCESS != st) return st;
    p.v3_ready = true;
    if (p.mixed) return mixed_set_buffer(p, stream);
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpu_bsrsv_getBuffer(tfqmrgpuHandle_t, tfqmrgpuBsrsvPlan_t plan, void* *pBuffer) {
    if (nullptr == plan || nullptr == pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    *pBuffer = P(plan)->pBuffer;
    if (nullptr == *pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    return TFQMRGPU_STATUS_SUCCESS;
}

// Blocks [b0, b0 + nb) of A: host -> device into the A window and layout conversion in place (+ the block maxima for the row scales of the
// fp16-pair operand, xop.cu).  valBlocks points to the first of these blocks in the caller's host array.  Large ranges go in chunks on
// a copy stream while the layout kernel converts the previous chunk on the caller's stream: the conversion hides behind the PCIe transfer.
static tfqmrgpuStatus_t upload_a_blocks(Plan &p, cudaStream_t stream, void const *valBlocks, uint32_t b0, uint32_t nb, bool is_double,
                                        tfqmrgpuDataLayout_t layout, bool trans, double scal_imag)
{
    size_t const s = is_double ? 8 : 4;
    size_t const blockBytes = 2*size_t(p.LM)*p.LM*s, bytes = size_t(nb)*blockBytes;
    char *const dst = p.pBuffer + p.off_A + size_t(b0)*blockBytes;
    size_t const chunkBytes = size_t(256) << 20;
    auto converted = [&](uint32_t c0, uint32_t cn) -> tfqmrgpuStatus_t {      // blocks [b0 + c0, b0 + c0 + cn)
        tfqmrgpuStatus_t const cst = convert_inplace(p, dst + size_t(c0)*blockBytes, cn, p.LM, p.LM, is_double, layout, trans, scal_imag, stream);
        if (cst || !p.use_tc16) return cst;
        return launch_aop_blockmax(p, b0 + c0, cn, stream);
    };
    if (bytes <= chunkBytes) {
        TFQ_CUDA(cudaMemcpyAsync(dst, valBlocks, bytes, cudaMemcpyHostToDevice, stream));
        return converted(0, nb);
    }
    if (nullptr == p.copy_stream) TFQ_CUDA(cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking));
    if (nullptr == p.chunk_ev[0]) {
        for (auto &e : p.chunk_ev) TFQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    TFQ_CUDA(cudaEventRecord(p.chunk_ev[0], stream));        // earlier work on the caller's stream may still read A (or must precede the upload)
    TFQ_CUDA(cudaStreamWaitEvent(p.copy_stream, p.chunk_ev[0], 0));
    uint32_t const blocksPerChunk = uint32_t(chunkBytes/blockBytes);
    tfqmrgpuStatus_t cst = TFQMRGPU_STATUS_SUCCESS;
    int turn = 0;
    for (uint32_t c0 = 0; c0 < nb && TFQMRGPU_STATUS_SUCCESS == cst; c0 += blocksPerChunk, ++turn) {
        uint32_t const cn = std::min(blocksPerChunk, nb - c0);
        size_t const off = size_t(c0)*blockBytes;
        TFQ_CUDA(cudaMemcpyAsync(dst + off, static_cast<char const*>(valBlocks) + off, size_t(cn)*blockBytes, cudaMemcpyHostToDevice, p.copy_stream));
        // (re-recording an event that a stream still waits for is fine: the wait took the earlier record)
        cudaEvent_t const ev = p.chunk_ev[1 + (turn & 1)];
        TFQ_CUDA(cudaEventRecord(ev, p.copy_stream));
        TFQ_CUDA(cudaStreamWaitEvent(stream, ev, 0));
        cst = converted(c0, cn);
    }
    return cst;
}

// ---- setMatrix / getMatrix: tfqmrgpu.cu:467-645 -----------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_setMatrix(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, char const var, void const *val,
    char const precision, int const, int const, char const transposition, tfqmrgpuDataLayout_t const layout)
{
    bool trans = false; double scal_imag = 1;
    tfqmrgpuStatus_t st = parse_layout_trans(layout, transposition, trans, scal_imag);
    if (st) return st;
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    bool const is_double = ('z' == p.precision);
    size_t const s = is_double ? 8 : 4;
    char const v = lower(var);
    uint32_t nnzb = 0;
    switch (v) {
        case 'a': nnzb = p.nnzbA; trans = !trans; break; // A is stored transposed [k][i] (tfqmrgpu.cu:509-520)
        case 'b': nnzb = p.nnzbB; break;
        case 'x': nnzb = p.nnzbX; break;
        default: return TFQ_ERRC(TFQMRGPU_VARIABLENAME_UNKNOWN, var);
    }
    if (nnzb < 1) return TFQMRGPU_STATUS_SUCCESS;
    if (nullptr == p.pBuffer || nullptr == val || !p.configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if ('z' != p.precision && 'c' != p.precision) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, p.precision);
    if (data_is_double(p, precision) != is_double) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, precision); // tfqmrgpu.cu:538-542
    if (p.multi) return multi_set_matrix(p, v, val, precision, transposition, layout, trans, scal_imag);
    if ('a' == v) {
        st = upload_a_blocks(p, stream, val, 0, nnzb, is_double, layout, trans, scal_imag);
        if (TFQMRGPU_STATUS_SUCCESS == st && p.mixed) return mixed_after_set_a(p, stream);
        if (st || !p.use_tc16) return st;
        return launch_aop_convert(p, stream);
    }
    if ('b' == v) {
        char *const dst = p.pBuffer + p.off_B;
        TFQ_CUDA(cudaMemcpyAsync(dst, val, size_t(nnzb)*2*p.LM*p.LN*s, cudaMemcpyHostToDevice, stream));
        return convert_inplace(p, dst, nnzb, p.LM, p.LN, is_double, layout, trans, scal_imag, stream);
    }
    // 'x': accepted like in the reference; solve() discards the initial guess (core.hxx:125) unless tfqmrgpux_bsrsv_setInitialGuess is on
    char *const scratch = p.pBuffer + p.off_v[9];
    TFQ_CUDA(cudaMemcpyAsync(scratch, val, p.vecBytes, cudaMemcpyHostToDevice, stream));
    st = convert_permuted(p, p.pBuffer + p.off_v[1], scratch, nnzb, p.LM, p.LN, is_double, layout, trans, scal_imag, true, stream);
    // the tensor-core product reads X as an fp16-pair operand: make it with the upload, like A's (tfqmrgpux_bsrsv_multiply then is the
    // bare product, the role of `bench_tfqmrgpu multi` whose operands are resident before the timed launches)
    p.xop_of_x = false;
    if (TFQMRGPU_STATUS_SUCCESS == st && p.use_tc16) { st = launch_xop(p, p.pBuffer + p.off_v[1], -1, stream); p.xop_of_x = (TFQMRGPU_STATUS_SUCCESS == st); }
    return st;
}

tfqmrgpuStatus_t tfqmrgpu_bsrsv_getMatrix(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, char const var, void *val,
    char const precision, int const, int const, char const transposition, tfqmrgpuDataLayout_t const layout)
{
    if ('x' != lower(var)) return TFQ_ERRC(TFQMRGPU_UNDOCUMENTED_ERROR, var); // only X can be downloaded (tfqmrgpu.cu:635-643)
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    return download_vector(static_cast<Handle*>(handle), *P(plan), P(plan)->off_v[1], val, precision, transposition, layout);
}

// ---- solve / getInfo: tfqmrgpu.cu:648-679 -----------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpu_bsrsv_solve(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, double const threshold, int const maxIterations) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (P(plan)->multi) {
        if (nullptr == P(plan)->pBuffer || !P(plan)->configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
        if (P(plan)->initial_guess || P(plan)->precond) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);   // single-device plans only
        return multi_solve(*P(plan), threshold, maxIterations);
    }
    if (P(plan)->mixed) return mixed_solve(*P(plan), static_cast<Handle*>(handle)->stream, threshold, maxIterations);
    return solve(*P(plan), static_cast<Handle*>(handle)->stream, threshold, maxIterations);
}

tfqmrgpuStatus_t tfqmrgpu_bsrsv_getInfo(tfqmrgpuHandle_t, tfqmrgpuBsrsvPlan_t plan, double *residuum_reached,
    int32_t *iterations_needed, double *flops_performed, double *flops_performed_all)
{
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    int any = 0;
    if (residuum_reached)    { ++any; *residuum_reached    = p.residuum_reached; }
    if (iterations_needed)   { ++any; *iterations_needed   = p.iterations_needed; }
    if (flops_performed)     { ++any; *flops_performed     = p.flops_performed; }
    if (flops_performed_all) { ++any; *flops_performed_all = p.flops_performed_all; }
    return any ? TFQMRGPU_STATUS_SUCCESS : TFQMRGPU_STATUS_NO_INFO_PASSED;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_setProfiling(tfqmrgpuBsrsvPlan_t plan, int on) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    P(plan)->profile = (0 != on);
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getSolveProfile(tfqmrgpuBsrsvPlan_t plan, double profile[8]) {
    if (nullptr == plan || nullptr == profile) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    profile[0] = p.prof_solve_ms; profile[1] = p.prof_spmm_ms; profile[2] = p.prof_spmm_launches; profile[3] = p.prof_iterations;
    profile[4] = p.stat_launches; profile[5] = p.stat_probes; profile[6] = 0; profile[7] = 0;
    return TFQMRGPU_STATUS_SUCCESS;
}

} // extern "C"

// ---- quick starters: tfqmrgpu.cu:702-821 ------------------------------------------------------------
namespace tfq {
template <typename real_t>
static tfqmrgpuStatus_t one_shot(int mb, int ldA, int ldB,
    int32_t const *rowPtrA, int nnzbA, int32_t const *colIndA, real_t const *Amat, char transA,
    int32_t const *rowPtrX, int nnzbX, int32_t const *colIndX, real_t *Xmat, char transX,
    int32_t const *rowPtrB, int nnzbB, int32_t const *colIndB, real_t const *Bmat, char transB,
    int32_t *iterations, float *residual, int indexOffset, int echo)
{
    char const zoc = (8 == sizeof(real_t)) ? 'z' : 'c';
    if (echo > 0) std::printf("# tfqmrgpu_bsrsv_%c: mb= %d, ldA= %d, ldB= %d, iterations= %d, residual= %.1e\n",
                              zoc, mb, ldA, ldB, iterations ? *iterations : -1, residual ? double(*residual) : -1.);
    tfqmrgpuHandle_t handle = nullptr;
    tfqmrgpuBsrsvPlan_t plan = nullptr;
    void *buffer = nullptr;
    tfqmrgpuStatus_t stat = TFQMRGPU_STATUS_SUCCESS;
    char const *where = "";
    // unlike the reference, every exit path releases what was acquired
#define STEP(name, call) if (TFQMRGPU_STATUS_SUCCESS == stat) { stat = (call); where = name; }
    STEP("tfqmrgpuCreateHandle", tfqmrgpuCreateHandle(&handle))
    STEP("tfqmrgpuSetStream", tfqmrgpuSetStream(handle, nullptr))
    STEP("tfqmrgpu_bsrsv_createPlan", tfqmrgpu_bsrsv_createPlan(handle, &plan, mb, rowPtrA, nnzbA, colIndA, rowPtrX, nnzbX, colIndX,
                                                                 rowPtrB, nnzbB, colIndB, indexOffset, echo))
    size_t bytes = 0;
    STEP("tfqmrgpu_bsrsv_bufferSize", tfqmrgpu_bsrsv_bufferSize(handle, plan, ldA, ldA, ldB, ldB, zoc, &bytes))
    STEP("tfqmrgpuCreateWorkspace", tfqmrgpuCreateWorkspace(&buffer, bytes, 'd'))
    STEP("tfqmrgpu_bsrsv_setBuffer", tfqmrgpu_bsrsv_setBuffer(handle, plan, buffer))
    STEP("tfqmrgpu_bsrsv_setMatrix('A')", tfqmrgpu_bsrsv_setMatrix(handle, plan, 'A', Amat, zoc, ldA, ldA, transA, TFQMRGPU_LAYOUT_RIRIRIRI))
    STEP("tfqmrgpu_bsrsv_setMatrix('B')", tfqmrgpu_bsrsv_setMatrix(handle, plan, 'B', Bmat, zoc, ldB, ldA, transB, TFQMRGPU_LAYOUT_RIRIRIRI))
    double const threshold = residual ? double(*residual) : 1e-9;
    int const maxiter = iterations ? *iterations : 200;
    STEP("tfqmrgpu_bsrsv_solve", tfqmrgpu_bsrsv_solve(handle, plan, threshold, maxiter))
    double residuum = 0, flops = 0, flops_all = 0;
    int32_t needed = 0;
    STEP("tfqmrgpu_bsrsv_getInfo", tfqmrgpu_bsrsv_getInfo(handle, plan, &residuum, &needed, &flops, &flops_all))
    if (TFQMRGPU_STATUS_SUCCESS == stat) {
        if (echo > 1) std::printf("# tfQMRgpu needed %d iterations to converge to %.1e using %g GFlop\n", needed, residuum, flops*1e-9);
        if (residual) *residual = float(residuum);
        if (iterations) *iterations = needed;
    }
    STEP("tfqmrgpu_bsrsv_getMatrix", tfqmrgpu_bsrsv_getMatrix(handle, plan, 'X', Xmat, zoc, ldB, ldA, transX, TFQMRGPU_LAYOUT_RIRIRIRI))
#undef STEP
    if (stat && echo > 0) std::printf("# tfqmrgpu_bsrsv_%c: %s returned %d\n", zoc, where, stat);
    if (buffer) tfqmrgpuDestroyWorkspace(buffer);
    if (plan) tfqmrgpu_bsrsv_destroyPlan(handle, plan);
    if (handle) tfqmrgpuDestroyHandle(handle);
    return stat;
}
} // namespace tfq

extern "C" {

tfqmrgpuStatus_t tfqmrgpu_bsrsv_z(int mb, int ldA, int ldB,
    int32_t const* rowPtrA, int nnzbA, int32_t const* colIndA, double const* Amat, char transA,
    int32_t const* rowPtrX, int nnzbX, int32_t const* colIndX, double* Xmat, char transX,
    int32_t const* rowPtrB, int nnzbB, int32_t const* colIndB, double const* Bmat, char transB,
    int32_t *iterations, float *residual, int indexOffset, int echo) {
    return one_shot<double>(mb, ldA, ldB, rowPtrA, nnzbA, colIndA, Amat, transA, rowPtrX, nnzbX, colIndX, Xmat, transX,
                            rowPtrB, nnzbB, colIndB, Bmat, transB, iterations, residual, indexOffset, echo);
}
tfqmrgpuStatus_t tfqmrgpu_bsrsv_c(int mb, int ldA, int ldB,
    int32_t const* rowPtrA, int nnzbA, int32_t const* colIndA, float const* Amat, char transA,
    int32_t const* rowPtrX, int nnzbX, int32_t const* colIndX, float* Xmat, char transX,
    int32_t const* rowPtrB, int nnzbB, int32_t const* colIndB, float const* Bmat, char transB,
    int32_t *iterations, float *residual, int indexOffset, int echo) {
    return one_shot<float>(mb, ldA, ldB, rowPtrA, nnzbA, colIndA, Amat, transA, rowPtrX, nnzbX, colIndX, Xmat, transX,
                           rowPtrB, nnzbB, colIndB, Bmat, transB, iterations, residual, indexOffset, echo);
}

// ---- extensions (tfqmrgpu_b200_ext.h) ---------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpux_getVersion(int *major, int *minor, int *patch) {
    if (major) *major = 0; if (minor) *minor = 1; if (patch) *patch = 0;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_setVerbosity(int level) { set_verbosity(level); return TFQMRGPU_STATUS_SUCCESS; }

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getPlanArray(tfqmrgpuBsrsvPlan_t plan, int kind, void *out, size_t *count) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    void const *src = nullptr; size_t n = 0, es = 4; bool host = false;
    switch (kind) {
        case 0: src = p.d_starts;  n = size_t(p.nnzbX) + 1; break;
        case 1: src = p.d_pairs;   n = 2*size_t(p.nPairs); break;
        case 2: src = p.d_subset;  n = size_t(p.nnzbB); break;
        case 3: src = p.d_colindx; n = size_t(p.nnzbX); es = 2; break;
        case 4: src = p.d_perm;    n = size_t(p.nnzbX); break;
        case 5: src = p.h_colstart.data(); n = p.h_colstart.size(); host = true; break;
        default: return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    }
    if (count) *count = n;
    if (out && n) {
        if (host) std::memcpy(out, src, n*es);
        else TFQ_CUDA(cudaMemcpy(out, src, n*es, cudaMemcpyDeviceToHost));
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getPlanInfo(tfqmrgpuBsrsvPlan_t plan, int64_t info[16]) {
    if (nullptr == plan || nullptr == info) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    int64_t const v[16] = {p.nnzbX, p.nnzbB, p.nnzbA, int64_t(p.nCols), int64_t(p.nPairs), p.LM, p.LN, p.mixed ? int64_t('m') : int64_t(p.precision),
                           int64_t(p.nTiles), int64_t(p.nUnits), int64_t(p.gmax), int64_t(p.nEntries), p.mb, int64_t(p.use_tc16), int64_t(p.use_dmma), int64_t(p.use_small)};
    std::memcpy(info, v, sizeof(v));
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_setV3(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, float const *v3, int onDevice) {
    if (nullptr == plan || nullptr == handle || nullptr == v3) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);       // (a multi-device plan slices the reference's stream itself)
    if (nullptr == p.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    size_t const bytes = size_t(p.nnzbX)*2*p.LM*p.LN*sizeof(float);
    float *const scratch = ws<float>(p, p.off_v[9]);
    TFQ_CUDA(cudaMemcpyAsync(scratch, v3, bytes, onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    tfqmrgpuStatus_t const st = permute_v3(p, ws<float>(p, p.off_v[3]), scratch, true, stream);
    if (st) return st;
    TFQ_CUDA(cudaStreamSynchronize(stream));
    p.v3_ready = true;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getV3(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, float *v3Host) {
    if (nullptr == plan || nullptr == handle || nullptr == v3Host) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    if (nullptr == p.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    float *const scratch = ws<float>(p, p.off_v[9]);
    tfqmrgpuStatus_t const st = permute_v3(p, scratch, ws<float const>(p, p.off_v[3]), false, stream);
    if (st) return st;
    TFQ_CUDA(cudaMemcpyAsync(v3Host, scratch, size_t(p.nnzbX)*2*p.LM*p.LN*sizeof(float), cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_randomShadow(tfqmrgpuHandle_t handle, float *devOut, size_t n) {
    if (nullptr == handle || nullptr == devOut) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    curandGenerator_t gen;
    if (CURAND_STATUS_SUCCESS != curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT)) return TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    tfqmrgpuStatus_t st = TFQMRGPU_STATUS_SUCCESS;
    if (CURAND_STATUS_SUCCESS != curandSetStream(gen, stream)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    else if (CURAND_STATUS_SUCCESS != curandSetPseudoRandomGeneratorSeed(gen, 1234ull)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    else if (CURAND_STATUS_SUCCESS != curandGenerateUniform(gen, devOut, n)) st = TFQ_ERR(TFQMRGPU_STATUS_RANDOM_GEN_FAILED);
    cudaStreamSynchronize(stream);
    curandDestroyGenerator(gen);
    return st;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_setOperator(tfqmrgpuBsrsvPlan_t plan, tfqmrgpuxOperator_t op, void *ctx) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    p.user_op = op; p.user_ctx = op ? ctx : nullptr;
    plan_drop_graph(p);          // a captured iteration body holds the built-in product
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_setPreconditioner(tfqmrgpuBsrsvPlan_t plan, tfqmrgpuxOperator_t op, void *ctx) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (op && (p.multi || p.mixed)) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    p.precond = op; p.precond_ctx = op ? ctx : nullptr;
    plan_drop_graph(p);          // a captured iteration body multiplies v6 itself
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_multiply(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int nrep) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    if (nullptr == p.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    for (int r = 0; r < nrep; ++r) {
        tfqmrgpuStatus_t const st = p.xop_of_x ? launch_spmm_operand_ready(p, p.pBuffer + p.off_v[9], p.pBuffer + p.off_v[1], -1, static_cast<Handle*>(handle)->stream)
                                                : launch_spmm(p, p.pBuffer + p.off_v[9], p.pBuffer + p.off_v[1], -1, static_cast<Handle*>(handle)->stream);
        if (st) return st;
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getVector(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, char var, void *val,
    char precision, char trans, tfqmrgpuDataLayout_t layout) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    size_t off;
    switch (lower(var)) {
        case 'x': off = p.off_v[1]; break;
        case 'y': off = p.off_v[9]; break;
        default: return TFQ_ERRC(TFQMRGPU_VARIABLENAME_UNKNOWN, var);
    }
    return download_vector(static_cast<Handle*>(handle), p, off, val, precision, trans, layout);
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getWindow(tfqmrgpuBsrsvPlan_t plan, char var, size_t *offset, size_t *length) {
    if (nullptr == plan || nullptr == offset || nullptr == length) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    size_t const s = ('z' == p.precision) ? 8 : 4;
    switch (lower(var)) {
        case 'x': *offset = p.off_v[1]; *length = p.vecBytes; break;
        case 'y': *offset = p.off_v[9]; *length = p.vecBytes; break;
        case '3': *offset = p.off_v[3]; *length = size_t(p.nnzbX)*2*p.LM*p.LN*4; break;
        case 'a': *offset = p.off_A; *length = size_t(p.nnzbA)*2*p.LM*p.LM*s + (p.use_tc16 ? size_t(p.mb)*4 : 0); break; // (+ row scales)
        case 'b': *offset = p.off_B; *length = size_t(p.nnzbB)*2*p.LM*p.LN*s; break;
        default: return TFQ_ERRC(TFQMRGPU_VARIABLENAME_UNKNOWN, var);
    }
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getRhsStatus(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int8_t *statusHost) {
    if (nullptr == plan || nullptr == handle || nullptr == statusHost) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (P(plan)->multi) return (nullptr == P(plan)->pBuffer) ? TFQ_ERR(TFQMRGPU_POINTER_INVALID) : multi_rhs_status(*P(plan), statusHost);
    Plan const &p = P(plan)->mixed ? *mixed_inner(*P(plan)) : *P(plan);     // (mixed precision: of the last fp32 pass)
    if (nullptr == p.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    // snap = the status array as the reference's host sees it after the last completed iteration / probe
    // (status itself already carries the dec35 verdict of the following iteration, which is fused into K4)
    TFQ_CUDA(cudaMemcpyAsync(statusHost, p.pBuffer + p.off_snap, size_t(p.nCols)*p.LN, cudaMemcpyDeviceToHost, stream));
    TFQ_CUDA(cudaStreamSynchronize(stream));
    return TFQMRGPU_STATUS_SUCCESS;
}

tfqmrgpuStatus_t tfqmrgpux_bsrsv_getSolveStats(tfqmrgpuBsrsvPlan_t plan, double stats[8]) {
    if (nullptr == plan || nullptr == stats) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    stats[0] = p.stat_probes; stats[1] = p.stat_launches; stats[2] = p.stat_bodies; stats[3] = p.stat_ms;
    stats[4] = p.stat_bound2; stats[5] = p.stat_target2; stats[6] = 0; stats[7] = 0;
    return TFQMRGPU_STATUS_SUCCESS;
}

// ---- several GPUs (tfqmrgpu_b200_ext.h) ------------------------------------------------------------------
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setDevices(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int nDevices, int const *devices) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    if (nDevices < 1 || nDevices > kMaxDevices) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    tfqmrgpuStatus_t const st = multi_set_devices(*P(plan), nDevices, devices);
    if (TFQMRGPU_STATUS_SUCCESS != st) multi_destroy(*P(plan));
    return st;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getDevices(tfqmrgpuBsrsvPlan_t plan, int *nDevices, int *devices, int arrayLength) {
    if (nullptr == plan || nullptr == nDevices) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    *nDevices = multi_get_devices(*P(plan), devices, arrayLength);
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setShardExchange(tfqmrgpuBsrsvPlan_t plan, int shard, int nShards, int64_t nRhsGlobal,
    double *slots, tfqmrgpuxExchange_t hook, void *ctx) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    plan_drop_graph(p);              // the captured iteration body decides alone
    if (nullptr == hook || nShards <= 1) { p.exch = Exchange(); return TFQMRGPU_STATUS_SUCCESS; }
    if (nullptr == slots || shard < 0 || shard >= nShards) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    p.exch.nshards = nShards; p.exch.shard = shard; p.exch.nrhs_global = nRhsGlobal; p.exch.slots = slots;
    p.exch.hook = hook; p.exch.hook_ctx = ctx; p.exch.parity = 0;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setShardHints(tfqmrgpuBsrsvPlan_t plan, int64_t tileBlocksHint, int32_t maxColsPerRowHint) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    P(plan)->tile_blocks_hint = (tileBlocksHint > 0) ? size_t(tileBlocksHint) : 0;     // take effect with the next bufferSize
    P(plan)->max_cols_hint = (maxColsPerRowHint > 0) ? maxColsPerRowHint : 0;
    return TFQMRGPU_STATUS_SUCCESS;
}
static tfqmrgpuStatus_t matrix_part(Plan const &p, int part, int nParts, int64_t info[6], int &row0, int &row1) {
    if (nParts < 1 || part < 0 || part >= nParts || nullptr == info) return TFQ_ERR(TFQMRGPU_UNDOCUMENTED_ERROR);
    if (!p.configured || p.multi || p.h_rpA.size() != size_t(p.mb) + 1) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    auto row_at = [&](int q) {      // first block row at or after q/nParts of the blocks
        int32_t const target = int32_t((int64_t(p.nnzbA)*q + nParts/2)/nParts);
        return int(std::lower_bound(p.h_rpA.begin(), p.h_rpA.end(), target) - p.h_rpA.begin());
    };
    row0 = (0 == part) ? 0 : std::min(p.mb, row_at(part));
    row1 = (nParts - 1 == part) ? p.mb : std::min(p.mb, row_at(part + 1));
    if (row1 < row0) row1 = row0;
    size_t const blockBytes = 2*size_t(p.LM)*p.LM*(('z' == p.precision) ? 8 : 4);
    int64_t const b0 = p.h_rpA[row0], b1 = p.h_rpA[row1];
    info[0] = int64_t(p.off_A + size_t(b0)*blockBytes); info[1] = (b1 - b0)*int64_t(blockBytes);
    info[2] = int64_t(p.off_ainv + size_t(row0)*4);     info[3] = p.use_tc16 ? int64_t(row1 - row0)*4 : 0;
    info[4] = b0; info[5] = b1 - b0;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getMatrixPartInfo(tfqmrgpuBsrsvPlan_t plan, int part, int nParts, int64_t info[6]) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    int r0, r1;
    return matrix_part(*P(plan), part, nParts, info, r0, r1);
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setMatrixPart(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, void const *valPart, char precision,
    char transposition, tfqmrgpuDataLayout_t layout, int part, int nParts, int64_t info[6]) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    bool trans = false; double scal_imag = 1;
    tfqmrgpuStatus_t st = parse_layout_trans(layout, transposition, trans, scal_imag);
    if (st) return st;
    trans = !trans;                      // A is stored transposed [k][i] (tfqmrgpu.cu:509-520)
    int row0, row1;
    st = matrix_part(p, part, nParts, info, row0, row1);
    if (st) return st;
    if (nullptr == p.pBuffer) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    bool const is_double = ('z' == p.precision);
    if (p.mixed) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    if (('z' == lower(precision)) != is_double) return TFQ_ERRC(TFQMRGPU_PRECISION_MISSMATCH, precision);
    if (info[5] < 1) return TFQMRGPU_STATUS_SUCCESS;
    if (nullptr == valPart) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    cudaStream_t const stream = static_cast<Handle*>(handle)->stream;
    uint32_t const b0 = uint32_t(info[4]), nb = uint32_t(info[5]);
    st = upload_a_blocks(p, stream, valPart, b0, nb, is_double, layout, trans, scal_imag);
    if (TFQMRGPU_STATUS_SUCCESS == st && p.use_tc16) st = launch_aop_convert_rows(p, row0, row1, stream);
    return st;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setRhsTrivial(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan) {
    if (nullptr == plan || nullptr == handle) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION);
    if (nullptr == p.pBuffer || !p.configured) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    return launch_unit_rhs(p, static_cast<Handle*>(handle)->stream);
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setEarlyFreeze(tfqmrgpuBsrsvPlan_t plan, int on) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    p.early_freeze = on ? 1 : 0;
    if (p.multi) multi_set_early_freeze(p);
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setInitialGuess(tfqmrgpuBsrsvPlan_t plan, int on) {
    if (nullptr == plan) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan &p = *P(plan);
    if (p.multi) return on ? TFQ_ERR(TFQMRGPU_NO_IMPLEMENTATION) : TFQMRGPU_STATUS_SUCCESS;
    p.initial_guess = on ? 1 : 0;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getMixedInfo(tfqmrgpuBsrsvPlan_t plan, double info[8]) {
    if (nullptr == plan || nullptr == info) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    Plan const &p = *P(plan);
    for (int i = 0; i < 8; ++i) info[i] = 0;
    Plan const *const q = mixed_inner(p);
    if (nullptr == q) return TFQMRGPU_STATUS_SUCCESS;
    info[0] = 1; info[1] = mixed_passes(p); info[2] = p.iterations_run; info[3] = double(q->bufferBytes);
    info[4] = q->use_tc16 ? (q->tc_planar ? 2 : 1) : 0; info[5] = p.use_dmma ? 1 : 0;
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_tileBlocksFor(int64_t nnzbX, int64_t blockBytes, int64_t *tileBlocks) {
    if (nullptr == tileBlocks || nnzbX < 0 || blockBytes < 1) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    *tileBlocks = int64_t(plan_tile_blocks(size_t(nnzbX), size_t(blockBytes), nsm));
    return TFQMRGPU_STATUS_SUCCESS;
}
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getTileBlocks(tfqmrgpuBsrsvPlan_t plan, int64_t *tileBlocks) {
    if (nullptr == plan || nullptr == tileBlocks) return TFQ_ERR(TFQMRGPU_POINTER_INVALID);
    *tileBlocks = int64_t(P(plan)->tile_blocks);
    return TFQMRGPU_STATUS_SUCCESS;
}

} // extern "C"
