/*
 * tfqmrgpu_b200_ext.h - non-breaking extensions of the tfqmrgpu C-ABI (prefix tfqmrgpux_).
 *
 * None of these symbols exist in the reference; callers that only use tfqmrgpu.h never see them.
 * They expose what the parity tests, the multi-GPU (RHS block-column sharded) driver and the
 * benchmark need:  the device-built plan lists (bit-exact comparison with the reference's
 * createPlan, tfqmrgpu.cu:183-314), injection of the random shadow vector v3 (the reference draws
 * it inside setBuffer, tfqmrgpu.cu:430-442), the bare block-sparse product Y = A*X (the
 * `bench_tfqmrgpu multi` role, bench_tfqmrgpu.cu:289-440) and counters.
 */
#ifndef TFQMRGPU_B200_EXT_H
#define TFQMRGPU_B200_EXT_H

#include "tfqmrgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* library version of this implementation (the reference ships SOVERSION 1, VERSION 0.0.1) */
tfqmrgpuStatus_t tfqmrgpux_getVersion(int *major, int *minor, int *patch);

/* chatter on stdout: 0 silent (default), 1 reference-like "# norms of B ..." lines, 2 debug.
 * Also settable with the environment variable TFQMRGPU_VERBOSE. */
tfqmrgpuStatus_t tfqmrgpux_setVerbosity(int level);

/* Plan lists, copied to HOST memory.  kind: 0 starts u32[nnzbX+1], 1 pairs u32[2*nPairs],
 * 2 subset u32[nnzbB], 3 colindx u16[nnzbX]  (the four lists of the reference plan,
 * tfqmrgpu_plan.hxx:20-50), 4 perm u32[nnzbX] (X block index in caller order -> internal,
 * column-sorted, storage index), 5 colstart u32[nCols+1].
 * *count receives the number of elements; out may be NULL to query the count only. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getPlanArray(tfqmrgpuBsrsvPlan_t plan, int kind, void *out, size_t *count);

/* info[0..15] = nnzbX, nnzbB, nnzbA, nCols, nPairs, LM, LN, precision char, number of vector tiles,
 * number of SpMM units, SpMM columns per unit, SpMM entries, mb, 1 if the tcgen05 (fp32) product is used, 1 if the DMMA (fp64) product is used,
 * 1 if the small-block (LM <= 8) register-staged product is used */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getPlanInfo(tfqmrgpuBsrsvPlan_t plan, int64_t info[16]);

/* Random shadow vector v3, float[nnzbX][2][LM][LN] in the caller's X block order (= the layout the
 * reference's cuRAND stream fills).  set: after setBuffer; onDevice != 0 if v3 is a device pointer. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setV3(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, float const *v3, int onDevice);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getV3(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, float *v3Host);

/* The reference's shadow-vector stream (cuRAND XORWOW, seed 1234, linalg.hxx:784-797): n uniform floats
 * written to DEVICE memory.  A column-sharded multi-GPU run slices the 1-GPU stream with this. */
tfqmrgpuStatus_t tfqmrgpux_randomShadow(tfqmrgpuHandle_t handle, float *devOut, size_t n);

/* Y := A*X on the plan's vectors (X = what setMatrix('X') uploaded or the last solution),
 * repeated nrep times on the handle's stream; asynchronous.  Fetch with getVector('Y'). */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_multiply(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int nrep);

/* like tfqmrgpu_bsrsv_getMatrix but for var 'X' or 'Y' (the product of tfqmrgpux_bsrsv_multiply /
 * the residual vector of the last probe).  Blocking. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getVector(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, char var,
    void *val, char precision, char trans, tfqmrgpuDataLayout_t layout);

/* byte window of an operand inside the caller's workspace.  var 'X' (solution, internal column-sorted
 * block order, real_t[nnzbX][2][LM][LN]), 'A', 'B', '3' (v3), 'Y'.  Lets a multi-GPU driver gather X
 * slices device-to-device (NCCL) without a host round trip. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getWindow(tfqmrgpuBsrsvPlan_t plan, char var, size_t *offset, size_t *length);

/* per right-hand-side status int8[nCols*LN] after a solve: 0 ok, -1/-2/-3 breakdown codes of
 * linalg.hxx:57,123,209, +1 exact zero residual (core.hxx:282-285).  Blocking. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getRhsStatus(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int8_t *statusHost);

/* stats[0..7] of the last solve: probes executed, kernels launched, iteration bodies enqueued,
 * host wall milliseconds inside solve, last max_bound^2, last target_bound^2, 0, 0 */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getSolveStats(tfqmrgpuBsrsvPlan_t plan, double stats[8]);

/* Device-side profile of solves (off by default).  When on, solve records CUDA events on the handle's
 * stream around the whole solve and around every A*v product of the iteration bodies.
 * profile[0..7] of the last solve: device ms of the solve, summed device ms of the block-sparse products
 * of the iterations that really ran, how many products that sum covers, iterations run, kernels launched,
 * probes executed, 0, 0 */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setProfiling(tfqmrgpuBsrsvPlan_t plan, int on);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getSolveProfile(tfqmrgpuBsrsvPlan_t plan, double profile[8]);

/* User-defined linear operator - the C-callable form of the reference's C++ `action_t` concept (README.md "User-defined
 * linear operators", tfqmrgpu_core.hxx:27-37: the solver only ever calls action.multiply(y, x, ...)).  When `op` is set, every
 * product Y = A*X of solve() calls it instead of the library's block-sparse product; setMatrix('A') is then not needed.
 *   y, x   device pointers to X-shaped vectors in the solver's storage: real_t [nnzbX][2 (Re|Im)][LM][LN], blocks in the
 *          column-sorted storage order (getPlanArray kind 4 maps the caller's X block index to the storage index)
 *   state, expect   device-resident solver control: when expect >= 0 the product is only wanted while *state == expect
 *          (speculatively enqueued iterations after convergence, and the residual probe, are skipped that way); kernels of
 *          the operator SHOULD return at once otherwise - ignoring it is correct but wastes up to one product per iteration
 *   stream the handle's stream: the call must only enqueue work on it and must not synchronise
 * Return 0 on success; any other value aborts the solve with that status.  With an operator set the iteration is launched
 * kernel by kernel (no CUDA graph).  `op` = NULL restores the built-in product.  getInfo's flop count keeps the formula of the
 * block-sparse product. */
typedef int32_t (*tfqmrgpuxOperator_t)(void *ctx, void *y, void const *x, int32_t const *state, int32_t expect, cudaStream_t stream);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setOperator(tfqmrgpuBsrsvPlan_t plan, tfqmrgpuxOperator_t op, void *ctx);

/* Right preconditioner - the slot the reference left commented out (tfqmrgpu_core.hxx:37,57: `action.has_preconditioner()`, `vP`).
 * `op` has the signature and the rules of a user-defined operator and computes z = P*x for X-shaped vectors in the solver's storage,
 * P being an approximate inverse of A that acts within every right-hand-side column.  With a preconditioner set, solve() runs tfQMR
 * on A*P (every product becomes A*(P*v) through one extra vector owned by the plan; the residual probe evaluates A*(P*v1) - b, so
 * threshold and residual keep their meaning) and returns X = P*v1.  Kernel-by-kernel launches (no CUDA graph, no resident solver);
 * not combined with setInitialGuess, setDevices or precision 'm'.  `op` = NULL removes it. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setPreconditioner(tfqmrgpuBsrsvPlan_t plan, tfqmrgpuxOperator_t op, void *ctx);

/* The right-hand sides of the reference's `rhs_trivial` mode (tfqmrgpu_core.hxx:27,140-147: "columns of the unit matrix"): every
 * B block becomes the unit block (Re b[j mod LM][j] = 1).  Instead of setMatrix('B'); the reference offers this only to C++ callers
 * of its solve() template. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setRhsTrivial(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan);
/* Per-right-hand-side early freeze (not in the reference, whose stopping rule wants ALL right-hand sides below the threshold at the
 * same residual probe, tfqmrgpu_core.hxx:274-298): with on != 0 a right-hand side whose true residual passes a probe keeps its X
 * from then on (status 2 in getRhsStatus; its updates stop like after a breakdown) while the others continue.  In fp32 with
 * hundreds of right-hand sides the recurrences of converged columns drift and the reference's rule may not be met at one probe;
 * the freeze keeps what was reached.  (It cannot help a right-hand side that never passes: the two fp32 cases of the block-size sweep that
 * stall at a residual of 2e-2 - in the oracle as well - stall with it too.)  Default off = the reference's behaviour.  Takes effect with
 * the next solve. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setEarlyFreeze(tfqmrgpuBsrsvPlan_t plan, int on);

/* ---- precision 'm' ("start with float and converge double", tfqmrgpu.h:72) ---------------------------------------------------
 * The reference accepts 'm' in bufferSize (tfqmrgpu.cu:386) but its solver dispatch has the case commented out (tfqmrgpu.cu:42-44:
 * solve returns PRECISION_MISSMATCH).  Here bufferSize(..., 'm', &bytes) configures an fp64-accurate solve whose iterations run in
 * complex fp32 (on the tensor cores for 16/32/64 blocks): iterative refinement, R = B - A*X and X += D in fp64, A*D = R by the
 * ordinary fp32 solver with every right-hand side scaled to unit size.  The caller passes DOUBLE data to setMatrix / getMatrix
 * (precision argument 'z' or 'm'); the single workspace holds the fp64 side, the fp32 plan and the fp32 copy of the operator
 * (made on the device by setMatrix('A')).  solve(threshold, maxIterations): threshold is met by the true fp64 residual of every
 * right-hand side, maxIterations bounds the SUM of the fp32 iterations; getInfo reports that sum, the residual reached and the flops
 * of both sides.  Status 9 (MAX_ITERATIONS) also when two passes in a row gain less than a factor of two.  Not combined with
 * setDevices, setOperator or setMatrixPart.  TFQMRGPU_MIXED=0 restores the reference's refusal; TFQMRGPU_MIXED_INNER_TOL (default
 * 1e-3), TFQMRGPU_MIXED_INNER_ITER (default max(40, maxIterations/4)) and TFQMRGPU_MIXED_FREEZE (default 1: the fp32 passes use the
 * per-right-hand-side freeze of setEarlyFreeze) tune the passes.  Measured on one GPU's share of BASELINE config 4 (32^3 block rows,
 * 32x32 blocks, 128 right-hand sides, sigma 1, tol 1e-9): 'z' 2125 ms (29 iterations), 'm' 759 ms (3 passes, 42 fp32 iterations),
 * solutions equal to 2e-12 (tests/tools/bench_mixed.py, profiles/r02_bench_mixed.json).
 *
 * setInitialGuess (every single-device plan): with on != 0 solve starts from the X uploaded with setMatrix('X') - or, on a second
 * solve, from the previous solution - instead of zero.  (The reference zeroes X at the start of every solve, core.hxx:125, and so
 * does this library by default.)  'c' / 'z' plans run tfQMR on r0 = b - A*x0: one extra product before the first iteration, v1
 * accumulates the corrections on top of x0, the residual probe and the convergence bound stay relative to |b|; small systems then
 * take the per-kernel path instead of the resident solver.  Mixed plans start their refinement from X.
 * getMixedInfo: info[0] = 1 for a mixed plan, [1] refinement passes of the last solve, [2] fp32 iterations of the last solve,
 * [3] bytes of the fp32 plan's window, [4] fp32 product: 0 SIMT, 1 tcgen05 direct form, 2 tcgen05 planar form, [5] 1 if the fp64
 * product runs on DMMA. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setInitialGuess(tfqmrgpuBsrsvPlan_t plan, int on);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getMixedInfo(tfqmrgpuBsrsvPlan_t plan, double info[8]);

/* ---- several GPUs: the independent right-hand-side block columns of X/B are sharded, A is replicated --------
 *
 * (1) One process, several devices - for C, Fortran and Julia callers.  Call setDevices after createPlan and before
 *     bufferSize (or set the environment variable TFQMRGPU_NUM_GPUS=N before createPlan: devices 0..N-1).  From then on the
 *     ordinary entry points (bufferSize, setBuffer, setMatrix, solve, getInfo, getMatrix) drive all devices: the caller's
 *     single workspace lives on the device that was current at createPlan and holds that device's shard plus the gathered X;
 *     the other devices' workspaces are allocated by the library.  A is uploaded in nDevices parts over every device's own
 *     PCIe link and exchanged device to device; the solve keeps the reference's GLOBAL iteration and probe rule
 *     (core.hxx:239-299) by exchanging three numbers per shard and iteration; X is gathered device to device when
 *     getMatrix asks for it.  devices == NULL: 0 .. nDevices-1.  nDevices == 1 restores the single-GPU plan. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setDevices(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, int nDevices, int const *devices);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getDevices(tfqmrgpuBsrsvPlan_t plan, int *nDevices, int *devices, int arrayLength);

/* (2) One process per GPU (MPI / torch.distributed callers): every rank creates an ordinary plan for ITS block columns and
 *     registers an exchange: `slots` is device memory of 2*2*nShards*4 doubles on this rank's GPU; whenever the solver needs
 *     the other shards' convergence monitors it calls hook(ctx, slots + offset, nShards*4, stream): the hook must all-gather,
 *     in place and on `stream`, the nShards blocks of 4 doubles starting at the pointer it is given (block `shard` is this
 *     rank's contribution; NCCL: ncclAllGather(ptr + 4*shard, ptr, 4, ncclDouble, comm, stream)).  nRhsGlobal = number of
 *     right-hand sides over all shards.  hook == NULL removes the exchange.
 *     setShardHints (before bufferSize): maxColsPerRowHint = block columns per block row of the UNSHARDED problem - with it a
 *     shard chooses its product kernel like the single-GPU plan (the vector tiles are cut per block column, so they agree
 *     anyway) and reproduces the single-GPU bits of its columns when every block row of X holds the same block columns (a
 *     dense X, the reference's use case); for ragged X patterns the products group a row's entries by the block columns that
 *     share a unit, and results agree to rounding (same iteration count unless a decision falls within that rounding).
 *     (The bit-for-bit statement is about the per-kernel path; a small system that ONE GPU solves with the resident solver,
 *     resident.cu, groups its sums by that solver's own tiles and agrees with its shards to rounding.)
 *     tileBlocksHint > 0 forces a tile size (experiments); 0 = the default per-column rule (tfqmrgpux_tileBlocksFor). */
typedef int32_t (*tfqmrgpuxExchange_t)(void *ctx, double *slots, int count, cudaStream_t stream);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setShardExchange(tfqmrgpuBsrsvPlan_t plan, int shard, int nShards, int64_t nRhsGlobal,
    double *slots, tfqmrgpuxExchange_t hook, void *ctx);
/* Replicated A without N uploads of the whole operator: the block rows of A are cut into nParts ranges (balanced by blocks);
 * rank `part` uploads and converts ITS range only - valPart points to the first block of that range in the caller's host array -
 * and the ranks then exchange the converted ranges device to device (NCCL broadcast / all-gather of the byte windows).
 * info[0..5] = byte offset and length (inside the workspace) of the converted blocks of the range, byte offset and length of its
 * row scales (0 length for plans without them), first block and number of blocks of the range.  getMatrixPartInfo only reports. */
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getMatrixPartInfo(tfqmrgpuBsrsvPlan_t plan, int part, int nParts, int64_t info[6]);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setMatrixPart(tfqmrgpuHandle_t handle, tfqmrgpuBsrsvPlan_t plan, void const *valPart, char precision,
    char transposition, tfqmrgpuDataLayout_t layout, int part, int nParts, int64_t info[6]);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_setShardHints(tfqmrgpuBsrsvPlan_t plan, int64_t tileBlocksHint, int32_t maxColsPerRowHint);
tfqmrgpuStatus_t tfqmrgpux_bsrsv_getTileBlocks(tfqmrgpuBsrsvPlan_t plan, int64_t *tileBlocks);
/* X blocks per vector tile that bufferSize chooses for a block column of nnzbX blocks of blockBytes on the current device */
tfqmrgpuStatus_t tfqmrgpux_tileBlocksFor(int64_t nnzbX, int64_t blockBytes, int64_t *tileBlocks);

#ifdef __cplusplus
}
#endif
#endif /* TFQMRGPU_B200_EXT_H */
