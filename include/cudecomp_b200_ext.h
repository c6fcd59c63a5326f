/*
 * cudecomp_b200_ext.h -- entry points of libcudecomp.so that have no counterpart in the reference ABI.
 * Nothing here is needed by a drop-in caller; they exist for tests, benchmarks and tuning.
 */
#ifndef CUDECOMP_B200_EXT_H
#define CUDECOMP_B200_EXT_H

#include "cudecomp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One transfer box of a plan, in elements; extents and strides are indexed by global axis (x, y, z). */
typedef struct {
  int32_t peer_rank;  /* global rank that owns the destination */
  int32_t is_unpack;  /* 1: local workspace -> destination buffer copy of the staged path */
  int32_t step;       /* step of a chunked (pipelined) schedule, 0 otherwise */
  int32_t reserved;
  int64_t src_offset; /* first element in the source buffer */
  int64_t dst_offset; /* first element in the destination buffer (of peer_rank) */
  int64_t extent[3];
  int64_t src_stride[3];
  int64_t dst_stride[3];
} cudecompB200Box_t;

/* Handle creation without an MPI type in the signature, for applications that run on a real MPI (see
 * cudecomp_b200_mpi.h, which wraps these two behind the application's own cudecompInit(&handle, comm) call): this
 * process is `rank` of `nranks`, rank 0 listens on root_addr:root_port (root_addr as every rank can reach it). The
 * library builds its control-plane mesh among exactly these processes. Collective. */
cudecompResult_t cudecompB200InitBootstrap(cudecompHandle_t* handle, int32_t rank, int32_t nranks, const char* root_addr,
                                           int32_t root_port);
/* A TCP port that is free on this host right now (rank 0 picks one and broadcasts it with its own MPI). */
cudecompResult_t cudecompB200PickBootstrapPort(int32_t* port);

/* Number of copy kernels this process has launched so far. */
cudecompResult_t cudecompB200GetLaunchCount(uint64_t* count);

/* Path of the last transpose/halo call on the descriptor: 0 none, 1 local, 2 direct peer stores, 3 staged. */
cudecompResult_t cudecompB200GetLastPath(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t* path);

/* Outcome of CUDECOMP_ENABLE_CUMEM (reference docs/env_vars.rst; src/cudecomp.cc:596-660) on this handle: 0 not
 * requested, 1 on (cudecompMalloc hands out cuMemCreate / cuMemMap allocations with POSIX-fd [+ fabric] handles and
 * peers map them through those handles instead of CUDA IPC), 2 requested but the ranks cannot pass file descriptors to
 * each other (pidfd_getfd), 3 requested but the device / driver lacks VMM support with POSIX-fd handles. In states 2
 * and 3 the library has printed the reference's "Disabling this feature" warning and allocates with cudaMalloc. */
cudecompResult_t cudecompB200GetCumemState(cudecompHandle_t handle, int32_t* state);

/* Host-only, collective over the handle's communicator: can every rank duplicate a file descriptor of its neighbour
 * (pidfd_getfd)? This is how POSIX-fd handles of cuMem allocations reach the importing rank. ok = 1 / 0 on every rank. */
cudecompResult_t cudecompB200ProbeFdPassing(cudecompHandle_t handle, int32_t* ok);

/* grid_ctas: CTAs per launch (0 = all resident CTAs); force_staged != 0 routes every exchange through the workspace. */
cudecompResult_t cudecompB200SetTuning(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t grid_ctas,
                                       int32_t force_staged);

/* Row-copy kernel variant: 0 = LDG/STG.128 (default), 1 = TMA bulk copies (cp.async.bulk through shared memory) for
 * launches whose rows are all 16-byte aligned and at least 2 KiB long; other launches keep variant 0. 2 = 256-bit loads and
 * stores (sm_100 LDG/STG.256) where every address and stride is 32-byte aligned. Also settable with
 * CUDECOMP_B200_KERNEL=bulk|wide. EXPERIMENTAL in round 1 (measured equal in bench/microbench_copy.cu, the product
 * kernel is not yet confirmed on hardware). */
cudecompResult_t cudecompB200SetKernelVariant(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t variant);

/* Launch schedule of the copy kernels. tile_bytes: size of a row-copy tile, 0 (= 32 KiB) or a power of two in
 * [4096, 262144] -- small grids want small tiles (more grid-stride rounds, shorter tail). peer_order: 0 = consecutive
 * tiles go to different peers for the whole launch ("one-shot"), 1 = the GPU works through its peers one after the
 * other, me+1 first ("pairwise rounds"; the local share stays interleaved). balance_grid != 0: the CTA count is reduced
 * by up to 20 % to the value whose last grid-stride round is fullest. Also CUDECOMP_B200_TILE_BYTES,
 * CUDECOMP_B200_PEER_ORDER=pairwise, CUDECOMP_B200_BALANCE_GRID=1. Any value gives identical results. */
cudecompResult_t cudecompB200SetSchedule(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t tile_bytes,
                                         int32_t peer_order, int32_t balance_grid);

/* Who drives a direct (out-of-place, peer-mappable) transpose: 0 = the sender stores its blocks into the peers' outputs
 * (default), 1 = the receiver loads its blocks from the peers' inputs and writes its own output locally. Same kernel,
 * same handshake, identical results; needs every member's INPUT to be peer-mappable, otherwise the call silently uses
 * mode 0. Staged calls (in place, forced staging) then load into the receiver's OWN workspace and unpack locally, so
 * nothing is written into a peer's memory and the workspace need not be peer-mappable (also with chunked staging). Same value on every rank. Also CUDECOMP_B200_TRANSFER=pull. EXPERIMENTAL until measured. */
cudecompResult_t cudecompB200SetTransferMode(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t mode);

/* Chunked schedule of staged transposes (in-place calls, NVSHMEM-family backends, non-exportable outputs): the pencil
 * is pushed in `nchunks` chunks and unpacking overlaps the next chunk's push (csrc/plan.h PipelinedPlan). 0 or 1 = off
 * (default; also settable for all descriptors with CUDECOMP_B200_PIPELINE_CHUNKS). Same value on every rank.
 * EXPERIMENTAL in round 1: the schedule is property-tested on the host, its device execution is not yet validated. */
cudecompResult_t cudecompB200SetPipelineChunks(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t nchunks);

/* How staged transposes (in-place calls, NVSHMEM-family backends, outputs that peers cannot map) run. mode 0 (default):
 * ONE phased launch pushes chunk s into the peers' workspaces while it unpacks chunk s - lag from the local one, with
 * per-chunk flags in the signal pad (the chunk count is cudecompB200SetPipelineChunks, 0 = chosen from the pencil size);
 * mode 1: separate push and unpack launches (chunked: unpacks on a side stream). lag in [1, 8], 0 keeps the current
 * value (default 1). Same values on every rank. Also CUDECOMP_B200_STAGED=launches, CUDECOMP_B200_FUSED_LAG, and
 * CUDECOMP_B200_PHASE_HEAD (percent of a step's pushes that run before the previous chunk's unpacks join, default 25). */
cudecompResult_t cudecompB200SetStagedMode(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t mode, int32_t lag);

/* Reports (and clears) a device-side handshake timeout of an earlier operation. */
cudecompResult_t cudecompB200CheckErrors(cudecompHandle_t handle, cudecompGridDesc_t grid_desc);

/* The sender-side plan of a transpose (ax, dir) / halo update as this rank would execute it; needs no GPU.
 * Returns the number of boxes (writes at most max_boxes), or -1 on error. staged != 0 describes the
 * workspace-staged variant (push boxes followed by the unpack boxes). */
int32_t cudecompB200DescribeTransposeBoxes(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t ax,
                                           int32_t dir, const int32_t input_halo_extents[],
                                           const int32_t output_halo_extents[], const int32_t input_padding[],
                                           const int32_t output_padding[], int32_t staged, cudecompB200Box_t* boxes,
                                           int32_t max_boxes);
int32_t cudecompB200DescribeHaloBoxes(cudecompHandle_t handle, cudecompGridDesc_t grid_desc, int32_t ax, int32_t dim,
                                      const int32_t halo_extents[], const bool halo_periods[], const int32_t padding[],
                                      int32_t staged, cudecompB200Box_t* boxes, int32_t max_boxes);

/* Handle-free variants: the plan any `rank` of a `pdims[0] x pdims[1]` grid would execute for `config` (only gdims,
 * gdims_dist, pdims, rank_order, transpose_axis_contiguous and transpose_mem_order are read). Pure host arithmetic,
 * used to property-test the planner over arbitrary decompositions in one process. Returns the number of boxes, or
 * minus the cudecompResult_t code on error (e.g. -2 for decompositions with empty pencils). For
 * cudecompB200PlanTransposeBoxes, staged == 2 returns the receiver-driven plan (cudecompB200SetTransferMode):
 * peer_rank is then the rank whose INPUT the box is read from, and the destination is the calling rank's output. */
int32_t cudecompB200PlanTransposeBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax, int32_t dir,
                                       const int32_t input_halo_extents[], const int32_t output_halo_extents[],
                                       const int32_t input_padding[], const int32_t output_padding[], int32_t staged,
                                       cudecompB200Box_t* boxes, int32_t max_boxes);
int32_t cudecompB200PlanHaloBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax, int32_t dim,
                                  const int32_t halo_extents[], const bool halo_periods[], const int32_t padding[],
                                  int32_t staged, cudecompB200Box_t* boxes, int32_t max_boxes);

/* The chunked (pipelined) variant of the staged schedule for `nchunks` chunks (csrc/plan.h PipelinedPlan): boxes carry
 * the step they run in. Returns 0 boxes when chunking does not apply. `inplace`: bit 0 = in place, bit 1 = receiver-
 * driven (cudecompB200SetTransferMode; peer_rank of a push box is then the rank whose input it is loaded from),
 * bits 8-15 = element size in bytes: non-zero allows column chunks (chunks along the fastest axis when it takes no part
 * in the transpose) the way the engine plans them, bit 2 = ... whatever the row length. */
int32_t cudecompB200PlanPipelinedTransposeBoxes(const cudecompGridDescConfig_t* config, int32_t rank, int32_t ax,
                                                int32_t dir, const int32_t input_halo_extents[],
                                                const int32_t output_halo_extents[], const int32_t input_padding[],
                                                const int32_t output_padding[], int32_t inplace, int32_t nchunks,
                                                cudecompB200Box_t* boxes, int32_t max_boxes);

/* What the autotuner would sweep for `options` on `nranks` ranks after the reference's environment filters
 * (CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS / _HALO_BACKENDS comma lists with '^' exclusion, CUDECOMP_AUTOTUNE_P_ROW_RANGE /
 * _P_COL_RANGE "min,max") and the disable_* flags: reference src/autotune.cc:108-273. Any output pair may be NULL. */
cudecompResult_t cudecompB200GetAutotuneCandidates(const cudecompGridDescAutotuneOptions_t* options, int32_t nranks,
                                                   cudecompRankOrder_t rank_order, int32_t transpose_backends[8],
                                                   int32_t* n_transpose, int32_t halo_backends[5], int32_t* n_halo,
                                                   int32_t pdims[][2], int32_t max_pdims, int32_t* n_pdims);

/* Host-only self test of the shared-memory descriptor mailbox (collective over the handle's communicator, no GPU
 * needed): `iterations` exchanges on alternating channels with rank groups of varying shape and randomised delays;
 * every received message is checked; then the shared-memory acknowledgement board behind deferred frees (every rank
 * acknowledges counts for every other rank, owners check what they see). Returns CUDECOMP_RESULT_SUCCESS or
 * INTERNAL_ERROR. */
cudecompResult_t cudecompB200SelfTestMailbox(cudecompHandle_t handle, int32_t iterations, uint32_t seed);

#ifdef __cplusplus
}
#endif

#endif
