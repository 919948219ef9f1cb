/*
 * splacu.h -- C ABI of the B200-native (sm_100a) backend for spla's masked semiring
 * matrix-vector hot path: exec_mxv_masked (pull) and exec_vxm_masked (push / SpMSpV).
 *
 * This is the drop-in boundary. Everything above it (spla's Matrix/Vector/Scalar objects, the
 * algorithm registry, the storage manager) stays the reference's own host code; the C++ plug-in
 * that a spla maintainer adds in src/cuda (shipped here under spla_b200/src/cuda) includes ONLY
 * this header and never a CUDA header. Plain pointers and sizes, no C++/torch types.
 *
 * Conventions
 *   - every value is 4 bytes (spla T_INT / T_UINT / T_FLOAT, reference src/type.cpp:32-35);
 *     scalars cross the ABI as raw uint32_t bit patterns, arrays as void* device pointers
 *   - `d_` pointers are device memory, `h_` pointers host memory
 *   - `stream` is a cudaStream_t passed as void* (NULL = the backend's own stream); calls only
 *     ENQUEUE work unless documented to synchronise, like the reference's in-order OpenCL queue
 *     (reference src/opencl/cl_accelerator.hpp:81)
 *   - return 0 on success, a SPLACU_E_* code (< 0) or a cudaError_t (> 0) otherwise;
 *     splacu_last_error() gives the message. Nothing throws across this boundary.
 *   - there is NO CPU fallback behind any entry point.
 */
#ifndef SPLACU_H
#define SPLACU_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define SPLACU_API __attribute__((visibility("default")))

/* ---- enums mirroring the reference's built-ins ------------------------------------------- */

/* value types, reference src/type.cpp:32-35 */
typedef enum splacu_dtype { SPLACU_INT = 0, SPLACU_UINT = 1, SPLACU_FLOAT = 2 } splacu_dtype;

/* built-in binary ops in the order of reference src/op.cpp:194-241 (TOpBinary names) */
typedef enum splacu_binop {
    SPLACU_PLUS = 0, SPLACU_MINUS, SPLACU_MULT, SPLACU_DIV, SPLACU_MINUS_POW2, SPLACU_FIRST, SPLACU_SECOND,
    SPLACU_BONE, SPLACU_MIN, SPLACU_MAX, SPLACU_LOR, SPLACU_LAND, SPLACU_BOR, SPLACU_BAND, SPLACU_BXOR,
    SPLACU_BINOP_COUNT
} splacu_binop;

/* built-in select ops in the order of reference src/op.cpp:243-266 (TOpSelect names) */
typedef enum splacu_selop {
    SPLACU_EQZERO = 0, SPLACU_NQZERO, SPLACU_GTZERO, SPLACU_GEZERO, SPLACU_LTZERO, SPLACU_LEZERO,
    SPLACU_ALWAYS, SPLACU_NEVER, SPLACU_SELOP_COUNT
} splacu_selop;

enum {
    SPLACU_OK               = 0,
    SPLACU_E_INVALID        = -1, /* bad argument (null pointer, unknown op, op not defined for dtype) */
    SPLACU_E_NOT_INIT       = -2, /* splacu_init not called / no CUDA device */
    SPLACU_E_CAPACITY       = -3, /* caller-provided output buffer too small; required size reported */
    SPLACU_E_NOT_IMPLEMENTED = -4, /* user-defined (non built-in) op and no NVRTC / driver library to compile it with */
    SPLACU_E_COMPILE        = -5  /* user-defined op: its source text failed to compile (message in splacu_last_error) */
};

/* ---- runtime: replaces reference src/opencl/cl_accelerator.{hpp,cpp} (CLAccelerator::init,
 *      set_device, queue creation :84-201) ---------------------------------------------------- */

SPLACU_API int         splacu_init(int device);              /* cudaSetDevice + backend stream + scratch pool */
SPLACU_API int         splacu_finalize(void);
SPLACU_API int         splacu_device_count(int* count);
SPLACU_API int         splacu_device_name(char* buffer, int length);
SPLACU_API int         splacu_sm_count(int* count);
SPLACU_API void*       splacu_default_stream(void);          /* the backend's own in-order stream */
SPLACU_API int         splacu_sync(void* stream);            /* cudaStreamSynchronize */
SPLACU_API const char* splacu_last_error(void);
SPLACU_API int         splacu_launch_count(uint64_t* count); /* kernels launched by this library so far */
/* tuning knobs, read when a matrix handle is created (the reference's counterpart are the vendor heuristics of
 * src/opencl/cl_accelerator.cpp:84-175): "mxv_hub" 0 = off, 1 = auto (default), 2 = always build the hub cache of
 * the pull kernel; "mxv_hub_min_count" = references a column needs to earn a hub slot (default 16);
 * "mxv_hub_total" / "mxv_hub_smem" = hub slots in total / staged in shared memory; "mxv_l2_persist" = L2 persisting window
 * on v; "vxm_selbits" = expand large frontiers against a select(mask) bitmap instead of the mask itself; "vxm_struct" = structure-only
 * push for large frontiers whose products are provably one value under an idempotent add (default 1);
 * "mxv_phases" / "mxv_phase_slots" = column classes of the pull product and the slots of one class (4 x 45056);
 * "mxv_row_classes" / "mxv_row_min_count" = row classes of the tail (1) and the tail entries a row needs for a slot (64);
 * "mxv_red" (read per call, default 1) = the class passes of a PLUS semiring add their row sums onto r with reductions at the L2
 * (red.global.add) instead of load + add + store. Integer results are unchanged; FLOAT results keep the same two-operand rounding and
 * the same fixed order, but red.global.add.f32 flushes subnormal operands and sums to zero (PTX ISA), where the reference keeps
 * them: set 0 when |r| < 1.2e-38 matters.
 * "mxv_fixup_merge" (per call, default 2) = how the partial sums of rows that span tiles are folded in: 2 = two plain launches (all chain
 * sums, then one thread per row in class order), 1 = one cooperative launch with a grid barrier per class, 0 = one launch per class;
 * 2 and 1 add in the same order (bit-identical results). "mxv_bank_order" (at handle creation, default 1) = permute the entries of a row
 * run inside a lane of a hub class against shared-memory bank conflicts. "mxv_reserve_sms" (per call, default 0) = SMs the persistent
 * class kernels leave free for the kernels of a collective running beside them. "mxv_pdl" (per call, default 1) = the class passes of a
 * whole product are launched with programmatic stream serialization (the next pass's CTAs load their hub table under the tail of the
 * previous pass and wait, griddepcontrol.wait, before they touch r: same results).
 * A product on a handle with column classes forks the handle's side stream from the caller's stream and joins it before it returns
 * control of the stream (the mask-first CSR pass runs there): to the caller the call is still ordered on the one stream it passed.
 * Also from the environment: SPLACU_OPTIONS="name=value,..." */
SPLACU_API int         splacu_set_option(const char* name, int64_t value);
SPLACU_API int         splacu_get_option(const char* name, int64_t* value);

/* profiling hooks (the reference: TIME_PROFILE_SCOPE labels with host / device times, src/profiling/time_profiler.hpp:44-100, dumped by
 * Library::time_profile_dump). Every entry point of this header opens an NVTX range "splacu/<entry>"; with profiling enabled it
 * also brackets its work with a cudaEvent pair on the launching stream. dump: "label, calls, device_ms, host_ms" lines. */
SPLACU_API int         splacu_profile_enable(int on);
SPLACU_API int         splacu_profile_reset(void);
SPLACU_API int         splacu_profile_dump(char* buffer, int length);

/* device memory: replaces cl::Buffer creation / enqueueRead / enqueueWrite in
 * reference src/opencl/cl_format_dense_vec.hpp:43-87, cl_format_coo_vec.hpp:43-125, cl_format_csr.hpp:40-96 */
SPLACU_API int splacu_malloc(void** d_ptr, size_t bytes);
SPLACU_API int splacu_free(void* d_ptr);
SPLACU_API int splacu_malloc_host(void** h_ptr, size_t bytes); /* pinned staging memory */
SPLACU_API int splacu_free_host(void* h_ptr);
SPLACU_API int splacu_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);  /* async */
SPLACU_API int splacu_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);  /* async; sync before reading */
SPLACU_API int splacu_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
/* replaces kernels fill_zero / fill_value, reference src/opencl/kernels/fill.cl:30,41 (cl_fill.hpp) */
SPLACU_API int splacu_fill(void* d_dst, uint32_t value_bits, size_t n, void* stream);

/* Multi-GPU (single box, net-new relative to the single-device reference): publish the window
 * [offset, offset + count) of this rank's copy of a 4-byte-element vector into the copies held by its peers, as one kernel of
 * 128-bit peer stores. peer_bases[q] = device pointer to rank q's copy, mapped into this process (symmetric allocation: same
 * layout everywhere), peer_bases[self] = the local copy. offset / count in elements, multiples of 4; the caller orders steps
 * with a cross-device barrier. Replaces the per-step ncclAllGather of the row-sharded pull. */
#define SPLACU_MAX_PEERS 16
SPLACU_API int splacu_publish_window(void* const* peer_bases, int n_peers, int self, size_t offset, size_t count, void* stream);

/* ---- device CSR matrix: replaces CLCsr + cl_csr_init, reference src/opencl/cl_formats.hpp:93-103,
 *      cl_format_csr.hpp:40-63. The handle does not own Ap/Aj/Ax; it owns the load-balancing
 *      metadata built once per matrix (decorations may carry extra data, core/accelerator.hpp:50-52). */

typedef struct splacu_csr_t* splacu_csr;

SPLACU_API int splacu_csr_create(splacu_csr* M, uint32_t n_rows, uint32_t n_cols, uint32_t nnz,
                                 const uint32_t* d_Ap, const uint32_t* d_Aj, const void* d_Ax, void* stream);
SPLACU_API int splacu_csr_destroy(splacu_csr M);
/* introspection for tests / logs: nnz tiles of the pull kernel and hub-cache slots (either pointer may be NULL) */
SPLACU_API int splacu_csr_info(splacu_csr M, uint32_t* n_tiles, uint32_t* n_hub);
/* column-class phases of the pull kernel (hub classes first, the tail class last): *n_phases = number of classes (0 when
 * the matrix is processed in one pass), nnz_per_phase[p] = entries of class p for p < min(*n_phases, cap) */
SPLACU_API int splacu_csr_phases(splacu_csr M, int* n_phases, uint32_t* nnz_per_phase, int cap);
/* row classes of the pull kernel's tail (the tail-column entries of the rows with the most of them, kept in column order and
 * accumulated in shared memory while v streams): number of classes, entries and rows of each (pointers may be NULL) */
SPLACU_API int splacu_csr_row_classes(splacu_csr M, int* n_classes, uint32_t* nnz_per_class, uint32_t* rows_per_class, int cap);

/* ---- the hot path --------------------------------------------------------------------------- */

/* Pull: r[i] = select(mask[i]) ? fold_{(j,a) in row i, stored order}(add, init, mult(a, v[j])) : init
 * Replaces Algo_mxv_masked_cl<T>::execute (reference src/opencl/cl_mxv.hpp:65-246, kernels/mxv.cl:43-170);
 * semantics are those of Algo_mxv_masked_cpu<T>::execute (reference src/cpu/cpu_mxv.hpp:56-106):
 * every r[i] written, mult(a_ij, v_j) argument order, early_exit = stop at the first position where
 * the running sum != init. d_v has n_cols entries, d_mask and d_r n_rows entries.
 * Results: bit-exact for INT/UINT, for MIN/MAX/logical/bitwise ops and for early_exit with any op;
 * FLOAT PLUS/MULT reductions differ from the sequential fold only by summation order. */
SPLACU_API int splacu_mxv_masked(splacu_csr M, int dtype, int op_mult, int op_add, int op_select,
                                 const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits,
                                 int early_exit, void* stream);

/* The pull product in two parts, for callers that assemble v from several owners (row-sharded multi-GPU runs): the entries of
 * the hub column classes read only the n_hub most referenced elements of v, so their passes can start as soon as THOSE values have
 * arrived, while the rest of v is still in flight.
 *   SPLACU_PART_PROLOGUE  r = init / the mask pass (reads nothing of v: can run before anything has arrived)
 *   SPLACU_PART_HUB       the hub classes. d_hub_vals[i] = v[hub_cols[i]] (splacu_csr_hub_cols) if the caller has gathered them, else
 *                         NULL and d_v is read. May be combined with the prologue (PROLOGUE | HUB).
 *   SPLACU_PART_REST      everything that reads v itself (row classes, tail classes, the mask-first CSR pass for sparse masks) and
 *                         the fix-ups. For matrices without column classes this is the whole product, the other parts are no-ops.
 * All parts on the same stream in this order, same ops / mask / r / init; op_add associative + commutative, no early exit.
 * The result equals splacu_mxv_masked up to the order in which the classes add onto r (FLOAT PLUS / MULT: rounding). */
#define SPLACU_PART_HUB 1
#define SPLACU_PART_REST 2
#define SPLACU_PART_PROLOGUE 4
SPLACU_API int splacu_csr_hub_cols(splacu_csr M, uint32_t* n_hub, const uint32_t** d_cols);
SPLACU_API int splacu_mxv_masked_part(splacu_csr M, int dtype, int op_mult, int op_add, int op_select, const void* d_v, const void* d_hub_vals,
                                      const void* d_mask, void* d_r, uint32_t init_bits, int part, void* stream);

/* Per-vector-length scratch for vxm / compaction (dense accumulator, touched bitmap, scan buffers).
 * Replaces the temp linear allocator + counters of reference src/opencl/cl_alloc_linear.hpp, cl_counter.hpp:42-84. */
typedef struct splacu_workspace_t* splacu_workspace;
SPLACU_API int splacu_workspace_create(splacu_workspace* ws);
SPLACU_API int splacu_workspace_destroy(splacu_workspace ws);
/* A workspace serves ONE call at a time: between *_begin and its *_emit no other call may use it (the pending state is checked),
 * and calls that share a workspace or a matrix handle must be ordered on one stream -- the per-call scratch of a handle (selection
 * bitmap, packed hub values, tile partials) is not re-entrant. splacu_workspace_reset drops a pending emit after an error between
 * begin and emit (an exception in the caller) and marks the scratch dirty, so that the workspace stays usable. */
SPLACU_API int splacu_workspace_reset(splacu_workspace ws, void* stream);
/* introspection for tests / the bench: *struct_only = 1 when the last vxm call on this workspace took the structure-only path
 * (uniform frontier values x uniform matrix values under an idempotent add: Ax, vx and the accumulator are never read) */
SPLACU_API int splacu_workspace_info(splacu_workspace ws, int* struct_only);

/* Ingest: row-major triplets (what Matrix::build leaves in its CpuCoo decoration, reference src/core/tmatrix.hpp:220-253) -> CSR on the
 * device, replacing the reference's host chain CpuCoo -> CpuLil -> CpuCsr -> AccCsr (src/storage/storage_manager_matrix.hpp:133-159).
 * d_Ai / d_Aj / d_Ax: nnz triplets in any order; d_Ap: n_rows + 1 row extents (out). Rows already non-decreasing (a loader, a sorted
 * build): only Ap is computed and d_Aj_out / d_Ax_out MAY be the input arrays themselves (no data movement); otherwise a stable sort
 * by row fills d_Aj_out / d_Ax_out (which must then be distinct buffers). Inside a row the input order is kept, like the reference's
 * stable counting sort (src/cpu/cpu_format_coo.hpp:58-76). *was_sorted reports the path taken. Synchronises once (the order check). */
SPLACU_API int splacu_coo_to_csr(uint32_t n_rows, uint32_t nnz, const uint32_t* d_Ai, const uint32_t* d_Aj, const void* d_Ax,
                                 uint32_t* d_Ap, uint32_t* d_Aj_out, void* d_Ax_out, splacu_workspace ws, int* was_sorted, void* stream);


/* Push (SpMSpV): for every stored (i,x) of sparse v, for (j,a) in row i with select(mask[j]):
 *   acc[j] = first ? mult(x,a) : add(acc[j], mult(x,a));  result = touched (j, acc[j]) ascending in j.
 * Replaces Algo_vxm_masked_cl<T>::execute_sparse (reference src/opencl/cl_vxm.hpp:73-180,
 * kernels/vxm.cl:30-95, cl_sort_by_key.hpp, cl_reduce_by_key.hpp); semantics are those of
 * Algo_vxm_masked_cpu<T>::execute (reference src/cpu/cpu_vxm.hpp:58-128): init is ignored, the output
 * pattern is structural (touched columns stay even if the value equals the fill value).
 * d_vi/d_vx: nv frontier entries; d_mask: n_cols entries.
 *
 * Two-phase form (what the spla plug-in uses, because TDecoration::values is a host uint,
 * reference src/core/tdecoration.hpp:58):
 *   splacu_vxm_masked_begin   accumulates on the device and returns the result count in *h_nr
 *                             (ONE 4-byte device->host synchronisation; the reference's OpenCL path has 2-3)
 *   splacu_vxm_masked_emit    writes the ordered (ri, rx) pairs into exactly-sized buffers and resets the scratch
 */
SPLACU_API int splacu_vxm_masked_begin(splacu_csr M, int dtype, int op_mult, int op_add, int op_select,
                                       uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                                       splacu_workspace ws, uint32_t* h_nr, void* stream);
SPLACU_API int splacu_vxm_masked_emit(splacu_workspace ws, uint32_t* d_ri, void* d_rx, void* stream);
/* begin, split at its host synchronisation: _begin_async only ENQUEUES (offsets, expand, count, the 4-byte copy), _begin_finish waits
 * for the stream and returns the count. Lets a caller that drives several devices enqueue on all of them before it waits on any
 * (the multi-GPU push below does). Op pairs that take the exact ordered path still synchronise inside _begin_async (sort sizing). */
SPLACU_API int splacu_vxm_masked_begin_async(splacu_csr M, int dtype, int op_mult, int op_add, int op_select,
                                             uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                                             splacu_workspace ws, void* stream);
SPLACU_API int splacu_vxm_masked_begin_finish(splacu_workspace ws, uint32_t* h_nr, void* stream);

/* One-call form over caller-provided buffers of `capacity` entries (n_cols always suffices). */
SPLACU_API int splacu_vxm_masked(splacu_csr M, int dtype, int op_mult, int op_add, int op_select,
                                 uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                                 uint32_t* d_ri, void* d_rx, uint32_t capacity, uint32_t* h_nr,
                                 splacu_workspace ws, void* stream);

/* ---- vector format glue: replaces kernels sparse_to_dense / dense_to_sparse,
 *      reference src/opencl/kernels/vector_formats.cl:30,42 (cl_format_coo_vec.hpp:127-161,
 *      cl_format_dense_vec.hpp:89-143). dense_to_coo keeps ascending index order like the CPU
 *      converter (reference src/cpu/cpu_format_dense_vec.hpp:53-67), which the OpenCL kernel does not. */
SPLACU_API int splacu_coo_to_dense(uint32_t n, uint32_t fill_bits, uint32_t nv, const uint32_t* d_vi, const void* d_vx,
                                   void* d_dense, void* stream);
SPLACU_API int splacu_dense_to_coo_count(int dtype, uint32_t n, uint32_t fill_bits, const void* d_dense,
                                         splacu_workspace ws, uint32_t* h_nr, void* stream);
SPLACU_API int splacu_dense_to_coo_emit(int dtype, uint32_t n, uint32_t fill_bits, const void* d_dense,
                                        splacu_workspace ws, uint32_t* d_ri, void* d_rx, void* stream);

/* ---- neighbours of the hot path inside bfs / sssp / pr loops -------------------- */

/* r[i] = select(mask[i]) ? assign(r[i], value) : r[i]; reference src/cpu/cpu_v_assign.hpp:95-127,
 * replaces src/opencl/cl_v_assign.hpp + kernels/vector_assign.cl:30 */
SPLACU_API int splacu_v_assign_masked_dense(int dtype, int op_assign, int op_select, uint32_t n,
                                            void* d_r, const void* d_mask, uint32_t value_bits, void* stream);
/* sparse mask: for every stored (i,x) with select(x): r[i] = assign(r[i], value);
 * reference src/cpu/cpu_v_assign.hpp:66-93, kernels/vector_assign.cl:44 */
SPLACU_API int splacu_v_assign_masked_sparse(int dtype, int op_assign, int op_select,
                                             void* d_r, uint32_t nm, const uint32_t* d_mi, const void* d_mx,
                                             uint32_t value_bits, void* stream);
/* structure-only form of a dense vector, for the frontier exchange of a multi-GPU traversal ("all-gather an n-bit
 * bitmap", 2 MB instead of 64 MB at scale 24): bit i of d_bits = op_select(v[i]) (ceil(n / 32) words, trailing bits 0), and back:
 * out[i] = bit i ? one_bits : zero_bits. No reference counterpart (the reference is single-device). */
/* d_dst[k] = d_src[d_idx[k]] / d_dst[d_idx[k]] = d_src[k] for k < n, 4-byte elements (the hub-value exchange of the two-part product) */
SPLACU_API int splacu_v_gather(uint32_t n, const uint32_t* d_idx, const void* d_src, void* d_dst, void* stream);
SPLACU_API int splacu_v_scatter(uint32_t n, const uint32_t* d_idx, const void* d_src, void* d_dst, uint32_t n_dst, void* stream);
/* element k of segment q (d_seg_off[q] <= k < d_seg_off[q + 1], n_peers segments): d_peer_dst[q][d_dst_idx[k]] = d_src[d_src_idx[k]].
 * d_peer_dst is a DEVICE array of n_peers device pointers (peer-mapped buffers of the other GPUs of the box, or local ones): the owners'
 * values of every rank's hub columns go straight into the peers' hub tables with NVLink stores -- one kernel instead of gather + n copies +
 * scatter (spla_b200/dist.py:PipelinedPull). No reference counterpart (the reference is single-device). */
SPLACU_API int splacu_v_push_peers(uint32_t n, const uint32_t* d_src_idx, const uint32_t* d_dst_idx, const uint32_t* d_seg_off, uint32_t n_peers,
                                   void* const* d_peer_dst, const void* d_src, void* stream);
SPLACU_API int splacu_v_pack_bits(int dtype, int op_select, uint32_t n, const void* d_v, uint32_t* d_bits, void* stream);
SPLACU_API int splacu_v_unpack_bits(uint32_t n, const uint32_t* d_bits, uint32_t one_bits, uint32_t zero_bits, void* d_out, void* stream);
/* number of entries != fill; reference src/cpu/cpu_v_count_mf.hpp:91-107, kernels/count.cl:46. Synchronises. */
SPLACU_API int splacu_v_count_mf_dense(int dtype, uint32_t n, const void* d_v, uint32_t fill_bits,
                                       splacu_workspace ws, uint32_t* h_count, void* stream);
/* r[i] = op(r[i], v[i]); fdb[i] = changed ? r[i] : fdb_fill; reference src/cpu/cpu_v_eadd_fdb.hpp:104-137 */
SPLACU_API int splacu_v_eadd_fdb_dense(int dtype, int op, uint32_t n, void* d_r, const void* d_v,
                                       void* d_fdb, uint32_t fdb_fill_bits, void* stream);
/* sparse v -> sparse feedback in the order of v; reference src/cpu/cpu_v_eadd_fdb.hpp:70-102.
 * v indices must be unique (they are: vxm output). Two-phase like vxm. */
SPLACU_API int splacu_v_eadd_fdb_sparse_begin(int dtype, int op, void* d_r, uint32_t nv, const uint32_t* d_vi,
                                              const void* d_vx, splacu_workspace ws, uint32_t* h_nf, void* stream);
SPLACU_API int splacu_v_eadd_fdb_sparse_emit(splacu_workspace ws, uint32_t* d_fi, void* d_fx, void* stream);
/* r[i] = op(u[i], v[i]); reference src/cpu/cpu_v_eadd.hpp:128-152 */
SPLACU_API int splacu_v_eadd_dense(int dtype, int op, uint32_t n, void* d_r, const void* d_u, const void* d_v, void* stream);
/* s = fold(op, init, v); reference src/cpu/cpu_v_reduce.hpp:90-114. Synchronises. op must be associative
 * and commutative (PLUS, MULT, MIN, MAX, LOR, LAND, BOR, BAND, BXOR); FLOAT PLUS/MULT differ by summation order. */
SPLACU_API int splacu_v_reduce_dense(int dtype, int op, uint32_t n, const void* d_v, uint32_t init_bits,
                                     splacu_workspace ws, uint32_t* h_result_bits, void* stream);

/* ---- multi-GPU, single box (net-new: the reference is single-device and only reserves the knob,
 *      Accelerator::set_queues_count, reference src/core/accelerator.hpp:58-69) -----------------------------------------------
 * A group is N shards, one per listed device (a device may be listed more than once: several shards then share it -- how the
 * sharding logic is tested on one GPU). device_ids[0] must be the home device (splacu_init): the caller's vectors live there, so
 * every other entry point of this header keeps working on them unchanged. One host thread drives all shards; the products only
 * ENQUEUE (events order the caller's stream before and after the shards' streams), except for the result count of the push.
 *   pull  M row-sharded into contiguous blocks of ~nnz / N entries; a product = v to every shard (ncclBroadcast over NVLink when the
 *         shards sit on distinct devices and libnccl.so.2 is present, peer copies otherwise), mask / result windows by peer copies,
 *         N local splacu_mxv_masked concurrently.
 *   push  M column-sharded (windows nnz-balanced on the column histogram, slices built at the first push); every shard expands the
 *         whole frontier against its slice under its mask window; the output is the concatenation of the shards' results.
 * Results are those of the single-device entry points: bit-exact for integer / order-independent work; FLOAT sums may differ in
 * summation order (the column classes are built per shard). Built-in ops only (user-defined ops: use the single-device handle). */
typedef struct splacu_dist_t* splacu_dist;
typedef struct splacu_dcsr_t* splacu_dcsr;
SPLACU_API int splacu_dist_create(splacu_dist* group, int n_shards, const int* device_ids);
SPLACU_API int splacu_dist_destroy(splacu_dist group);
SPLACU_API int splacu_dist_info(splacu_dist group, int* n_shards, int* uses_nccl);
/* d_Ap / d_Aj / d_Ax: the CSR on the home device (not owned; shard 0 works on them in place) */
SPLACU_API int splacu_dcsr_create(splacu_dcsr* M, splacu_dist group, uint32_t n_rows, uint32_t n_cols, uint32_t nnz,
                                  const uint32_t* d_Ap, const uint32_t* d_Aj, const void* d_Ax, void* stream);
SPLACU_API int splacu_dcsr_destroy(splacu_dcsr M);
/* row_bounds / col_bounds: n_shards + 1 entries each (either may be NULL; col_bounds are all zero before the first push) */
SPLACU_API int splacu_dcsr_bounds(splacu_dcsr M, int* n_shards, uint32_t* row_bounds, uint32_t* col_bounds);
/* same arguments and semantics as splacu_mxv_masked; d_v, d_mask, d_r on the home device */
SPLACU_API int splacu_dist_mxv_masked(splacu_dcsr M, int dtype, int op_mult, int op_add, int op_select,
                                      const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits, int early_exit, void* stream);
/* same arguments and semantics as splacu_vxm_masked_begin / _emit (the scratch is the group's: one workspace per shard) */
SPLACU_API int splacu_dist_vxm_masked_begin(splacu_dcsr M, int dtype, int op_mult, int op_add, int op_select,
                                            uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask, uint32_t* h_nr, void* stream);
SPLACU_API int splacu_dist_vxm_masked_emit(splacu_dcsr M, uint32_t* d_ri, void* d_rx, void* stream);

/* ---- user-defined ops -----------------------------------------------------------------------------
 * The reference compiles the SOURCE TEXT of a user op, "(T a, T b) { ... }" (OpBinary::make_int/uint/float, OpSelect::make_*,
 * reference src/op.cpp:294-342, e.g. tests/test_vector.cpp:299-302), into its OpenCL kernels at first use
 * (src/opencl/cl_program_builder.cpp:65-120). Here an op is either a built-in (id >= 0: splacu_binop / splacu_selop, name and
 * source ignored) or user-defined (id < 0: `name` = the op's name, `source` = its text in exactly that form; `uint` is defined).
 * A call whose ops are all built-ins forwards to the ahead-of-time specialised entry point above. Otherwise the library compiles
 * generic kernels for (dtype, mult, add, select) with NVRTC at first use (cached by name + source) and keeps the reference CPU
 * backend's SEQUENTIAL semantics, since nothing is known about a user op: mxv folds a row strictly left to right (one thread per
 * row, early exit included), vxm sorts the (column, product) pairs stably and folds every column left to right. Bit-exact
 * against the CPU backend for any op (compiled with --fmad=false: mult and add round separately, as two std::function calls do).
 * Errors: SPLACU_E_COMPILE (the text does not compile; splacu_last_error() carries the NVRTC log), SPLACU_E_NOT_IMPLEMENTED
 * (no libnvrtc / libcuda on the machine). Never a CPU fallback.
 * v_reduce with a user op is the one neighbour task without device code (a fold of n elements by an op of unknown associativity
 * is sequential): the spla plug-in hands that task to spla's own CPU algorithm, exactly what Dispatcher::dispatch does for a key
 * an accelerator does not provide (reference src/core/dispatcher.cpp:57-60). */
typedef struct splacu_op {
    int         id;     /* built-in id, or < 0 for a user-defined op */
    const char* name;   /* user op: its name (part of the cache key) */
    const char* source; /* user op: "(T a, T b) { ... }" / "(T a) { ... }" */
} splacu_op;

SPLACU_API int splacu_mxv_masked_ops(splacu_csr M, int dtype, const splacu_op* op_mult, const splacu_op* op_add, const splacu_op* op_select,
                                     const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits, int early_exit, void* stream);
/* emit with splacu_vxm_masked_emit */
SPLACU_API int splacu_vxm_masked_begin_ops(splacu_csr M, int dtype, const splacu_op* op_mult, const splacu_op* op_add, const splacu_op* op_select,
                                           uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                                           splacu_workspace ws, uint32_t* h_nr, void* stream);
SPLACU_API int splacu_v_assign_masked_dense_ops(int dtype, const splacu_op* op_assign, const splacu_op* op_select, uint32_t n,
                                                void* d_r, const void* d_mask, uint32_t value_bits, void* stream);
SPLACU_API int splacu_v_assign_masked_sparse_ops(int dtype, const splacu_op* op_assign, const splacu_op* op_select,
                                                 void* d_r, uint32_t nm, const uint32_t* d_mi, const void* d_mx, uint32_t value_bits, void* stream);
SPLACU_API int splacu_v_eadd_dense_op(int dtype, const splacu_op* op, uint32_t n, void* d_r, const void* d_u, const void* d_v, void* stream);
SPLACU_API int splacu_v_eadd_fdb_dense_op(int dtype, const splacu_op* op, uint32_t n, void* d_r, const void* d_v,
                                          void* d_fdb, uint32_t fdb_fill_bits, void* stream);
/* emit with splacu_v_eadd_fdb_sparse_emit */
SPLACU_API int splacu_v_eadd_fdb_sparse_begin_op(int dtype, const splacu_op* op, void* d_r, uint32_t nv, const uint32_t* d_vi,
                                                 const void* d_vx, splacu_workspace ws, uint32_t* h_nf, void* stream);
/* compile the module for (dtype, mult, add, select) without running anything (needs no GPU: NVRTC cross-compiles for sm_100a);
 * any op pointer may be NULL (= unused slot). *image_bytes = size of the device image. For tests and for warming the cache. */
SPLACU_API int splacu_jit_compile(int dtype, const splacu_op* op_mult, const splacu_op* op_add, const splacu_op* op_select, size_t* image_bytes);
SPLACU_API int splacu_jit_compile_count(uint64_t* count); /* modules compiled so far (cache misses) */

#if defined(__cplusplus)
}
#endif
#endif /* SPLACU_H */
