// common.cuh -- shared host/device helpers of the splacu backend (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/splacu.h"

namespace splacu {

    // ---- error plumbing -------------------------------------------------------------------
    void        set_error(const char* fmt, ...);
    int         cuda_fail(cudaError_t e, const char* what, const char* file, int line);
    extern bool g_initialised;
    void        count_launch(int n = 1);
    void        count_jit_compile();

#define SPLACU_CUDA(expr)                                                          \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) return ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SPLACU_REQUIRE(cond, msg)                       \
    do {                                                \
        if (!(cond)) {                                  \
            ::splacu::set_error("%s: %s", __func__, msg); \
            return SPLACU_E_INVALID;                    \
        }                                               \
    } while (0)

#define SPLACU_CHECK_INIT()                                               \
    do {                                                                  \
        if (!::splacu::g_initialised) {                                   \
            ::splacu::set_error("%s: splacu_init() has not been called", __func__); \
            return SPLACU_E_NOT_INIT;                                     \
        }                                                                 \
    } while (0)

// check the launch itself (bad config etc.); execution errors surface at the next sync
#define SPLACU_LAUNCH_CHECK()                                   \
    do {                                                        \
        ::splacu::count_launch();                               \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return ::splacu::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

    // stream == NULL selects the backend's own in-order stream of the CURRENT device (one per initialised device)
    cudaStream_t resolve_stream(void* stream);
    int          sm_count();
    int          current_device();
    int          ensure_device(int device);// create the backend stream of a further device (multi-GPU group, dist.cu)

    // grid sized as a multiple of the SM count for grid-stride kernels
    inline int grid_for(size_t work_items, int block, int ctas_per_sm) {
        size_t want = (work_items + block - 1) / block;
        size_t cap  = (size_t) sm_count() * ctas_per_sm;
        if (want < 1) want = 1;
        return (int) (want < cap ? want : cap);
    }

    // ---- device csr handle ---------------------------------------------------------------
    // One column class of the matrix (mxv_pull.cu, "column-class phases"): the entries whose column belongs to the class, as a
    // CSR over all rows. Hub classes index a table of <= 65536 values of v that the kernel keeps in shared memory, so their
    // column ids are 16-bit slots; the tail class keeps the original 32-bit column ids.
    struct CsrPhase {
        uint32_t  nnz = 0, n_tiles = 0;
        uint32_t  slot_base = 0, n_slots = 0;// hub classes: slots [slot_base, slot_base + n_slots) of hub_cols / hub_vals
        bool      idx16     = false;
        uint32_t* Ap        = nullptr;       // [n_rows + 1]
        void*     Aj        = nullptr;       // uint16 slots (idx16) or uint32 column ids, padded to a whole tile
        uint32_t* Ax        = nullptr;
        uint2*    tile_rows = nullptr;
        uint32_t* carry     = nullptr;
        // segmented-tile format (mxv_seg.cu): Aj / Ax hold the entries lane-blocked per 512-entry tile (a coalesced 128-bit
        // load hands every lane 16 consecutive entries of the row order), Ap / tile_rows / carry are not kept
        bool      seg       = false;
        uint32_t  n_segs    = 0;             // non-empty rows of the class = row segments
        uint32_t* flags     = nullptr;       // [n_tiles * 16] bit (lane * 16 + i): entry (lane, i) is the last entry of its row
        uint32_t* seg_base  = nullptr;       // [n_tiles + 1] segments that end before tile t
        uint32_t* seg_row   = nullptr;       // [n_segs] row of every segment, ascending
        uint32_t* chain     = nullptr;       // [n_tiles] bit 31: the tile starts inside a row of the previous tile; low bits: tiles
                                             //           before t that hold the head of the row ending at t's first flag (0: none)
        uint32_t* chain_row = nullptr;       // [n_tiles] the row that ends at t's first flag (valid where chain has a length)
        uint32_t* head      = nullptr;       // [n_tiles] sum of the tile's first segment when it continues a row (per call)
        uint32_t* tail      = nullptr;       // [n_tiles] sum after the tile's last flag (per call)
    };
    static constexpr int kMaxHubPhases = 16;

    // One ROW class of the tail (mxv_scat.cu): the tail-column entries of the n_slots rows with the most such entries, stored in
    // COLUMN order in the same lane-blocked 512-entry tiles. The roles of the hub classes are swapped: the segments are columns
    // (v[seg_col] is read once per segment, in ascending order -- a stream), the 16-bit index is the row's slot in a table of
    // partial results that the CTA keeps in shared memory, and no gather of v ever leaves the SM for these entries.
    struct CsrScat {
        uint32_t  nnz = 0, n_tiles = 0, n_slots = 0, n_segs = 0;
        uint32_t* rows     = nullptr;// [n_slots] the row of every slot
        void*     slot     = nullptr;// uint16 [n_tiles * 512] row slot of every entry, lane-blocked
        uint32_t* Ax       = nullptr;// [n_tiles * 512] lane-blocked
        uint32_t* flags    = nullptr;// [n_tiles * 16] entry is the last one of its column
        uint32_t* seg_base = nullptr;// [n_tiles + 1] column segments that end before tile t
        uint32_t* seg_col  = nullptr;// [n_segs + 1] column of every segment, ascending
        uint32_t* partial  = nullptr;// [grid * n_slots] per-CTA tables of a call, folded in CTA order by the merge kernel
        uint32_t  grid     = 0;
    };
    static constexpr int kMaxScat = 4;

    // storage position of entry e (row order inside a class) in the lane-blocked tile layout: 4-byte items / 2-byte items
    __host__ __device__ __forceinline__ uint32_t seg_pos32(uint32_t e) {
        return (e & ~511u) + (((((e & 15u) >> 2) * 32u) + ((e & 511u) >> 4)) << 2) + (e & 3u);
    }
    __host__ __device__ __forceinline__ uint32_t seg_pos16(uint32_t e) {
        return (e & ~511u) + (((((e & 15u) >> 3) * 32u) + ((e & 511u) >> 4)) << 3) + (e & 7u);
    }
    struct Select;
    struct Csr;
    // mxv_seg.cu: build the segment metadata of a class from its row extents (ph.Ap) and row counts; run all classes
    int seg_build(const Csr* M, CsrPhase& ph, const uint32_t* d_row_count, cudaStream_t s);
    int seg_build_fixlist(Csr* M, cudaStream_t s);
    int seg_structure(const uint32_t* d_ext, const uint32_t* d_count, uint32_t n_units, uint32_t n_tiles, uint32_t** flags, uint32_t** seg_base,
                      uint32_t** seg_unit, uint32_t* n_segs, cudaStream_t s);
    // mxv_scat.cu: the row classes of the tail (built from the CSR + the column slot map + the rows chosen by build_phases)
    int  scat_build(Csr* M, const uint32_t* d_col_slot, const uint32_t* d_rows_sorted, uint32_t n_hub_rows, uint32_t slots_per_class, cudaStream_t s);
    void scat_free(Csr* M);
    int  scat_mxv(const Csr* M, int dtype, int op_mult, int op_add, const Select& sel, const void* d_v, void* d_r, const uint32_t* gate,
                  uint32_t gate_min, cudaStream_t s);
    // parts: bit 0 = the hub classes (they read only the packed hub values), bit 1 = everything that reads v itself (row classes, tail
    // classes) + the fix-ups, bit 2 = the prologue (r = init when no mask pass has filled it); 7 = the whole product
    int seg_mxv(const Csr* M, int dtype, int op_mult, int op_add, const Select& sel, const void* d_v, const void* d_mask, void* d_r,
                uint32_t init_bits, const uint32_t* gate, uint32_t gate_min, cudaStream_t s, int parts = 7);

    struct Csr {
        uint32_t        n_rows = 0, n_cols = 0, nnz = 0;
        const uint32_t* Ap = nullptr;
        const uint32_t* Aj = nullptr;
        const uint32_t* Ax = nullptr;
        // load-balancing metadata for the streaming pull kernel (mxv_pull.cu), owned by the handle
        uint32_t  tile         = 0;      // entries per nnz tile (kMxvTile)
        uint32_t  n_tiles      = 0;
        uint2*    tile_rows    = nullptr;// [n_tiles] (first row with entries in the tile, one past the last row starting in it)
        uint32_t* carry        = nullptr;// [2*n_tiles] (head, tail) partials of rows crossing tile borders
        bool      vec_ok       = false;  // Aj / Ax 16-byte aligned: 128-bit streaming loads
        float     avg_row_nnz  = 0.f;
        // hub cache of the pull kernel: the n_hub most referenced columns live in shared memory
        uint32_t  n_hub        = 0;
        uint32_t  n_hub_smem   = 0;      // slots [0, n_hub_smem) are staged in shared memory, the rest is served by L1
        uint32_t* hub_cols     = nullptr;// [n_hub] column ids, most referenced first
        uint32_t* hub_vals     = nullptr;// [n_hub] v[hub_cols[s]], packed per call
        uint32_t* Aj_hub       = nullptr;// [nnz] Aj with hub columns replaced by (0x80000000 | slot)
        // column-class phases (hub classes first, the tail class last); n_phases == 0: single-pass kernel on Ap / Aj / Ax
        int       n_phases     = 0;
        CsrPhase  phase[kMaxHubPhases + 1];
        int       n_scat       = 0;      // row classes of the tail (mxv_scat.cu)
        CsrScat   scat[kMaxScat];
        uint32_t* sel_count    = nullptr;// device counter: rows the mask of the current call selects (chooses the masked path on the device)
        uint32_t* sel_bits     = nullptr;// [n_rows / 32] bit i = select(mask[i]) of the current call: what the class passes read
        // rows that span tiles in any class: (row << 8 | class) sorted, and the tile the row ends in (the two-launch fix-up, mxv_seg.cu)
        uint64_t* fix_key      = nullptr;
        uint32_t* fix_tile     = nullptr;
        uint32_t  n_fix        = 0;
        // side stream of a product (the gated CSR pass runs beside the class passes), created at first use
        mutable cudaStream_t side    = nullptr;
        mutable cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
        // every stored value has the same bit pattern (adjacency matrices): lets the push product run structure-only (vxm_push.cu)
        bool      ax_uniform   = false;
        uint32_t  ax_value     = 0;
    };

    static constexpr int kMxvTile = 512;// nnz per warp tile of the streaming pull kernel

    // ---- tuning options (splacu_set_option) -------------------------------------------------
    enum Option { OPT_MXV_HUB = 0, OPT_MXV_HUB_MIN_COUNT, OPT_MXV_HUB_TOTAL, OPT_MXV_HUB_SMEM, OPT_MXV_L2_PERSIST, OPT_VXM_SELBITS, OPT_MXV_PHASES, OPT_MXV_PHASE_SLOTS, OPT_MXV_PHASE_ONLY, OPT_MXV_SEG, OPT_MXV_SEG_MIN_DENSITY, OPT_MXV_TAIL_RANGE_LOG2, OPT_SMALL_FRONT, OPT_VXM_STRUCT, OPT_MXV_RED, OPT_MXV_ROW_CLASSES, OPT_MXV_ROW_MIN_COUNT, OPT_MXV_FIXUP_MERGE, OPT_MXV_ROW_MIN_NNZ, OPT_MXV_BANK_ORDER, OPT_MXV_RESERVE_SMS, OPT_MXV_PDL, OPT_COUNT };
    int64_t get_option(int opt);
    // CTAs of a persistent one-CTA-per-SM kernel of the pull product: all SMs but the ones option mxv_reserve_sms keeps free
    inline uint32_t persistent_grid_cap() {
        const int64_t sms = sm_count(), keep = get_option(OPT_MXV_RESERVE_SMS);
        return (uint32_t) (keep > 0 && keep < sms ? sms - keep : sms);
    }

    // ---- workspace ------------------------------------------------------------------------
    struct Workspace {
        // dense accumulator + touched bitmap (vxm), sized for the largest vector seen
        uint32_t* acc          = nullptr;
        uint32_t* bitmap       = nullptr;
        uint32_t* sel_bits     = nullptr;// select(mask[j]) bitmap for large frontiers (vxm)
        uint32_t  cap_sel      = 0;
        uint32_t  cap_n        = 0;      // capacity in elements of acc
        uint32_t  acc_identity = 0;      // bit pattern acc[] is currently filled with
        bool      acc_clean    = false;  // acc[] == identity everywhere and bitmap == 0
        // scan scratch
        uint32_t* block_sums   = nullptr;
        uint32_t  cap_blocks   = 0;
        uint32_t* d_scalars    = nullptr;// small device scalars (counters, totals): 64 words
        uint32_t* h_scalars    = nullptr;// pinned mirror
        // generic exact vxm path buffers
        uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr, *offsets = nullptr;
        size_t    cap_pairs = 0, cap_offsets = 0;
        uint32_t* chunk_first = nullptr;// coarse index of the expanded frontier (vxm_push.cu)
        size_t    cap_chunks  = 0;
        void*     sort_tmp = nullptr;
        size_t    cap_sort_tmp = 0;
        // small-front scratch (launch-latency paths): touched-column list of the push expand [kSmallList], compacted feedback
        // indices / values of the sparse eadd_fdb [2 * kSmallFront]
        uint32_t* small        = nullptr;
        bool      pend_small   = false;  // the pending emit reads the small-front scratch instead of the bitmap
        bool      pend_const   = false;  // structure-only push: every result value is pend_value, acc[] was never touched
        uint32_t  pend_value   = 0;
        bool      last_struct  = false;  // the last vxm call took the structure-only path (introspection for tests / the bench)
        // an enqueued push whose count has not been read yet (vxm_begin_typed -> vxm_finish)
        bool      fin_strct = false, fin_small = false;
        uint32_t  fin_n = 0, fin_identity = 0;
        // pending emit state between *_begin and *_emit
        int       pending      = 0;      // 0 none, 1 vxm (acc/bitmap), 2 eadd_fdb sparse, 3 dense_to_coo, 4 vxm enqueued (begin_async)
        uint32_t  pend_n       = 0;      // length of the bitmap domain
        uint32_t  pend_count   = 0;
        uint32_t  pend_identity = 0;
        const uint32_t* pend_vi = nullptr;
        const uint32_t* pend_src = nullptr;
    };

    static constexpr uint32_t kSmallFront = 8192;// frontier entries one CTA turns into offsets / filters in one launch
    static constexpr uint32_t kSmallList  = 4096;// touched columns a single CTA sorts and emits
    int ws_reserve_vector(Workspace* ws, uint32_t n, cudaStream_t s);
    int ws_reserve_blocks(Workspace* ws, uint32_t n_blocks);
    int ws_reserve_selbits(Workspace* ws, uint32_t n);
    int ws_reserve_pairs(Workspace* ws, size_t n_pairs, size_t n_offsets);
    int ws_reserve_chunks(Workspace* ws, size_t n_chunks);

    // ---- shared device-side building blocks (defined in vector_ops.cu) -----------------------
    // exclusive scan of n uint32 (in place allowed); total written to d_total (may be null)
    int scan_exclusive_u32(Workspace* ws, const uint32_t* d_in, uint32_t* d_out, uint32_t n, uint32_t* d_total, cudaStream_t s);

    enum EmitMode {
        EMIT_ACC_RESET = 0,// (j, src[j]); afterwards src[j] = identity, bitmap word = 0     (vxm)
        EMIT_DENSE     = 1,// (j, src[j]); nothing reset                                      (dense -> coo)
        EMIT_INDIRECT  = 2,// (vi[k], src[vi[k]]) for set bit k; bitmap word = 0              (eadd_fdb sparse)
        EMIT_CONST     = 3 // (j, identity) -- one value for every entry; bitmap word = 0     (structure-only vxm)
    };
    // popcount the bitmap over [0, n) -> per-block sums scanned in ws->block_sums, total in ws->d_scalars[0]
    int bitmap_count(Workspace* ws, const uint32_t* d_bitmap, uint32_t n, cudaStream_t s);
    // ordered emit; requires bitmap_count() on the same bitmap before
    int bitmap_emit(Workspace* ws, uint32_t* d_bitmap, uint32_t n, int mode, uint32_t* d_src, const uint32_t* d_vi,
                    uint32_t identity, uint32_t* d_ri, uint32_t* d_rx, cudaStream_t s);

}// namespace splacu
