// pg_abi.cu — extern "C" entry points of libpegasus_b200.so (see include/pegasus_b200.h) and the
// per-frame launch sequence:
//
//   memset(workspace head)                      clear counters, histograms, look-back status
//   preprocess_kernel                           A.1-A.5, writes GeomRec / depth key / radii
//   compact_hist_kernel + scan_rows_kernel      (depth key, index) of the visible Gaussians in index order + the digit
//                                               histograms of those keys (4 x 8 bit)
//   onesweep_pass_kernel x4                     visible Gaussians by depth (stable)
//   count_kernel + pair_scan_kernel             tile-row runs of every visible Gaussian (row-parallel, kept for emit), pairs per
//                                               Gaussian -> output offsets, pair count, overflow
//   emit_kernel                                 (tile, index) pairs in depth order (pure writer) + digit histograms of the
//                                               tile sort
//   tile_scan_kernel                            digit bases of the tile sort
//   onesweep_pass_kernel x2                     stored pairs by tile id (stable)  => reference order; the last
//                                               pass also reduces the tile ranges (identifyTileRanges)
//   tile_order_kernel                           normalises the ranges, tiles by descending list length
//   composite3_kernel<MASKS>                    A.7 (+ fused K+3 passes); composite2_kernel for counting runs
//
// Around the frame: pose_kernel (pg_pose_apply), pack_kernel / pack_masks_kernel, png_size_kernel + png_write_kernel
// (pg_png_encode).
//
// Nothing here synchronises with the host: R stays on the device (grids are sized by the pair
// capacity and surplus CTAs exit), overflow is reported through pg_read_status.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "pg_common.cuh"

namespace pg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// (kernel, device) -> dynamic shared memory already granted with cudaFuncSetAttribute
int smem_granted(const void* kernel, int device, int bytes, bool record) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> granted;
    std::lock_guard<std::mutex> lock(mu);
    int& g = granted[std::make_pair(kernel, device)];
    if (record && bytes > g) g = bytes;
    return g;
}

struct CompArgs;
struct PoseDev;

int launch_preprocess(const pg_raster_settings* s, const pg_gaussians* g, const pg_object_table* objs,
                      int32_t* radii, GeomRec* recs, uint32_t* dkey, Counters* counters, cudaStream_t stream);
int launch_mark_visible(int P, const float* means, const float* view, uint8_t* present, cudaStream_t stream);
int launch_compact_hist(const uint32_t* keys_in, uint32_t n, uint32_t* keys_out, uint32_t* vals_out, uint32_t* hist,
                        uint32_t* status, uint32_t* ticket, cudaStream_t stream);
int launch_onesweep_pass(bool iota, bool write_keys, const uint32_t* keys_in, uint32_t* keys_out,
                         const uint32_t* vals_in, uint32_t* vals_out, const uint32_t* n_ptr,
                         uint32_t n_imm, uint32_t max_tiles, int begin_bit, int num_bits,
                         const uint32_t* bin_base, uint32_t* status, uint32_t* ticket, uint2* ranges_raw,
                         uint32_t n_env, uint32_t* tile_obj_count, cudaStream_t stream);
int launch_count(bool keep_all, const uint32_t* perm, ushort4* srect, const GeomRec* recs, uint32_t P, int W, int H,
                 uint32_t* runs_fix, uint32_t* runs_ovf, uint32_t* ovf_base, uint32_t R_cap, uint32_t* rnd_off,
                 uint32_t* grp_loc, uint32_t* cta_pairs, uint32_t* cta_base, Counters* counters, Sticky* sticky,
                 cudaStream_t stream);
int launch_emit(bool keep_all, const uint32_t* perm, const uint32_t* cta_base, const uint32_t* grp_loc, const uint32_t* rnd_off,
                const uint32_t* runs_fix, const uint32_t* runs_ovf, const uint32_t* ovf_base, const ushort4* srect, uint32_t P,
                uint32_t gx, int gy, uint32_t* tkeys, uint32_t* tvals, uint32_t R_cap, Counters* counters, int bits_lo,
                uint32_t* hist_tile, cudaStream_t stream);
int launch_tile_scan(const uint32_t* hist_tile, uint32_t* bins, Counters* counters, cudaStream_t stream);
int launch_tile_order(uint2* ranges, uint32_t tiles, uint32_t* order, cudaStream_t stream);
int launch_export_keys(const uint2* ranges, uint32_t tiles, const uint32_t* point_list, const GeomRec* recs,
                       uint64_t* keys, uint32_t* point_list_out, uint32_t* ranges_out, cudaStream_t stream);
int launch_pose(int K, const int32_t* first, const PoseDev* poses_dev, const pg_canonical* canon,
                int scene_offset, const pg_scene* scene, cudaStream_t stream);
int launch_pack(int W, int H, const float* color, const float* depth, uint8_t* rgb_u8, uint16_t* depth_u16,
                cudaStream_t stream);
int launch_pack_masks(int W, int H, int n_planes, const uint8_t* masks, uint8_t* bits, cudaStream_t stream);
int launch_png_encode(int n_images, const pg_png_image* images, int width, int height, cudaStream_t stream);
int launch_composite_from_abi(const uint2* ranges, const uint32_t* tile_order, const uint32_t* point_list, const GeomRec* recs, int W,
                              int H, const float* bg, const pg_raster_outputs* ro, const pg_frame_outputs* fo,
                              const pg_object_table* objs, uint32_t n_env, const uint32_t* tile_obj_count,
                              unsigned long long* stats, bool fast, cudaStream_t stream);

// ---- opt-in profiling (bench / tests): CUDA events at stage boundaries, launch counter ----------
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

constexpr int kStageEvents = PG_NUM_STAGES + 1;
// One profiling session = an event table [max_frames][kStageEvents], shared by reference: a forward in flight on
// another thread keeps the table it started with alive while pg_profile_enable installs the next one.
struct ProfSession {
    std::vector<cudaEvent_t> events;
    int max_frames = 0;
    std::atomic<int> frames{0};
    ~ProfSession() {
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};
static std::mutex g_prof_mu;
static std::shared_ptr<ProfSession> g_prof;                 // guarded by g_prof_mu
static thread_local std::shared_ptr<ProfSession> t_prof;    // session of the forward in flight on this thread
static thread_local int t_prof_frame = -1;                  // its slot in that session

static int check_opts(const pg_launch_opts* o) {
    if (!o) return PG_OK;
    if (o->composite_stream && (!o->fork_event || !o->join_event)) {
        set_error("pg_launch_opts.composite_stream needs a fork and a join event");
        return PG_ERR_INVALID;
    }
    if (o->numerics != PG_NUMERICS_EXACT && o->numerics != PG_NUMERICS_FAST) {
        set_error("pg_launch_opts.numerics must be PG_NUMERICS_EXACT or PG_NUMERICS_FAST");
        return PG_ERR_INVALID;
    }
    return PG_OK;
}
// The stream compositing runs on; with a composite_stream it first waits for everything enqueued on `stream`.
static int comp_fork(const pg_launch_opts* o, cudaStream_t stream, cudaStream_t* cs) {
    *cs = stream;
    if (!o || !o->composite_stream) return PG_OK;
    PG_CUDA_CHECK(cudaEventRecord((cudaEvent_t)o->fork_event, stream));
    PG_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)o->composite_stream, (cudaEvent_t)o->fork_event, 0));
    *cs = (cudaStream_t)o->composite_stream;
    return PG_OK;
}
static int comp_join(const pg_launch_opts* o, cudaStream_t stream) {
    if (!o || !o->composite_stream) return PG_OK;
    PG_CUDA_CHECK(cudaEventRecord((cudaEvent_t)o->join_event, (cudaStream_t)o->composite_stream));
    PG_CUDA_CHECK(cudaStreamWaitEvent(stream, (cudaEvent_t)o->join_event, 0));
    return PG_OK;
}

static void prof_begin_frame() {
    t_prof_frame = -1;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        t_prof = g_prof;
    }
    if (!t_prof) return;
    const int f = t_prof->frames.fetch_add(1);
    if (f < t_prof->max_frames) t_prof_frame = f;
}
static void prof_mark(int stage_boundary, cudaStream_t stream) {
    if (t_prof_frame >= 0) cudaEventRecord(t_prof->events[(size_t)t_prof_frame * kStageEvents + stage_boundary], stream);
}

template <typename T>
static inline T* at(void* base, size_t off) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off);
}

static int check_common(const pg_raster_settings* s, const pg_gaussians* g, uint64_t pair_capacity) {
    if (!s || !g) { set_error("null settings/gaussians"); return PG_ERR_INVALID; }
    if (g->P < 0 || s->image_width <= 0 || s->image_height <= 0) { set_error("bad sizes"); return PG_ERR_INVALID; }
    if ((g->shs == nullptr) == (g->colors_precomp == nullptr)) {
        set_error("Please provide excatly one of either SHs or precomputed colors!");
        return PG_ERR_INVALID;
    }
    bool has_sr = g->scales != nullptr && g->rotations != nullptr;
    bool any_sr = g->scales != nullptr || g->rotations != nullptr;
    if ((!has_sr && g->cov3D_precomp == nullptr) || (any_sr && g->cov3D_precomp != nullptr)) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return PG_ERR_INVALID;
    }
    if (g->shs && (s->sh_degree < 0 || s->sh_degree > 3 || g->sh_coeffs < (s->sh_degree + 1) * (s->sh_degree + 1))) {
        set_error("sh_degree %d needs %d coefficients, got %d", s->sh_degree, (s->sh_degree + 1) * (s->sh_degree + 1), g->sh_coeffs);
        return PG_ERR_INVALID;
    }
    if (pair_capacity == 0 || pair_capacity > (1ull << 30)) { set_error("pair_capacity must be in [1, 2^30]"); return PG_ERR_INVALID; }
    uint32_t gx = (s->image_width + PG_TILE - 1) / PG_TILE, gy = (s->image_height + PG_TILE - 1) / PG_TILE;
    if (gx > 2047 || gy > 65535 || (uint64_t)gx * gy > (1u << 16)) {
        set_error("image too large: at most 65536 tiles of 16x16 and 2047 tile columns are supported");
        return PG_ERR_INVALID;
    }
    return PG_OK;
}

// everything up to (and including) the tile sort
static int run_binning(const pg_raster_settings* s, const pg_gaussians* g, const pg_object_table* objs,
                       int32_t* radii, void* ws, const Layout& L, uint64_t R_cap, bool keep_all,
                       const pg_launch_opts* opts, cudaStream_t stream) {
    const int P = g->P;
    const int W = s->image_width, H = s->image_height;
    const uint32_t gx = (W + PG_TILE - 1) / PG_TILE;
    Counters* counters = at<Counters>(ws, L.counters);
    prof_begin_frame();
    prof_mark(0, stream);
    PG_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.zero_begin), 0, L.zero_end - L.zero_begin, stream));
    prof_mark(1, stream);
    int rc = launch_preprocess(s, g, objs, radii, at<GeomRec>(ws, L.recs), at<uint32_t>(ws, L.dkey_a), counters, stream);
    if (rc) return rc;
    if (opts && opts->scene_read_event)  // nothing after this point reads the caller's scene arrays
        PG_CUDA_CHECK(cudaEventRecord((cudaEvent_t)opts->scene_read_event, stream));
    prof_mark(2, stream);
    if (s->debug & 1) PG_CUDA_CHECK(cudaStreamSynchronize(stream));
    // depth sort of the VISIBLE Gaussians: compaction a -> b (+ digit histograms), then b -> a -> b -> a -> b
    uint32_t* hist = at<uint32_t>(ws, L.hist_depth);
    uint32_t* ka = at<uint32_t>(ws, L.dkey_b); uint32_t* kb = at<uint32_t>(ws, L.dkey_a);
    uint32_t* va = at<uint32_t>(ws, L.dval_b); uint32_t* vb = at<uint32_t>(ws, L.dval_a);
    rc = launch_compact_hist(at<uint32_t>(ws, L.dkey_a), (uint32_t)P, ka, va, hist, at<uint32_t>(ws, L.status_compact),
                             &counters->tile_counter[7], stream);
    if (rc) return rc;
    uint32_t* st = at<uint32_t>(ws, L.status_depth);
    for (int p = 0; p < 4; ++p) {
        rc = launch_onesweep_pass(false, true, ka, kb, va, vb, &counters->num_visible, 0, L.tilesP, 8 * p, 8,
                                  hist + p * RADIX, st + (size_t)p * L.tilesP * RADIX,
                                  &counters->tile_counter[p], nullptr, 0, nullptr, stream);
        if (rc) return rc;
        uint32_t* t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    prof_mark(3, stream);
    if (s->debug & 1) PG_CUDA_CHECK(cudaStreamSynchronize(stream));
    // after 4 passes the sorted keys / permutation are back in (dkey_b, dval_b) == (ka, va)
    const uint32_t n_env = (objs && objs->num_objects > 0) ? (uint32_t)objs->first[0] : (uint32_t)P;
    const int bits = tile_bits(L.tiles);
    const int bits_lo = (bits + 1) / 2, bits_hi = bits - bits_lo;
    rc = launch_count(keep_all, va, at<ushort4>(ws, L.srect), at<GeomRec>(ws, L.recs), (uint32_t)P, W, H,
                      at<uint32_t>(ws, L.runs_fix), at<uint32_t>(ws, L.runs_ovf), at<uint32_t>(ws, L.ovf_base), (uint32_t)R_cap,
                      at<uint32_t>(ws, L.rnd_off), at<uint32_t>(ws, L.grp_loc), at<uint32_t>(ws, L.cta_pairs),
                      at<uint32_t>(ws, L.cta_base), counters, at<Sticky>(ws, L.sticky), stream);
    if (rc) return rc;
    rc = launch_emit(keep_all, va, at<uint32_t>(ws, L.cta_base), at<uint32_t>(ws, L.grp_loc), at<uint32_t>(ws, L.rnd_off),
                     at<uint32_t>(ws, L.runs_fix), at<uint32_t>(ws, L.runs_ovf), at<uint32_t>(ws, L.ovf_base),
                     at<ushort4>(ws, L.srect), (uint32_t)P, gx, (int)((H + PG_TILE - 1) / PG_TILE), at<uint32_t>(ws, L.tkey_a),
                     at<uint32_t>(ws, L.tval_a), (uint32_t)R_cap, counters, bits_lo, at<uint32_t>(ws, L.hist_tile), stream);
    if (rc) return rc;
    prof_mark(4, stream);
    rc = launch_tile_scan(at<uint32_t>(ws, L.hist_tile), at<uint32_t>(ws, L.bins_tile), counters, stream);
    if (rc) return rc;
    prof_mark(5, stream);
    uint32_t* bins = at<uint32_t>(ws, L.bins_tile);
    uint32_t* stt = at<uint32_t>(ws, L.status_tile);
    const bool two = bits_hi > 0;
    // pass lo: a -> b ; pass hi: b -> a.  With a single pass the result lands in b.  The LAST pass does not
    // write the sorted tile ids: it reduces the tile ranges instead (raw form, normalised by tile_order_kernel).
    uint2* ranges_raw = at<uint2>(ws, L.ranges);
    rc = launch_onesweep_pass(false, two, at<uint32_t>(ws, L.tkey_a), at<uint32_t>(ws, L.tkey_b),
                              at<uint32_t>(ws, L.tval_a), at<uint32_t>(ws, L.tval_b), &counters->sort_n, 0,
                              L.tilesR, 0, bits_lo, bins, stt, &counters->tile_counter[5], two ? nullptr : ranges_raw, n_env,
                              two ? nullptr : at<uint32_t>(ws, L.tile_obj_count), stream);
    if (rc) return rc;
    if (two) {
        rc = launch_onesweep_pass(false, false, at<uint32_t>(ws, L.tkey_b), at<uint32_t>(ws, L.tkey_a),
                                  at<uint32_t>(ws, L.tval_b), at<uint32_t>(ws, L.tval_a), &counters->sort_n, 0,
                                  L.tilesR, bits_lo, bits_hi, bins + RADIX, stt + (size_t)L.tilesR * RADIX,
                                  &counters->tile_counter[6], ranges_raw, n_env, at<uint32_t>(ws, L.tile_obj_count), stream);
        if (rc) return rc;
    }
    rc = launch_tile_order(at<uint2>(ws, L.ranges), L.tiles, at<uint32_t>(ws, L.tile_order), stream);
    if (rc) return rc;
    prof_mark(6, stream);
    if (s->debug & 1) PG_CUDA_CHECK(cudaStreamSynchronize(stream));
    return PG_OK;
}

static inline const uint32_t* sorted_point_list(const void* ws, const Layout& L) {
    const int bits = tile_bits(L.tiles);
    const bool two = bits - (bits + 1) / 2 > 0;
    return reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(ws) + (two ? L.tval_a : L.tval_b));
}

}  // namespace pg

using namespace pg;

extern "C" {

const char* pg_version(void) { return "pegasus_b200 0.1 (sm_100a)"; }
const char* pg_last_error(void) { return g_err; }

size_t pg_workspace_bytes(int32_t P, int32_t width, int32_t height, uint64_t pair_capacity) {
    if (P < 0 || width <= 0 || height <= 0) return 0;
    return make_layout(P, width, height, pair_capacity).total;
}

static int copy_status(const void* ws, pg_status* host, cudaStream_t stream) {
    const Layout L = make_layout(0, 16, 16, 1);  // the head of the workspace does not depend on the problem size
    const char* b = reinterpret_cast<const char*>(ws);
    static_assert(sizeof(pg_status) == 16 + sizeof(Sticky), "pg_status = 4 counter words + Sticky");
    PG_CUDA_CHECK(cudaMemcpyAsync(host, b + L.counters, 16, cudaMemcpyDeviceToHost, stream));
    PG_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char*>(host) + 16, b + L.sticky, sizeof(Sticky), cudaMemcpyDeviceToHost, stream));
    return PG_OK;
}

int pg_workspace_init(void* ws, size_t ws_bytes, pg_stream_t stream) {
    if (!ws || ws_bytes < make_layout(0, 16, 16, 1).counters + sizeof(Counters)) { set_error("workspace too small"); return PG_ERR_WORKSPACE; }
    PG_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, make_layout(0, 16, 16, 1).sticky), 0, sizeof(Sticky), (cudaStream_t)stream));
    return PG_OK;
}

int pg_rasterize_forward(const pg_raster_settings* s, const pg_gaussians* g, const pg_raster_outputs* out,
                         void* ws, size_t ws_bytes, uint64_t pair_capacity, const pg_launch_opts* opts,
                         pg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_common(s, g, pair_capacity);
    if (rc) return rc;
    if ((rc = check_opts(opts))) return rc;
    if (!out || !out->color || !out->radii || !out->depth || !ws) { set_error("null output/workspace"); return PG_ERR_INVALID; }
    Layout L = make_layout(g->P, s->image_width, s->image_height, pair_capacity);
    if (ws_bytes < L.total) { set_error("workspace too small: %zu < %zu", ws_bytes, L.total); return PG_ERR_WORKSPACE; }
    // the reference's complete pair lists are kept when asked for (debug bit 2) or needed (n_contrib
    // is a position in the reference's list)
    const bool keep_all = (s->debug & 4) != 0 || out->n_contrib != nullptr;
    rc = run_binning(s, g, nullptr, out->radii, ws, L, pair_capacity, keep_all, opts, stream);
    if (rc) return rc;
    cudaStream_t cs;
    rc = comp_fork(opts, stream, &cs);
    if (rc) return rc;
    rc = launch_composite_from_abi(at<uint2>(ws, L.ranges), at<uint32_t>(ws, L.tile_order), sorted_point_list(ws, L), at<GeomRec>(ws, L.recs),
                                   s->image_width, s->image_height, s->bg, out, nullptr, nullptr,
                                   (uint32_t)g->P, at<uint32_t>(ws, L.tile_obj_count),
                                   (s->debug & 2) ? at<Counters>(ws, L.counters)->stats : nullptr,
                                   opts && opts->numerics == PG_NUMERICS_FAST, cs);
    const int rj = comp_join(opts, stream);
    prof_mark(7, stream);
    if (!rc && !rj && opts && opts->status_host) return copy_status(ws, opts->status_host, stream);
    return rc ? rc : rj;
}

int pg_render_composed(const pg_raster_settings* s, const pg_gaussians* g, const pg_object_table* objs,
                       const pg_frame_outputs* out, void* ws, size_t ws_bytes, uint64_t pair_capacity,
                       const pg_launch_opts* opts, pg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_common(s, g, pair_capacity);
    if (rc) return rc;
    if ((rc = check_opts(opts))) return rc;
    if (!objs || !out || !out->color || !out->radii || !out->depth || !ws) { set_error("null argument"); return PG_ERR_INVALID; }
    if (objs->num_objects < 0 || objs->num_objects > PG_MAX_OBJECTS || objs->num_colors < 0 || objs->num_colors > PG_MAX_COLORS) {
        set_error("too many objects/colours"); return PG_ERR_INVALID;
    }
    for (int k = 0; k < objs->num_objects; ++k) {
        if (objs->first[k] > objs->first[k + 1] || objs->first[k] < 0 || objs->first[k + 1] > g->P ||
            objs->color_index[k] < 0 || objs->color_index[k] >= objs->num_colors) {
            set_error("bad object table entry %d", k); return PG_ERR_INVALID;
        }
    }
    Layout L = make_layout(g->P, s->image_width, s->image_height, pair_capacity);
    if (ws_bytes < L.total) { set_error("workspace too small: %zu < %zu", ws_bytes, L.total); return PG_ERR_WORKSPACE; }
    rc = run_binning(s, g, objs, out->radii, ws, L, pair_capacity, (s->debug & 4) != 0, opts, stream);
    if (rc) return rc;
    const uint32_t n_env = objs->num_objects > 0 ? (uint32_t)objs->first[0] : (uint32_t)g->P;
    cudaStream_t cs;
    rc = comp_fork(opts, stream, &cs);
    if (rc) return rc;
    if (out->silhouette && objs->num_colors > 0)
        PG_CUDA_CHECK(cudaMemsetAsync(out->silhouette, 0, (size_t)objs->num_colors * s->image_width * s->image_height, cs));
    rc = launch_composite_from_abi(at<uint2>(ws, L.ranges), at<uint32_t>(ws, L.tile_order), sorted_point_list(ws, L), at<GeomRec>(ws, L.recs),
                                   s->image_width, s->image_height, s->bg, nullptr, out, objs, n_env,
                                   at<uint32_t>(ws, L.tile_obj_count),
                                   (s->debug & 2) ? at<Counters>(ws, L.counters)->stats : nullptr,
                                   opts && opts->numerics == PG_NUMERICS_FAST, cs);
    const int rj = comp_join(opts, stream);
    prof_mark(7, stream);
    if (!rc && !rj && opts && opts->status_host) return copy_status(ws, opts->status_host, stream);
    return rc ? rc : rj;
}

int pg_read_status(const void* ws, pg_status* host_status, pg_stream_t stream) {
    if (!ws || !host_status) { set_error("null argument"); return PG_ERR_INVALID; }
    return copy_status(ws, host_status, (cudaStream_t)stream);
}

int pg_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present, pg_stream_t stream) {
    if (P < 0 || !means3D || !viewmatrix || !present) { set_error("null argument"); return PG_ERR_INVALID; }
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int pg_pose_apply(int32_t K, const int32_t* first, const pg_pose* poses_dev, const pg_canonical* canon,
                  int32_t scene_offset, const pg_scene* scene, pg_stream_t stream) {
    if (K < 0 || K > PG_MAX_OBJECTS || !first || !poses_dev || !canon || !scene) { set_error("bad argument"); return PG_ERR_INVALID; }
    for (int k = 0; k < K; ++k)
        if (first[k] > first[k + 1] || first[k] < 0 || first[k + 1] > canon->n_total || scene_offset + first[k + 1] > scene->P) {
            set_error("object %d does not fit", k); return PG_ERR_INVALID;
        }
    return launch_pose(K, first, reinterpret_cast<const PoseDev*>(poses_dev), canon, scene_offset, scene, (cudaStream_t)stream);
}

int pg_export_binning(const void* ws, int32_t P, int32_t width, int32_t height, uint64_t pair_capacity,
                      uint64_t* keys, uint32_t* point_list, uint32_t* ranges, pg_stream_t stream) {
    if (!ws || !keys || !point_list || !ranges) { set_error("null argument"); return PG_ERR_INVALID; }
    Layout L = make_layout(P, width, height, pair_capacity);
    const char* b = reinterpret_cast<const char*>(ws);
    return launch_export_keys(reinterpret_cast<const uint2*>(b + L.ranges), L.tiles, sorted_point_list(ws, L),
                              reinterpret_cast<const GeomRec*>(b + L.recs), keys, point_list, ranges,
                              (cudaStream_t)stream);
}

int pg_read_stats(const void* ws, uint64_t* host_stats8, pg_stream_t stream) {
    if (!ws || !host_stats8) { set_error("null argument"); return PG_ERR_INVALID; }
    const char* src = reinterpret_cast<const char*>(ws) + make_layout(0, 16, 16, 1).counters + offsetof(Counters, stats);
    PG_CUDA_CHECK(cudaMemcpyAsync(host_stats8, src, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PG_OK;
}

uint64_t pg_launch_count(void) { return (uint64_t)g_launches.load(); }

int pg_profile_enable(int32_t max_frames) {
    std::shared_ptr<ProfSession> next;
    if (max_frames > 0) {
        next = std::make_shared<ProfSession>();
        next->events.resize((size_t)max_frames * kStageEvents);
        for (auto& e : next->events) {
            e = nullptr;
            PG_CUDA_CHECK(cudaEventCreate(&e));
        }
        next->max_frames = max_frames;
    }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof = next;  // the previous session's events are destroyed when its last forward lets go of it
    return PG_OK;
}

static std::shared_ptr<ProfSession> prof_current() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return g_prof;
}

int32_t pg_profile_frames(void) {
    const auto p = prof_current();
    if (!p) return 0;
    const int f = p->frames.load();
    return f < p->max_frames ? f : p->max_frames;
}

int pg_profile_read(int32_t frame, float* stage_ms) {
    const auto p = prof_current();
    if (!p || frame < 0 || frame >= pg_profile_frames() || !stage_ms) { set_error("bad profile frame"); return PG_ERR_INVALID; }
    cudaEvent_t* ev = &p->events[(size_t)frame * kStageEvents];
    PG_CUDA_CHECK(cudaEventSynchronize(ev[PG_NUM_STAGES]));
    for (int i = 0; i < PG_NUM_STAGES; ++i) PG_CUDA_CHECK(cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
    return PG_OK;
}

int pg_pack_frame(int32_t width, int32_t height, const float* color, const float* depth, uint8_t* rgb_u8,
                  uint16_t* depth_u16, pg_stream_t stream) {
    if (width <= 0 || height <= 0 || (rgb_u8 && !color) || (depth_u16 && !depth)) { set_error("bad argument"); return PG_ERR_INVALID; }
    return launch_pack(width, height, color, depth, rgb_u8, depth_u16, (cudaStream_t)stream);
}

int pg_pack_masks(int32_t width, int32_t height, int32_t n_planes, const uint8_t* masks, uint8_t* bits,
                  pg_stream_t stream) {
    if (width <= 0 || height <= 0 || n_planes < 0 || (n_planes > 0 && (!masks || !bits))) { set_error("bad argument"); return PG_ERR_INVALID; }
    return launch_pack_masks(width, height, n_planes, masks, bits, (cudaStream_t)stream);
}

size_t pg_png_scratch_bytes(int32_t height) {
    if (height <= 0) return 0;
    // per row: bit count (u32), Adler partial sums (2 x u64), per-thread bit counts (256 x u16); H + 1 row offsets (u32)
    return ((size_t)height * 4 + 15) / 16 * 16 + (size_t)height * 16 + (size_t)height * 512 + (((size_t)height + 1) * 4 + 15) / 16 * 16;
}

size_t pg_png_worst_case_bytes(int32_t kind, int32_t width, int32_t height) {
    if (kind < PG_PNG_RGB8 || kind > PG_PNG_MASK8 || width <= 0 || height <= 0) return 0;
    // 2 bytes zlib header, block header, <= 15 bits per stream byte (a match token of <= 21 bits stands for >= 3
    // bytes), end of block, padding, Adler-32; rounded up to the 16-byte granularity of out_capacity
    const size_t bpp = kind == PG_PNG_RGB8 ? 3 : (kind == PG_PNG_GRAY16 ? 2 : 1);
    const size_t row = bpp * (size_t)width + 1;
    const size_t b = 2 + (32 * (size_t)(PG_PNG_TABLE_WORDS - 514) + 15 * row * (size_t)height + 15 + 7) / 8 + 4 + 8;
    return (b + 15) / 16 * 16;
}

int pg_png_encode(int32_t n_images, const pg_png_image* images, int32_t width, int32_t height, pg_stream_t stream) {
    if (n_images < 0 || width <= 0 || height <= 0 || (n_images > 0 && !images)) { set_error("bad argument"); return PG_ERR_INVALID; }
    for (int i = 0; i < n_images; ++i) {
        const pg_png_image& im = images[i];
        if (im.kind < PG_PNG_RGB8 || im.kind > PG_PNG_MASK8 || !im.src || !im.out || !im.table || !im.scratch || !im.result) {
            set_error("pg_png_encode: image %d: bad kind or NULL pointer", i);
            return PG_ERR_INVALID;
        }
        const int bpp = im.kind == PG_PNG_RGB8 ? 3 : (im.kind == PG_PNG_GRAY16 ? 2 : 1);
        if (im.src_pitch < bpp * width || (im.kind == PG_PNG_GRAY16 && (((uintptr_t)im.src | (uintptr_t)im.src_pitch) & 1))) {
            set_error("pg_png_encode: image %d: src_pitch %d too small or 16-bit source misaligned", i, im.src_pitch);
            return PG_ERR_INVALID;
        }
        if (((uintptr_t)im.out & 15) || ((uintptr_t)im.scratch & 15) || (im.out_capacity & 15) || im.out_capacity < 64) {
            set_error("pg_png_encode: image %d: out / scratch must be 16-byte aligned, out_capacity a multiple of 16 >= 64", i);
            return PG_ERR_INVALID;
        }
    }
    return launch_png_encode(n_images, images, width, height, (cudaStream_t)stream);
}

}  // extern "C"
