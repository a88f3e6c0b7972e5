// Host side of the C ABI declared in include/simclr_b200.h: argument validation, workspace carving,
// TMA tensor-map encoding and kernel launches.  No host synchronisation, no allocation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

#include "../../include/simclr_b200.h"
#include "aux_kernels.cuh"
#include "head_kernels.cuh"
#if SIMCLR_TRACE
// Diagnostics (timelines, rate probes, primitive self-test) exist only in the tracing build of the library
// (lib/libsimclr_b200_trace.so, include/simclr_b200_debug.h); the product library carries none of it.
#include "../../include/simclr_b200_debug.h"
#include "probes.cuh"
#include "selftest.cuh"
#endif

using namespace simclr;

namespace {

#if SIMCLR_TRACE
// debug timeline target (simclr_debug_set_trace); applies to subsequent launches of this process
long long* g_trace_ptr = nullptr;
int g_trace_cta = 0;
unsigned long long* g_ktrace_ptr = nullptr;
#else
constexpr long long* g_trace_ptr = nullptr;
constexpr int g_trace_cta = 0;
constexpr unsigned long long* g_ktrace_ptr = nullptr;
#endif
// Which kernels a staged call launches: a per-call argument of the measurement entry points simclr_forward_stages /
// simclr_backward_stages (bench.py times ONE kernel of the step by itself); every other entry point passes kAllStages.
enum StageBit : unsigned { kStagePrepare = SIMCLR_STAGE_PREPARE, kStageFwdTile = SIMCLR_STAGE_FORWARD_TILE,
                           kStageFwdFin = SIMCLR_STAGE_FORWARD_FINALIZE, kStageBwdTile = SIMCLR_STAGE_BACKWARD_TILE,
                           kStageBwdFin = SIMCLR_STAGE_BACKWARD_FINALIZE, kStageBwdPrepare = SIMCLR_STAGE_BACKWARD_PREPARE };
constexpr unsigned kAllStages = ~0u;

// ------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// A training loop calls with the same buffers step after step (and a CUDA graph is captured once), so the encoded maps
// are kept in a small per-process cache keyed by everything that determines them; cuTensorMapEncodeTiled is a driver
// call of a microsecond or two and a step needs five maps.  A CUtensorMap holds no reference to the memory it
// describes: a stale entry (buffer freed and reallocated elsewhere) simply never matches again.
struct MapKey {
    const void* base;
    int64_t rows, d_pad;
    int kind;      // 0 operand (bf16), 1 accumulator (f32)
    bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && d_pad == o.d_pad && kind == o.kind; }
};
constexpr int kMapCacheSize = 32;
struct MapCache {
    std::mutex mu;
    MapKey key[kMapCacheSize] = {};
    CUtensorMap map[kMapCacheSize];
    bool used[kMapCacheSize] = {};
    int next = 0;
    bool find(const MapKey& k, CUtensorMap* out) {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < kMapCacheSize; ++i)
            if (used[i] && key[i] == k) {
                *out = map[i];
                return true;
            }
        return false;
    }
    void put(const MapKey& k, const CUtensorMap& m) {
        std::lock_guard<std::mutex> lock(mu);
        key[next] = k;
        map[next] = m;
        used[next] = true;
        next = (next + 1) % kMapCacheSize;
    }
};
MapCache& map_cache() {
    static MapCache c;
    return c;
}

// bf16 [rows][d_pad] row-major, box = 128 rows x 64 elements (128 B), 128-byte swizzle
int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t d_pad) {
    const MapKey key{base, rows, d_pad, 0};
    if (map_cache().find(key, map)) return SIMCLR_OK;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SIMCLR_ERR_DRIVER_ENTRY;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d_pad) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kAtomK), static_cast<cuuint32_t>(kBlockM)};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SIMCLR_ERR_TENSOR_MAP;
    map_cache().put(key, *map);
    return SIMCLR_OK;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait() in sm100_ptx.cuh): the kernel may be
// staged while its predecessor in the stream is still running.  Works in eager streams and under stream capture
// (the graph records a programmatic dependency edge).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

struct DeviceInfo {
    int sm_count = 148;
    int cc_major = 0;
    bool valid = false;
};

constexpr int kMaxDevices = 64;

const DeviceInfo& device_info() {
    static DeviceInfo info[kMaxDevices];
    static std::once_flag once[kMaxDevices];
    static const DeviceInfo fallback;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fallback;
    std::call_once(once[dev], [dev] {
        DeviceInfo d;
        d.valid = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess;
        info[dev] = d;
    });
    return info[dev];
}

// Watchdog of the cross-GPU flag waits (peer_flag_wait): SIMCLR_B200_PEER_TIMEOUT_S seconds, default 600, 0 = none.
// Read once per process.
unsigned long long peer_timeout_ns() {
    static const unsigned long long ns = [] {
        double seconds = 600.0;
        if (const char* env = std::getenv("SIMCLR_B200_PEER_TIMEOUT_S")) {
            char* end = nullptr;
            const double v = std::strtod(env, &end);
            if (end != env && v >= 0.0) seconds = v;
        }
        return static_cast<unsigned long long>(seconds * 1e9);
    }();
    return ns;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct Geometry {
    int64_t bl_pad, bg_pad, d_pad;
    int n_row_blocks, n_col_tiles, grid, max_segs;
    long long total_tiles;
    int loss, tiles_per_view, col_start, col_cnt;
};
void set_window(Geometry* g, int start, int cnt);

// Column window of one tile-kernel launch: per view `cnt` tiles starting at view-tile `start` (cyclic).
void set_window(Geometry* g, int start, int cnt) {
    g->col_start = start;
    g->col_cnt = cnt;
    g->n_col_tiles = (g->loss == SIMCLR_LOSS_NTXENT ? 2 : 1) * cnt;
    g->total_tiles = static_cast<long long>(g->n_row_blocks) * g->n_col_tiles;
    const int sms = device_info().sm_count;
    g->grid = static_cast<int>(g->total_tiles < sms ? g->total_tiles : sms);
    const long long per_cta = (g->total_tiles + g->grid - 1) / g->grid;
    g->max_segs = static_cast<int>((per_cta + g->n_col_tiles - 2) / g->n_col_tiles + 1);
}

int make_geometry(int loss, int64_t b_local, int64_t b_global, int64_t row_offset, int64_t d, Geometry* g) {
    if (loss != SIMCLR_LOSS_NTXENT && loss != SIMCLR_LOSS_MODIFIED) return SIMCLR_ERR_BAD_LOSS;
    if (b_local < 1 || b_global < b_local || d < 1 || row_offset < 0 || row_offset + b_local > b_global)
        return SIMCLR_ERR_BAD_SHAPE;
    if (b_global > (int64_t(1) << 29)) return SIMCLR_ERR_BAD_SHAPE;
    g->d_pad = simclr_pad_dim(d);
    if (g->d_pad == 0) return SIMCLR_ERR_UNSUPPORTED_DIM;
    g->bl_pad = round_up(b_local, kBlockM);
    g->bg_pad = round_up(b_global, kBlockM);
    g->n_row_blocks = static_cast<int>(2 * g->bl_pad / kBlockM);
    g->loss = loss;
    g->tiles_per_view = static_cast<int>(g->bg_pad / kBlockN);
    set_window(g, 0, g->tiles_per_view);
    return SIMCLR_OK;
}

// The forward workspace starts with a 256-byte header holding the ticket of the finalize kernel's last-block
// reduction.  It must be zero when simclr_forward starts and is left zero; simclr_prepare zeroes it as well.
inline size_t header_bytes(const Geometry&) { return 256; }

struct FwdWorkspace {
    unsigned int* ticket;
    unsigned int* cand_cnt;   // [2*bl_pad] candidate counters of the exact accuracy count (zero between calls)
    int* cand;                // [2*bl_pad][kCandMax]
    int* amb_list;            // [2*bl_pad] rows whose accuracy decision is left to the backward finalize kernel (fused step)
    float* part;
    float* part2;      // partials of the second launch of the overlapped row-sharded forward
    float* block_part;
    size_t bytes, part_bytes;
};
FwdWorkspace carve_forward(const Geometry& g, void* base) {
    FwdWorkspace w;
    size_t off = 0;
    w.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(base) + off);
    off += header_bytes(g);
    w.block_part = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    off += align256(static_cast<size_t>(g.n_row_blocks) * 4 * sizeof(float));
    // (everything up to here depends on the LOCAL rows only: the prepare kernel, which does not know the window of the
    // forward launch, zeroes the ticket header and the candidate counters)
    w.cand_cnt = reinterpret_cast<unsigned int*>(static_cast<char*>(base) + off);
    off += align256(static_cast<size_t>(2) * g.bl_pad * sizeof(unsigned int));
    w.cand = reinterpret_cast<int*>(static_cast<char*>(base) + off);
    off += align256(static_cast<size_t>(2) * g.bl_pad * kCandMax * sizeof(int));
    w.amb_list = reinterpret_cast<int*>(static_cast<char*>(base) + off);
    off += align256(static_cast<size_t>(2) * g.bl_pad * sizeof(int));
    // sized for the full column window; a narrower window never needs more (fewer tiles per CTA, at most as many CTAs),
    // except that max_segs can grow by the row blocks a CTA additionally spans: bound it by n_row_blocks
    const int segs = g.n_row_blocks < 8 ? g.n_row_blocks : (g.max_segs + 6 < g.n_row_blocks ? g.max_segs + 6 : g.n_row_blocks);
    const size_t part_bytes = align256(static_cast<size_t>(device_info().sm_count) * segs * kFwdFields * kBlockM * sizeof(float));
    w.part = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    off += part_bytes;
    w.part2 = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    off += part_bytes;
    w.part_bytes = part_bytes;
    w.bytes = off;
    return w;
}
struct BwdWorkspace {
    float* colvec;
    float* dacc;
    float* det_part;       // deterministic mode: [grid * max_segs * 128][d_pad] per-(CTA, segment) accumulator slots
    size_t dacc_floats;
    int64_t det_rows;
    size_t bytes;
};
BwdWorkspace carve_backward(const Geometry& g, void* base, bool deterministic = false) {
    BwdWorkspace w;
    size_t off = 0;
    w.colvec = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    off += align256(static_cast<size_t>(4) * g.bg_pad * sizeof(float));
    w.dacc = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    w.dacc_floats = static_cast<size_t>(2) * g.bl_pad * g.d_pad;
    off += align256(w.dacc_floats * sizeof(float));
    w.det_part = nullptr;
    w.det_rows = 0;
    if (deterministic) {
        w.det_part = reinterpret_cast<float*>(static_cast<char*>(base) + off);
        w.det_rows = static_cast<int64_t>(g.grid) * g.max_segs * kBlockM;
        off += align256(static_cast<size_t>(w.det_rows) * g.d_pad * sizeof(float));
    }
    w.bytes = off;
    return w;
}

// log2-domain constants shared by forward and backward
struct Scales {
    float k2, inv_tau, tau, m2, qscale, op_scale;
    int const_shift;
    int pow;       // modified loss: 1 / 2 when 1/tau is exactly that (plain-power path), else 0
};
Scales make_scales(int loss, float temperature, int normalize, int64_t b_global) {
    Scales s;
    s.inv_tau = 1.0f / temperature;
    s.tau = temperature;
    s.qscale = static_cast<float>(b_global);
    if (loss == SIMCLR_LOSS_NTXENT) {
        s.k2 = 1.4426950408889634f * s.inv_tau;
        // the operands carry sqrt(k2) each: the MMA yields log2-domain logits S' = k2 * S
        s.op_scale = std::sqrt(s.k2);
        // |S'| <= k2 (+ bf16 rounding) when rows are normalised: with k2 <= 40, exp2(S') <= 2^41 and
        // a_r = g 2^(-lse2_r) >= 2^-13 2^-(41+17) stay far inside the fp32 range without any shift
        s.m2 = kConstShiftRaw;
        s.const_shift = (normalize && 2.0f * s.k2 <= 80.0f) ? 1 : 0;
        s.pow = 0;
    } else {
        s.k2 = s.inv_tau;
        s.op_scale = 1.0f;
        // y = log2(q) * (1/tau - 1) with q in [1e-4, B]
        const float lo = std::log2(kClampMin) * (s.k2 - 1.0f), hi = std::log2(s.qscale) * (s.k2 - 1.0f);
        s.m2 = std::fmax(lo, hi);
        // a_r = g 2^(m2 - lse2_r) with lse2_r >= log2(1e-4)/tau
        const float worst = s.m2 - std::log2(kClampMin) * s.k2;
        s.const_shift = (worst <= 80.0f) ? 1 : 0;
        // the reference's default temperature (1.0, objective.py:68) and the training default (0.5, utils/configs.json):
        // (B P)^(1/tau) needs no exponential; sums of at most B terms of at most B^2 stay far inside fp32
        s.pow = std::fabs(s.k2 - 1.0f) < 1e-6f ? 1 : (std::fabs(s.k2 - 2.0f) < 1e-6f ? 2 : 0);
#ifdef SIMCLR_NO_POW_PATH
        s.pow = 0;
#endif
    }
    return s;
}

// fp32 [rows][d_pad] row-major (the gradient accumulation buffer), box = 128 rows x 32 floats (128 B), 128-byte
// swizzle: target of the backward kernel's TMA reduce-add
int make_dacc_map(CUtensorMap* map, const void* base, int64_t rows, int64_t d_pad) {
    const MapKey key{base, rows, d_pad, 1};
    if (map_cache().find(key, map)) return SIMCLR_OK;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SIMCLR_ERR_DRIVER_ENTRY;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d_pad) * 4};
    cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(kBlockM)};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SIMCLR_ERR_TENSOR_MAP;
    map_cache().put(key, *map);
    return SIMCLR_OK;
}

template <int D, int kLoss, bool kBackward, bool kConst, int kPrec, bool kDet, bool kWindows = false>
int launch_tile_k(const CUtensorMap& rows, const CUtensorMap& cols, const CUtensorMap& dacc, const TileParams& p, int grid,
                  cudaStream_t st) {
    auto kern = contrastive_tile_kernel<D, kLoss, kBackward, kConst, kPrec, kDet, kWindows>;
    static std::once_flag configured[kMaxDevices];
    static cudaError_t configure_rc[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= kMaxDevices - 1;
    std::call_once(configured[dev], [&] {
        configure_rc[dev] = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 SmemLayout<D, kPrec>::kDynamicBytes);
    });
    if (configure_rc[dev] != cudaSuccess) return static_cast<int>(configure_rc[dev]);
    return static_cast<int>(launch_pdl(kern, dim3(grid), dim3(kBackward ? kThreadsBackward : kThreadsForward),
                                       SmemLayout<D, kPrec>::kDynamicBytes, st, rows, cols, dacc, p));
}

template <int D, int kLoss, bool kBackward, bool kConst, int kPrec>
int launch_tile(const CUtensorMap& rows, const CUtensorMap& cols, const CUtensorMap& dacc, const TileParams& p, int grid,
                cudaStream_t st) {
    if constexpr (kBackward) {
        if (p.deterministic) return launch_tile_k<D, kLoss, true, kConst, kPrec, true>(rows, cols, dacc, p, grid, st);
    } else if constexpr (kPrec == 0) {
        // the two-window forward of the row-sharded fused step (bf16 operands only) is its own instantiation: the window
        // loops cost the one-window kernel a microsecond per launch when they are a run-time switch
        if (p.n_windows == 2) return launch_tile_k<D, kLoss, false, kConst, 0, false, true>(rows, cols, dacc, p, grid, st);
    }
    return launch_tile_k<D, kLoss, kBackward, kConst, kPrec, false>(rows, cols, dacc, p, grid, st);
}

// kConst = p.const_shift (bounded scores: one exponential per element, no running maximum).  The forward kernel of
// the modified loss has no constant-shift variant.
template <int D, bool kBackward, int kPrec>
int launch_tile_d(int loss, const CUtensorMap& rows, const CUtensorMap& cols, const CUtensorMap& dacc, const TileParams& p,
                  int grid, cudaStream_t st) {
    const bool c = p.const_shift != 0;
    if (loss == SIMCLR_LOSS_NTXENT)
        return c ? launch_tile<D, kNtXent, kBackward, true, kPrec>(rows, cols, dacc, p, grid, st)
                 : launch_tile<D, kNtXent, kBackward, false, kPrec>(rows, cols, dacc, p, grid, st);
    if constexpr (kBackward) {
        return c ? launch_tile<D, kModified, true, true, kPrec>(rows, cols, dacc, p, grid, st)
                 : launch_tile<D, kModified, true, false, kPrec>(rows, cols, dacc, p, grid, st);
    } else {
        // forward: the constant-shift instantiation of the modified loss is its plain-power path (TileParams::pow)
        return p.pow != 0 ? launch_tile<D, kModified, false, true, kPrec>(rows, cols, dacc, p, grid, st)
                          : launch_tile<D, kModified, false, false, kPrec>(rows, cols, dacc, p, grid, st);
    }
}

template <bool kBackward>
int dispatch_tile(int loss, int64_t d_pad, int precision, const CUtensorMap& rows, const CUtensorMap& cols,
                  const CUtensorMap& dacc, const TileParams& p_in, int grid, cudaStream_t st) {
    TileParams p = p_in;
    if (p.n_windows != 2 || kBackward || precision != SIMCLR_PRECISION_BF16) {        // one window: the launch's own column range
        p.n_windows = 1;
        p.win0_local = 0;
        p.win[0] = ColWindow{p.col_start, p.col_cnt, p.n_col_tiles, p.max_segs, p.total_tiles, p.part};
    }
    if (precision == SIMCLR_PRECISION_SPLIT) {
        switch (d_pad) {
            case 64: return launch_tile_d<64, kBackward, 1>(loss, rows, cols, dacc, p, grid, st);
            case 128: return launch_tile_d<128, kBackward, 1>(loss, rows, cols, dacc, p, grid, st);
        }
        return SIMCLR_ERR_UNSUPPORTED_DIM;      // split operands of d > 128 do not fit shared memory
    }
    switch (d_pad) {
        case 64: return launch_tile_d<64, kBackward, 0>(loss, rows, cols, dacc, p, grid, st);
        case 128: return launch_tile_d<128, kBackward, 0>(loss, rows, cols, dacc, p, grid, st);
        case 256: return launch_tile_d<256, kBackward, 0>(loss, rows, cols, dacc, p, grid, st);
    }
    return SIMCLR_ERR_UNSUPPORTED_DIM;
}

inline bool bad_precision(int precision) { return precision != SIMCLR_PRECISION_BF16 && precision != SIMCLR_PRECISION_SPLIT; }

AuxParams make_aux(const Geometry& g, const Scales& s, int64_t b_local, int64_t b_global, int64_t row_offset,
                   int64_t d, int normalize) {
    AuxParams a;
    a.b_loc = static_cast<int>(b_local);
    a.b_glob = static_cast<int>(b_global);
    a.row_off = static_cast<int>(row_offset);
    a.bl_pad = static_cast<int>(g.bl_pad);
    a.bg_pad = static_cast<int>(g.bg_pad);
    a.d = static_cast<int>(d);
    a.d_pad = static_cast<int>(g.d_pad);
    a.normalize = normalize;
    a.k2 = s.k2;
    a.inv_tau = s.inv_tau;
    a.m2 = s.m2;
    a.const_shift = s.const_shift;
    a.qscale = s.qscale;
    a.op_scale = s.op_scale;
    a.split = 0;
    a.bn_state = nullptr;
    return a;
}

TileParams make_tile_params(const Geometry& g, const Scales& s, int64_t b_local, int64_t b_global, int64_t row_offset) {
    TileParams p;
    std::memset(&p, 0, sizeof(p));
    p.b_loc = static_cast<int>(b_local);
    p.b_glob = static_cast<int>(b_global);
    p.row_off = static_cast<int>(row_offset);
    p.bl_pad = static_cast<int>(g.bl_pad);
    p.bg_pad = static_cast<int>(g.bg_pad);
    p.n_row_blocks = g.n_row_blocks;
    p.n_col_tiles = g.n_col_tiles;
    p.col_start = g.col_start;
    p.col_cnt = g.col_cnt;
    p.tiles_per_view = g.tiles_per_view;
    p.max_segs = g.max_segs;
    p.total_tiles = g.total_tiles;
    p.k2 = s.k2;
    p.m2 = s.m2;
    p.const_shift = s.const_shift;
    p.pow = s.pow;
    p.qscale = s.qscale;
    p.inv_tau = s.inv_tau;
    p.tau = s.tau;
    p.acc_scale = 1.0f / s.op_scale;
    p.trace = g_trace_ptr;
    p.trace_cta = g_trace_cta;
    p.ktrace = g_ktrace_ptr;
    p.tile_grid = g.grid;
    p.peer_timeout_ns = peer_timeout_ns();
    return p;
}

inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

int check_device() {
    const DeviceInfo& di = device_info();
    if (!di.valid) return static_cast<int>(cudaErrorNoDevice);
    if (di.cc_major != 10) return SIMCLR_ERR_NOT_SM100;
    return SIMCLR_OK;
}

// PeerTable from the C-ABI arguments (world == 0 / NULL: no peers)
int make_peer_table(int world, int rank, void* const* ptrs, PeerTable* t) {
    std::memset(t, 0, sizeof(*t));
    if (world == 0 || ptrs == nullptr) return SIMCLR_OK;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return SIMCLR_ERR_BAD_PEERS;
    for (int r = 0; r < world; ++r) {
        if (ptrs[r] == nullptr) return SIMCLR_ERR_NULL_POINTER;
        if (misaligned(ptrs[r])) return SIMCLR_ERR_MISALIGNED;
        t->ptr[r] = ptrs[r];
    }
    t->world = world;
    t->rank = rank;
    return SIMCLR_OK;
}

}  // namespace

extern "C" {

int simclr_abi_version(void) { return SIMCLR_ABI_VERSION; }

const char* simclr_error_string(int code) {
    switch (code) {
        case SIMCLR_OK: return "ok";
        case SIMCLR_ERR_NULL_POINTER: return "null pointer argument";
        case SIMCLR_ERR_BAD_SHAPE: return "bad shape (need 1 <= b_local <= b_global, d >= 1, shard inside the batch)";
        case SIMCLR_ERR_UNSUPPORTED_DIM: return "embedding dimension above 256 is not supported";
        case SIMCLR_ERR_BAD_DTYPE: return "unsupported element type (f32 or bf16)";
        case SIMCLR_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case SIMCLR_ERR_MISALIGNED: return "pointer not 16-byte aligned";
        case SIMCLR_ERR_BAD_TEMPERATURE: return "temperature must be finite and > 0";
        case SIMCLR_ERR_NOT_SM100: return "device is not compute capability 10.x (B200)";
        case SIMCLR_ERR_DRIVER_ENTRY: return "cuTensorMapEncodeTiled not available from the driver";
        case SIMCLR_ERR_TENSOR_MAP: return "cuTensorMapEncodeTiled failed";
        case SIMCLR_ERR_BAD_LOSS: return "unknown loss kind";
        case SIMCLR_ERR_BAD_PEERS: return "bad peer table (1 <= world <= 16, 0 <= rank < world, shard = rank * b_local)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

int64_t simclr_pad_rows(int64_t b) { return b < 1 ? 0 : round_up(b, kBlockM); }

int64_t simclr_pad_dim(int64_t d) {
    if (d < 1 || d > 256) return 0;
    return d <= 64 ? 64 : (d <= 128 ? 128 : 256);
}

size_t simclr_forward_workspace_bytes(int loss, int64_t b_local, int64_t b_global, int64_t d) {
    Geometry g;
    if (make_geometry(loss, b_local, b_global, 0, d, &g) != SIMCLR_OK) return 0;
    return carve_forward(g, nullptr).bytes;
}

size_t simclr_backward_workspace_bytes(int loss, int64_t b_local, int64_t b_global, int64_t d) {
    return simclr_backward_workspace_bytes_flags(loss, b_local, b_global, d, 0);
}

size_t simclr_backward_workspace_bytes_flags(int loss, int64_t b_local, int64_t b_global, int64_t d, int flags) {
    Geometry g;
    if (make_geometry(loss, b_local, b_global, 0, d, &g) != SIMCLR_OK) return 0;
    return carve_backward(g, nullptr, (flags & SIMCLR_FLAG_DETERMINISTIC) != 0).bytes;
}

size_t simclr_operand_bytes(int64_t b, int64_t d, int precision) {
    const int64_t bp = simclr_pad_rows(b), dp = simclr_pad_dim(d);
    if (bp == 0 || dp == 0 || bad_precision(precision)) return 0;
    if (precision == SIMCLR_PRECISION_SPLIT && dp > 128) return 0;
    return static_cast<size_t>(precision == SIMCLR_PRECISION_SPLIT ? 2 : 1) * 2 * bp * dp * sizeof(__nv_bfloat16);
}

}  // extern "C"

namespace {

// In-kernel barriers of the fused row-sharded step (TileParams::sync_flags)
struct FusedSync {
    int world, rank;
    void* const* flag_peers;
    unsigned int* epoch;
    const float* stats_all;
    // two-window forward (TileParams::n_windows): the tile kernel pushes the operand rows itself
    bool windows = false;
    void* const* operand_peers = nullptr;
    void* operand_multicast = nullptr;
};

// Can the forward of the fused row-sharded step run as one launch over two column windows (own columns while the operand
// rows travel, then the peers')?  Needs tile-aligned shards and enough tiles in either window for every CTA.
// SIMCLR_B200_PEER_WINDOWS=0 switches it off (A/B measurements).
// (The same two windows in the BACKWARD tile kernel -- own columns first, the wait for the peers' column vectors in front of
// the rest -- were built and measured on 8 GPUs at 2N = 65536: 0.393 / 0.395 ms per step against 0.3865 with the forward
// windows alone and 0.402 with none; the second barrier exposes less than the extra accumulator flush per CTA and window
// costs.  Not instantiated; the kernel's window loops are written for either direction.)
bool peer_windows_ok(int loss, int64_t b_local, int world) {
    static const bool enabled = [] {
        const char* e = std::getenv("SIMCLR_B200_PEER_WINDOWS");
        return !(e && e[0] == '0');
    }();
    if (!enabled || world < 2 || b_local % kBlockM != 0) return false;
    const long long tpr = b_local / kBlockM;
    const long long row_blocks = 2 * tpr;
    const long long views = loss == SIMCLR_LOSS_NTXENT ? 2 : 1;
    const long long sms = device_info().sm_count;
    return row_blocks * views * tpr >= sms && row_blocks * views * tpr * (world - 1) >= sms;
}

int prepare_impl(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d, int in_dtype,
                 int normalize, float temperature, int precision, void* operand, float* inv_norm, float* pos_dot,
                 void* forward_workspace, int world, int rank, void* const* operand_global_peers,
                 void* operand_global_multicast, void* stream, unsigned int* bump_epoch, float* zrows_local,
                 const float* bn_state = nullptr, void* const* signal_flag_peers = nullptr, int shard_world = 0) {
    // shard_world > 0: this rank's batch is shard `rank` of `shard_world`, but the kernel pushes nothing (world == 0: the
    // two-window forward does the push)
    if (bad_precision(precision)) return SIMCLR_ERR_BAD_DTYPE;
    if (zrows_local != nullptr && misaligned(zrows_local)) return SIMCLR_ERR_MISALIGNED;
    if (!x_batch1 || !x_batch2 || !operand || !inv_norm || !pos_dot) return SIMCLR_ERR_NULL_POINTER;
    if (in_dtype != SIMCLR_DTYPE_F32 && in_dtype != SIMCLR_DTYPE_BF16) return SIMCLR_ERR_BAD_DTYPE;
    if (!(temperature > 0.f) || !std::isfinite(temperature)) return SIMCLR_ERR_BAD_TEMPERATURE;
    PeerTable peers;
    int rc = make_peer_table(world, rank, operand_global_peers, &peers);
    if (rc) return rc;
    if (peers.world > 0 && precision != SIMCLR_PRECISION_BF16) return SIMCLR_ERR_BAD_PEERS;   // bf16 operands only
    if (peers.world > 0 && operand_global_multicast != nullptr) {
        if (misaligned(operand_global_multicast)) return SIMCLR_ERR_MISALIGNED;
        peers.mc = operand_global_multicast;
    }
    if (shard_world > 0 && (peers.world > 0 || rank < 0 || rank >= shard_world)) return SIMCLR_ERR_BAD_PEERS;
    const int geom_world = shard_world > 0 ? shard_world : peers.world;
    const int64_t b_global = geom_world > 0 ? b_local * geom_world : b_local;
    const int64_t row_offset = geom_world > 0 ? b_local * (shard_world > 0 ? rank : peers.rank) : 0;
    Geometry g;
    if ((rc = make_geometry(loss, b_local, b_global, row_offset, d, &g))) return rc;
    if (misaligned(operand) || misaligned(forward_workspace)) return SIMCLR_ERR_MISALIGNED;
    if ((rc = check_device())) return rc;
    Scales s = make_scales(loss, temperature, normalize, b_global);
    unsigned int* zero_ptr = static_cast<unsigned int*>(forward_workspace);
    const int zero_words = static_cast<int>(header_bytes(g) / 4);
    unsigned int* cand_cnt = forward_workspace ? carve_forward(g, forward_workspace).cand_cnt : nullptr;
    if (precision == SIMCLR_PRECISION_SPLIT && g.d_pad > 128) return SIMCLR_ERR_UNSUPPORTED_DIM;
    AuxParams a = make_aux(g, s, b_local, b_global, row_offset, d, normalize);
    a.split = precision == SIMCLR_PRECISION_SPLIT ? 1 : 0;
    a.bn_state = bn_state;
    PeerTable signal_flags;
    if ((rc = make_peer_table(signal_flag_peers ? world : 0, rank, signal_flag_peers, &signal_flags))) return rc;
    if (signal_flags.world > 0 && bump_epoch == nullptr) return SIMCLR_ERR_NULL_POINTER;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int warps = 8;
    const int blocks = static_cast<int>((g.bl_pad + warps - 1) / warps);
    auto* op = static_cast<__nv_bfloat16*>(operand);
    cudaError_t launch_rc = cudaSuccess;
#define SIMCLR_PREP2(T, LOSS, PER) \
    launch_rc = launch_pdl(prepare_kernel<T, LOSS, PER>, dim3(blocks), dim3(warps * 32), 0, st, static_cast<const T*>(x_batch1), static_cast<const T*>(x_batch2), a, op, inv_norm, pos_dot, zero_ptr, zero_words, g_ktrace_ptr, peers, bump_epoch, cand_cnt, zrows_local, signal_flags)
#define SIMCLR_PREP(T, LOSS)                          \
    switch (g.d_pad) {                                \
        case 64: SIMCLR_PREP2(T, LOSS, 2); break;     \
        case 128: SIMCLR_PREP2(T, LOSS, 4); break;    \
        default: SIMCLR_PREP2(T, LOSS, 8); break;     \
    }
    if (loss == SIMCLR_LOSS_NTXENT) {
        if (in_dtype == SIMCLR_DTYPE_F32) SIMCLR_PREP(float, kNtXent)
        else SIMCLR_PREP(__nv_bfloat16, kNtXent)
    } else {
        if (in_dtype == SIMCLR_DTYPE_F32) SIMCLR_PREP(float, kModified)
        else SIMCLR_PREP(__nv_bfloat16, kModified)
    }
#undef SIMCLR_PREP
#undef SIMCLR_PREP2
    return static_cast<int>(launch_rc);
}

}  // namespace

extern "C" {

int simclr_prepare_peer(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d, int in_dtype,
                        int normalize, float temperature, int precision, void* operand, float* inv_norm, float* pos_dot,
                        void* forward_workspace, int world, int rank, void* const* operand_global_peers,
                        void* operand_global_multicast, float* zrows_local, void* stream) {
    return prepare_impl(loss, x_batch1, x_batch2, b_local, d, in_dtype, normalize, temperature, precision, operand, inv_norm,
                        pos_dot, forward_workspace, world, rank, operand_global_peers, operand_global_multicast, stream,
                        nullptr, zrows_local);
}

int simclr_prepare(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d, int in_dtype,
                   int normalize, float temperature, void* operand, float* inv_norm, float* pos_dot,
                   void* forward_workspace, void* stream) {
    return simclr_prepare_peer(loss, x_batch1, x_batch2, b_local, d, in_dtype, normalize, temperature,
                               SIMCLR_PRECISION_BF16, operand, inv_norm, pos_dot, forward_workspace, 0, 0, nullptr, nullptr,
                               nullptr, stream);
}

int simclr_peer_barrier(int world, int rank, void* const* flag_peers, unsigned int* epoch_local, const float* stats_all,
                        float* stats_out, float* loss_out, void* stream) {
    if (!flag_peers || !epoch_local) return SIMCLR_ERR_NULL_POINTER;
    if (stats_all != nullptr && stats_out == nullptr) return SIMCLR_ERR_NULL_POINTER;
    if (world < 1) return SIMCLR_ERR_BAD_PEERS;
    PeerTable flags;
    int rc = make_peer_table(world, rank, flag_peers, &flags);
    if (rc) return rc;
    if ((rc = check_device())) return rc;
    return static_cast<int>(launch_pdl(peer_barrier_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), flags,
                                       epoch_local, stats_all, stats_out, loss_out, peer_timeout_ns()));
}

}  // extern "C"

namespace {

// Where the forward finalize kernel finds exact fp32 rows for the accuracy candidates (TileParams::cand_cnt): the caller's
// inputs on one GPU, a gathered fp32 copy, or the ranks' symmetric copies.  All NULL: tensor-core decision.
struct ExactSource {
    const void* x1 = nullptr;
    const void* x2 = nullptr;
    int in_dtype = SIMCLR_DTYPE_F32;
    const float* inv_norm = nullptr;
    const float* zrows = nullptr;            // [2*Bgpad][Dpad]
    void* const* zrows_peers = nullptr;      // world pointers to [2*Blpad][Dpad]
    const float* bn_state = nullptr;         // projection-head tail: x1 / x2 are pre-BatchNorm activations
};

// defer_stats: the backward of the same fused step finishes the loss statistics (simclr_forward_backward)
int forward_impl(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                 int64_t row_offset, int64_t d, float temperature, int normalize, int precision, const float* pos_dot,
                 const float* row_weight, float* lse2, float* row_loss, float* stats, float* loss_out, void* workspace,
                 size_t workspace_bytes, void* backward_workspace, size_t backward_workspace_bytes, int world, int rank,
                 void* const* colvec_peers, void* const* stats_peers, void* const* flag_peers, unsigned int* epoch_local,
                 void* stream, bool defer_stats, const FusedSync* fused = nullptr, unsigned stages = kAllStages,
                 const ExactSource& exact = ExactSource()) {
    if (!operand_rows || !operand_cols || !pos_dot || !lse2 || !row_loss || !stats || !workspace)
        return SIMCLR_ERR_NULL_POINTER;
    if (!(temperature > 0.f) || !std::isfinite(temperature)) return SIMCLR_ERR_BAD_TEMPERATURE;
    Geometry g;
    int rc = make_geometry(loss, b_local, b_global, row_offset, d, &g);
    if (rc) return rc;
    if (misaligned(operand_rows) || misaligned(operand_cols) || misaligned(workspace)) return SIMCLR_ERR_MISALIGNED;
    FwdWorkspace w = carve_forward(g, workspace);
    if (workspace_bytes < w.bytes) return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = check_device())) return rc;

    if (bad_precision(precision)) return SIMCLR_ERR_BAD_DTYPE;
    const int planes = precision == SIMCLR_PRECISION_SPLIT ? 2 : 1;       // hi (+ lo) operand planes
    CUtensorMap map_rows, map_cols;
    if ((rc = make_operand_map(&map_rows, operand_rows, planes * 2 * g.bl_pad, g.d_pad))) return rc;
    if ((rc = make_operand_map(&map_cols, operand_cols, planes * 2 * g.bg_pad, g.d_pad))) return rc;

    // normalised NT-Xent rows bound the scores: constant softmax shift (no running maximum) in the tile kernel
    Scales s = make_scales(loss, temperature, normalize, b_global);
    TileParams p = make_tile_params(g, s, b_local, b_global, row_offset);
    p.d = static_cast<int>(d);
    p.ticket = w.ticket;
    p.part = w.part;
    p.pos_dot = pos_dot;
    p.row_weight = row_weight;
    p.lse2 = lse2;
    p.row_loss = row_loss;
    p.block_part = w.block_part;
    p.stats = stats;
    p.loss_out = loss_out;
    p.defer_stats = defer_stats ? 1 : 0;
    p.normalize = normalize;
    // Exact accuracy count (bf16 operands; normalised rows bound the error of a tensor-core score): needs a source of
    // exact fp32 rows for every column -- the inputs themselves when this call covers the whole batch
    if (precision == SIMCLR_PRECISION_BF16 && (normalize || loss == SIMCLR_LOSS_MODIFIED)) {
        if (exact.in_dtype != SIMCLR_DTYPE_F32 && exact.in_dtype != SIMCLR_DTYPE_BF16) return SIMCLR_ERR_BAD_DTYPE;
        bool have = false;
        if (exact.zrows_peers != nullptr && world > 0) {
            if ((rc = make_peer_table(world, rank, exact.zrows_peers, &p.zrows_peers))) return rc;
            have = true;
        } else if (exact.zrows != nullptr) {
            if (misaligned(exact.zrows)) return SIMCLR_ERR_MISALIGNED;
            p.zrows = exact.zrows;
            have = true;
        } else if (exact.x1 && exact.x2 && exact.inv_norm && b_local == b_global) {
            p.x1 = exact.x1;
            p.x2 = exact.x2;
            p.bn_state = exact.bn_state;
            p.inv_norm = exact.inv_norm;
            p.in_bf16 = exact.in_dtype == SIMCLR_DTYPE_BF16 ? 1 : 0;
            have = true;
        }
        if (have) {
            p.cand_cnt = w.cand_cnt;
            p.cand = w.cand;
            p.band = (loss == SIMCLR_LOSS_NTXENT ? s.k2 : 1.0f) * kBandRel;
            if (defer_stats && fused == nullptr && p.x1 != nullptr) {
                // fused one-GPU step: the exact re-scoring is left to the backward finalize kernel (header words 1 - 3 of the
                // workspace: zeroed by the prepare kernel with the ticket)
                p.defer_accuracy = 1;
                p.amb_cnt = w.ticket + 1;
                p.amb_done = w.ticket + 3;
                p.amb_list = w.amb_list;
            }
        }
    }
    if (fused != nullptr) {
        // the barrier after the operand push runs inside the tile kernel; the finalize kernel bumps the epoch for the
        // barrier inside the backward tile kernel
        if ((rc = make_peer_table(fused->world, fused->rank, fused->flag_peers, &p.sync_flags))) return rc;
        if (p.sync_flags.world < 1 || fused->epoch == nullptr) return SIMCLR_ERR_BAD_PEERS;
        p.sync_epoch = fused->epoch;
        p.bump_epoch = fused->epoch;
        p.sync_presignaled = 1;       // the prepare kernel's last warp published the epoch (windows: this kernel's spare warps)
    }
    if ((rc = make_peer_table(world, rank, colvec_peers, &p.colvec_peers))) return rc;
    if ((rc = make_peer_table(world, rank, stats_peers, &p.stats_peers))) return rc;
    if (p.colvec_peers.world > 0 && (b_global != b_local * world || row_offset != b_local * rank)) return SIMCLR_ERR_BAD_PEERS;
    if (backward_workspace != nullptr) {
        // prime the backward: zeroed accumulation buffer, column vectors (weighted losses need sum(w): not primed)
        if (row_weight != nullptr) return SIMCLR_ERR_BAD_PEERS;
        if (misaligned(backward_workspace)) return SIMCLR_ERR_MISALIGNED;
        BwdWorkspace bw = carve_backward(g, backward_workspace);
        if (backward_workspace_bytes < bw.bytes) return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
        p.prime_dacc = reinterpret_cast<float4*>(bw.dacc);
        p.prime_dacc_vec4 = bw.dacc_floats / 4;
        if (p.colvec_peers.world == 0) {
            if (b_local != b_global) return SIMCLR_ERR_BAD_PEERS;     // a shard cannot know the other ranks' lse2
            p.prime_colvec = bw.colvec;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tpr = static_cast<int>(b_local / kBlockM);                   // column tiles per rank and view
    const bool overlap = flag_peers != nullptr && epoch_local != nullptr && p.colvec_peers.world > 1 &&
                         b_local % kBlockM == 0 && g.tiles_per_view == tpr * world;
    if (flag_peers != nullptr && !overlap) {
        // the caller relies on this call to order the operand exchange: do it up front
        if ((rc = simclr_peer_barrier(world, rank, flag_peers, epoch_local, nullptr, nullptr, nullptr, stream))) return rc;
    }
    if (fused != nullptr && fused->windows) {
        // One launch, two column windows: own columns from the local operand rows while this kernel's spare warps push
        // those rows to the peers, then everybody else's columns behind the in-kernel wait (TileParams::n_windows).
        Geometry ga = g, gb = g;
        set_window(&ga, rank * tpr, tpr);
        set_window(&gb, ((rank + 1) % world) * tpr, g.tiles_per_view - tpr);
        if (ga.grid != g.grid || gb.grid != g.grid) return SIMCLR_ERR_BAD_PEERS;       // peer_windows_ok() guarantees it
        if (static_cast<size_t>(ga.grid) * ga.max_segs * kFwdFields * kBlockM * sizeof(float) > w.part_bytes ||
            static_cast<size_t>(gb.grid) * gb.max_segs * kFwdFields * kBlockM * sizeof(float) > w.part_bytes)
            return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
        p.n_windows = 2;
        p.win0_local = 1;
        p.win[0] = ColWindow{ga.col_start, ga.col_cnt, ga.n_col_tiles, ga.max_segs, ga.total_tiles, w.part};
        p.win[1] = ColWindow{gb.col_start, gb.col_cnt, gb.n_col_tiles, gb.max_segs, gb.total_tiles, w.part2};
        p.push_src = operand_rows;
        if ((rc = make_peer_table(world, rank, fused->operand_peers, &p.push_peers))) return rc;
        p.push_peers.mc = fused->operand_multicast;
        p.push_ticket = fused->epoch + 1;
        p.fin_set[0] = PartSet{w.part, ga.total_tiles, ga.n_col_tiles, ga.max_segs, ga.grid};
        p.fin_set[1] = PartSet{w.part2, gb.total_tiles, gb.n_col_tiles, gb.max_segs, gb.grid};
        p.n_fin_sets = 2;
        if ((stages & kStageFwdTile) && (rc = dispatch_tile<false>(loss, g.d_pad, precision, map_rows, map_cols, map_rows, p, g.grid, st))) return rc;
    } else if (!overlap) {
        p.fin_set[0] = PartSet{w.part, g.total_tiles, g.n_col_tiles, g.max_segs, g.grid};
        p.n_fin_sets = 1;
        if ((stages & kStageFwdTile) && (rc = dispatch_tile<false>(loss, g.d_pad, precision, map_rows, map_cols, map_rows, p, g.grid, st))) return rc;
    } else {
        // Overlapped exchange: the columns this rank produced itself need no peer, so their tiles run while the other
        // ranks' operand rows are still crossing NVLink; the barrier follows, then the remote columns.
        Geometry ga = g, gb = g;
        set_window(&ga, rank * tpr, tpr);
        set_window(&gb, ((rank + 1) % world) * tpr, g.tiles_per_view - tpr);
        if (static_cast<size_t>(ga.grid) * ga.max_segs * kFwdFields * kBlockM * sizeof(float) > w.part_bytes ||
            static_cast<size_t>(gb.grid) * gb.max_segs * kFwdFields * kBlockM * sizeof(float) > w.part_bytes)
            return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
        TileParams pa = p, pb = p;
        auto window = [](TileParams& q, const Geometry& gw, float* part) {
            q.n_col_tiles = gw.n_col_tiles;
            q.col_start = gw.col_start;
            q.col_cnt = gw.col_cnt;
            q.total_tiles = gw.total_tiles;
            q.max_segs = gw.max_segs;
            q.tile_grid = gw.grid;
            q.part = part;
        };
        window(pa, ga, w.part);
        window(pb, gb, w.part2);
        pb.prime_dacc = nullptr;                                  // zeroed once, by the first launch
        if ((rc = dispatch_tile<false>(loss, g.d_pad, precision, map_rows, map_cols, map_rows, pa, ga.grid, st))) return rc;
        if ((rc = simclr_peer_barrier(world, rank, flag_peers, epoch_local, nullptr, nullptr, nullptr, stream))) return rc;
        if ((rc = dispatch_tile<false>(loss, g.d_pad, precision, map_rows, map_cols, map_rows, pb, gb.grid, st))) return rc;
        p.fin_set[0] = PartSet{w.part, ga.total_tiles, ga.n_col_tiles, ga.max_segs, ga.grid};
        p.fin_set[1] = PartSet{w.part2, gb.total_tiles, gb.n_col_tiles, gb.max_segs, gb.grid};
        p.n_fin_sets = 2;
    }
    cudaError_t e;
    if (!(stages & kStageFwdFin)) return SIMCLR_OK;
    // (the fused one-GPU step has its own, leaner instantiation: see forward_finalize_rowblock)
    const bool lean = p.defer_stats != 0 && (p.cand_cnt == nullptr || p.defer_accuracy != 0);
    if (loss == SIMCLR_LOSS_NTXENT) {
        e = lean ? launch_pdl(forward_finalize_kernel<kNtXent, true>, dim3(g.n_row_blocks), dim3(kBlockM), 0, st, p)
                 : launch_pdl(forward_finalize_kernel<kNtXent, false>, dim3(g.n_row_blocks), dim3(kBlockM), 0, st, p);
    } else {
        e = lean ? launch_pdl(forward_finalize_kernel<kModified, true>, dim3(g.n_row_blocks), dim3(kBlockM), 0, st, p)
                 : launch_pdl(forward_finalize_kernel<kModified, false>, dim3(g.n_row_blocks), dim3(kBlockM), 0, st, p);
    }
    return static_cast<int>(e);
}

// finish_ws: forward workspace of the same fused step whose finalize kernel deferred the loss statistics, or nullptr
int backward_impl(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t b_global,
                  int64_t row_offset, int64_t d, int in_dtype, int normalize, float temperature, int precision,
                  const void* operand_rows, const void* operand_cols, const float* inv_norm, const float* pos_dot,
                  const float* lse2_cols, const float* col_scale, const float* grad_out, void* grad1, void* grad2,
                  void* workspace, size_t workspace_bytes, const float* primed_colvec, void* stream, void* finish_ws,
                  float* finish_stats, float* finish_loss, const FusedSync* fused = nullptr, unsigned stages = kAllStages,
                  int flags = 0, const simclr_bn_t* bn = nullptr) {
    if (!x_batch1 || !x_batch2 || !operand_rows || !operand_cols || !inv_norm || !pos_dot || !grad1 || !grad2 || !workspace)
        return SIMCLR_ERR_NULL_POINTER;
    if (!lse2_cols && !primed_colvec) return SIMCLR_ERR_NULL_POINTER;
    if (primed_colvec && col_scale) return SIMCLR_ERR_BAD_PEERS;      // weighted losses are never primed
    if (in_dtype != SIMCLR_DTYPE_F32 && in_dtype != SIMCLR_DTYPE_BF16) return SIMCLR_ERR_BAD_DTYPE;
    if (!(temperature > 0.f) || !std::isfinite(temperature)) return SIMCLR_ERR_BAD_TEMPERATURE;
    Geometry g;
    int rc = make_geometry(loss, b_local, b_global, row_offset, d, &g);
    if (rc) return rc;
    if (misaligned(operand_rows) || misaligned(operand_cols) || misaligned(workspace)) return SIMCLR_ERR_MISALIGNED;
    const bool deterministic = (flags & SIMCLR_FLAG_DETERMINISTIC) != 0;
    BwdWorkspace w = carve_backward(g, workspace, deterministic);
    if (workspace_bytes < w.bytes) return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = check_device())) return rc;

    if (bad_precision(precision)) return SIMCLR_ERR_BAD_DTYPE;
    const int planes = precision == SIMCLR_PRECISION_SPLIT ? 2 : 1;       // hi (+ lo) operand planes
    CUtensorMap map_rows, map_cols;
    if ((rc = make_operand_map(&map_rows, operand_rows, planes * 2 * g.bl_pad, g.d_pad))) return rc;
    if ((rc = make_operand_map(&map_cols, operand_cols, planes * 2 * g.bg_pad, g.d_pad))) return rc;
    // the accumulator flush targets the shared accumulation buffer (TMA reduce-add) or, in deterministic mode, the
    // (CTA, segment) slots (TMA store)
    CUtensorMap map_dacc;
    if ((rc = deterministic ? make_dacc_map(&map_dacc, w.det_part, w.det_rows, g.d_pad)
                            : make_dacc_map(&map_dacc, w.dacc, 2 * g.bl_pad, g.d_pad)))
        return rc;

    Scales s = make_scales(loss, temperature, normalize, b_global);
    AuxParams a = make_aux(g, s, b_local, b_global, row_offset, d, normalize);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    if (primed_colvec == nullptr && (stages & kStageBwdPrepare)) {
        if ((rc = static_cast<int>(launch_pdl(backward_prepare_kernel, dim3(device_info().sm_count * 2), dim3(256), 0, st, a,
                                              lse2_cols, col_scale, w.colvec, reinterpret_cast<float4*>(w.dacc),
                                              w.dacc_floats / 4, g_ktrace_ptr))))
            return rc;
    }

    TileParams p = make_tile_params(g, s, b_local, b_global, row_offset);
    p.d = static_cast<int>(d);
    p.in_bf16 = in_dtype == SIMCLR_DTYPE_BF16 ? 1 : 0;
    p.normalize = normalize;
    p.colvec = primed_colvec ? primed_colvec : w.colvec;
    p.dacc = w.dacc;
    p.det_part = w.det_part;
    p.deterministic = deterministic ? 1 : 0;
    if (bn != nullptr) {
        if (!bn->state || !bn->partial) return SIMCLR_ERR_NULL_POINTER;
        p.bn_state = bn->state;
        p.bn_partial = bn->partial;
    }
    p.x1 = x_batch1;
    p.x2 = x_batch2;
    p.g1 = grad1;
    p.g2 = grad2;
    p.inv_norm = inv_norm;
    p.pos_dot = pos_dot;
    p.col_scale = col_scale;
    p.grad_out = grad_out;
    // A primed backward follows the forward of the same operands; when the tile kernels fill the device (1 CTA per SM
    // by shared memory and TMEM), no CTA of this launch can become resident before a forward tile CTA -- which waited
    // for the complete operand matrix -- has exited: the operand loads and score MMAs need not wait for the finalize
    // kernel in between (TileParams::early_operand).
    p.early_operand = (primed_colvec != nullptr && g.grid == device_info().sm_count) ? 1 : 0;
#ifdef SIMCLR_NO_EARLY_OPERAND
    p.early_operand = 0;
#endif
    if (finish_ws != nullptr) {
        FwdWorkspace fw = carve_forward(g, finish_ws);
        p.block_part = fw.block_part;
        p.stats = finish_stats;
        p.loss_out = finish_loss;
        p.finish_stats = 1;
        if (precision == SIMCLR_PRECISION_BF16 && (normalize || loss == SIMCLR_LOSS_MODIFIED) && b_local == b_global) {
            // the forward finalize kernel of this fused step listed the rows it could not decide (forward_impl)
            p.resolve_ambiguous = 1;
            p.amb_cnt = fw.ticket + 1;
            p.amb_ticket = fw.ticket + 2;
            p.amb_done = fw.ticket + 3;
            p.amb_list = fw.amb_list;
            p.cand_cnt = fw.cand_cnt;
            p.cand = fw.cand;
        }
    }
    if (fused != nullptr) {
        if ((rc = make_peer_table(fused->world, fused->rank, fused->flag_peers, &p.sync_flags))) return rc;
        if (p.sync_flags.world < 1 || fused->epoch == nullptr || fused->stats_all == nullptr || finish_stats == nullptr)
            return SIMCLR_ERR_BAD_PEERS;
        p.sync_epoch = fused->epoch;
        p.sync_presignaled = 1;       // the forward finalize kernel's last block published the epoch
        p.stats_all = fused->stats_all;
        p.stats_world = fused->world;
        p.stats = finish_stats;
        p.loss_out = finish_loss;
        p.finish_stats = 1;
    }
    if ((stages & kStageBwdTile) &&
        (rc = dispatch_tile<true>(loss, g.d_pad, precision, map_rows, map_cols, map_dacc, p, g.grid, st)))
        return rc;
    cudaError_t fin_rc = cudaSuccess;
    if (!(stages & kStageBwdFin)) return SIMCLR_OK;
    // (+ the blocks that re-score the rows the forward finalize kernel listed, see resolve_ambiguous_rows)
    const int fin_blocks = g.n_row_blocks * kBwdFinBlocksPerRowBlock + (p.resolve_ambiguous ? kResolveBlocks : 0);
#define SIMCLR_BFIN(DV)                                                                                  \
    case DV:                                                                                             \
        if (loss == SIMCLR_LOSS_NTXENT) fin_rc = deterministic ? launch_pdl(backward_finalize_kernel<DV, kNtXent, true>, dim3(fin_blocks), dim3(512), 0, st, p) : launch_pdl(backward_finalize_kernel<DV, kNtXent, false>, dim3(fin_blocks), dim3(512), 0, st, p); \
        else fin_rc = deterministic ? launch_pdl(backward_finalize_kernel<DV, kModified, true>, dim3(fin_blocks), dim3(512), 0, st, p) : launch_pdl(backward_finalize_kernel<DV, kModified, false>, dim3(fin_blocks), dim3(512), 0, st, p); \
        break;
    switch (g.d_pad) {
        SIMCLR_BFIN(64)
        SIMCLR_BFIN(128)
        SIMCLR_BFIN(256)
    }
#undef SIMCLR_BFIN
    return static_cast<int>(fin_rc);
}

}  // namespace

extern "C" {

int simclr_forward_peer(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                        int64_t row_offset, int64_t d, float temperature, int normalize, int precision,
                        const float* pos_dot,
                        const float* row_weight, float* lse2, float* row_loss, float* stats, float* loss_out,
                        void* workspace, size_t workspace_bytes, void* backward_workspace,
                        size_t backward_workspace_bytes, int world, int rank, void* const* colvec_peers,
                        void* const* stats_peers, void* const* flag_peers, unsigned int* epoch_local,
                        const void* x_batch1, const void* x_batch2, int in_dtype, const float* inv_norm,
                        const float* zrows_global, void* const* zrows_peers, void* stream) {
    ExactSource ex;
    ex.x1 = x_batch1;
    ex.x2 = x_batch2;
    ex.in_dtype = in_dtype;
    ex.inv_norm = inv_norm;
    ex.zrows = zrows_global;
    ex.zrows_peers = zrows_peers;
    return forward_impl(loss, operand_rows, operand_cols, b_local, b_global, row_offset, d, temperature, normalize, precision,
                        pos_dot, row_weight, lse2, row_loss, stats, loss_out, workspace, workspace_bytes, backward_workspace,
                        backward_workspace_bytes, world, rank, colvec_peers, stats_peers, flag_peers, epoch_local, stream,
                        false, nullptr, kAllStages, ex);
}

int simclr_forward(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                   int64_t row_offset, int64_t d, float temperature, int normalize, const float* pos_dot,
                   const float* row_weight, float* lse2, float* row_loss, float* stats, float* loss_out, void* workspace,
                   size_t workspace_bytes, const void* x_batch1, const void* x_batch2, int in_dtype, const float* inv_norm,
                   void* stream) {
    return simclr_forward_peer(loss, operand_rows, operand_cols, b_local, b_global, row_offset, d, temperature, normalize,
                               SIMCLR_PRECISION_BF16, pos_dot, row_weight, lse2, row_loss, stats, loss_out, workspace,
                               workspace_bytes, nullptr, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, x_batch1, x_batch2,
                               in_dtype, inv_norm, nullptr, nullptr, stream);
}

int simclr_backward(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t b_global,
                    int64_t row_offset, int64_t d, int in_dtype, int normalize, float temperature, int precision,
                    const void* operand_rows, const void* operand_cols, const float* inv_norm, const float* pos_dot,
                    const float* lse2_cols, const float* col_scale, const float* grad_out, void* grad1, void* grad2,
                    void* workspace, size_t workspace_bytes, const float* primed_colvec, int flags, void* stream) {
    return backward_impl(loss, x_batch1, x_batch2, b_local, b_global, row_offset, d, in_dtype, normalize, temperature,
                         precision, operand_rows, operand_cols, inv_norm, pos_dot, lse2_cols, col_scale, grad_out, grad1,
                         grad2, workspace, workspace_bytes, primed_colvec, stream, nullptr, nullptr, nullptr, nullptr,
                         kAllStages, flags);
}

int simclr_forward_stages(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                          int64_t row_offset, int64_t d, float temperature, int normalize, int precision,
                          const float* pos_dot, const float* row_weight, float* lse2, float* row_loss, float* stats,
                          float* loss_out, void* workspace, size_t workspace_bytes, void* backward_workspace,
                          size_t backward_workspace_bytes, const void* x_batch1, const void* x_batch2, int in_dtype,
                          const float* inv_norm, void* stream, unsigned int stage_mask) {
    ExactSource ex;
    ex.x1 = x_batch1;
    ex.x2 = x_batch2;
    ex.in_dtype = in_dtype;
    ex.inv_norm = inv_norm;
    return forward_impl(loss, operand_rows, operand_cols, b_local, b_global, row_offset, d, temperature, normalize, precision,
                        pos_dot, row_weight, lse2, row_loss, stats, loss_out, workspace, workspace_bytes, backward_workspace,
                        backward_workspace_bytes, 0, 0, nullptr, nullptr, nullptr, nullptr, stream, false, nullptr,
                        stage_mask, ex);
}

int simclr_backward_stages(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t b_global,
                           int64_t row_offset, int64_t d, int in_dtype, int normalize, float temperature, int precision,
                           const void* operand_rows, const void* operand_cols, const float* inv_norm, const float* pos_dot,
                           const float* lse2_cols, const float* col_scale, const float* grad_out, void* grad1, void* grad2,
                           void* workspace, size_t workspace_bytes, const float* primed_colvec, int flags, void* stream,
                           unsigned int stage_mask) {
    return backward_impl(loss, x_batch1, x_batch2, b_local, b_global, row_offset, d, in_dtype, normalize, temperature,
                         precision, operand_rows, operand_cols, inv_norm, pos_dot, lse2_cols, col_scale, grad_out, grad1,
                         grad2, workspace, workspace_bytes, primed_colvec, stream, nullptr, nullptr, nullptr, nullptr,
                         stage_mask, flags);
}

namespace {
// The fused single-GPU step, whole (begin && finish) or split in two calls around the backward finalize kernel.
int fused_step_impl(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype, int normalize,
                    float temperature, int precision, const float* grad_out, void* operand, float* rowvec, float* stats,
                    float* loss_out, void* grad1, void* grad2, void* forward_workspace, size_t forward_workspace_bytes,
                    void* backward_workspace, size_t backward_workspace_bytes, void* stream, bool begin, bool finish,
                    int flags, const simclr_bn_t* bn = nullptr) {
    if (!rowvec || !backward_workspace) return SIMCLR_ERR_NULL_POINTER;
    if (begin && !stats) return SIMCLR_ERR_NULL_POINTER;
    const int64_t bp = simclr_pad_rows(b);
    if (bp == 0) return SIMCLR_ERR_BAD_SHAPE;
    float* inv_norm = rowvec;
    float* pos_dot = rowvec + 2 * bp;
    float* lse2 = rowvec + 4 * bp;
    float* row_loss = rowvec + 6 * bp;
    int rc;
    if (begin) {
        rc = prepare_impl(loss, x_batch1, x_batch2, b, d, in_dtype, normalize, temperature, precision, operand, inv_norm,
                          pos_dot, forward_workspace, 0, 0, nullptr, nullptr, stream, nullptr, nullptr,
                          bn ? bn->state : nullptr);
        if (rc) return rc;
        ExactSource ex;
        ex.x1 = x_batch1;
        ex.x2 = x_batch2;
        ex.in_dtype = in_dtype;
        ex.inv_norm = inv_norm;
        ex.bn_state = bn ? bn->state : nullptr;
        // The forward primes the backward workspace.  Whole step: the reduction of the loss statistics is left to the
        // backward finalize kernel (five launches, nothing but the column vectors between the two tile kernels).  Split
        // step: the forward finalize kernel completes them, so that the caller can read loss / accuracy while the
        // backward tile kernel is still running.
        rc = forward_impl(loss, operand, operand, b, b, 0, d, temperature, normalize, precision, pos_dot, nullptr, lse2,
                          row_loss, stats, loss_out, forward_workspace, forward_workspace_bytes, backward_workspace,
                          backward_workspace_bytes, 0, 0, nullptr, nullptr, nullptr, nullptr, stream, finish, nullptr,
                          kAllStages, ex);
        if (rc) return rc;
    }
    const unsigned stages = (begin ? kStageBwdTile : 0u) | (finish ? kStageBwdFin : 0u);
    // grad1 / grad2 are only written by the finalize kernel; a `begin` call may pass NULL
    void* g1 = grad1 ? grad1 : const_cast<void*>(x_batch1);
    void* g2 = grad2 ? grad2 : const_cast<void*>(x_batch2);
    if (finish && (!grad1 || !grad2)) return SIMCLR_ERR_NULL_POINTER;
    return backward_impl(loss, x_batch1, x_batch2, b, b, 0, d, in_dtype, normalize, temperature, precision, operand, operand,
                         inv_norm, pos_dot, nullptr, nullptr, grad_out, g1, g2, backward_workspace,
                         backward_workspace_bytes, static_cast<const float*>(backward_workspace), stream,
                         (begin && finish) ? forward_workspace : nullptr, stats, loss_out, nullptr, stages, flags, bn);
}
}  // namespace

int simclr_forward_backward(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                            int normalize, float temperature, int precision, const float* grad_out, void* operand,
                            float* rowvec, float* stats, float* loss_out, void* grad1, void* grad2,
                            void* forward_workspace, size_t forward_workspace_bytes, void* backward_workspace,
                            size_t backward_workspace_bytes, int flags, void* stream) {
    return fused_step_impl(loss, x_batch1, x_batch2, b, d, in_dtype, normalize, temperature, precision, grad_out, operand,
                           rowvec, stats, loss_out, grad1, grad2, forward_workspace, forward_workspace_bytes,
                           backward_workspace, backward_workspace_bytes, stream, true, true, flags);
}

int simclr_forward_backward_begin(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                                  int normalize, float temperature, int precision, void* operand, float* rowvec,
                                  float* stats, float* loss_out, void* forward_workspace, size_t forward_workspace_bytes,
                                  void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream) {
    return fused_step_impl(loss, x_batch1, x_batch2, b, d, in_dtype, normalize, temperature, precision, nullptr, operand,
                           rowvec, stats, loss_out, nullptr, nullptr, forward_workspace, forward_workspace_bytes,
                           backward_workspace, backward_workspace_bytes, stream, true, false, flags);
}

int simclr_forward_backward_finish(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                                   int normalize, float temperature, int precision, const float* grad_out,
                                   const void* operand, const float* rowvec, void* grad1, void* grad2,
                                   void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream) {
    return fused_step_impl(loss, x_batch1, x_batch2, b, d, in_dtype, normalize, temperature, precision, grad_out,
                           const_cast<void*>(operand), const_cast<float*>(rowvec), nullptr, nullptr, grad1, grad2, nullptr, 0,
                           backward_workspace, backward_workspace_bytes, stream, false, true, flags);
}

// ---- projection-head tail (include/simclr_b200.h, "Projection-head tail") ----
size_t simclr_bn_state_floats(int64_t d) {
    const int64_t dp = simclr_pad_dim(d);
    return dp == 0 ? 0 : static_cast<size_t>(2) * kBnPlanes * dp;
}

size_t simclr_bn_workspace_bytes(int64_t b, int64_t d) {
    const int64_t bp = simclr_pad_rows(b), dp = simclr_pad_dim(d);
    if (bp == 0 || dp == 0) return 0;
    const size_t stats_parts = static_cast<size_t>(2) * ((b + kBnRowsPerBlock - 1) / kBnRowsPerBlock) * 2 * dp;   // bn_stats partials
    const size_t bwd_parts = static_cast<size_t>(2 * bp / 16) * 2 * dp;                                        // finalize CTAs
    return 256 + sizeof(float) * (stats_parts > bwd_parts ? stats_parts : bwd_parts);
}

int simclr_bn_stats(const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype, const float* gamma, const float* beta,
                    float eps, float* bn_state, void* workspace, size_t workspace_bytes, void* stream) {
    if (!u1 || !u2 || !gamma || !beta || !bn_state || !workspace) return SIMCLR_ERR_NULL_POINTER;
    if (in_dtype != SIMCLR_DTYPE_F32 && in_dtype != SIMCLR_DTYPE_BF16) return SIMCLR_ERR_BAD_DTYPE;
    const int64_t dp = simclr_pad_dim(d);
    if (b < 1 || d < 1) return SIMCLR_ERR_BAD_SHAPE;
    if (dp == 0) return SIMCLR_ERR_UNSUPPORTED_DIM;
    if (workspace_bytes < simclr_bn_workspace_bytes(b, d)) return SIMCLR_ERR_WORKSPACE_TOO_SMALL;
    if (misaligned(workspace) || misaligned(bn_state)) return SIMCLR_ERR_MISALIGNED;
    int rc = check_device();
    if (rc) return rc;
    // the first 256 bytes of the workspace hold the ticket of the last-block reduction: zero on entry (the caller zeroes
    // the workspace once), left zero
    unsigned int* ticket = static_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    const int blocks = static_cast<int>(2 * ((b + kBnRowsPerBlock - 1) / kBnRowsPerBlock));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (in_dtype == SIMCLR_DTYPE_F32)
        e = launch_pdl(bn_stats_kernel<float>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(u1),
                       static_cast<const float*>(u2), static_cast<int>(b), static_cast<int>(d), static_cast<int>(dp), gamma, beta,
                       eps, bn_state, partial, ticket);
    else
        e = launch_pdl(bn_stats_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(u1),
                       static_cast<const __nv_bfloat16*>(u2), static_cast<int>(b), static_cast<int>(d), static_cast<int>(dp), gamma,
                       beta, eps, bn_state, partial, ticket);
    return static_cast<int>(e);
}

int simclr_head_forward(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype, int normalize,
                        float temperature, int precision, const float* bn_state, void* operand, float* rowvec, float* stats,
                        float* loss_out, void* forward_workspace, size_t forward_workspace_bytes, void* stream) {
    if (!rowvec || !stats || !bn_state) return SIMCLR_ERR_NULL_POINTER;
    const int64_t bp = simclr_pad_rows(b);
    if (bp == 0) return SIMCLR_ERR_BAD_SHAPE;
    float* inv_norm = rowvec;
    float* pos_dot = rowvec + 2 * bp;
    int rc = prepare_impl(loss, u1, u2, b, d, in_dtype, normalize, temperature, precision, operand, inv_norm, pos_dot,
                          forward_workspace, 0, 0, nullptr, nullptr, stream, nullptr, nullptr, bn_state);
    if (rc) return rc;
    ExactSource ex;
    ex.x1 = u1;
    ex.x2 = u2;
    ex.in_dtype = in_dtype;
    ex.inv_norm = inv_norm;
    ex.bn_state = bn_state;
    return forward_impl(loss, operand, operand, b, b, 0, d, temperature, normalize, precision, pos_dot, nullptr, rowvec + 4 * bp,
                        rowvec + 6 * bp, stats, loss_out, forward_workspace, forward_workspace_bytes, nullptr, 0, 0, 0, nullptr,
                        nullptr, nullptr, nullptr, stream, false, nullptr, kAllStages, ex);
}

int simclr_head_forward_backward_begin(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype,
                                       int normalize, float temperature, int precision, const simclr_bn_t* bn, void* operand,
                                       float* rowvec, float* stats, float* loss_out, void* forward_workspace,
                                       size_t forward_workspace_bytes, void* backward_workspace,
                                       size_t backward_workspace_bytes, int flags, void* stream) {
    if (!bn) return SIMCLR_ERR_NULL_POINTER;
    return fused_step_impl(loss, u1, u2, b, d, in_dtype, normalize, temperature, precision, nullptr, operand, rowvec, stats,
                           loss_out, nullptr, nullptr, forward_workspace, forward_workspace_bytes, backward_workspace,
                           backward_workspace_bytes, stream, true, false, flags, bn);
}

int simclr_head_forward_backward_finish(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype,
                                        int normalize, float temperature, int precision, const float* grad_out,
                                        const simclr_bn_t* bn, int bn_training, const void* operand, const float* rowvec,
                                        void* grad_u1, void* grad_u2, float* grad_gamma, float* grad_beta,
                                        void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream) {
    if (!bn || !bn->state || !bn->partial) return SIMCLR_ERR_NULL_POINTER;
    int rc = fused_step_impl(loss, u1, u2, b, d, in_dtype, normalize, temperature, precision, grad_out,
                             const_cast<void*>(operand), const_cast<float*>(rowvec), nullptr, nullptr, grad_u1, grad_u2, nullptr,
                             0, backward_workspace, backward_workspace_bytes, stream, false, true, flags, bn);
    if (rc) return rc;
    // dL/dz (just written to grad_u1 / grad_u2) -> dL/du, dL/dgamma, dL/dbeta
    const int64_t dp = simclr_pad_dim(d), bp = simclr_pad_rows(b);
    const int blocks = static_cast<int>(2 * ((b + kBnRowsPerBlock - 1) / kBnRowsPerBlock));
    const int parts_per_view = static_cast<int>(bp / 16);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (in_dtype == SIMCLR_DTYPE_F32)
        e = launch_pdl(bn_backward_kernel<float>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(u1),
                       static_cast<const float*>(u2), static_cast<float*>(grad_u1), static_cast<float*>(grad_u2),
                       static_cast<int>(b), static_cast<int>(d), static_cast<int>(dp), bn->state,
                       static_cast<const float*>(bn->partial), parts_per_view, bn_training, grad_gamma, grad_beta);
    else
        e = launch_pdl(bn_backward_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(u1),
                       static_cast<const __nv_bfloat16*>(u2), static_cast<__nv_bfloat16*>(grad_u1),
                       static_cast<__nv_bfloat16*>(grad_u2), static_cast<int>(b), static_cast<int>(d), static_cast<int>(dp),
                       bn->state, static_cast<const float*>(bn->partial), parts_per_view, bn_training, grad_gamma, grad_beta);
    return static_cast<int>(e);
}

int simclr_forward_backward_peer(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d,
                                 int in_dtype, int normalize, float temperature, const float* grad_out, void* operand,
                                 float* rowvec, float* stats_local, float* stats_global, float* loss_out, void* grad1,
                                 void* grad2, void* forward_workspace, size_t forward_workspace_bytes,
                                 void* backward_workspace, size_t backward_workspace_bytes, int world, int rank,
                                 void* const* operand_global_peers, void* operand_global_multicast,
                                 void* const* colvec_peers, void* const* stats_peers, void* const* flag_peers,
                                 unsigned int* epoch_local, void* const* zrows_peers, int flags, void* stream) {
    if (!rowvec || !stats_local || !stats_global || !backward_workspace || !operand_global_peers || !colvec_peers ||
        !stats_peers || !flag_peers || !epoch_local)
        return SIMCLR_ERR_NULL_POINTER;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return SIMCLR_ERR_BAD_PEERS;
    const int64_t bp = simclr_pad_rows(b_local);
    if (bp == 0) return SIMCLR_ERR_BAD_SHAPE;
    float* inv_norm = rowvec;
    float* pos_dot = rowvec + 2 * bp;
    float* lse2 = rowvec + 4 * bp;
    float* row_loss = rowvec + 6 * bp;
    const int64_t b_global = b_local * world, row_offset = b_local * rank;
    const void* operand_cols = operand_global_peers[rank];
    const float* colvec_local = static_cast<const float*>(colvec_peers[rank]);
    FusedSync fs{world, rank, flag_peers, epoch_local, static_cast<const float*>(stats_peers[rank])};
    if (check_device() == SIMCLR_OK && peer_windows_ok(loss, b_local, world)) {
        fs.windows = true;
        fs.operand_peers = operand_global_peers;
        fs.operand_multicast = operand_global_multicast;
    }
    // prepare bumps the epoch and pushes the operand rows, publishing the epoch when its last warp is done -- or, with the
    // two-window forward, leaves push and signal to the forward tile kernel, which overlaps them with its own columns'
    // tiles.  The exact fp32 rows for the accuracy candidates stay on the rank that owns them (zrows_peers[rank]); the
    // finalize kernels of the other ranks read the few rows they need over NVLink
    int rc = prepare_impl(loss, x_batch1, x_batch2, b_local, d, in_dtype, normalize, temperature, SIMCLR_PRECISION_BF16,
                          operand, inv_norm, pos_dot, forward_workspace, fs.windows ? 0 : world, rank,
                          fs.windows ? nullptr : operand_global_peers, fs.windows ? nullptr : operand_global_multicast, stream,
                          epoch_local, zrows_peers ? static_cast<float*>(zrows_peers[rank]) : nullptr, nullptr,
                          fs.windows ? nullptr : flag_peers, fs.windows ? world : 0);
    if (rc) return rc;
    ExactSource ex;
    ex.zrows_peers = zrows_peers;
    rc = forward_impl(loss, operand, operand_cols, b_local, b_global, row_offset, d, temperature, normalize,
                      SIMCLR_PRECISION_BF16, pos_dot, nullptr, lse2, row_loss, stats_local, nullptr, forward_workspace,
                      forward_workspace_bytes, backward_workspace, backward_workspace_bytes, world, rank, colvec_peers,
                      stats_peers, nullptr, nullptr, stream, false, &fs, kAllStages, ex);
    if (rc) return rc;
    return backward_impl(loss, x_batch1, x_batch2, b_local, b_global, row_offset, d, in_dtype, normalize, temperature,
                         SIMCLR_PRECISION_BF16, operand, operand_cols, inv_norm, pos_dot, nullptr, nullptr, grad_out, grad1,
                         grad2, backward_workspace, backward_workspace_bytes, colvec_local, stream, nullptr, stats_global,
                         loss_out, &fs, kAllStages, flags);
}

#if SIMCLR_TRACE
int simclr_debug_set_trace(void* device_buffer, int cta) {
    g_trace_ptr = static_cast<long long*>(device_buffer);
    g_trace_cta = cta;
    return SIMCLR_OK;
}

int simclr_debug_set_kernel_trace(void* device_buffer) {
    g_ktrace_ptr = static_cast<unsigned long long*>(device_buffer);
    return SIMCLR_OK;
}

int simclr_debug_mma_rate(long long* out_device, int batches, int grid, int mode, float* sink, void* stream) {
    if (!out_device || !sink) return SIMCLR_ERR_NULL_POINTER;
    int rc = check_device();
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRateSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    mma_rate_kernel<<<grid, 640, kRateSmemBytes, static_cast<cudaStream_t>(stream)>>>(out_device, batches, mode, sink);
    return static_cast<int>(cudaGetLastError());
}

int simclr_debug_chunk_rate(long long* out_device, int iters, int grid, int nwarps, float k2, float* sink, void* stream) {
    if (!out_device || !sink) return SIMCLR_ERR_NULL_POINTER;
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define SIMCLR_RATE(V) chunk_rate_kernel<V><<<grid, kThreadsForward, 0, st>>>(out_device, iters, nwarps, k2, sink);
    SIMCLR_RATE(0) SIMCLR_RATE(1) SIMCLR_RATE(2) SIMCLR_RATE(3) SIMCLR_RATE(4) SIMCLR_RATE(5) SIMCLR_RATE(6) SIMCLR_RATE(7)
    SIMCLR_RATE(8) SIMCLR_RATE(9)
#undef SIMCLR_RATE
    return static_cast<int>(cudaGetLastError());
}

int simclr_debug_pipe_rate(long long* out_device, int iters, int grid, int nwarps, float* sink, void* stream) {
    if (!out_device || !sink) return SIMCLR_ERR_NULL_POINTER;
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define SIMCLR_PIPE(V) pipe_rate_kernel<V><<<grid, kThreadsForward, 0, st>>>(out_device, iters, nwarps, 1.0f, sink);
    SIMCLR_PIPE(0) SIMCLR_PIPE(1) SIMCLR_PIPE(2) SIMCLR_PIPE(3) SIMCLR_PIPE(4) SIMCLR_PIPE(5) SIMCLR_PIPE(6) SIMCLR_PIPE(7)
    SIMCLR_PIPE(8) SIMCLR_PIPE(9) SIMCLR_PIPE(10) SIMCLR_PIPE(11) SIMCLR_PIPE(12) SIMCLR_PIPE(13) SIMCLR_PIPE(14) SIMCLR_PIPE(15)
    SIMCLR_PIPE(16) SIMCLR_PIPE(17) SIMCLR_PIPE(18) SIMCLR_PIPE(19) SIMCLR_PIPE(20) SIMCLR_PIPE(21) SIMCLR_PIPE(22) SIMCLR_PIPE(23)
#undef SIMCLR_PIPE
    return static_cast<int>(cudaGetLastError());
}

int simclr_selftest_umma(const void* a_bf16, const void* b_bf16, float* out_f32, void* stream) {
    if (!a_bf16 || !b_bf16 || !out_f32) return SIMCLR_ERR_NULL_POINTER;
    int rc = check_device();
    if (rc) return rc;
    CUtensorMap map_a, map_b;
    if ((rc = make_operand_map(&map_a, a_bf16, 128, 128))) return rc;
    if ((rc = make_operand_map(&map_b, b_bf16, 128, 128))) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSelftestSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = true;
    }
    selftest_umma_kernel<<<1, 128, kSelftestSmemBytes, static_cast<cudaStream_t>(stream)>>>(map_a, map_b, out_f32);
    return static_cast<int>(cudaGetLastError());
}

#endif  // SIMCLR_TRACE

}  // extern "C"
