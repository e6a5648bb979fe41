// Error reporting + trivial entry points of the C ABI (include/stove_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void stove_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* stove_last_error(void) { return g_err; }
extern "C" int stove_abi_version(void) { return 1; }

// ---------------------------------------------------------------------------------------
// launch accounting + optional per-kernel event timing
// ---------------------------------------------------------------------------------------
#include <vector>
#include <mutex>
#include "kernel_names.inc"

struct ProfRec {
    int id;
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static bool g_prof_on = false;
static int64_t g_launches = 0;

int stove_prof_begin(int id, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ++g_launches;
    if (!g_prof_on) return -1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return -1;
    ProfRec r;
    r.id = id;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!g_pool.empty()) {
            *e = g_pool.back();
            g_pool.pop_back();
        } else if (cudaEventCreate(e) != cudaSuccess) {
            return -1;
        }
    }
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
    return (int)g_recs.size() - 1;
}

void stove_prof_end(int slot, cudaStream_t s) {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, s);
}

extern "C" int stove_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return STOVE_OK;
}

// host arrays; waits for the recorded events, returns the number of records (and clears them)
extern "C" int stove_profile_read(int32_t* ids, float* ms, int max_n) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (ProfRec& r : g_recs) {
        float t = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&t, r.a, r.b);
        if (n < max_n && ids && ms) {
            ids[n] = r.id;
            ms[n] = t;
            ++n;
        }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_recs.clear();
    return n;
}

// ---- fork/join helper ---------------------------------------------------------------------
static StoveFork g_forks[16][STOVE_FORK_FAMILIES];
static bool g_fork_ok[16][STOVE_FORK_FAMILIES];

StoveFork* stove_fork_get(int family) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16 || family < 0 || family >= STOVE_FORK_FAMILIES) {
        stove_set_error("stove_fork_get: bad device / family");
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_prof_mu);
    StoveFork* f = &g_forks[dev][family];
    if (!g_fork_ok[dev][family]) {
        bool ok = true;
        for (int i = 0; i < 2 && ok; ++i) {
            ok = cudaStreamCreateWithFlags(&f->side[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&f->join_ev[i], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreateWithFlags(&f->fork_ev, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            stove_set_error("stove_fork_get: cannot create side streams");
            return nullptr;
        }
        g_fork_ok[dev][family] = true;
    }
    return f;
}

int stove_fork(StoveFork* f, cudaStream_t s, int nside) {
    STOVE_CUDA(cudaEventRecord(f->fork_ev, s));
    for (int i = 0; i < nside; ++i) STOVE_CUDA(cudaStreamWaitEvent(f->side[i], f->fork_ev, 0));
    return STOVE_OK;
}

int stove_join(StoveFork* f, cudaStream_t s, int nside) {
    for (int i = 0; i < nside; ++i) {
        STOVE_CUDA(cudaEventRecord(f->join_ev[i], f->side[i]));
        STOVE_CUDA(cudaStreamWaitEvent(s, f->join_ev[i], 0));
    }
    return STOVE_OK;
}

static int g_options[OPT_COUNT] = {1, 1, 0, 2, 0, 0, 2, 0, 0, 0, 74, 0};
static const char* const kOptionNames[OPT_COUNT] = {"fork", "spn2_nodes_stage", "dynloop_generic", "dynloop_nw",
                                                    "dynloop_recompute", "rollout_cta", "rollout_nw", "gnn_seq_fwd",
                                                    "gnn_seq_bwd", "head_par_ctas", "wgrad_ctas", "gnn_threads"};
int stove_opt(int id) { return g_options[id]; }
extern "C" int stove_set_option(const char* name, int value) {
    for (int i = 0; i < OPT_COUNT; ++i)
        if (name && strcmp(name, kOptionNames[i]) == 0) {
            const int prev = g_options[i];
            g_options[i] = value;
            return prev;
        }
    stove_set_error("stove_set_option: unknown option '%s'", name ? name : "(null)");
    return STOVE_ERR_ARG;
}
extern "C" int stove_get_option(const char* name) {
    for (int i = 0; i < OPT_COUNT; ++i)
        if (name && strcmp(name, kOptionNames[i]) == 0) return g_options[i];
    stove_set_error("stove_get_option: unknown option '%s'", name ? name : "(null)");
    return STOVE_ERR_ARG;
}

extern "C" const char* stove_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : "?"; }
extern "C" int stove_kernel_count(void) { return K_COUNT; }
extern "C" int64_t stove_launch_count(int reset) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

// ---------------------------------------------------------------------------------------
// bw_transform (model/utils/utils.py:10-15): y = clamp(sum_c x[:, c], 0, 1).
// Pure streaming: reads C*4 B (or C bytes for uint8 frames, scaled by 1/255 on the fly), writes 4 B per
// pixel; four pixels per thread when hw % 4 == 0.  Optionally also writes the (hi, lo) TF32 operand planes of
// y (the left operand of the recognition LSTM's input GEMM, csrc/lstm_tc.cu), saving that kernel a pass.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4(const uint8_t* p) {
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
    const float s = 1.0f / 255.0f;
    return make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
}
__device__ __forceinline__ float load1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load1(const uint8_t* p) { return __ldg(p) * (1.0f / 255.0f); }
__device__ __forceinline__ float tf32_hi_(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

template <typename TI, bool V4>
__global__ void bw_transform_kernel(const TI* __restrict__ x, const void* const* __restrict__ x_cell,
                                    float* __restrict__ y, float* __restrict__ pl, int64_t n, int C, int64_t hw) {
    constexpr int W = V4 ? 4 : 1;
    // indirect input: the frames' address is read from a device cell at run time, so a captured graph can be
    // pointed at a different batch without copying it into a static buffer
    if (x_cell) x = reinterpret_cast<const TI*>(*x_cell);
    const int64_t hww = hw / W, total = n * hww, plane = n * hw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / hww, p = (i - b * hww) * W;
        const TI* src = x + (b * C) * hw + p;
        const int64_t o = b * hw + p;
        if (V4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < C; ++c) {
                const float4 v = load4(src + c * hw);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            acc.x = fminf(fmaxf(acc.x, 0.f), 1.f);
            acc.y = fminf(fmaxf(acc.y, 0.f), 1.f);
            acc.z = fminf(fmaxf(acc.z, 0.f), 1.f);
            acc.w = fminf(fmaxf(acc.w, 0.f), 1.f);
            *reinterpret_cast<float4*>(y + o) = acc;
            if (pl) {
                const float4 hi = make_float4(tf32_hi_(acc.x), tf32_hi_(acc.y), tf32_hi_(acc.z), tf32_hi_(acc.w));
                *reinterpret_cast<float4*>(pl + o) = hi;
                *reinterpret_cast<float4*>(pl + plane + o) = make_float4(acc.x - hi.x, acc.y - hi.y, acc.z - hi.z, acc.w - hi.w);
            }
        } else {
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc += load1(src + c * hw);
            acc = fminf(fmaxf(acc, 0.f), 1.f);
            y[o] = acc;
            if (pl) {
                const float hi = tf32_hi_(acc);
                pl[o] = hi;
                pl[plane + o] = acc - hi;
            }
        }
    }
}

template <typename TI>
static int bw_launch(const TI* x, const void* const* x_cell, float* y, float* pl, int64_t n, int channels, int64_t hw,
                     cudaStream_t st) {
    const int threads = 256;
    // (an indirect source must be 16-byte aligned: the caller guarantees it)
    const bool v4 = (hw % 4 == 0) && (x_cell || ((uintptr_t)x % (4 * sizeof(TI))) == 0) && (((uintptr_t)y | (uintptr_t)pl) % 16 == 0);
    const int64_t items = v4 ? n * (hw / 4) : n * hw;
    int blocks = (int)((items + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (v4)
        STOVE_KERNEL(K_BW_TRANSFORM, st, bw_transform_kernel<TI, true><<<blocks, threads, 0, st>>>(x, x_cell, y, pl, n, channels, hw));
    else
        STOVE_KERNEL(K_BW_TRANSFORM, st, bw_transform_kernel<TI, false><<<blocks, threads, 0, st>>>(x, x_cell, y, pl, n, channels, hw));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_bw_transform(const float* x, float* y, int64_t n, int channels, int64_t hw,
                                  void* stream) {
    STOVE_CHECK_ARG(x && y && n >= 0 && channels > 0 && hw > 0, "bad argument");
    if (n == 0) return STOVE_OK;
    return bw_launch<float>(x, nullptr, y, nullptr, n, channels, hw, (cudaStream_t)stream);
}

extern "C" int stove_bw_transform_ex(const void* x, int x_is_u8, int x_is_cell, float* y, float* y_planes, int64_t n,
                                     int channels, int64_t hw, void* stream) {
    STOVE_CHECK_ARG(x && y && n >= 0 && channels > 0 && hw > 0, "bad argument");
    if (n == 0) return STOVE_OK;
    const void* const* cell = x_is_cell ? (const void* const*)x : nullptr;
    return x_is_u8 ? bw_launch<uint8_t>(x_is_cell ? nullptr : (const uint8_t*)x, cell, y, y_planes, n, channels, hw, (cudaStream_t)stream)
                   : bw_launch<float>(x_is_cell ? nullptr : (const float*)x, cell, y, y_planes, n, channels, hw, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
// Gradient bucket: gather many device tensors into one flat fp32 buffer in one or two launches (the
// library concatenation of 110 tensors takes ~32 us for 5.6 MB).  Addresses travel as kernel
// parameters, so the call is capturable in a CUDA graph without any table upload.
// ---------------------------------------------------------------------------------------
constexpr int GATHER_MAX = 128, GATHER_CHUNK = 4096;
struct GatherArgs {
    const float* src[GATHER_MAX];
    int64_t off[GATHER_MAX];
    int64_t num[GATHER_MAX];
};

__global__ void gather_flat_kernel(const __grid_constant__ GatherArgs a, float* __restrict__ dst, float scale) {
    const int t = blockIdx.y;
    const int64_t num = a.num[t];
    const int64_t lo = (int64_t)blockIdx.x * GATHER_CHUNK;
    if (lo >= num) return;
    const int64_t hi = lo + GATHER_CHUNK < num ? lo + GATHER_CHUNK : num;
    const float* __restrict__ s = a.src[t];
    float* __restrict__ d = dst + a.off[t];
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int64_t v_hi = lo + ((hi - lo) & ~(int64_t)3);
        for (int64_t i = lo + 4 * threadIdx.x; i < v_hi; i += 4 * blockDim.x) {
            float4 v = __ldg(reinterpret_cast<const float4*>(s + i));
            v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;      // scale = 1 is exact
            *reinterpret_cast<float4*>(d + i) = v;
        }
        for (int64_t i = v_hi + threadIdx.x; i < hi; i += blockDim.x) d[i] = __ldg(s + i) * scale;
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) d[i] = __ldg(s + i) * scale;
    }
}

extern "C" int stove_gather_flat(const void* const* srcs, const int64_t* offsets, const int64_t* numels, int count,
                                 float* dst, float scale, void* stream) {
    STOVE_CHECK_ARG(count >= 0 && (count == 0 || (srcs && offsets && numels && dst)), "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    // two classes so that a few large tensors do not force thousands of empty CTAs on the small ones
    for (int pass = 0; pass < 2; ++pass) {
        GatherArgs a;
        int k = 0;
        int64_t mx = 0;
        auto flush = [&]() -> int {
            if (k == 0) return STOVE_OK;
            const dim3 grid((unsigned)((mx + GATHER_CHUNK - 1) / GATHER_CHUNK), (unsigned)k);
            STOVE_KERNEL(K_GATHER_FLAT, st, gather_flat_kernel<<<grid, 256, 0, st>>>(a, dst, scale));
            STOVE_LAUNCH_CHECK();
            k = 0;
            mx = 0;
            return STOVE_OK;
        };
        for (int i = 0; i < count; ++i) {
            const bool big = numels[i] > 4 * GATHER_CHUNK;
            if (numels[i] <= 0 || big != (pass == 1)) continue;
            STOVE_CHECK_ARG(srcs[i] != nullptr && offsets[i] >= 0, "null source or negative offset");
            a.src[k] = (const float*)srcs[i];
            a.off[k] = offsets[i];
            a.num[k] = numels[i];
            if (numels[i] > mx) mx = numels[i];
            if (++k == GATHER_MAX) {
                const int rc = flush();
                if (rc != STOVE_OK) return rc;
            }
        }
        const int rc = flush();
        if (rc != STOVE_OK) return rc;
    }
    return STOVE_OK;
}
