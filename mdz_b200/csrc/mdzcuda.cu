// mdzcuda.cu -- host side of libmdzcuda's plain C ABI (include/mdzcuda.h):
// the O(W+H) coordinate prologue, device buffers, kernel dispatch by limb
// count, progress / cancel plumbing and the IMAD peak microbenchmark.
//
// There is deliberately no CPU implementation of the escape-time loop in this
// library: if CUDA is unavailable every entry point here fails with an error text.
// (The rth_* layer, rth.cpp, then runs the line callback the HOST installed -- the
// reference's own interface, src/render_threads.c:375-377 -- and counts those lines in
// mdzcuda_fallback_lines().)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <float.h>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <map>

#include <atomic>
#include <algorithm>

#include "../../include/mdzcuda.h"
#include "escape_params.cuh"
#include "mp_convert.h"
#include "band_grants.h"
#include "mdz_run.h"

using namespace mdz;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static void set_err(const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
}
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    set_err("%s failed: %s", #call, cudaGetErrorString(e_)); return 0; } } while (0)
#define CUDA_OKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    set_err("%s failed: %s", #call, cudaGetErrorString(e_)); goto fail; } } while (0)

extern "C" const char* mdzcuda_last_error(void) { return g_err.c_str(); }

extern "C" int mdzcuda_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_err("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return 0; }
    return n;
}

// ---------------------------------------------------------------------------
// Device memory / stream / event pool.
//
// A render is 60 ms of kernel; cudaMalloc + cudaFree of its buffers (and the
// implicit device synchronisation of cudaFree) cost between 1 and tens of
// milliseconds per call and vary from box to box.  MDZ re-renders constantly (the
// Julia preview on every mouse move, main_gui.c:786-793), so freed blocks, side
// streams, events and pinned staging words are kept per device and handed to the
// next plan.  mdzcuda_trim() gives everything back to the driver.
// ---------------------------------------------------------------------------
namespace {
struct DevPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;       // capacity -> block
    std::map<void*, size_t> live;                   // blocks handed out -> capacity
    size_t cached = 0;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> events;
    std::vector<unsigned int*> pinned;              // 16-word pinned staging chunks
    std::multimap<size_t, void*> pinned_bufs;       // larger pinned staging buffers (table uploads), by capacity
};
constexpr int kMaxDev = 64;
DevPool g_pool[kMaxDev];

size_t pool_limit_bytes()
{
    static size_t lim = [] {
        const char* e = getenv("MDZCUDA_POOL_MB");
        return (size_t)(e && *e ? strtoull(e, nullptr, 10) : 4096) << 20;
    }();
    return lim;
}

size_t pool_round(size_t bytes)
{
    const size_t g = bytes <= (1u << 20) ? (64u << 10) : (2u << 20);
    return (bytes + g - 1) / g * g;
}

// the current device must be `dev`
cudaError_t pool_alloc(int dev, void** out, size_t bytes)
{
    DevPool& P = g_pool[dev];
    const size_t want = pool_round(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.free_blocks.lower_bound(want);
        if (it != P.free_blocks.end() && it->first <= 2 * want) {
            *out = it->second;
            P.live[it->second] = it->first;
            P.cached -= it->first;
            P.free_blocks.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e != cudaSuccess) {                           // give the cache back and retry once
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lock(P.mu);
            for (auto& kv : P.free_blocks) drop.push_back(kv.second);
            P.free_blocks.clear(); P.cached = 0;
        }
        for (void* q : drop) cudaFree(q);
        (void)cudaGetLastError();
        e = cudaMalloc(out, want);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lock(P.mu);
    P.live[*out] = want;
    return cudaSuccess;
}

// the caller guarantees no work that touches the block is still in flight
void pool_free(int dev, void* ptr)
{
    if (!ptr) return;
    DevPool& P = g_pool[dev];
    size_t cap = 0;
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.live.find(ptr);
        if (it == P.live.end()) { drop.push_back(ptr); }
        else {
            cap = it->second;
            P.live.erase(it);
            if (cap > pool_limit_bytes()) drop.push_back(ptr);
            else {
                P.free_blocks.insert(std::make_pair(cap, ptr));
                P.cached += cap;
                while (P.cached > pool_limit_bytes() && !P.free_blocks.empty()) {
                    auto big = std::prev(P.free_blocks.end());
                    P.cached -= big->first;
                    drop.push_back(big->second);
                    P.free_blocks.erase(big);
                }
            }
        }
    }
    for (void* q : drop) cudaFree(q);
}

cudaError_t pool_stream(int dev, cudaStream_t* out)
{
    DevPool& P = g_pool[dev];
    {
        std::lock_guard<std::mutex> lock(P.mu);
        if (!P.streams.empty()) { *out = P.streams.back(); P.streams.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
cudaError_t pool_event(int dev, cudaEvent_t* out)
{
    DevPool& P = g_pool[dev];
    {
        std::lock_guard<std::mutex> lock(P.mu);
        if (!P.events.empty()) { *out = P.events.back(); P.events.pop_back(); return cudaSuccess; }
    }
    return cudaEventCreateWithFlags(out, cudaEventDisableTiming);
}
cudaError_t pool_pinned(int dev, unsigned int** out)
{
    DevPool& P = g_pool[dev];
    {
        std::lock_guard<std::mutex> lock(P.mu);
        if (!P.pinned.empty()) { *out = P.pinned.back(); P.pinned.pop_back(); memset(*out, 0, 64); return cudaSuccess; }
    }
    cudaError_t e = cudaMallocHost((void**)out, 64);
    if (e == cudaSuccess) memset(*out, 0, 64);
    return e;
}
// pinned staging buffer of at least `bytes` (the H2D image of a plan's tables)
cudaError_t pool_pinned_buf(int dev, void** out, size_t* cap, size_t bytes)
{
    DevPool& P = g_pool[dev];
    const size_t want = (bytes + 65535) / 65536 * 65536;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.pinned_bufs.lower_bound(want);
        if (it != P.pinned_bufs.end()) { *out = it->second; *cap = it->first; P.pinned_bufs.erase(it); return cudaSuccess; }
    }
    *cap = want;
    return cudaMallocHost(out, want);
}
void pool_pinned_buf_free(int dev, void* ptr, size_t cap)
{
    if (!ptr) return;
    DevPool& P = g_pool[dev];
    std::lock_guard<std::mutex> lock(P.mu);
    P.pinned_bufs.insert(std::make_pair(cap, ptr));
}
}  // namespace

extern "C" void mdzcuda_trim(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return;
    for (int dev = 0; dev < n && dev < kMaxDev; ++dev) {
        DevPool& P = g_pool[dev];
        std::vector<void*> drop, hostbufs; std::vector<cudaStream_t> ss; std::vector<cudaEvent_t> es; std::vector<unsigned int*> ps;
        {
            std::lock_guard<std::mutex> lock(P.mu);
            for (auto& kv : P.free_blocks) drop.push_back(kv.second);
            P.free_blocks.clear(); P.cached = 0;
            ss.swap(P.streams); es.swap(P.events); ps.swap(P.pinned);
            for (auto& kv : P.pinned_bufs) hostbufs.push_back(kv.second);
            P.pinned_bufs.clear();
        }
        if (drop.empty() && ss.empty() && es.empty() && ps.empty() && hostbufs.empty()) continue;
        if (cudaSetDevice(dev) != cudaSuccess) continue;
        for (void* q : drop) cudaFree(q);
        for (auto s : ss) cudaStreamDestroy(s);
        for (auto e : es) cudaEventDestroy(e);
        for (auto q : ps) cudaFreeHost(q);
        for (auto q : hostbufs) cudaFreeHost(q);
    }
}

// ---------------------------------------------------------------------------
// host tables
// ---------------------------------------------------------------------------
struct HostTable {
    int n32 = 0, count = 0;
    std::vector<uint32_t> m;    // limb-major [n32][count]
    std::vector<int32_t>  e;
    std::vector<uint32_t> s;
    void init(int n32_, int count_)
    {
        n32 = n32_; count = count_;
        m.assign((size_t)n32 * count, 0u); e.assign(count, E_ZERO); s.assign(count, 0u);
    }
    void set_zero_entry(int i) { for (int k = 0; k < n32; ++k) m[(size_t)k * count + i] = 0; e[i] = E_ZERO; s[i] = 0; }
    // from an mpfr value already rounded to `prec` bits
    void set_mpfr(int i, const __mpfr_struct* v, long prec)
    {
        if (v->_mpfr_exp == MDZ_MPFR_EXP_ZERO) { set_zero_entry(i); return; }
        // the ceil(prec/32) significant limbs go to the top of the entry; a wider entry (the warp-per-pixel
        // kernels pad to 32 K limbs) is zero below
        const int n = limbs32_for_prec(prec);
        std::vector<uint32_t> tmp((size_t)n);
        sig64_to_sig32((const uint64_t*)v->_mpfr_d, prec, tmp.data(), n);
        for (int k = 0; k < n32 - n; ++k) m[(size_t)k * count + i] = 0u;
        for (int k = 0; k < n; ++k) m[(size_t)(n32 - n + k) * count + i] = tmp[(size_t)k];
        long ex = v->_mpfr_exp;
        if (ex < E_MIN) { set_zero_entry(i); return; }
        if (ex > (1 << 28)) ex = (1 << 28);
        e[i] = (int32_t)ex;
        s[i] = v->_mpfr_sign < 0 ? 1u : 0u;
    }
    // from an mpf_t: |size| limbs top-aligned into nl64 = n32/2 limbs, zero padded below
    void set_mpf(int i, const __mpf_struct* v)
    {
        const int nl = n32 / 2;
        int size = v->_mp_size < 0 ? -v->_mp_size : v->_mp_size;
        set_zero_entry(i);
        e[i] = 0;
        if (size == 0) return;
        const mp_limb_t* d = v->_mp_d;
        int skip = 0;
        if (size > nl) { skip = size - nl; size = nl; }     // cannot happen for values of this precision
        for (int k = 0; k < size; ++k) {
            const uint64_t w = d[skip + k];
            const int at = nl - size + k;
            m[(size_t)(2 * at) * count + i] = (uint32_t)w;
            m[(size_t)(2 * at + 1) * count + i] = (uint32_t)(w >> 32);
        }
        e[i] = (int32_t)v->_mp_exp;
        s[i] = v->_mp_size < 0 ? 1u : 0u;
    }
    // from an x87 long double: 64-bit significand, same value as MPFR p=64
    void set_ld(int i, long double v)
    {
        if (v == 0.0L || v != v) { set_zero_entry(i); return; }
        int ex;
        long double f = frexpl(fabsl(v), &ex);          // f in [0.5,1)
        if (ex < -16000 || ex > 16000) {                // inf / beyond anything a view can hold
            set_zero_entry(i); if (ex > 16000) { e[i] = 1 << 20; m[(size_t)(n32 - 1) * count + i] = 0x80000000u; }
            return;
        }
        long double sc = ldexpl(f, 64);                  // exact: integer < 2^64
        uint64_t mant = (uint64_t)sc;
        m[(size_t)0 * count + i] = (uint32_t)mant;
        m[(size_t)1 * count + i] = (uint32_t)(mant >> 32);
        e[i] = ex;
        s[i] = v < 0 ? 1u : 0u;
    }
};

struct DevTable {
    uint32_t* m = nullptr; int32_t* e = nullptr; uint32_t* s = nullptr; int count = 0;
    CoordTable view() const { CoordTable t; t.m = m; t.e = e; t.s = s; t.count = count; return t; }
};

// All small device state of a plan lives in ONE allocation (cudaMalloc / cudaFree cost
// ~0.1-0.5 ms each and they used to be a dozen per render): the three coordinate tables,
// the control words, the per-band counters and flags.  One H2D copy fills the tables.
struct Arena {
    std::vector<uint32_t> host;     // image of the table part
    size_t words = 0;
    size_t reserve(size_t n) { const size_t at = words; words += (n + 3) & ~(size_t)3; return at; }
};

static size_t arena_put(Arena& a, const HostTable& h, size_t& om, size_t& oe, size_t& os)
{
    om = a.reserve(h.m.size()); oe = a.reserve(h.e.size()); os = a.reserve(h.s.size());
    a.host.resize(a.words, 0u);
    if (!h.m.empty()) memcpy(&a.host[om], h.m.data(), h.m.size() * 4);
    if (!h.e.empty()) memcpy(&a.host[oe], h.e.data(), h.e.size() * 4);
    if (!h.s.empty()) memcpy(&a.host[os], h.s.data(), h.s.size() * 4);
    return a.words;
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
constexpr int kMaxTilesPerBand = 16;     // a band of a centre-out static plan is cut into at most this many tiles

struct mdzcuda_plan {
    mdzcuda_view view;          // scalars only are used after create
    int device = 0;
    int n32 = 0;                // limbs per table entry (for the warp-per-pixel kernels: padded to 32 * coop_k)
    int coop_k = 0;             // 0: one thread per pixel; else K * 100 + T: T lanes per pixel, K limbs per lane (coop_kernel.cuh)
    int band_first = 0, band_stride = 1, nbands = 0, local_lines = 0;
    std::vector<int> line_map;  // local line -> global line
    DevTable xs, ys, jc;
    RoundCfg rc;
    int32_t* d_raw = nullptr;
    uint32_t* d_arena = nullptr;        // tables + everything below
    unsigned int* d_ctrl = nullptr;     // [0] queue, [1] bands_done, [2] parked pixels, [3] phase 1's queue
    unsigned int* d_band_count = nullptr;
    unsigned int* d_band_flag = nullptr;    // [nbands] generation of the launch that completed the band
    unsigned int gen = 0;                   // generation of the current launch (1, 2, ...; never 0)
    std::vector<unsigned int> h_flags;      // host mirror for mdzcuda_plan_poll_bands
    size_t reset_words = 0;             // ctrl + band_count: reset by every launch, in stream order
    size_t zero_words = 0;              // ctrl .. cancel word: zeroed once at create
    unsigned int* d_cancel = nullptr;   // holds the generation of the launch that is to stop (never reset)
    cudaStream_t side = nullptr;        // progress / cancel traffic while the kernel runs
    cudaEvent_t done_ev = nullptr;
    int chunk = 0, blocks_per_sm = 0, spec = 1;
    int cycle = 0;                      // exact periodicity check (mdzcuda_plan_set_cycle_detection)
    uint32_t* d_cycle = nullptr;        // its saved-state columns, allocated at the first launch that needs them
    size_t cycle_words = 0;
    int kernels_launched = 0;           // kernels this plan has put on a stream so far (mdzcuda_plan_kernels_launched)
    int park = -1;                      // tail compaction (escape_kernel.cuh "Parking"): -1 automatic, 0 off, 1 on
    uint32_t* d_park = nullptr;         // parked pixel states + reading order + claim words, allocated at the first launch that parks
    size_t park_words = 0;
    bool gmp = false;
    // fed plan (band scheduler): queue slots filled by the host while the kernel runs
    bool fed = false;
    int order_mode = 0;                 // MDZCUDA_ORDER_*: the sequence in which the queue visits the plan's bands / tiles
    bool order_uploaded = false;
    int tiles_per_band = 1;             // static plans in centre-out order cut their bands into this many tiles
    unsigned int* d_order = nullptr;    // [nbands] slot -> band
    unsigned int* d_feed = nullptr;     // [0] slots filled, [1] generation of the launch that is closed
    unsigned int* h_stage = nullptr;    // pinned staging: [nbands] order entries + [4] control words
    size_t h_stage_cap = 0;
    int granted = 0;                    // bands fed so far in the current launch
    int granted_slots = 0;              // ... as queue slots (bands x tiles_per_band)
    bool closed = false;
    unsigned int claimed = 0;           // queue counter as of the last poll
    cudaStream_t own = nullptr;         // the plan's own non-blocking launch stream (mdzcuda_plan_stream)
    mdzcuda_kernel_info info;
    unsigned int* h_pinned = nullptr;   // 4 words of pinned staging
    uint32_t* d_palette = nullptr;      // 256 entries
    uint32_t* d_rgb = nullptr;          // [nbands][user_width]
    ColourParams colour;                // enabled = 0 until mdzcuda_plan_set_colour
};

typedef void (*kernel_fn)(const EscapeParams);

// The kernels are instantiated in separate translation units (kernels_*.cu) so that the
// build parallelises: one unrolled kernel per limb count is seconds to minutes of ptxas.
kernel_fn mdz_kernel_mpfr(int n32, int cyc);    // N = 2..32 words, without / with the periodicity check (kernels_mpfr_*.cu)
int       mdz_smem_words_mpfr(int n32);
void      mdz_coop_shape(int n32, int* k, int* t);     // T lanes x K limbs per value for 33 .. 256 limbs  (kernels_coop.cu)
kernel_fn mdz_kernel_coop(int k, int t);
kernel_fn mdz_kernel_coop_gmp(int k, int t);                // GMP mpf above 512 bits  (kernels_coop_gmp.cu)
int       mdz_smem_words_coop(int k, int t);
kernel_fn mdz_kernel_gmp_clear(int nl);         // NL = 3..10 limbs  (kernels_gmp.cu)
kernel_fn mdz_kernel_gmp_fast(int nl);          // NL = 4..10 limbs  (kernels_gmpf_*.cu)
int       mdz_smem_words_gmp_fast(int nl);

// GMP mode: nl = P+1 64-bit limbs, P = mpf_init2's precision in limbs.  The fast kernel
// (mpf_fast.cuh) covers P >= 3; P = 2 (precision below 65 bits, which MDZ's settings
// cannot select: image_info.c:535 keeps precision >= 80) uses the clear version, as does
// everything when MDZCUDA_GMP_CLEAR is set (A/B measurements).
static kernel_fn gmp_kernel_for_limbs(int nl)
{
    if (!getenv("MDZCUDA_GMP_CLEAR")) { kernel_fn f = mdz_kernel_gmp_fast(nl); if (f) return f; }
    return mdz_kernel_gmp_clear(nl);
}
static int gmp_smem_words(int nl)
{
    if (getenv("MDZCUDA_GMP_CLEAR")) return 0;
    return mdz_smem_words_gmp_fast(nl);
}
static kernel_fn kernel_for_limbs(int n, int cyc = 0) { return mdz_kernel_mpfr(n, cyc); }
static int smem_words_for_limbs(int n) { return mdz_smem_words_mpfr(n); }

// the kernel that renders this mode / precision (nullptr + error text: none is instantiated)
constexpr long kMaxMpfrBits = 8192;
constexpr long kMaxGmpBits = 8000;      // P = (p + 127) / 64 limbs of precision, P + 2 limbs of 128 in the widest lane-group shape
static kernel_fn kernel_for_view(const mdzcuda_view* v, int* n32_out, int* coop_k_out = nullptr)
{
    int n32;
    if (coop_k_out) *coop_k_out = 0;
    if (v->mode == MDZCUDA_MODE_LD) n32 = 2;
    else if (v->mode == MDZCUDA_MODE_MPFR) {
        if (v->precision < 33) { set_err("MPFR precision below 33 bits is not supported"); return nullptr; }
        if (v->precision > kMaxMpfrBits) { set_err("MPFR precision %ld: no GPU kernel above %d bits", v->precision, kMaxMpfrBits); return nullptr; }
        n32 = limbs32_for_prec(v->precision);
        if (n32 > 32) {
            // a group of T lanes per pixel, K limbs per lane (coop_ops.cuh), the entry padded to T K limbs
            int k = 0, t = 0;
            mdz_coop_shape(n32, &k, &t);
            kernel_fn cf = mdz_kernel_coop(k, t);
            if (!cf) { set_err("MPFR precision %ld: no lane-group kernel for %d x %d limbs", v->precision, t, k); return nullptr; }
            *n32_out = t * k;
            if (coop_k_out) *coop_k_out = k * 100 + t;
            return cf;
        }
    } else if (v->mode == MDZCUDA_MODE_GMP) {
        // mpf_init2(p): precision in limbs P = (max(53,p)+127)/64, storage P+1 limbs
        const long pb = v->precision < 53 ? 53 : v->precision;
        if (pb > kMaxGmpBits) { set_err("GMP mpf precision %ld: no GPU kernel above %ld bits", v->precision, kMaxGmpBits); return nullptr; }
        const int nl = (int)((pb + 127) / 64 + 1);
        n32 = 2 * nl;
        if (!gmp_kernel_for_limbs(nl)) {
            // a group of T lanes per pixel (coop_mpf.cuh): the NL limbs top aligned in T K words with one limb to spare
            int k = 0, t = 0;
            mdz_coop_shape(2 * (nl + 1), &k, &t);
            kernel_fn cf = (2 * (nl + 1) <= t * k) ? mdz_kernel_coop_gmp(k, t) : nullptr;
            if (!cf) { set_err("GMP mpf precision %ld: no lane-group kernel for %d limbs", v->precision, nl); return nullptr; }
            *n32_out = t * k;
            if (coop_k_out) *coop_k_out = k * 100 + t;
            return cf;
        }
    }
    else { set_err("unknown mode %d", v->mode); return nullptr; }
    kernel_fn fn = v->mode == MDZCUDA_MODE_GMP ? gmp_kernel_for_limbs(n32 / 2) : kernel_for_limbs(n32);
    if (!fn) { set_err("%s precision %ld needs %d limbs: no GPU kernel instantiated", v->mode == MDZCUDA_MODE_GMP ? "GMP mpf" : "MPFR", v->precision, n32); return nullptr; }
    *n32_out = n32;
    return fn;
}

extern "C" int mdzcuda_view_supported(const mdzcuda_view* v)
{
    g_err.clear();
    if (!v) { set_err("null view"); return 0; }
    int n32 = 0;
    return kernel_for_view(v, &n32) != nullptr;
}

// ---- prologue: MPFR mode (reference src/fractal.c:143-188) ------------------
static int prologue_mpfr(const mdzcuda_view* v, const std::vector<int>& lines,
                         HostTable& xs, HostTable& ys, HostTable& jc, int n32)
{
    const long p = v->precision;
    mpfr_t img_rw, img_xmin, width, t1, x, y;
    mpfr_init2(img_rw, p); mpfr_init2(img_xmin, p); mpfr_init2(width, p);
    mpfr_init2(t1, p); mpfr_init2(x, p); mpfr_init2(y, p);
    mpfr_set_si(img_rw, v->real_width, MPFR_RNDN);          // fractal.c:160
    mpfr_set(img_xmin, v->xmin, MPFR_RNDN);                 // :161
    mpfr_set(width, v->width, MPFR_RNDN);                   // :162

    xs.init(n32, v->real_width);
    for (int ix = 0; ix < v->real_width; ++ix) {
        mpfr_si_div(t1, ix, img_rw, MPFR_RNDN);             // :183
        mpfr_mul(x, t1, width, MPFR_RNDN);                  // :185
        mpfr_add(x, x, img_xmin, MPFR_RNDN);                // :186
        xs.set_mpfr(ix, x, p);
    }
    ys.init(n32, (int)lines.size());
    for (size_t i = 0; i < lines.size(); ++i) {
        mpfr_div(t1, width, img_rw, MPFR_RNDN);             // :167
        mpfr_mul_si(t1, t1, lines[i], MPFR_RNDN);           // :169
        mpfr_sub(y, v->ymax, t1, MPFR_RNDN);                // :170 (ymax at its own precision)
        ys.set_mpfr((int)i, y, p);
    }
    jc.init(n32, 2);
    if (v->family == MDZCUDA_FAMILY_JULIA) {
        if (!v->julia_re || !v->julia_im) { set_err("julia family needs julia_re/julia_im"); return 0; }
        mpfr_set(x, v->julia_re, MPFR_RNDN);                // :197
        mpfr_set(y, v->julia_im, MPFR_RNDN);                // :198
        jc.set_mpfr(0, x, p); jc.set_mpfr(1, y, p);
    }
    mpfr_clear(img_rw); mpfr_clear(img_xmin); mpfr_clear(width);
    mpfr_clear(t1); mpfr_clear(x); mpfr_clear(y);
    return 1;
}

// ---- prologue: GMP mpf mode (reference src/fractal.c:283-328, :341-342) -------
static int prologue_gmp(const mdzcuda_view* v, const std::vector<int>& lines,
                        HostTable& xs, HostTable& ys, HostTable& jc, int n32)
{
    if (!v->gxmin || !v->gymax || !v->gwidth) { set_err("GMP mode needs gxmin/gymax/gwidth"); return 0; }
    const unsigned long p = (unsigned long)v->precision;
    mpf_t img_rw, img_xmin, width, t1, x, y;
    mpf_init2(img_rw, p); mpf_init2(img_xmin, p); mpf_init2(width, p);
    mpf_init2(t1, p); mpf_init2(x, p); mpf_init2(y, p);
    mpf_set_si(img_rw, v->real_width);                      // fractal.c:300
    mpf_set(img_xmin, v->gxmin);                            // :301
    mpf_set(width, v->gwidth);                              // :302
    xs.init(n32, v->real_width);
    for (int ix = 0; ix < v->real_width; ++ix) {
        mpf_ui_div(t1, (unsigned long)ix, img_rw);          // :323
        mpf_mul(x, t1, width);                              // :325
        mpf_add(x, x, img_xmin);                            // :326
        xs.set_mpf(ix, x);
    }
    ys.init(n32, (int)lines.size());
    for (size_t i = 0; i < lines.size(); ++i) {
        mpf_div(t1, width, img_rw);                         // :307
        mpf_mul_ui(t1, t1, (unsigned long)lines[i]);        // :309
        mpf_sub(y, v->gymax, t1);                           // :310
        ys.set_mpf((int)i, y);
    }
    jc.init(n32, 2);
    if (v->family == MDZCUDA_FAMILY_JULIA) {
        if (!(v->gjulia_re && v->gjulia_im) && (!v->julia_re || !v->julia_im)) { set_err("julia family needs julia_re/julia_im"); return 0; }
        // mpfr_to_gmp (coords.c:13-18): through the decimal text my_mpfr_to_str prints,
        // i.e. mpfr_snprintf "%.Re" (my_mpfr_to_str.c:68) -- one significant digit with
        // MPFR >= 4 (SURVEY finding 3); reproduced literally, as the drop-in must
        // A host that keeps the constant as mpf (or wants another conversion) passes it directly.
        if (v->gjulia_re && v->gjulia_im) { mpf_set(x, v->gjulia_re); mpf_set(y, v->gjulia_im); }
        else {
            // MDZCUDA_RE_FORMAT=full: the host was built with the format fixed to "%Re" (the oracle's mdz_fixre
            // binaries), whose own my_mpfr_to_str the rth_* layer cannot reach
            const char* e = getenv("MDZCUDA_RE_FORMAT");
            const bool full = e && !strcmp(e, "full");
            char buf[4097];
            if (full) mpfr_snprintf(buf, 4096, "%Re", v->julia_re); else mpfr_snprintf(buf, 4096, "%.Re", v->julia_re);
            mpf_set_str(x, buf, 10);
            if (full) mpfr_snprintf(buf, 4096, "%Re", v->julia_im); else mpfr_snprintf(buf, 4096, "%.Re", v->julia_im);
            mpf_set_str(y, buf, 10);
        }
        jc.set_mpf(0, x); jc.set_mpf(1, y);
    }
    mpf_clear(img_rw); mpf_clear(img_xmin); mpf_clear(width);
    mpf_clear(t1); mpf_clear(x); mpf_clear(y);
    return 1;
}

// ---- prologue: long double mode (reference src/fractal.c:50-73) -------------
static int prologue_ld(const mdzcuda_view* v, const std::vector<int>& lines,
                       HostTable& xs, HostTable& ys, HostTable& jc)
{
#if LDBL_MANT_DIG != 64
#error "MODE_LD reproduces x87 extended precision; this host's long double is not x87"
#endif
    const int img_width = v->real_width;
    volatile long double xmin = mpfr_get_ld(v->xmin, MPFR_RNDN);    // :50
    volatile long double xmax = mpfr_get_ld(v->xmax, MPFR_RNDN);    // :51
    volatile long double ymax = mpfr_get_ld(v->ymax, MPFR_RNDN);    // :52
    volatile long double width = xmax - xmin;                        // :53
    xs.init(2, img_width);
    for (int ix = 0; ix < img_width; ++ix) {
        volatile long double q = ix / (long double)img_width;
        volatile long double x = q * width;
        x = x + xmin;                                                // :72
        xs.set_ld(ix, x);
    }
    ys.init(2, (int)lines.size());
    for (size_t i = 0; i < lines.size(); ++i) {
        volatile long double q = width / (long double)img_width;
        volatile long double t = q * (long double)lines[i];
        volatile long double y = ymax - t;                           // :61-62
        ys.set_ld((int)i, y);
    }
    jc.init(2, 2);
    if (v->family == MDZCUDA_FAMILY_JULIA) {
        if (!v->julia_re || !v->julia_im) { set_err("julia family needs julia_re/julia_im"); return 0; }
        jc.set_ld(0, mpfr_get_ld(v->julia_re, MPFR_RNDN));           // :57
        jc.set_ld(1, mpfr_get_ld(v->julia_im, MPFR_RNDN));           // :58
    }
    return 1;
}

// kernel facts per (device, kernel) are cached: cudaGetDeviceProperties and the
// occupancy query cost milliseconds, which is a visible share of a 60 ms render
static int kernel_facts(int device, kernel_fn fn, int smem, int n32, mdzcuda_kernel_info* out)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, mdzcuda_kernel_info> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(device, (const void*)fn);
    auto it = cache.find(key);
    if (it == cache.end()) {
        cudaFuncAttributes fa;
        CUDA_OK(cudaFuncGetAttributes(&fa, (const void*)fn));
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        if (smem > 48 * 1024)
            CUDA_OK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0;
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, kBlock, smem));
        if (occ < 1) { set_err("kernel for %d limbs does not fit on an SM", n32); return 0; }
        mdzcuda_kernel_info ki;
        ki.limbs = n32;
        ki.lanes_per_pixel = 1;
        ki.regs_per_thread = fa.numRegs;
        ki.local_bytes = (int)fa.localSizeBytes;
        ki.shared_bytes = smem;
        ki.block_threads = kBlock;
        ki.blocks_per_sm = occ;
        ki.sm_count = sms;
        ki.grid_blocks = occ * sms;
        it = cache.insert(std::make_pair(key, ki)).first;
    }
    *out = it->second;
    return 1;
}

extern "C" mdzcuda_plan* mdzcuda_plan_create(const mdzcuda_view* v, int device,
                                             int band_first, int band_stride)
{
    g_err.clear();
    if (!v) { set_err("null view"); return nullptr; }
    if (v->real_width < 1 || v->real_height < 1 || v->aa_factor < 1 ||
        v->real_height % v->aa_factor != 0) { set_err("bad image size / aa factor"); return nullptr; }
    if (v->depth < 1 || v->depth > 2147483647L) { set_err("depth out of range"); return nullptr; }
    if ((long long)v->real_width * v->real_height >= (1LL << 31)) { set_err("image too large (2^31 supersamples or more)"); return nullptr; }
    if (band_stride < 1 || band_first < 0) { set_err("bad band partition"); return nullptr; }
    if (!v->xmin || !v->ymax) { set_err("view rect missing"); return nullptr; }
    int ndev = mdzcuda_device_count();
    if (ndev <= 0) { if (g_err.empty()) set_err("no CUDA device"); return nullptr; }
    if (device < 0 || device >= ndev) { set_err("device %d out of range (%d visible)", device, ndev); return nullptr; }
    if (device >= kMaxDev) { set_err("device %d: this build pools at most %d devices", device, kMaxDev); return nullptr; }

    int n32 = 0, coop_k = 0;
    kernel_fn fn = kernel_for_view(v, &n32, &coop_k);
    if (!fn) return nullptr;
    const bool gmp = v->mode == MDZCUDA_MODE_GMP;

    mdzcuda_plan* pl = new mdzcuda_plan();
    pl->view = *v;
    pl->device = device;
    pl->n32 = n32;
    pl->coop_k = coop_k;
    pl->band_first = band_first; pl->band_stride = band_stride;
    memset(&pl->colour, 0, sizeof pl->colour);
    const int total_bands = v->real_height / v->aa_factor;
    for (int b = band_first; b < total_bands; b += band_stride)
        for (int k = 0; k < v->aa_factor; ++k) pl->line_map.push_back(b * v->aa_factor + k);
    pl->local_lines = (int)pl->line_map.size();
    pl->nbands = pl->local_lines / v->aa_factor;
    pl->gmp = gmp;
    { const char* e = getenv("MDZCUDA_CYCLE_DETECT"); pl->cycle = (e && *e && *e != '0') ? 1 : 0; }
    pl->rc = make_round_cfg(n32, v->mode == MDZCUDA_MODE_LD ? 64 : (gmp ? 32 * n32 : (int)v->precision));

    std::vector<uint32_t> table_image; size_t table_bytes = 0;
    HostTable xs, ys, jc;
    int ok = (v->mode == MDZCUDA_MODE_LD) ? prologue_ld(v, pl->line_map, xs, ys, jc)
           : gmp ? prologue_gmp(v, pl->line_map, xs, ys, jc, n32)
                 : prologue_mpfr(v, pl->line_map, xs, ys, jc, n32);
    if (!ok) { delete pl; return nullptr; }

    {
        CUDA_OKP(cudaSetDevice(device));
        {
            Arena ar;
            size_t o[9];
            arena_put(ar, xs, o[0], o[1], o[2]);
            arena_put(ar, ys, o[3], o[4], o[5]);
            const size_t table_words = arena_put(ar, jc, o[6], o[7], o[8]);
            const size_t o_ctrl = ar.reserve(8);
            const size_t o_count = ar.reserve((size_t)pl->nbands + 1);
            const size_t o_flag = ar.reserve((size_t)pl->nbands + 1);
            const size_t o_cancel = ar.reserve(4);
            const size_t o_feed = ar.reserve(4);
            const size_t o_order = ar.reserve((size_t)pl->nbands * kMaxTilesPerBand + 1);
            CUDA_OKP(pool_alloc(device, (void**)&pl->d_arena, ar.words * sizeof(uint32_t)));
            table_image.swap(ar.host); table_bytes = table_words * sizeof(uint32_t);
            uint32_t* b = pl->d_arena;
            pl->xs.m = b + o[0]; pl->xs.e = (int32_t*)(b + o[1]); pl->xs.s = b + o[2]; pl->xs.count = xs.count;
            pl->ys.m = b + o[3]; pl->ys.e = (int32_t*)(b + o[4]); pl->ys.s = b + o[5]; pl->ys.count = ys.count;
            pl->jc.m = b + o[6]; pl->jc.e = (int32_t*)(b + o[7]); pl->jc.s = b + o[8]; pl->jc.count = jc.count;
            pl->d_ctrl = b + o_ctrl;
            pl->d_band_count = b + o_count;
            pl->d_band_flag = b + o_flag;
            pl->d_cancel = b + o_cancel;
            pl->d_feed = b + o_feed;
            pl->d_order = b + o_order;
            pl->zero_words = ar.words - o_ctrl;     // everything from the control words on, once, at create
            pl->reset_words = o_flag - o_ctrl;      // per launch: queue counter and band counters only
        }
        size_t npx = (size_t)pl->local_lines * v->real_width;
        CUDA_OKP(pool_alloc(device, (void**)&pl->d_raw, (npx ? npx : 1) * sizeof(int32_t)));
        CUDA_OKP(pool_stream(device, &pl->side));
        CUDA_OKP(pool_stream(device, &pl->own));
        // Control words, band counters and flags start at zero.  On the side stream, and waited
        // for: the polls run on that stream, which is not ordered against the caller's, and a
        // recycled arena still holds the previous plan's flags (a cudaMemset on the default
        // stream could still be queued when the first poll reads them -- seen once per ~300
        // strided renders as a frame delivered before it was rendered).
        CUDA_OKP(cudaMemsetAsync(pl->d_ctrl, 0, pl->zero_words * sizeof(uint32_t), pl->side));
        {
            // the tables go up from pinned staging (pooled: cudaMallocHost costs a millisecond)
            void* stage = nullptr; size_t cap = 0;
            CUDA_OKP(pool_pinned_buf(device, &stage, &cap, table_bytes));
            memcpy(stage, table_image.data(), table_bytes);
            cudaError_t e1 = cudaMemcpyAsync(pl->d_arena, stage, table_bytes, cudaMemcpyHostToDevice, pl->side);
            cudaError_t e2 = cudaStreamSynchronize(pl->side);
            pool_pinned_buf_free(device, stage, cap);
            CUDA_OKP(e1); CUDA_OKP(e2);
        }
        pl->h_flags.assign((size_t)pl->nbands + 1, 0u);
        CUDA_OKP(pool_event(device, &pl->done_ev));
        CUDA_OKP(pool_pinned(device, &pl->h_pinned));                    // staging words for progress / cancel traffic

        const int smem = coop_k ? mdz_smem_words_coop(coop_k / 100, coop_k % 100) * (int)sizeof(uint32_t)                // one shifter strip per group
                                : (gmp ? gmp_smem_words(n32 / 2) : smem_words_for_limbs(n32)) * kBlock * (int)sizeof(uint32_t);   // c_re, c_im, shifter scratch, checkpoint
        if (!kernel_facts(device, fn, smem, n32, &pl->info)) goto fail;
        pl->info.limbs = !coop_k ? n32 : gmp ? 2 * (int)(((v->precision < 53 ? 53 : v->precision) + 127) / 64 + 1) : limbs32_for_prec(v->precision);
        pl->info.lanes_per_pixel = coop_k ? coop_k % 100 : 1;
    }
    return pl;
fail:
    mdzcuda_plan_destroy(pl);
    return nullptr;
}

extern "C" int mdzcuda_plan_set_cycle_detection(mdzcuda_plan* pl, int on)
{
    if (!pl) { set_err("null plan"); return 0; }
    pl->cycle = on ? 1 : 0;
    return 1;
}

// ---------------------------------------------------------------------------
// Fed plans: the device side is escape_kernel.cuh "The pixel queue"; the policy is band_grants.h.
// ---------------------------------------------------------------------------
extern "C" void* mdzcuda_plan_stream(mdzcuda_plan* pl) { return pl ? (void*)pl->own : nullptr; }

// position -> band for a sequence over n bands: raster, or from the middle outwards (mid, mid+1, mid-1, ...)
static inline int band_at(int mode, int n, int pos)
{
    if (mode != MDZCUDA_ORDER_CENTRE_OUT) return pos;
    const int mid = n / 2;
    const int k = (pos + 1) / 2;
    int b = (pos & 1) ? mid + k : mid - k;
    // one side runs out first when n is even / at the ends: the rest continues on the other side
    if (b < 0) b = pos;                 // positions beyond 2*mid: only the upper side is left
    if (b >= n) b = n - 1 - pos;        // (cannot happen for mid = n / 2, kept for safety)
    return b;
}

// tiles a band is cut into in centre-out order: as many as divide the width evenly, at most 16, at least 32 columns each
static int tiles_for_width(int width)
{
    for (int t = kMaxTilesPerBand; t >= 1; --t) if (width % t == 0 && width / t >= 32) return t;
    return 1;
}

extern "C" int mdzcuda_plan_set_order(mdzcuda_plan* pl, int mode)
{
    if (!pl) { set_err("null plan"); return 0; }
    if (mode != MDZCUDA_ORDER_RASTER && mode != MDZCUDA_ORDER_CENTRE_OUT) { set_err("unknown order %d", mode); return 0; }
    if (pl->order_mode != mode) pl->order_uploaded = false;
    pl->order_mode = mode;
    return 1;
}

extern "C" int mdzcuda_plan_set_fed(mdzcuda_plan* pl, int on)
{
    if (!pl) { set_err("null plan"); return 0; }
    if (on && (pl->band_first != 0 || pl->band_stride != 1)) { set_err("a fed plan must cover the whole image (band_first 0, band_stride 1)"); return 0; }
    if (on && !pl->h_stage) {
        void* q = nullptr; size_t cap = 0;
        CUDA_OK(cudaSetDevice(pl->device));
        CUDA_OK(pool_pinned_buf(pl->device, &q, &cap, ((size_t)pl->nbands * kMaxTilesPerBand + 8) * sizeof(unsigned int)));
        pl->h_stage = (unsigned int*)q; pl->h_stage_cap = cap;
    }
    pl->fed = on != 0;
    pl->order_uploaded = false;
    return 1;
}

extern "C" int mdzcuda_plan_feed(mdzcuda_plan* pl, const int* bands, int count, int close)
{
    if (!pl) { set_err("null plan"); return 0; }
    if (!pl->fed) { set_err("not a fed plan"); return 0; }
    if (count < 0 || pl->granted + count > pl->nbands) { set_err("feed: more bands than the plan has"); return 0; }
    if (pl->closed) { if (count == 0) return 1; set_err("feed: the launch is closed"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    // The staging buffer holds the slot table (nbands x tiles_per_band entries at most) and, behind it, the control words.
    unsigned int* ctl = pl->h_stage + (size_t)pl->nbands * kMaxTilesPerBand;
    if (count > 0) {
        // A chunk of bands becomes `tiles_per_band` runs of slots: the chunk's tiles column by column, the columns
        // nearest the middle of the image first (centre-out plans; one tile per band otherwise), so that what is
        // dense in the chunk -- deep zooms keep it in the middle -- is started first (mdzcuda_plan_set_order).
        const int nt = pl->tiles_per_band;
        unsigned int* dst = pl->h_stage + pl->granted_slots;
        for (int k = 0; k < count; ++k)
            if (bands[k] < 0 || bands[k] >= pl->nbands) { set_err("feed: band %d out of range", bands[k]); return 0; }
        for (int c = 0; c < nt; ++c) {
            const int t = band_at(MDZCUDA_ORDER_CENTRE_OUT, nt, c);      // column c of the sequence: middle outwards
            for (int k = 0; k < count; ++k) *dst++ = (unsigned int)(bands[k] * nt + t);
        }
        // order[] first, then the limit, then the close word: the kernel reads them in the opposite order
        CUDA_OK(cudaMemcpyAsync(pl->d_order + pl->granted_slots, pl->h_stage + pl->granted_slots, (size_t)count * nt * sizeof(unsigned int),
                                cudaMemcpyHostToDevice, pl->side));
        pl->granted += count;
        pl->granted_slots += count * nt;
        ctl[0] = (unsigned int)pl->granted_slots;
        CUDA_OK(cudaMemcpyAsync(pl->d_feed + 0, ctl + 0, sizeof(unsigned int), cudaMemcpyHostToDevice, pl->side));
    }
    if (close) {
        ctl[1] = pl->gen;
        CUDA_OK(cudaMemcpyAsync(pl->d_feed + 1, ctl + 1, sizeof(unsigned int), cudaMemcpyHostToDevice, pl->side));
        pl->closed = true;
    }
    CUDA_OK(cudaStreamSynchronize(pl->side));
    return 1;
}

// pixels granted to the current launch that no lane had claimed at the last mdzcuda_plan_poll_bands
extern "C" long long mdzcuda_plan_backlog(mdzcuda_plan* pl)
{
    if (!pl) { set_err("null plan"); return -1; }
    const long long band_px = (long long)pl->view.real_width * pl->view.aa_factor;
    const long long granted = (pl->fed ? (long long)pl->granted : (long long)pl->nbands) * band_px;
    const long long b = granted - (long long)pl->claimed;
    return b > 0 ? b : 0;
}

static std::atomic<long> g_fallback_lines(0);
void mdz_count_fallback_lines(long n) { g_fallback_lines += n; }
extern "C" long mdzcuda_fallback_lines(void) { return g_fallback_lines.load(); }

// Order in which phase 1 reads the parked list: a counting sort by iteration count (2048 buckets over
// 0..depth), one block.  Pixels that entered the list with about the same count have about the same
// number of iterations left -- exactly so for those that never escape -- and end up in the same warps,
// which then finish, and leave the SM to the others, as a whole.  Also zeroes phase 1's claim words.
__global__ void __launch_bounds__(1024)
park_sort_kernel(const unsigned int* park_count, const uint32_t* iters, unsigned int* perm, int depth,
                 unsigned int* zero, unsigned int zero_words)
{
    __shared__ unsigned int hist[2048];
    __shared__ unsigned int part[1024];
    const unsigned n = park_count[0];
    const unsigned t = threadIdx.x;
    for (unsigned i = t; i < zero_words; i += 1024) zero[i] = 0u;
    hist[t] = 0u; hist[t + 1024] = 0u;
    __syncthreads();
    const unsigned long long scale = ((unsigned long long)2048 << 32) / (unsigned long long)(depth > 0 ? depth : 1);
    for (unsigned i = t; i < n; i += 1024) {
        const unsigned b = (unsigned)(((unsigned long long)iters[i] * scale) >> 32);
        atomicAdd(&hist[b < 2047u ? b : 2047u], 1u);
    }
    __syncthreads();
    const unsigned a0 = hist[2 * t], a1 = hist[2 * t + 1];
    part[t] = a0 + a1;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {
        const unsigned v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const unsigned before = part[t] - (a0 + a1);
    hist[2 * t] = before; hist[2 * t + 1] = before + a0;
    __syncthreads();
    for (unsigned i = t; i < n; i += 1024) {
        const unsigned b = (unsigned)(((unsigned long long)iters[i] * scale) >> 32);
        perm[atomicAdd(&hist[b < 2047u ? b : 2047u], 1u)] = i;
    }
}

extern "C" int mdzcuda_plan_kernels_launched(mdzcuda_plan* pl)
{
    if (!pl) { set_err("null plan"); return -1; }
    return pl->kernels_launched;
}

extern "C" int mdzcuda_plan_set_parking(mdzcuda_plan* pl, int mode)
{
    if (!pl) { set_err("null plan"); return 0; }
    pl->park = mode < 0 ? -1 : (mode ? 1 : 0);
    return 1;
}

extern "C" int mdzcuda_plan_tune(mdzcuda_plan* pl, int chunk_iters, int blocks_per_sm)
{
    if (!pl) return 0;
    pl->spec = chunk_iters < 0 ? 0 : 1;      // negative chunk: speculative body off (A/B measurements)
    pl->chunk = chunk_iters < 0 ? -chunk_iters : chunk_iters;
    pl->blocks_per_sm = blocks_per_sm > 0 ? blocks_per_sm : 0;
    return 1;
}

static int default_chunk(int n32)
{
    // refill cost is roughly two squarings plus an atomic round trip; keep it
    // to a few percent of a chunk for every limb count
    if (n32 <= 2) return 32;
    if (n32 <= 4) return 16;
    if (n32 > 32) return 4;             // one warp per pixel: an iteration is microseconds, the poll a few hundred cycles
    return 8;
}

extern "C" int mdzcuda_plan_launch(mdzcuda_plan* pl, void* cuda_stream)
{
    if (!pl) { set_err("null plan"); return 0; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CUDA_OK(cudaSetDevice(pl->device));
    CUDA_OK(cudaMemsetAsync(pl->d_ctrl, 0, pl->reset_words * sizeof(uint32_t), st));   // queue, counters, flags
    pl->claimed = 0;
    if (pl->fed) {
        // the fill count restarts at zero; on the side stream, where the feeds will follow it in order
        // (the close word carries the launch generation and needs no reset)
        pl->granted = 0; pl->granted_slots = 0; pl->closed = false;
        pl->tiles_per_band = pl->order_mode == MDZCUDA_ORDER_CENTRE_OUT ? tiles_for_width(pl->view.real_width) : 1;
        CUDA_OK(cudaMemsetAsync(pl->d_feed, 0, sizeof(unsigned int), pl->side));
        CUDA_OK(cudaStreamSynchronize(pl->side));
    }
    {
        // test hook: fill the iteration buffer with a pattern no render produces, so that a band
        // delivered before it was complete shows up deterministically (tests/test_parity_gpu.py)
        static const bool poison = [] { const char* e = getenv("MDZCUDA_DEBUG_POISON"); return e && *e && *e != '0'; }();
        if (poison && pl->local_lines > 0)
            CUDA_OK(cudaMemsetAsync(pl->d_raw, 0x7f, (size_t)pl->local_lines * pl->view.real_width * sizeof(int32_t), st));
    }
    if (pl->local_lines > 0) {
        EscapeParams p;
        p.xs = pl->xs.view(); p.ys = pl->ys.view(); p.jc = pl->jc.view();
        p.rc = pl->rc;
        p.raw = pl->d_raw;
        p.queue = pl->d_ctrl + 0;
        p.bands_done = pl->d_ctrl + 1;
        p.cancel = (const volatile unsigned int*)pl->d_cancel;
        p.band_count = pl->d_band_count;
        p.band_flag = pl->d_band_flag;
        if (++pl->gen == 0) pl->gen = 1;
        p.gen = pl->gen;
        p.width = pl->view.real_width;
        p.lines = pl->local_lines;
        p.aa = pl->view.aa_factor;
        p.depth = (int)pl->view.depth;
        p.family = pl->view.family;
        p.fractal = pl->view.fractal;
        p.chunk = pl->chunk ? pl->chunk : default_chunk(pl->n32);
        p.prec_bits = (int)pl->view.precision;
        if (pl->gmp && pl->coop_k) p.prec_bits = (int)(((pl->view.precision < 53 ? 53 : pl->view.precision) + 127) / 64 + 1);  // NL = P + 1 limbs (coop_mpf.cuh)
        p.spec = pl->spec;
        if (p.spec == 1) { static const int forced = [] { const char* e = getenv("MDZCUDA_SPEC_LEVEL"); return e && *e ? atoi(e) : 1; }(); p.spec = forced; }   // A/B: 2 / 3 pin level 1 / 2
        p.colour = pl->colour;
        if (!pl->fed && pl->order_mode != MDZCUDA_ORDER_RASTER && !pl->order_uploaded) {
            // A static plan that starts in the middle of the image: its bands are cut into tiles (as many as divide
            // the width evenly, at most 16) and the tiles sorted by their distance from the image centre, in units
            // of the image's half extent -- the slot table, uploaded once.
            const int ntile = tiles_for_width(pl->view.real_width);
            pl->tiles_per_band = ntile;
            const int total_bands = pl->view.real_height / pl->view.aa_factor;
            std::vector<std::pair<double, unsigned int> > key((size_t)pl->nbands * ntile);
            for (int b = 0; b < pl->nbands; ++b) {
                const double gy = ((pl->band_first + (double)b * pl->band_stride + 0.5) / total_bands - 0.5) * 2.0;
                for (int t = 0; t < ntile; ++t) {
                    const double gx = ((t + 0.5) / ntile - 0.5) * 2.0;
                    key[(size_t)b * ntile + t] = std::make_pair(gx * gx + gy * gy, (unsigned int)(b * ntile + t));
                }
            }
            std::sort(key.begin(), key.end());
            std::vector<unsigned int> seq(key.size());
            for (size_t k = 0; k < key.size(); ++k) seq[k] = key[k].second;
            CUDA_OK(cudaMemcpyAsync(pl->d_order, seq.data(), seq.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, pl->side));
            CUDA_OK(cudaStreamSynchronize(pl->side));
            pl->order_uploaded = true;
        }
        if (!pl->fed && pl->order_mode == MDZCUDA_ORDER_RASTER) pl->tiles_per_band = 1;
        p.tiles_per_band = pl->tiles_per_band;
        p.tile_w = pl->view.real_width / pl->tiles_per_band;
        p.order = (pl->fed || pl->order_mode != MDZCUDA_ORDER_RASTER) ? pl->d_order : nullptr;
        p.feed = pl->fed ? (const volatile unsigned int*)pl->d_feed : nullptr;
        p.ld_masks.im_keep = p.fractal == FRACTAL_BURNING_SHIP ? 0u : 1u;           // ld64_step.cuh: ld64_masks
        p.ld_masks.re_and = p.fractal == FRACTAL_VARIANT ? 1u : 0u;
        p.ld_masks.re_xor = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 0u : 1u;
        const int cyc = (pl->cycle && !pl->gmp && !pl->coop_k) ? 1 : 0;
        kernel_fn fn = pl->coop_k ? (pl->gmp ? mdz_kernel_coop_gmp : mdz_kernel_coop)(pl->coop_k / 100, pl->coop_k % 100) : pl->gmp ? gmp_kernel_for_limbs(pl->n32 / 2) : kernel_for_limbs(pl->n32, cyc);
        if (!fn) { set_err("no kernel for %d limbs", pl->n32); return 0; }
        mdzcuda_kernel_info ki;
        if (!kernel_facts(pl->device, fn, pl->info.shared_bytes, pl->n32, &ki)) return 0;
        int bps = pl->blocks_per_sm ? pl->blocks_per_sm : ki.blocks_per_sm;
        if (bps > ki.blocks_per_sm) bps = ki.blocks_per_sm;
        long long npx = (long long)pl->local_lines * pl->view.real_width;
        long long grid = (long long)bps * ki.sm_count;
        const int px_per_block = pl->coop_k ? kBlock / (pl->coop_k % 100) : kBlock;
        long long need = (npx + px_per_block - 1) / px_per_block;
        if (grid > need) grid = need;
        ki.grid_blocks = (int)grid;
        ki.limbs = pl->info.limbs; ki.lanes_per_pixel = pl->info.lanes_per_pixel;
        pl->info = ki;
        p.cycle = cyc;
        p.cycle_scratch = nullptr;
        if (cyc) {
            const size_t words = (size_t)(2 * pl->n32 + 3) * (size_t)grid * kBlock;
            if (words > pl->cycle_words) {
                if (pl->d_cycle) { CUDA_OK(cudaStreamSynchronize(st)); pool_free(pl->device, pl->d_cycle); pl->d_cycle = nullptr; }
                CUDA_OK(pool_alloc(pl->device, (void**)&pl->d_cycle, words * sizeof(uint32_t)));
                pl->cycle_words = words;
            }
            p.cycle_scratch = pl->d_cycle;
        }
        // Tail compaction: a second launch for the pixels still in flight when the queue runs dry
        // (DESIGN.md 4.3).  It pays when this plan's last generation of long-running pixels is sparse --
        // one GPU's share of a strong-scaled render: 14.5 -> 9.8 ms on an eighth of config 2 -- and is
        // neutral when it is dense; not worth the launches for images below twice the grid.
        // MDZCUDA_PARK=0 / 1 or mdzcuda_plan_set_parking force it off / on.
        p.phase = 0; p.park_cap = 0; p.park_count = pl->d_ctrl + 2; p.park_buf = nullptr; p.park_perm = nullptr;
        p.park_smslot = nullptr; p.park_claimed = nullptr; p.park_sms = 1;
        static const int park_env = [] { const char* e = getenv("MDZCUDA_PARK"); return e && *e ? atoi(e) : -1; }();
        const int park_mode = pl->park >= 0 ? pl->park : park_env;
        const bool park = !pl->gmp && !pl->fed && pl->n32 <= kParkMaxLimbs && (park_mode > 0 || (park_mode < 0 && npx >= 2 * grid * kBlock));
        if (park) {
            const size_t cap = (size_t)grid * kBlock;
            const size_t state_words = (size_t)(6 * pl->n32 + 9) * cap;
            const size_t claim_words = 4096 + (cap + 31) / 32;
            const size_t words = state_words + cap + claim_words;
            if (words > pl->park_words) {
                if (pl->d_park) { CUDA_OK(cudaStreamSynchronize(st)); pool_free(pl->device, pl->d_park); pl->d_park = nullptr; pl->park_words = 0; }
                CUDA_OK(pool_alloc(pl->device, (void**)&pl->d_park, words * sizeof(uint32_t)));
                pl->park_words = words;
            }
            p.park_cap = (unsigned int)cap;
            p.park_buf = pl->d_park;
            p.park_perm = pl->d_park + state_words;
            p.park_smslot = pl->d_park + state_words + cap;
            p.park_claimed = p.park_smslot + 4096;
            p.park_sms = ki.sm_count > 0 ? ki.sm_count : 1;
        }
        fn<<<(unsigned)grid, kBlock, ki.shared_bytes, st>>>(p);
        CUDA_OK(cudaGetLastError());
        pl->kernels_launched += 1;
        if (park) {
            p.phase = 1;
            park_sort_kernel<<<1, 1024, 0, st>>>(p.park_count, p.park_buf + (size_t)(6 * pl->n32 + 7) * p.park_cap,
                                                 (unsigned int*)p.park_perm, p.depth, p.park_smslot,
                                                 (unsigned int)(4096 + ((size_t)p.park_cap + 31) / 32));
            CUDA_OK(cudaGetLastError());
            fn<<<(unsigned)grid, kBlock, ki.shared_bytes, st>>>(p);
            CUDA_OK(cudaGetLastError());
            pl->kernels_launched += 2;
        }
    }
    CUDA_OK(cudaEventRecord(pl->done_ev, st));
    return 1;
}

extern "C" int mdzcuda_plan_wait(mdzcuda_plan* pl)
{
    if (!pl) { set_err("null plan"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    CUDA_OK(cudaEventSynchronize(pl->done_ev));
    return 1;
}

extern "C" int mdzcuda_plan_cancel(mdzcuda_plan* pl)
{
    if (!pl) { set_err("null plan"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    // the kernel stops when the word equals its own generation; nothing ever resets the word, so
    // a stop that arrives before the launch's reset has executed cannot be wiped out by it
    pl->h_pinned[0] = pl->gen;
    CUDA_OK(cudaMemcpyAsync(pl->d_cancel, pl->h_pinned, sizeof(unsigned int), cudaMemcpyHostToDevice, pl->side));
    CUDA_OK(cudaStreamSynchronize(pl->side));
    return 1;
}

extern "C" int mdzcuda_plan_bands_done(mdzcuda_plan* pl)
{
    if (!pl) { set_err("null plan"); return -1; }
    if (pl->nbands == 0) return 0;
    std::vector<unsigned char> f((size_t)pl->nbands);
    return mdzcuda_plan_poll_bands(pl, f.data());
}

extern "C" int mdzcuda_plan_bands_total(mdzcuda_plan* pl) { return pl ? pl->nbands : -1; }

extern "C" int mdzcuda_plan_fetch(mdzcuda_plan* pl, int32_t* raw_host)
{
    if (!pl || !raw_host) { set_err("null argument"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    CUDA_OK(cudaEventSynchronize(pl->done_ev));
    const int W = pl->view.real_width, aa = pl->view.aa_factor;
    if (pl->band_stride == 1 && pl->band_first == 0) {
        CUDA_OK(cudaMemcpy(raw_host, pl->d_raw, (size_t)pl->local_lines * W * sizeof(int32_t), cudaMemcpyDeviceToHost));
        return 1;
    }
    // strided bands: one 2D copy (band = aa*W contiguous ints on both sides)
    const size_t band_bytes = (size_t)aa * W * sizeof(int32_t);
    if (pl->nbands > 0)
        CUDA_OK(cudaMemcpy2D(raw_host + (size_t)pl->band_first * aa * W, band_bytes * pl->band_stride,
                             pl->d_raw, band_bytes, band_bytes, pl->nbands, cudaMemcpyDeviceToHost));
    return 1;
}


extern "C" int mdzcuda_plan_poll_bands(mdzcuda_plan* pl, unsigned char* flags_host)
{
    if (!pl || !flags_host) { set_err("null argument"); return -1; }
    if (cudaSetDevice(pl->device) != cudaSuccess) return -1;
    if (pl->nbands == 0) return 0;
    // A band is complete when its flag holds the generation of the current launch; whatever an
    // earlier launch (or an earlier plan that owned this memory) left there does not match, so
    // the poll needs no ordering against the reset that the launch enqueues on the caller's stream.
    if (cudaMemcpyAsync(pl->h_flags.data(), pl->d_band_flag, (size_t)pl->nbands * sizeof(unsigned int),
                        cudaMemcpyDeviceToHost, pl->side) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(pl->h_pinned + 4, pl->d_ctrl, sizeof(unsigned int), cudaMemcpyDeviceToHost, pl->side) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(pl->side) != cudaSuccess) return -1;
    pl->claimed = pl->h_pinned[4];
    int done = 0;
    const unsigned int gen = pl->gen;
    for (int i = 0; i < pl->nbands; ++i) { const int c = gen != 0 && pl->h_flags[i] == gen; flags_host[i] = (unsigned char)c; done += c; }
    return done;
}

extern "C" int mdzcuda_plan_fetch_bands(mdzcuda_plan* pl, int32_t* raw_host, int first_local_band, int count)
{
    if (!pl || !raw_host) { set_err("null argument"); return 0; }
    if (first_local_band < 0 || count < 0 || first_local_band + count > pl->nbands) { set_err("band range"); return 0; }
    if (count == 0) return 1;
    CUDA_OK(cudaSetDevice(pl->device));
    const int W = pl->view.real_width, aa = pl->view.aa_factor;
    const size_t band_bytes = (size_t)aa * W * sizeof(int32_t);
    const int gband = pl->band_first + first_local_band * pl->band_stride;
    CUDA_OK(cudaMemcpy2DAsync(raw_host + (size_t)gband * aa * W, band_bytes * pl->band_stride,
                              pl->d_raw + (size_t)first_local_band * aa * W, band_bytes, band_bytes, count,
                              cudaMemcpyDeviceToHost, pl->side));
    CUDA_OK(cudaStreamSynchronize(pl->side));
    return 1;
}


// ---- colour epilogue -------------------------------------------------------------
extern "C" int mdzcuda_plan_set_colour(mdzcuda_plan* pl, const mdzcuda_colour* c)
{
    if (!pl) { set_err("null plan"); return 0; }
    if (!c) { pl->colour.enabled = 0; return 1; }
    if (!c->palette || c->pal_indexes < 2 || c->pal_indexes > 256) { set_err("palette needs 2..256 entries"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    const int uw = pl->view.real_width / pl->view.aa_factor;
    if (!pl->d_palette) CUDA_OK(pool_alloc(pl->device, (void**)&pl->d_palette, 256 * sizeof(uint32_t)));
    if (!pl->d_rgb) CUDA_OK(pool_alloc(pl->device, (void**)&pl->d_rgb, ((size_t)pl->nbands * uw + 1) * sizeof(uint32_t)));
    CUDA_OK(cudaMemcpy(pl->d_palette, c->palette, (size_t)c->pal_indexes * sizeof(uint32_t), cudaMemcpyHostToDevice));
    pl->colour.palette = pl->d_palette;
    pl->colour.rgb = pl->d_rgb;
    pl->colour.scale = c->colour_scale;
    pl->colour.pal_indexes = c->pal_indexes;
    pl->colour.pal_offset = c->pal_offset;
    pl->colour.interpolate = c->palette_ip ? 1 : 0;
    pl->colour.enabled = 1;
    return 1;
}

extern "C" int mdzcuda_plan_recolour(mdzcuda_plan* pl, void* cuda_stream)
{
    if (!pl) { set_err("null plan"); return 0; }
    if (!pl->colour.enabled) { set_err("no colour parameters set"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    if (pl->nbands > 0) {
        const int blocks = (pl->nbands + 3) / 4;
        recolour_kernel<<<blocks, 128, 0, (cudaStream_t)cuda_stream>>>(pl->d_raw, pl->view.real_width,
                                                                      pl->view.aa_factor, pl->nbands, pl->colour);
        CUDA_OK(cudaGetLastError());
        pl->kernels_launched += 1;
    }
    CUDA_OK(cudaEventRecord(pl->done_ev, (cudaStream_t)cuda_stream));
    return 1;
}

extern "C" int mdzcuda_plan_fetch_rgb(mdzcuda_plan* pl, uint32_t* rgb_host)
{
    if (!pl || !rgb_host) { set_err("null argument"); return 0; }
    if (!pl->colour.enabled) { set_err("no colour parameters set"); return 0; }
    CUDA_OK(cudaSetDevice(pl->device));
    CUDA_OK(cudaEventSynchronize(pl->done_ev));
    const int uw = pl->view.real_width / pl->view.aa_factor;
    const size_t line_bytes = (size_t)uw * sizeof(uint32_t);
    if (pl->nbands > 0)
        CUDA_OK(cudaMemcpy2D(rgb_host + (size_t)pl->band_first * uw, line_bytes * pl->band_stride,
                             pl->d_rgb, line_bytes, line_bytes, pl->nbands, cudaMemcpyDeviceToHost));
    return 1;
}

extern "C" void* mdzcuda_plan_device_raw(mdzcuda_plan* pl) { return pl ? pl->d_raw : nullptr; }
extern "C" int mdzcuda_plan_local_lines(mdzcuda_plan* pl) { return pl ? pl->local_lines : -1; }

extern "C" int mdzcuda_plan_kernel_info(mdzcuda_plan* pl, mdzcuda_kernel_info* out)
{
    if (!pl || !out) return 0;
    *out = pl->info;
    return 1;
}

extern "C" void mdzcuda_plan_destroy(mdzcuda_plan* pl)
{
    if (!pl) return;
    cudaSetDevice(pl->device);
    // the blocks go back to the pool, so nothing of this plan may still be running
    if (pl->done_ev) cudaEventSynchronize(pl->done_ev);
    if (pl->side) cudaStreamSynchronize(pl->side);
    DevPool& P = g_pool[pl->device];
    pool_free(pl->device, pl->d_palette); pool_free(pl->device, pl->d_rgb);
    pool_free(pl->device, pl->d_raw); pool_free(pl->device, pl->d_arena); pool_free(pl->device, pl->d_cycle); pool_free(pl->device, pl->d_park);
    if (pl->own) cudaStreamSynchronize(pl->own);
    pool_pinned_buf_free(pl->device, pl->h_stage, pl->h_stage_cap);
    {
        std::lock_guard<std::mutex> lock(P.mu);
        if (pl->own) P.streams.push_back(pl->own);
        if (pl->side) P.streams.push_back(pl->side);
        if (pl->done_ev) P.events.push_back(pl->done_ev);
        if (pl->h_pinned) P.pinned.push_back(pl->h_pinned);
    }
    delete pl;
}

// ---------------------------------------------------------------------------
// The render driver: launched plans -> the caller's raw_data.
//
// Finished bands are copied into raw_host while the kernels are still running, so that by the time they
// end only the last few are still on the device (the D2H of an 8 MB raw_data otherwise adds 2-3 ms to a
// 60 ms render), and -- rth.cpp -- so that MDZ's consumer loop sees lines arrive progressively.
//
// Band scheduler (several devices, one image; SURVEY 8e "Partitioning").  Every device holds a *fed* plan
// of the whole image whose persistent kernel eats a queue of band slots that this loop fills while it
// runs: a device is given the next chunk of bands (band_grants.h: large first, small at the end) whenever
// the unclaimed part of its queue falls below one grid's worth of pixels.  A device that is slow -- shared
// with another job, or holding the deep part of the image -- asks less often; nothing is decided up front.
// The reference does the same with lines and a mutex (src/render_threads.c:360-393).
// ---------------------------------------------------------------------------
static int run_plans(std::vector<mdzcuda_plan*>& plans, int32_t* raw_host, const mdz_run_hooks* hooks, mdz::BandGrants* grants)
{
    const int order_mode = hooks ? hooks->order : MDZCUDA_ORDER_RASTER;
    const int n = (int)plans.size();
    std::vector<std::vector<unsigned char> > flags(n), seen(n);
    std::vector<int> delivered(n, 0);
    int remaining = 0;
    for (int i = 0; i < n; ++i) {
        const int nb = plans[i]->nbands;
        flags[i].assign((size_t)nb + 1, 0); seen[i].assign((size_t)nb + 1, 0);
        if (!grants) remaining += nb;
    }
    if (grants) remaining = grants->total;
    const int min_run_cfg = hooks && hooks->min_run > 0 ? hooks->min_run : 16;
    const struct timespec nap = { 0, 200 * 1000 };
    int rc = 1;
    while (remaining > 0) {
        if (hooks && hooks->should_stop && hooks->should_stop(hooks->user)) {
            for (int i = 0; i < n; ++i) mdzcuda_plan_cancel(plans[i]);
            rc = 2;
            break;
        }
        bool progress = false;
        for (int i = 0; i < n; ++i) {
            mdzcuda_plan* pl = plans[i];
            const int nb = pl->nbands;
            const int expect = pl->fed ? pl->granted : nb;       // bands this plan is to deliver (so far)
            if (!pl->fed && delivered[i] == nb) continue;
            CUDA_OK(cudaSetDevice(pl->device));
            const cudaError_t q = cudaEventQuery(pl->done_ev);
            if (q != cudaSuccess && q != cudaErrorNotReady) { set_err("kernel failed: %s", cudaGetErrorString(q)); return 0; }
            const bool finished = q == cudaSuccess;
            if (mdzcuda_plan_poll_bands(pl, flags[i].data()) < 0) return 0;
            // whole runs only, and not in crumbs: each copy costs ~20 us of driver time
            const int min_run = finished ? 1 : min_run_cfg;
            for (int b = 0; b < nb;) {
                if (!flags[i][b] || seen[i][b]) { ++b; continue; }
                int e = b;
                while (e < nb && flags[i][e] && !seen[i][e]) ++e;
                if (e - b >= min_run) {
                    if (!mdzcuda_plan_fetch_bands(pl, raw_host, b, e - b)) return 0;
                    for (int k = b; k < e; ++k) seen[i][k] = 1;
                    delivered[i] += e - b; remaining -= e - b;
                    if (hooks && hooks->bands_ready)
                        hooks->bands_ready(hooks->user, pl->band_first + b * pl->band_stride, e - b, pl->band_stride);
                    progress = true;
                }
                b = e;
            }
            if (finished && delivered[i] < expect && (!pl->fed || pl->closed)) {
                // the launch is over and a band it owed is not flagged: a kernel that died without a sticky error
                set_err("device %d: the kernel ended with %d of %d bands missing", pl->device, expect - delivered[i], expect);
                return 0;
            }
            if (grants && !pl->closed) {
                const long long band_px = (long long)pl->view.real_width * pl->view.aa_factor;
                const long long low_water = (long long)pl->info.grid_blocks * kBlock;
                long long backlog = mdzcuda_plan_backlog(pl);
                grants->progress(i, (double)pl->claimed / (double)band_px);
                while (backlog < low_water && grants->remaining() > 0) {
                    int first = 0;
                    const int cnt = grants->take(i, &first);
                    std::vector<int> list((size_t)cnt);
                    for (int k = 0; k < cnt; ++k) list[k] = band_at(order_mode, grants->total, first + k);
                    if (!mdzcuda_plan_feed(pl, list.data(), cnt, 0)) return 0;
                    backlog += cnt * band_px;
                    progress = true;
                }
                if (grants->remaining() == 0)
                    for (int j = 0; j < n; ++j)
                        if (!plans[j]->closed && !mdzcuda_plan_feed(plans[j], nullptr, 0, 1)) return 0;
            }
        }
        if (!progress && remaining > 0) nanosleep(&nap, 0);
    }
    for (int i = 0; i < n; ++i) if (!mdzcuda_plan_wait(plans[i])) return 0;
    return rc;
}

extern "C" int mdzcuda_plan_run(mdzcuda_plan* pl, void* cuda_stream, int32_t* raw_host)
{
    if (!pl || !raw_host) { set_err("null argument"); return 0; }
    if (pl->fed) { set_err("a fed plan is driven by mdzcuda_render"); return 0; }
    if (!mdzcuda_plan_launch(pl, cuda_stream)) return 0;
    std::vector<mdzcuda_plan*> one(1, pl);
    return run_plans(one, raw_host, nullptr, nullptr) == 1;
}

int mdz_run_view(const mdzcuda_view* view, int32_t* raw_host, const int* devices, int ndev, const mdz_run_hooks* hooks)
{
    g_err.clear();
    if (!view || !raw_host) { set_err("null argument"); return 0; }
    if (ndev < 1) ndev = 1;
    const int total_bands = view->aa_factor > 0 ? view->real_height / view->aa_factor : 0;
    if (ndev > total_bands) ndev = total_bands > 0 ? total_bands : 1;
    const char* sched_env = getenv("MDZCUDA_SCHED");
    const bool force_static = sched_env && !strcmp(sched_env, "static");
    const bool dynamic = ndev > 1 && !force_static;
    std::vector<mdzcuda_plan*> plans(ndev, nullptr);
    int ok = 1;
    {
        // plan creation runs the O(W+H) prologue; per device in parallel host threads
        std::vector<std::thread> th;
        std::vector<std::string> errs(ndev);
        auto make = [&](int i) {
            const int dev = devices ? devices[i] : i;
            plans[i] = dynamic ? mdzcuda_plan_create(view, dev, 0, 1) : mdzcuda_plan_create(view, dev, i, ndev);
            if (plans[i] && dynamic && !mdzcuda_plan_set_fed(plans[i], 1)) { mdzcuda_plan_destroy(plans[i]); plans[i] = nullptr; }
            if (!plans[i]) errs[i] = mdzcuda_last_error();
        };
        if (ndev == 1) make(0);
        else {
            for (int i = 0; i < ndev; ++i) th.emplace_back(make, i);
            for (auto& t : th) t.join();
        }
        for (int i = 0; i < ndev; ++i) if (!plans[i]) { ok = 0; set_err("%s", errs[i].c_str()); }
    }
    if (ok && hooks && hooks->cycle_detection >= 0)
        for (int i = 0; i < ndev; ++i) mdzcuda_plan_set_cycle_detection(plans[i], hooks->cycle_detection);
    if (ok && hooks)
        for (int i = 0; i < ndev; ++i) mdzcuda_plan_set_order(plans[i], hooks->order);
    for (int i = 0; ok && i < ndev; ++i) ok = mdzcuda_plan_launch(plans[i], plans[i]->own);
    int rc = 0;
    if (ok) {
        if (dynamic) {
            // at least a quarter of a grid's worth of pixels per chunk, so that a grant is worth its two copies
            const long long band_px = (long long)view->real_width * view->aa_factor;
            const long long grid_px = (long long)plans[0]->info.grid_blocks * kBlock;
            long long mc = (grid_px / 4 + band_px - 1) / band_px;
            mdz::BandGrants grants(total_bands, ndev, (int)(mc < 1 ? 1 : mc));
            rc = run_plans(plans, raw_host, hooks, &grants);
        } else rc = run_plans(plans, raw_host, hooks, nullptr);
        if (rc == 0) for (int i = 0; i < ndev; ++i) mdzcuda_plan_cancel(plans[i]);     // do not leave a fed kernel waiting for bands
    }
    std::string keep = g_err;
    for (int i = 0; i < ndev; ++i) mdzcuda_plan_destroy(plans[i]);
    g_err = keep;
    return rc;
}

extern "C" int mdzcuda_render(const mdzcuda_view* view, int32_t* raw_host, int ndev, const int* devices)
{
    mdz_run_hooks h;
    memset(&h, 0, sizeof h);
    h.min_run = 16;
    h.cycle_detection = -1;         // as the plans' default (MDZCUDA_CYCLE_DETECT)
    // nobody watches the lines arrive here, so the bands in the middle of the image go first (see mdzcuda_plan_set_order)
    { const char* e = getenv("MDZCUDA_ORDER"); h.order = (e && !strcmp(e, "raster")) ? MDZCUDA_ORDER_RASTER : MDZCUDA_ORDER_CENTRE_OUT; }
    return mdz_run_view(view, raw_host, devices, ndev, &h) == 1;
}

// ---------------------------------------------------------------------------
// Integer-multiply peak microbenchmarks (register-only).
//
//  wide:  IMAD.WIDE.U32(.X) carry chains exactly as mul_full / mul_hi emit them
//         (two interleaved chains of four 32x32->64 multiply-accumulates) -- the
//         instruction that performs the unit SURVEY 8(d) counts.
//  lo32:  independent 32-bit IMAD (low half only), the pipe's nominal issue
//         rate that SURVEY 8(d) quotes as "64 IMAD/clk/SM".
// The multiplier operand is data dependent so ptxas cannot hoist the product
// (it does, and turns a naive loop into plain 64-bit adds).
// On B200 the first runs at half the rate of the second (profiles/r1_pipe_microbench.txt).
// ---------------------------------------------------------------------------
template <int WIDE>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t seed, int iters, unsigned long long* out)
{
    uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
    uint32_t r0 = a, r1 = b, r2 = a ^ b, r3 = a + b, r4 = a * 3, r5 = b * 5, r6 = a * 7, r7 = b * 9;
    uint32_t r8 = a + 1, r9 = b + 2, r10 = a + 3, r11 = b + 4, r12 = a + 5, r13 = b + 6, r14 = a + 7, r15 = b + 8;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (WIDE) {
                asm volatile(
                    "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %1,%16,%17,%1; madc.lo.cc.u32 %2,%16,%17,%2; madc.hi.cc.u32 %3,%16,%17,%3;"
                    "madc.lo.cc.u32 %4,%16,%17,%4; madc.hi.cc.u32 %5,%16,%17,%5; madc.lo.cc.u32 %6,%16,%17,%6; madc.hi.u32 %7,%16,%17,%7;"
                    "mad.lo.cc.u32 %8,%16,%17,%8; madc.hi.cc.u32 %9,%16,%17,%9; madc.lo.cc.u32 %10,%16,%17,%10; madc.hi.cc.u32 %11,%16,%17,%11;"
                    "madc.lo.cc.u32 %12,%16,%17,%12; madc.hi.cc.u32 %13,%16,%17,%13; madc.lo.cc.u32 %14,%16,%17,%14; madc.hi.u32 %15,%16,%17,%15;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                      "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15) : "r"(a), "r"(b));
            } else {
                asm volatile("mad.lo.u32 %0,%1,%8,%0; mad.lo.u32 %1,%2,%8,%1; mad.lo.u32 %2,%3,%8,%2; mad.lo.u32 %3,%4,%8,%3;"
                             "mad.lo.u32 %4,%5,%8,%4; mad.lo.u32 %5,%6,%8,%5; mad.lo.u32 %6,%7,%8,%6; mad.lo.u32 %7,%0,%8,%7;"
                    : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(b));
            }
        }
    }
    uint32_t x = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ r8 ^ r9 ^ r10 ^ r11 ^ r12 ^ r13 ^ r14 ^ r15;
    if (x == 0x1234567u) out[0] = x;      // keep the chains alive
}

template <int WIDE>
static double run_imad_peak(int device, int ms)
{
    g_err.clear();
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return 0.0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_err("cudaGetDeviceProperties failed"); return 0.0; }
    unsigned long long* d_out = nullptr;
    if (cudaMalloc(&d_out, 8) != cudaSuccess) { set_err("cudaMalloc failed"); return 0.0; }
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000;
    double best = 0.0, spent = 0.0;
    imad_peak_kernel<WIDE><<<blocks, threads>>>(1u, 200, d_out);        // warm-up
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 64 && spent < (double)ms; ++rep) {
        cudaEventRecord(e0);
        imad_peak_kernel<WIDE><<<blocks, threads>>>(rep + 2u, iters, d_out);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { set_err("imad kernel failed"); best = 0.0; break; }
        float t = 0; cudaEventElapsedTime(&t, e0, e1);
        spent += t;
        // multiply-accumulates per thread per iteration: 8 x (8 wide) or 8 x (8 lo)
        const double macs = (double)blocks * threads * (double)iters * 64.0;
        const double rate = macs / (t * 1e-3);
        if (rate > best) best = rate;
        if (t < 5.0f) iters *= 2;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

extern "C" double mdzcuda_imad_peak(int device, int ms) { return run_imad_peak<1>(device, ms); }
extern "C" double mdzcuda_imad32_peak(int device, int ms) { return run_imad_peak<0>(device, ms); }

// ---------------------------------------------------------------------------
// Test hook (tests/test_multi_gpu.py): occupy `blocks` SMs of a device for `ms` milliseconds -- one block
// per SM, each claiming all of the SM's shared memory so that nothing else fits beside it -- to play a
// device that is busy with someone else's work.  Asynchronous; returns once the kernel is running.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) occupy_kernel(unsigned long long ns, unsigned int* started)
{
    extern __shared__ uint32_t hog[];
    hog[threadIdx.x] = threadIdx.x;
    if (threadIdx.x == 0) atomicAdd(started, 1u);
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(20000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
    if (hog[threadIdx.x] == 0xffffffffu) started[1] = 1u;
}

extern "C" int mdzcuda_debug_occupy(int device, int blocks, int ms)
{
    g_err.clear();
    CUDA_OK(cudaSetDevice(device));
    int smem = 0;
    CUDA_OK(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_OK(cudaFuncSetAttribute((const void*)occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    static unsigned int* d_started[kMaxDev];
    static cudaStream_t hog_stream[kMaxDev];
    if (device >= kMaxDev) { set_err("device out of range"); return 0; }
    if (!d_started[device]) {
        CUDA_OK(cudaMalloc((void**)&d_started[device], 8));
        CUDA_OK(cudaStreamCreateWithFlags(&hog_stream[device], cudaStreamNonBlocking));
    }
    CUDA_OK(cudaMemsetAsync(d_started[device], 0, 8, hog_stream[device]));
    occupy_kernel<<<blocks, 1024, smem, hog_stream[device]>>>((unsigned long long)ms * 1000000ull, d_started[device]);
    CUDA_OK(cudaGetLastError());
    // wait until every block is resident, so that what is launched next finds those SMs taken
    for (int spin = 0; spin < 20000; ++spin) {
        unsigned int n = 0;
        cudaStream_t s2 = nullptr;
        CUDA_OK(pool_stream(device, &s2));
        cudaError_t e = cudaMemcpyAsync(&n, d_started[device], 4, cudaMemcpyDeviceToHost, s2);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s2);
        { std::lock_guard<std::mutex> lock(g_pool[device].mu); g_pool[device].streams.push_back(s2); }
        CUDA_OK(e);
        if ((int)n >= blocks) return 1;
        const struct timespec nap = { 0, 100 * 1000 };
        nanosleep(&nap, 0);
    }
    set_err("occupy kernel did not become resident");
    return 0;
}
