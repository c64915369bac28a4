// escape_kernel.cuh -- the per-pixel escape-time loop, MPFR-faithful mode.
//
// Restates, for one pixel per thread, the loop bodies of the reference's
// frac_mandel_mpfr (src/frac_mandel.c:25-52), frac_burning_ship_mpfr
// (src/frac_burning_ship.c:27-55), frac_generalized_celtic_mpfr
// (src/frac_generalized_celtic.c:27-55) and frac_variant_mpfr
// (src/frac_variant.c:26-55), and -- at p = 64 -- their long double twins
// (frac_mandel.c:5-21 etc.), with the per-pixel set-up of
// fractal_mpfr_calculate_line (src/fractal.c:183-203) fed from host-built
// column / row tables.
//
// Scheduling: persistent warps pull pixels from a device-side atomic queue
// (the reference hands out lines under a mutex, src/render_threads.c:360-393;
// escape times inside a line are wildly uneven, so the unit here is a pixel).
// Every `chunk` iterations a warp refills its finished lanes with one
// warp-aggregated atomicAdd.
#pragma once
#include "escape_params.cuh"
#include "escape_step.cuh"
#include "ld64_step.cuh"
#include "mpf_sf.cuh"
#include "mpf_fast.cuh"
#include "colour.cuh"

namespace mdz {


// Warp-level completion of bands: every lane passes the band it just completed (or
// -1).  All writers fenced before bumping the band counter, so after this fence the
// band's raw values are visible; they are read past L1 (__ldcg) by colour_band.
__device__ __forceinline__ bool publish_bands(const EscapeParams& p, int finished_band, unsigned lane)
{
    unsigned fin = __ballot_sync(0xffffffffu, finished_band >= 0);
    if (!fin) return false;
    __threadfence();
    while (fin) {
        const int src = __ffs(fin) - 1;
        fin &= fin - 1;
        const int band = __shfl_sync(0xffffffffu, finished_band, src);
        if (p.colour.enabled)
            colour_band(p.raw + (size_t)band * p.aa * p.width, p.width, p.aa, band, p.colour, lane);
        __syncwarp();
        __threadfence();
        if (lane == 0) {
            p.band_flag[band] = p.gen;
            atomicAdd((unsigned int*)p.bands_done, 1u);
        }
    }
    return true;
}


// ---------------------------------------------------------------------------
// The pixel queue.
//
// Static plans: position q of the queue is pixel q of the plan's own lines (raster order).
// Ordered plans (p.order != nullptr): the queue runs over *slots* of one tile each (aa lines x tile_w
// columns; a whole band when tiles_per_band is 1) and p.order[slot] says which tile that is -- a static
// plan that starts in the middle of the image (mdzcuda_plan_set_order), or a
// fed plan (several devices sharing one image, mdzcuda.cu "band scheduler"), whose slots -- whole bands --
// the host fills while the kernel runs: p.order[slot] is the band it put there, p.feed[0] the number of slots filled so
// far, p.feed[1] == p.gen once no more will come.  A lane whose claim lies beyond the filled part
// keeps it as a reservation (the index stays in `idx` with kReserved set) and looks again at the next
// refill; after the close, reservations beyond the final limit are void.  The host writes order[],
// then feed[0], then feed[1], in stream order; the device reads them in the opposite order.
// ---------------------------------------------------------------------------
constexpr unsigned kReserved = 0x80000000u;      // top bit of a lane's pixel index: it is a reservation (images hold < 2^31 pixels)

__device__ __forceinline__ bool claim_pixels(const EscapeParams& p, unsigned lane, unsigned total, bool active,
                                             unsigned& idx, bool& exhausted)
{
    bool start = false;
    unsigned limit = total;
    bool closed = true;
    bool pending = false;
    if (p.feed) {
        unsigned f0 = 0, f1 = 0;
        if (lane == 0) { f1 = p.feed[1]; __threadfence(); f0 = p.feed[0]; }
        f0 = __shfl_sync(0xffffffffu, f0, 0);
        f1 = __shfl_sync(0xffffffffu, f1, 0);
        closed = f1 == p.gen;
        limit = f0 * ((unsigned)p.tile_w * (unsigned)p.aa);
        pending = !active && (idx & kReserved) != 0u;
        if (pending) {
            const unsigned want = idx & ~kReserved;
            if (want < limit) { idx = want; start = true; pending = false; }
            else if (closed) { idx = 0u; pending = false; }
        }
    }
    if (!exhausted) {
        const unsigned need = __ballot_sync(0xffffffffu, !active && !pending && !start);
        if (need) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(p.queue, (unsigned)__popc(need));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned end = base + (unsigned)__popc(need);
            if (end >= total || (closed && end >= limit)) exhausted = true;
            if (!active && !pending && !start) {
                const unsigned i = base + (unsigned)__popc(need & ((1u << lane) - 1u));
                if (i < limit) { idx = i; start = true; }
                else if (!closed && i < total) idx = i | kReserved;
            }
        }
    }
    return start;
}

// queue position -> pixel (index into raw, line of the row table, column)
__device__ __forceinline__ void pixel_of_claim(const EscapeParams& p, unsigned idx, unsigned& pix, int& line, int& ix)
{
    if (p.order) {
        const unsigned tw = (unsigned)p.tile_w, slot_px = tw * (unsigned)p.aa;
        const unsigned slot = idx / slot_px, rem = idx - slot * slot_px;
        const unsigned code = __ldcg(&p.order[slot]);
        const unsigned band = code / (unsigned)p.tiles_per_band, tile = code - band * (unsigned)p.tiles_per_band;
        const unsigned l = rem / tw;
        line = (int)(band * (unsigned)p.aa + l);
        ix = (int)(tile * tw + rem - l * tw);
        pix = (unsigned)line * (unsigned)p.width + (unsigned)ix;
    } else {
        pix = idx;
        line = (int)(idx / (unsigned)p.width);
        ix = (int)(idx - (unsigned)line * (unsigned)p.width);
    }
}

// nobody is iterating: leave, or -- fed plan with reservations outstanding -- wait for the host
__device__ __forceinline__ bool queue_idle_wait(const EscapeParams& p, unsigned idx)
{
    if (!p.feed || !__any_sync(0xffffffffu, (idx & kReserved) != 0u)) return false;
    __nanosleep(2000);
    return true;
}

template <int N>
__device__ __forceinline__ void load_entry(const CoordTable& t, int i, Num<N>& v)
{
#pragma unroll
    for (int k = 0; k < N; ++k) v.m[k] = __ldg(&t.m[(size_t)k * t.count + i]);
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

// limb counts for which the speculative iteration (escape_step.cuh) is compiled in:
// it keeps the previous state alive for the fall-back, 4N extra registers
#ifndef MDZ_HYBRID_MIN
#define MDZ_HYBRID_MIN 11       // smallest limb count that runs the hybrid iteration (A/B builds: make EXTRA=-DMDZ_HYBRID_MIN=3)
#endif
template <int N> struct SpecLimbs { static constexpr bool value = N >= 2 && N < (N == 2 ? 3 : MDZ_HYBRID_MIN); };   // N = 2: ld64_step.cuh
// ... and for which the hybrid iteration (escape_step.cuh pixel_step_hybrid: fall-backs inside the step, no checkpoint)
template <int N> struct HybridLimbs { static constexpr bool value = N >= MDZ_HYBRID_MIN && N <= 16 && N > 2; };
// From 11 limbs up the hybrid iteration is the only one a warp runs (a second, general copy of the unrolled step next
// to it is what the instruction cache cannot hold, and the adaptive machinery alone cost 3 % at 512 bits); below that
// -- A/B builds -- a warp may give it up for the general step when most of its iterations decline.
template <int N> struct HybridAdapts { static constexpr bool value = HybridLimbs<N>::value && N <= 10; };
// ... and from where on its checkpoint lives in shared memory instead of registers
template <int N> struct SpecSmemCkpt { static constexpr bool value = N > 10 && SpecLimbs<N>::value; };
// shared-memory words per thread: c_re, c_im, limb-shifter scratch, checkpoint
template <int N> struct SmemWords {
    static constexpr int value = 2 * N + ScratchWords<N>::value + (SpecSmemCkpt<N>::value ? CkptWords<N>::value : 0);
};

// resident blocks per SM the register allocator is asked to make room for
#ifndef MDZ_MINBLOCKS_13_16
// 3 blocks of 168 registers.  4 blocks of 128 (make EXTRA=-DMDZ_MINBLOCKS_13_16=4) measured +1 ... +4 % on one GPU at 512 bits
// (13.66 -> 14.05 G it/s on the target frame) but a pixel that runs to depth then shares its scheduler with a fourth warp:
// 100 000 iterations take 0.54 s instead of 0.45, and the same frame split over eight GPUs went from 600 to 652 ms (99.6 ->
// 89 % of eight times one GPU).  Latency of the deep pixels is what bounds strong scaling, so the lower occupancy stays.
#define MDZ_MINBLOCKS_13_16 3
#endif
#ifndef MDZ_MINBLOCKS_BUMP
#define MDZ_MINBLOCKS_BUMP 0        // A/B builds: one more resident block for 5..12 limbs
#endif
template <int N> struct MinBlocks {
    static constexpr int value = N <= 2 ? 6 : N <= 3 ? 8 : N <= 4 ? 6 : N <= 5 ? 6 + MDZ_MINBLOCKS_BUMP : N <= 8 ? 5 + MDZ_MINBLOCKS_BUMP
                               : N <= 12 ? 4 + MDZ_MINBLOCKS_BUMP : N <= 16 ? MDZ_MINBLOCKS_13_16 : N <= 24 ? 2 : 1;
};

// ---------------------------------------------------------------------------
// Parking: tail compaction (EscapeParams::park_cap).
//
// The queue hands out pixels in raster order; when it runs dry every warp still holds some
// pixels that need thousands of iterations more next to lanes with nothing left to do.  When the
// image (or this GPU's share of it) is only a few times the grid, that last generation is sparse:
// on an eighth of BASELINE configs[1] 41 % of the lanes are busy, in every warp, at the pace of a
// fully occupied SM (13 ms per 10 000 iterations in long double mode, against 5.4 ms with two warps
// per scheduler).  With parking on, phase 0 stops at that point: every lane writes the complete
// state of its pixel (6N + 9 words) to HBM and the kernel ends; a one-block counting sort orders the
// entries by iteration count; phase 1 -- a second launch of this kernel, next in the stream -- reads
// the list instead of the queue, 32 pixels with about the same number of iterations left per warp, on
// as few warps as that takes, spread evenly over the SMs.  The state is carried over bit for bit, so
// raw_data does not depend on any of this.
// ---------------------------------------------------------------------------
template <int N> struct ParkWords { static constexpr int value = 6 * N + 9; };
// Compiled in for N <= 4 only (long double, MPFR to 128 bits): in the wider kernels the two extra
// paths cost registers the hot loop has no room for (N = 10: spill stores 32 -> 136 bytes, MPFR-320
// 15.1 -> 13.6 G it/s on the B200), and the deep views those precisions are for keep every lane busy
// to the end (configs[3] splits over eight GPUs at 99 %).  mdzcuda.cu parks only plans with n32 <= kParkMaxLimbs.
template <int N> struct Parkable { static constexpr bool value = N <= kParkMaxLimbs; };

template <int N>
__device__ __forceinline__ void park_store(uint32_t* col, size_t stride, const PixelState<N>& st,
                                           const uint32_t* cre_m, const uint32_t* cim_m, unsigned pix)
{
#pragma unroll
    for (int k = 0; k < N; ++k) {
        col[(size_t)(0 * N + k) * stride] = st.wre.m[k];
        col[(size_t)(1 * N + k) * stride] = st.wim.m[k];
        col[(size_t)(2 * N + k) * stride] = st.wre2.m[k];
        col[(size_t)(3 * N + k) * stride] = st.wim2.m[k];
        col[(size_t)(4 * N + k) * stride] = cre_m[k * kBlock];
        col[(size_t)(5 * N + k) * stride] = cim_m[k * kBlock];
    }
    uint32_t* t = col + (size_t)(6 * N) * stride;
    t[0 * stride] = (uint32_t)st.wre.e;  t[1 * stride] = (uint32_t)st.wim.e;
    t[2 * stride] = (uint32_t)st.wre2.e; t[3 * stride] = (uint32_t)st.wim2.e;
    t[4 * stride] = (uint32_t)st.cre_e;  t[5 * stride] = (uint32_t)st.cim_e;
    t[6 * stride] = (st.wre.s & 1u) | ((st.wim.s & 1u) << 1) | ((st.wre2.s & 1u) << 2) | ((st.wim2.s & 1u) << 3)
                  | ((st.cre_s & 1u) << 4) | ((st.cim_s & 1u) << 5);
    t[7 * stride] = (uint32_t)st.iter;
    t[8 * stride] = pix;
}

template <int N>
__device__ __forceinline__ void park_load(const uint32_t* col, size_t stride, PixelState<N>& st,
                                          uint32_t* cre_m, uint32_t* cim_m, unsigned& pix)
{
#pragma unroll
    for (int k = 0; k < N; ++k) {
        st.wre.m[k]  = __ldcg(&col[(size_t)(0 * N + k) * stride]);
        st.wim.m[k]  = __ldcg(&col[(size_t)(1 * N + k) * stride]);
        st.wre2.m[k] = __ldcg(&col[(size_t)(2 * N + k) * stride]);
        st.wim2.m[k] = __ldcg(&col[(size_t)(3 * N + k) * stride]);
        cre_m[k * kBlock] = __ldcg(&col[(size_t)(4 * N + k) * stride]);
        cim_m[k * kBlock] = __ldcg(&col[(size_t)(5 * N + k) * stride]);
    }
    const uint32_t* t = col + (size_t)(6 * N) * stride;
    st.wre.e  = (int32_t)__ldcg(&t[0 * stride]); st.wim.e  = (int32_t)__ldcg(&t[1 * stride]);
    st.wre2.e = (int32_t)__ldcg(&t[2 * stride]); st.wim2.e = (int32_t)__ldcg(&t[3 * stride]);
    st.cre_e  = (int32_t)__ldcg(&t[4 * stride]); st.cim_e  = (int32_t)__ldcg(&t[5 * stride]);
    const uint32_t f = __ldcg(&t[6 * stride]);
    st.wre.s = f & 1u; st.wim.s = (f >> 1) & 1u; st.wre2.s = (f >> 2) & 1u; st.wim2.s = (f >> 3) & 1u;
    st.cre_s = (f >> 4) & 1u; st.cim_s = (f >> 5) & 1u;
    st.iter = (int)__ldcg(&t[7 * stride]);
    pix = __ldcg(&t[8 * stride]);
}

// every active lane of the warp appends its pixel to the list (one warp-aggregated atomicAdd)
template <int N>
__device__ __forceinline__ void park_lanes(const EscapeParams& p, bool who, unsigned lane,
                                           const PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m, unsigned pix)
{
    const unsigned m = __ballot_sync(0xffffffffu, who);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&p.park_count[0], (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (who) park_store<N>(p.park_buf + base + (unsigned)__popc(m & ((1u << lane) - 1u)), (size_t)p.park_cap, st, cre_m, cim_m, pix);
}

// ---------------------------------------------------------------------------
// Exact periodicity check (optional, EscapeParams::cycle).  The recurrence is
// deterministic, so if (wre, wim) at iteration j equals, bit for bit, the state saved
// at an earlier iteration i (same parity of j - i for the hybrid fractal, whose step
// depends on the iteration number's parity, src/frac_variant.c:42-43), the orbit repeats
// the iterations i..j for ever; none of them escaped, so none ever will, and the
// reference's loop would run to `depth` and return 0.  The pixel is finished with 0
// right there: identical raw_data, without the remaining iterations.  Brent's schedule:
// the state is saved at iterations 1, 2, 4, 8, ...  The saved state lives in a
// per-thread column of global memory (touched ~log2(depth) times per pixel); two
// filter words stay in registers, so an iteration pays two compares.
// ---------------------------------------------------------------------------
struct CycleState {
    uint32_t f0, f1;        // wre.m[0], wim.m[0] of the saved state
    int next;               // iteration at which to save again
};

template <int N>
__device__ __forceinline__ void cycle_save(const EscapeParams& p, const PixelState<N>& st, CycleState& cs)
{
    uint32_t* col = p.cycle_scratch + (size_t)blockIdx.x * kBlock + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * kBlock;
#pragma unroll
    for (int k = 0; k < N; ++k) { col[(size_t)k * stride] = st.wre.m[k]; col[(size_t)(N + k) * stride] = st.wim.m[k]; }
    col[(size_t)(2 * N) * stride] = (uint32_t)st.wre.e;
    col[(size_t)(2 * N + 1) * stride] = (uint32_t)st.wim.e;
    col[(size_t)(2 * N + 2) * stride] = st.wre.s | (st.wim.s << 1) | ((uint32_t)(st.iter & 1) << 2);
    cs.f0 = st.wre.m[0]; cs.f1 = st.wim.m[0];
    cs.next = st.iter < (1 << 30) ? (st.iter > 0 ? st.iter * 2 : 1) : 0x7fffffff;
}

// full comparison after the filter words matched
template <int N>
__device__ __noinline__ bool cycle_match(const EscapeParams& p, const PixelState<N>& st)
{
    const uint32_t* col = p.cycle_scratch + (size_t)blockIdx.x * kBlock + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * kBlock;
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        diff |= col[(size_t)k * stride] ^ st.wre.m[k];
        diff |= col[(size_t)(N + k) * stride] ^ st.wim.m[k];
    }
    diff |= col[(size_t)(2 * N) * stride] ^ (uint32_t)st.wre.e;
    diff |= col[(size_t)(2 * N + 1) * stride] ^ (uint32_t)st.wim.e;
    uint32_t tag = st.wre.s | (st.wim.s << 1) | ((uint32_t)(st.iter & 1) << 2);
    uint32_t dt = col[(size_t)(2 * N + 2) * stride] ^ tag;
    if (p.fractal != FRACTAL_VARIANT) dt &= 3u;        // parity only matters for the hybrid
    return (diff | dt) == 0;
}

// One chunk of the long double mode's event-driven loop (ld64_step.cuh).  The iteration is one
// branch-free block that every lane runs, finished lanes included (their results are ignored); the
// warp leaves that block only when some lane has an event -- escaped, reached depth, or met a case
// the fast step declines -- so an interior pixel's ten thousand iterations cost one vote and one
// branch each on top of the arithmetic.  WIDE: level 2, the additions of add64_core<true>.
template <bool CYC, bool WIDE>
__device__ __forceinline__ void ld64_event_chunk(const EscapeParams& p, PixelState<2>& st, uint32_t* cre_m, uint32_t* cim_m,
                                                 uint32_t* scr, bool& active, const unsigned& pix, int& finished_band,
                                                 CycleState& cyc, uint32_t& rare_seen, int& warp_steps, int& warp_fell,
                                                 const unsigned lane, const bool abs_im, const int abs_re)
{
    Num<2> cre, cim;                         // c stays in registers across the chunk
    cim.m[0] = cim_m[0]; cim.m[1] = cim_m[kScratchStride]; cim.e = st.cim_e; cim.s = st.cim_s;
    cre.m[0] = cre_m[0]; cre.m[1] = cre_m[kScratchStride]; cre.e = st.cre_e; cre.s = st.cre_s;
    PixelState<2> nx;
    const Ld64Masks mk = p.ld_masks;
    for (int k = 0; k < p.chunk; ++k) {
        // two iterations per trip, st -> nx -> st, so that no state is copied back
        bool rare = false;
        bool esc = ld64_step<WIDE>(st, nx, cre, cim, scr, p.rc, mk, rare);
        bool cyc_ev = CYC && ((nx.wre.m[0] == cyc.f0 && nx.wim.m[0] == cyc.f1) || nx.iter == cyc.next);
        bool ev = active && (rare || esc || nx.iter >= p.depth || cyc_ev);
        warp_steps += 1;
        if (!__any_sync(0xffffffffu, ev)) {
            if (++k >= p.chunk) { st = nx; break; }
            rare = false;
            esc = ld64_step<WIDE>(nx, st, cre, cim, scr, p.rc, mk, rare);
            cyc_ev = CYC && ((st.wre.m[0] == cyc.f0 && st.wim.m[0] == cyc.f1) || st.iter == cyc.next);
            ev = active && (rare || esc || st.iter >= p.depth || cyc_ev);
            warp_steps += 1;
            if (!__any_sync(0xffffffffu, ev)) continue;
            // event in the second half: present it to the handler as (old = st, new = nx)
            const PixelState<2> tmp = st; st = nx; nx = tmp;
        }
        warp_fell += __any_sync(0xffffffffu, active && rare) ? 1 : 0;
        if (active) {
            if (rare) { rare_seen += 1; esc = pixel_step<2>(st, cre_m, cim_m, scr, p.rc, abs_im, abs_re); }
            else st = nx;
            bool periodic = false;
            if (CYC && !esc) {
                if (st.wre.m[0] == cyc.f0 && st.wim.m[0] == cyc.f1) periodic = cycle_match<2>(p, st);
                if (!periodic && st.iter >= cyc.next) cycle_save<2>(p, st, cyc);
            }
            if (esc || periodic || st.iter >= p.depth) {
                p.raw[pix] = esc ? st.iter : 0;
                __threadfence();
                active = false;
                const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
            }
        }
        if (publish_bands(p, finished_band, lane)) finished_band = -1;
        if (!__any_sync(0xffffffffu, active)) break;
    }
}

// CYC: compiled with the exact periodicity check.  A separate instantiation, because the
// extra state and cold paths cost the plain kernels 7-15 % when merely present.
template <int N, bool CYC>
__global__ void __launch_bounds__(kBlock, MinBlocks<N>::value)
escape_mpfr_kernel(const EscapeParams p)
{
    // per-thread copy of c (2N limbs), limb-major: conflict-free
    extern __shared__ uint32_t csm[];
    uint32_t* cre_m = csm + threadIdx.x;
    uint32_t* cim_m = csm + N * kBlock + threadIdx.x;
    // limb-shifter scratch column (mpfr_sf.cuh shift_right_far): upper half stays zero
    uint32_t* scr = csm + 2 * N * kBlock + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ScratchWords<N>::value; ++k) scr[k * kBlock] = 0u;
    static_assert(kScratchStride == kBlock, "scratch stride must equal the block size");
    uint32_t* ckpt = csm + (2 * N + ScratchWords<N>::value) * kBlock + threadIdx.x;   // only used when SpecSmemCkpt<N>

    const unsigned lane = threadIdx.x & 31u;
    constexpr bool kPark = Parkable<N>::value;
    const bool phase1 = kPark && p.phase != 0;
    const unsigned total = phase1 ? __ldcg(&p.park_count[0]) : (unsigned)p.width * (unsigned)p.lines;
    const bool parking = kPark && p.park_cap != 0u && p.phase == 0;

    PixelState<N> st;
    set_zero(st.wre); set_zero(st.wim); set_zero(st.wre2); set_zero(st.wim2);
    st.cre_e = E_ZERO; st.cim_e = E_ZERO; st.cre_s = 0; st.cim_s = 0; st.iter = 0;

    bool active = false;
    int finished_band = -1;
    bool exhausted = false;         // warp-uniform
    bool first_claim = true;        // phase 1: the first group is chosen by position on the SM
    CycleState cyc; cyc.f0 = 0; cyc.f1 = 0; cyc.next = 0x7fffffff;
    // warp-uniform: 0 general step, 1 speculative, 2 (long double mode only) speculative with the level-2 additions
    // (p.spec: 0 off, 1 adaptive; 2 / 3 pin level 1 / 2 for A/B measurements)
    int spec_level = ((SpecLimbs<N>::value || HybridLimbs<N>::value) && p.spec != 0) ? (p.spec == 3 ? 2 : 1) : 0;
    int spec_pause = 0, spec_backoff = 8;
    unsigned pix = 0;

    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;

    for (;;) {
        // ---- cooperative cancel (reference polls every 64 px, fractal.c:113) --
        {
            unsigned ctl = 0;
            if (lane == 0) {
                ctl = *p.cancel == p.gen ? 1u : 0u;
                if (parking && *(volatile unsigned int*)p.queue >= total) ctl |= 2u;
            }
            ctl = __shfl_sync(0xffffffffu, ctl, 0);
            if (ctl & 1u) break;
            if constexpr (kPark) if (ctl & 2u) {
                // the queue is dry: hand what is still in flight to phase 1
                park_lanes<N>(p, active, lane, st, cre_m, cim_m, pix);
                break;
            }
        }
        // ---- refill finished lanes from the queue ------------------------
        if (phase1) { if constexpr (kPark) {
            // Phase 1: a warp takes 32 consecutive entries of the sorted list at a time and the next 32
            // once all of them are done.  A short list must not end up on the SMs whose blocks happened
            // to start first, so the first claim goes by position: the k-th warp to arrive on the d-th SM
            // asks for group k * SMs + d (SM ids need not be dense: the first warp to arrive on an SM
            // gives it the next number).  Groups nobody asked for are picked up, after a pause, by the
            // warps whose claim was beyond the list; claimed[] makes every group go out once.
            if (!exhausted && !__any_sync(0xffffffffu, active)) {
                const unsigned groups = (total + 31u) >> 5;
                unsigned g = 0xffffffffu;
                if (lane == 0) {
                    if (first_claim) {
                        unsigned smid;
                        asm("mov.u32 %0, %%smid;" : "=r"(smid));
                        smid &= 1023u;
                        // [0..1023] arrivals per SM id, [1024..2047] number + 1 of that SM, [2048] next number
                        const unsigned k = atomicAdd(&p.park_smslot[smid], 1u);
                        volatile unsigned int* number = p.park_smslot + 1024 + smid;
                        unsigned d;
                        if (k == 0u) { d = atomicAdd(&p.park_smslot[2048], 1u) + 1u; *number = d; }
                        else while ((d = *number) == 0u) { }           // written by a warp that is already running
                        const unsigned want = k * (unsigned)p.park_sms + (d - 1u);
                        if (want < groups && atomicExch(&p.park_claimed[want], 1u) == 0u) g = want;
                        else __nanosleep(20000);
                    }
                    while (g == 0xffffffffu) {
                        const unsigned q = atomicAdd(&p.park_count[1], 1u);
                        if (q >= groups) break;
                        if (atomicExch(&p.park_claimed[q], 1u) == 0u) g = q;
                    }
                }
                first_claim = false;
                g = __shfl_sync(0xffffffffu, g, 0);
                if (g == 0xffffffffu) exhausted = true;
                else {
                    const unsigned idx = (g << 5) + lane;
                    if (idx < total) {
                        park_load<N>(p.park_buf + __ldg(&p.park_perm[idx]), (size_t)p.park_cap, st, cre_m, cim_m, pix);
                        active = true;
                        if (CYC) cycle_save<N>(p, st, cyc);
                    }
                }
            }
        } } else
        if (!exhausted || (p.feed && __any_sync(0xffffffffu, !active && (pix & kReserved) != 0u))) {
            if (claim_pixels(p, lane, total, active, pix, exhausted)) {
                int line, ix;
                pixel_of_claim(p, pix, pix, line, ix);
                Num<N> x, y, cx, cy;
                load_entry<N>(p.xs, ix, x);
                load_entry<N>(p.ys, line, y);
                if (p.family == FAMILY_JULIA) {
                    load_entry<N>(p.jc, 0, cx);
                    load_entry<N>(p.jc, 1, cy);
                } else { cx = x; cy = y; }
                pixel_init<N>(st, x, y, cx, cy, p.rc, cre_m, cim_m);
                active = true;
                if (CYC) cycle_save<N>(p, st, cyc);
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (queue_idle_wait(p, pix)) continue; break; }

        // ---- iterate ------------------------------------------------------
        uint32_t rare_seen = 0;
        int warp_steps = 0, warp_fell = 0;       // iterations of this chunk, and those in which some lane fell back
        bool event_loop = false;
        if constexpr (N == 2) event_loop = spec_level != 0 && p.rc.ulp == 1u;   // the 64-bit step needs p = 64 exactly
        if (event_loop) {
            if constexpr (N == 2) {
                if (spec_level == 2)
                    ld64_event_chunk<CYC, true>(p, st, cre_m, cim_m, scr, active, pix, finished_band, cyc,
                                                rare_seen, warp_steps, warp_fell, lane, abs_im, abs_re);
                else
                    ld64_event_chunk<CYC, false>(p, st, cre_m, cim_m, scr, active, pix, finished_band, cyc,
                                                 rare_seen, warp_steps, warp_fell, lane, abs_im, abs_re);
            }
        } else
        for (int k = 0; k < p.chunk; ++k) {
            const uint32_t rare_before = rare_seen;
            if (active) {
                bool esc;
                if (SpecLimbs<N>::value)
                    esc = pixel_step_auto<N, SpecSmemCkpt<N>::value>(st, cre_m, cim_m, scr, ckpt, p.rc, abs_im, abs_re, spec_level, rare_seen);
                else if (HybridLimbs<N>::value && spec_level != 0)
                    esc = pixel_step_hybrid<N>(st, cre_m, cim_m, scr, p.rc, abs_im, abs_re, rare_seen);
                else
                    esc = pixel_step<N>(st, cre_m, cim_m, scr, p.rc, abs_im, abs_re);
                const int iter = st.iter;
                bool periodic = false;
                if (CYC && !esc) {
                    if (st.wre.m[0] == cyc.f0 && st.wim.m[0] == cyc.f1) periodic = cycle_match<N>(p, st);
                    if (!periodic && iter >= cyc.next) cycle_save<N>(p, st, cyc);
                }
                if (esc || periodic || iter >= p.depth) {
                    p.raw[pix] = esc ? iter : 0;
                    __threadfence();            // raw visible before the band counter moves
                    active = false;
                    const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                    const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                    if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
                }
            }
            // one lane falling back makes the whole warp wait for the general step
            if ((SpecLimbs<N>::value || HybridAdapts<N>::value) && spec_level != 0) {
                warp_steps += 1;
                warp_fell += __any_sync(0xffffffffu, rare_seen != rare_before) ? 1 : 0;
            }
            // a lane that completed a band hands it to the whole warp: colour it (fused
            // epilogue), then publish the band flag the host polls
            if (publish_bands(p, finished_band, lane)) finished_band = -1;
            if (!__any_sync(0xffffffffu, active)) break;
        }
        // ---- adapt: speculation is only worth it while fall-backs are scarce ----
        // One lane that falls back makes its whole warp run the general step as well, and an event that a
        // lane meets once in three hundred iterations a warp of unsynchronised lanes meets in every tenth.
        if ((SpecLimbs<N>::value || HybridAdapts<N>::value) && p.spec == 1) {
            if constexpr (N == 2) {
                // Long double mode, three levels (ld64_step.cuh): when more than a quarter of a chunk's
                // iterations had a fall-back, level 1 gives way to level 2 (far smaller or zero second
                // operand, 32..62 cancelled bits; ~10 instructions more per addition) and level 2 to the
                // general step, each for an exponentially growing number of chunks (16 ... 4096) before the
                // cheaper one is retried.
                if (spec_level != 0) {
                    if (warp_fell * 4 > warp_steps) {
                        spec_backoff = spec_backoff < 4096 ? spec_backoff * 2 : 4096;
                        spec_pause = spec_backoff;
                        spec_level = spec_level == 1 ? 2 : 0;
                    } else if (spec_level == 2) {
                        if (--spec_pause <= 0) spec_level = 1;          // see whether the narrow one will do again
                    } else if (warp_fell == 0) spec_backoff = 8;
                } else if (--spec_pause <= 0) {
                    spec_level = 1;
                }
            } else {
                // Multi-limb kernels, two modes: the hybrid iteration (escape_step.cuh pixel_step_hybrid: speculative
                // additions, redone with the general fadd when one declines) or the general step.  Trying costs the
                // speculative additions (~6N instructions each against ~10N), so it pays while fewer than about a third
                // of a warp's iterations decline; judged over windows of 64 iterations, the general step then stays for
                // an exponentially growing number of chunks (16 ... 4096) before the warp tries again.  (Kernels built
                // with the older whole-iteration fall-back, pixel_step_auto -- A/B builds with MDZ_HYBRID_MIN above the
                // limb count -- pay a full general step per fall-back and give up at 1/32.)
                // (one register: spec_pause counts the window's iterations in its low half and its fall-backs in
                // the high half while speculating, and the chunks left to sit out while not)
                if (spec_level != 0) {
                    spec_pause += warp_steps + (warp_fell << 16);
                    if ((spec_pause & 0xffff) >= 64) {
                        const int fell = spec_pause >> 16, steps = spec_pause & 0xffff;
                        spec_pause = 0;
                        if (HybridLimbs<N>::value ? fell * 3 > steps : fell * 32 > steps) {
                            spec_backoff = spec_backoff < 4096 ? spec_backoff * 2 : 4096;
                            spec_pause = spec_backoff;
                            spec_level = 0;
                        } else if (fell == 0) spec_backoff = 8;
                    }
                } else if (--spec_pause <= 0) {
                    spec_level = 1;
                    spec_pause = 0;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------
// GMP mpf mode (reference src/fractal.c:260-397 + frac_*_gmp): same persistent
// scheduler, arithmetic from mpf_sf.cuh.  Tables hold NL 64-bit limbs per entry
// as 2*NL 32-bit words (low word first), the limb exponent, the sign.
// ---------------------------------------------------------------------------
template <int NL>
__device__ __forceinline__ void load_mpf_entry(const CoordTable& t, int i, Mpf<NL>& v)
{
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const uint32_t lo = __ldg(&t.m[(size_t)(2 * k) * t.count + i]);
        const uint32_t hi = __ldg(&t.m[(size_t)(2 * k + 1) * t.count + i]);
        v.l[k] = ((uint64_t)hi << 32) | lo;
    }
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

template <int NL>
__global__ void __launch_bounds__(kBlock)
escape_gmp_kernel(const EscapeParams p)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)p.width * (unsigned)p.lines;
    GmpPixel<NL> st;
    st.iter = 0;
    bool active = false;
    int finished_band = -1;
    bool exhausted = false;
    unsigned pix = 0;
    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;
    for (;;) {
        {
            int stop = 0;
            if (lane == 0) stop = *p.cancel == p.gen;
            if (__shfl_sync(0xffffffffu, stop, 0)) break;
        }
        if (!exhausted || (p.feed && __any_sync(0xffffffffu, !active && (pix & kReserved) != 0u))) {
            if (claim_pixels(p, lane, total, active, pix, exhausted)) {
                int line, ix;
                pixel_of_claim(p, pix, pix, line, ix);
                Mpf<NL> x, y, cx, cy;
                load_mpf_entry<NL>(p.xs, ix, x);
                load_mpf_entry<NL>(p.ys, line, y);
                if (p.family == FAMILY_JULIA) {
                    load_mpf_entry<NL>(p.jc, 0, cx);
                    load_mpf_entry<NL>(p.jc, 1, cy);
                } else { cx = x; cy = y; }
                gmp_pixel_init<NL>(st, x, y, cx, cy);
                active = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (queue_idle_wait(p, pix)) continue; break; }
        for (int k = 0; k < p.chunk; ++k) {
            if (active) {
                const bool esc = gmp_pixel_step<NL>(st, abs_im, abs_re);
                if (esc || st.iter >= p.depth) {
                    p.raw[pix] = esc ? st.iter : 0;
                    __threadfence();
                    active = false;
                    const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                    const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                    if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
                }
            }
            // a lane that completed a band hands it to the whole warp: colour it (fused
            // epilogue), then publish the band flag the host polls
            if (publish_bands(p, finished_band, lane)) finished_band = -1;
            if (!__any_sync(0xffffffffu, active)) break;
        }
    }
}


// ---------------------------------------------------------------------------
// GMP mpf mode, fast kernel: mpf_fast.cuh (register-resident 32-bit words,
// IMAD.WIDE high products, shared-memory limb shifter).  NW = 2*(P+1) words.
// Shared memory per thread: c_re, c_im (2*NW) + shifter column (3*NW).
// ---------------------------------------------------------------------------
template <int NW> struct GSmemWords { static constexpr int value = 2 * NW + GScratchWords<NW>::value; };
// (measured on the B200: 14 words -- 320 bits -- 15.1 -> 16.4 G it/s with 4 blocks of 128 registers instead of 3 of 168;
// 20 words -- 512 bits -- 8.9 -> 8.0, the spills outweigh the fourth block)
template <int NW> struct GMinBlocks { static constexpr int value = NW <= 8 ? 6 : NW <= 14 ? 4 : NW <= 20 ? 3 : NW <= 24 ? 2 : 1; };

template <int NW>
__device__ __forceinline__ void load_gf_entry(const CoordTable& t, int i, GF<NW>& v)
{
#pragma unroll
    for (int k = 0; k < NW; ++k) v.m[k] = __ldg(&t.m[(size_t)k * t.count + i]);
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

template <int NW>
__global__ void __launch_bounds__(kBlock, GMinBlocks<NW>::value)
escape_gmpf_kernel(const EscapeParams p)
{
    extern __shared__ uint32_t csm[];
    uint32_t* cre_m = csm + threadIdx.x;
    uint32_t* cim_m = csm + NW * kBlock + threadIdx.x;
    uint32_t* scr = csm + 2 * NW * kBlock + threadIdx.x;
#pragma unroll
    for (int k = 0; k < GScratchWords<NW>::value; ++k) scr[k * kBlock] = 0u;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)p.width * (unsigned)p.lines;
    GFPixel<NW> st;
    st.iter = 0;
    bool active = false;
    int finished_band = -1;
    bool exhausted = false;
    unsigned pix = 0;
    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;
    for (;;) {
        {
            int stop = 0;
            if (lane == 0) stop = *p.cancel == p.gen;
            if (__shfl_sync(0xffffffffu, stop, 0)) break;
        }
        if (!exhausted || (p.feed && __any_sync(0xffffffffu, !active && (pix & kReserved) != 0u))) {
            if (claim_pixels(p, lane, total, active, pix, exhausted)) {
                int line, ix;
                pixel_of_claim(p, pix, pix, line, ix);
                GF<NW> x, y, cx, cy;
                load_gf_entry<NW>(p.xs, ix, x);
                load_gf_entry<NW>(p.ys, line, y);
                if (p.family == FAMILY_JULIA) {
                    load_gf_entry<NW>(p.jc, 0, cx);
                    load_gf_entry<NW>(p.jc, 1, cy);
                } else { cx = x; cy = y; }
                gf_pixel_init<NW>(st, x, y, cx, cy, cre_m, cim_m);
                active = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (queue_idle_wait(p, pix)) continue; break; }
        for (int k = 0; k < p.chunk; ++k) {
            if (active) {
                const bool esc = gf_pixel_step<NW>(st, cre_m, cim_m, scr, abs_im, abs_re);
                if (esc || st.iter >= p.depth) {
                    p.raw[pix] = esc ? st.iter : 0;
                    __threadfence();
                    active = false;
                    const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                    const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                    if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
                }
            }
            if (publish_bands(p, finished_band, lane)) finished_band = -1;
            if (!__any_sync(0xffffffffu, active)) break;
        }
    }
}

}  // namespace mdz
