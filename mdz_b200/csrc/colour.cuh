// colour.cuh -- the colour epilogue on the device: iteration counts -> packed RGB.
//
// Restates reference src/palette.c:406-463 (palette_apply, aa_factor == 1),
// src/palette.c:466-503 (get_pixel_colour) and src/render.c:108-158
// (do_anti_aliasing), including their quirks: the two paths convert to integer
// differently (int vs guint32), the AA path adds pal_offset only when the scaled
// value is non-zero and paints "inside" with palette[0] when interpolating, and
// every index is taken modulo pal_indexes - 1.  The host code is x86-64 without
// FMA contraction, so every double operation here is an explicit round-to-nearest
// multiply or add (__dmul_rn / __dadd_rn), never a fused one.
// Defined for scaled values below 2^31 (the reference's own conversions are
// undefined beyond that).
#pragma once
#include <stdint.h>

namespace mdz {

struct ColourParams {
    const uint32_t* palette;    // device copy, up to 256 entries, R | G<<8 | B<<16 (palette.h:16)
    uint32_t* rgb;              // [bands][user_width]
    double scale;               // img->colour_scale
    int pal_indexes;
    int pal_offset;
    int interpolate;            // img->palette_ip
    int enabled;
};

__device__ __forceinline__ uint32_t c_red(uint32_t c)   { return c & 0xffu; }
__device__ __forceinline__ uint32_t c_green(uint32_t c) { return (c >> 8) & 0xffu; }
__device__ __forceinline__ uint32_t c_blue(uint32_t c)  { return (c >> 16) & 0xffu; }
__device__ __forceinline__ uint32_t c_rgb(uint32_t r, uint32_t g, uint32_t b) { return r | (g << 8) | (b << 16); }

// (guint32)double as gcc does it on x86-64: truncate to 64 bits, keep the low 32
__device__ __forceinline__ uint32_t to_u32(double v) { return (uint32_t)(unsigned long long)__double2ll_rz(v); }
// (guint8)double
__device__ __forceinline__ uint32_t to_u8(double v) { return (uint32_t)__double2int_rz(v) & 0xffu; }

__device__ __forceinline__ uint32_t mix(double diff, double rdiff, uint32_t c1, uint32_t c2)
{
    const uint32_t r = to_u8(__dadd_rn(__dmul_rn(diff, (double)c_red(c1)),   __dmul_rn(rdiff, (double)c_red(c2))));
    const uint32_t g = to_u8(__dadd_rn(__dmul_rn(diff, (double)c_green(c1)), __dmul_rn(rdiff, (double)c_green(c2))));
    const uint32_t b = to_u8(__dadd_rn(__dmul_rn(diff, (double)c_blue(c1)),  __dmul_rn(rdiff, (double)c_blue(c2))));
    return c_rgb(r, g, b);
}

// palette.c:466-503
__device__ __forceinline__ uint32_t get_pixel_colour(double val, const ColourParams& cp)
{
    if (val != 0.0) val = __dadd_rn(val, (double)cp.pal_offset);
    const uint32_t m = (uint32_t)(cp.pal_indexes - 1);
    if (cp.interpolate) {
        const double cval = ceil(val);
        const double diff = __dadd_rn(cval, -val);
        const double rdiff = __dadd_rn(1.0, -diff);
        const uint32_t ind1 = to_u32(floor(val)) % m;
        const uint32_t ind2 = to_u32(cval) % m;
        return mix(diff, rdiff, cp.palette[ind1], cp.palette[ind2]);
    }
    if (val == 0.0) return 0u;
    return cp.palette[to_u32(val) % m];
}

// palette.c:406-463, one pixel of an aa_factor == 1 image
__device__ __forceinline__ uint32_t palette_apply_px(int raw, const ColourParams& cp)
{
    if (!raw) return 0u;
    const double val = __dadd_rn(__dmul_rn((double)raw, cp.scale), (double)cp.pal_offset);
    if (cp.interpolate) {
        const int cval = __double2int_rz(ceil(val));
        const double diff = __dadd_rn((double)cval, -val);
        const double rdiff = __dadd_rn(1.0, -diff);
        const int m = cp.pal_indexes - 1;
        const int ind1 = __double2int_rz(floor(val)) % m;
        const int ind2 = cval % m;
        return mix(diff, rdiff, cp.palette[ind1], cp.palette[ind2]);
    }
    return cp.palette[to_u32(val) % (uint32_t)(cp.pal_indexes - 1)];
}

// Colour one band (aa lines of `width` supersamples -> one line of width/aa pixels)
// with the 32 lanes of a warp.  raw points at the band's first supersample.
__device__ __forceinline__ void colour_band(const int* raw, int width, int aa, int band,
                                            const ColourParams& cp, unsigned lane)
{
    const int uw = width / aa;
    uint32_t* out = cp.rgb + (size_t)band * uw;
    if (aa == 1) {
        for (int x = (int)lane; x < uw; x += 32) out[x] = palette_apply_px(__ldcg(raw + x), cp);
        return;
    }
    const uint32_t aaaa = (uint32_t)(aa * aa);
    for (int x = (int)lane; x < uw; x += 32) {
        uint32_t r = 0, g = 0, b = 0;
        for (int yi = 0; yi < aa; ++yi) {
            const int* row = raw + (size_t)yi * width + (size_t)x * aa;
            for (int xi = 0; xi < aa; ++xi) {
                const uint32_t c = get_pixel_colour(__dmul_rn((double)__ldcg(row + xi), cp.scale), cp);
                r += c_red(c); g += c_green(c); b += c_blue(c);
            }
        }
        out[x] = c_rgb(r / aaaa, g / aaaa, b / aaaa);
    }
}

// recolour-only kernel (palette cycling from resident raw_data): one warp per band
static __global__ void recolour_kernel(const int* raw, int width, int aa, int bands, ColourParams cp)
{
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    for (int band = blockIdx.x * warps_per_block + (threadIdx.x >> 5); band < bands; band += gridDim.x * warps_per_block)
        colour_band(raw + (size_t)band * aa * width, width, aa, band, cp, lane);
}

}  // namespace mdz
