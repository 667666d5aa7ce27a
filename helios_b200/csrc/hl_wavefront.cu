// hl_wavefront.cu — the wavefront path integrator that replaces vkCmdTraceRaysKHR on the reference's
// recursive RT pipeline (PathIntegrator::launch_rays, src/engine/gfx/path_integrator.cpp:125-200).
//
// One frame = generate, then per bounce { extend, shade, connect }, then resolve:
//   generate  path_trace_rgen.glsl:180-215   RNG seed, thin-lens primary ray, payload init
//   extend    traceRayEXT (rgen:205, rchit:523) + path_trace_rahit.glsl   closest hit per extension ray
//   shade     path_trace_rchit.glsl:542-580 / path_trace_rmiss.glsl:60-65  NEE sample, BRDF sample, RR,
//             queue compaction with warp ballot + one atomic per warp
//   connect   traceRayEXT (rchit:438) + path_trace_shadow.{rchit,rmiss}     visibility, L += direct term
//   resolve   rgen:217-248 (+ tone_map.frag when fused)                      clamp, progressive blend
// No host synchronisation inside a frame: queue sizes live in device memory, extend/connect are persistent
// kernels that pull 32 rays per warp from an atomic cursor, shade is a grid-stride loop.
#include "hl_bvh.h"
#include "hl_camera.h"
#include "hl_debug.h"
#include "hl_film.h"
#include "hl_internal.h"
#include "hl_shade.h"

namespace hl
{
// register cap of the persistent trace kernels.  Measured on configs[1] (tools/tune_trace.py, set "occ"): the
// compiler's own choice under __launch_bounds__(128) — 72 registers, 7 CTAs/SM — 1.81 ms/frame; capped to 64
// (8 CTAs/SM, spills) 1.89 ms; uncapped (113 registers, 4 CTAs/SM) 2.21 ms.
#ifdef HL_TRACE_MIN_BLOCKS
#define HL_TRACE_BOUNDS __launch_bounds__(HL_TRACE_BLOCK, HL_TRACE_MIN_BLOCKS)
#else
#define HL_TRACE_BOUNDS __launch_bounds__(HL_TRACE_BLOCK)
#endif
#ifndef HL_TRACE_GRID_MULT
#define HL_TRACE_GRID_MULT 6 /* persistent trace kernels: CTAs launched per SM (7 fit; measured with four frames in flight, tools/tune_trace.py set "grid": 4 / 5 / 6 / 7 / 8 / 12 give 1.461 / 1.431 / 1.444 / 1.489 / 1.468 / 1.479 ms on configs[1] and 4.69 / 4.47 / 4.22 / 4.31 / 4.34 / 4.34 ms on configs[2]: one CTA slot per SM left to the other frames' kernels) */
#endif
#ifndef HL_SHADE_BLOCK
#define HL_SHADE_BLOCK 128
#endif
#ifdef HL_SHADE_MIN_BLOCKS
#define HL_SHADE_BOUNDS __launch_bounds__(HL_SHADE_BLOCK, HL_SHADE_MIN_BLOCKS)
#else
#define HL_SHADE_BOUNDS __launch_bounds__(HL_SHADE_BLOCK)
#endif
#ifndef HL_SHADE_GRID_MULT
#define HL_SHADE_GRID_MULT 8
#endif

struct FrameParams
{
    hl_push_constants pc;
    uint32_t          lw, lh; // launch rectangle (already clipped to the image)
    uint32_t          tiled;  // 1: path i covers pixel launch_pixel(i) in 8 x 4 tiles (lw % 8 == 0 and lh % 4 == 0), 0: row-major
};
// Path index -> pixel of the launch rectangle.  A warp's 32 consecutive paths cover an 8 x 4 pixel tile instead of a 32 x 1
// strip whenever the rectangle allows it: the primary rays of a warp (and, because the queues are compacted in order, the
// rays their paths spawn) stay close in BOTH image dimensions, so they visit the same nodes and the same materials.
// The image does not depend on the mapping: every pixel's sample is a function of (pixel, frame index) alone.
#ifndef HL_TILE_PATHS
#define HL_TILE_PATHS 1
#endif
__host__ __device__ __forceinline__ void launch_pixel(uint32_t lw, uint32_t tiled, uint32_t i, uint32_t& x, uint32_t& y)
{
    if (tiled)
    {
        const uint32_t tile = i >> 5, w = i & 31u, tiles_x = lw >> 3;
        x = (tile % tiles_x) * 8u + (w & 7u), y = (tile / tiles_x) * 4u + (w >> 3);
    }
    else
        x = i % lw, y = i / lw;
}
static inline uint32_t launch_is_tiled(uint32_t lw, uint32_t lh) { return HL_TILE_PATHS && lw % 8u == 0u && lh % 4u == 0u ? 1u : 0u; }

__device__ __forceinline__ float4 ld4(const float4* p) { return *p; }

// shared memory of a kernel that traces: HL_STACK_FAST stack entries per thread + the pair table of the warp-cooperative
// triangle phase (hl_bvh.h coop_triangles) per warp
#define HL_TRACE_SHARED(st)                                                                   \
    __shared__ uint8_t pair_mem[HL_COOP_TABLE * (HL_TRACE_BLOCK / 32)];                       \
    u2                 spill_mem[HL_STACK_SPILL];                                             \
    TravStack          st;                                                                    \
    st.init(spill_mem, nullptr), st.pairs = pair_mem + HL_COOP_TABLE * (threadIdx.x >> 5)

// ---- generate --------------------------------------------------------------------------------------
__global__ void k_generate(FrameParams fp, float4* state_a, float4* state_b, float4* ext_o, float4* ext_d, uint32_t* counters)
{
    const uint32_t n = fp.lw * fp.lh;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counters[CTR_EXT_COUNT] = n;
    if (i >= n) return;
    uint32_t lx, ly;
    launch_pixel(fp.lw, fp.tiled, i, lx, ly);
    const uint32_t px = fp.pc.launch_id_size[0] + lx, py = fp.pc.launch_id_size[1] + ly;
    Rng            rng = rng_seed(px, py, fp.pc.num_frames);
    f3             o, d;
    primary_ray(fp.pc, px, py, rng, o, d);
    state_a[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng.x));
    state_b[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(rng.y));
    ext_o[i]   = make_float4(o.x, o.y, o.z, __uint_as_float(i));
    ext_d[i]   = make_float4(d.x, d.y, d.z, 0.0f);
}

// ---- persistent trace loop -------------------------------------------------------------------------
// Shared by extend and connect.  Every lane owns one traversal (hl_bvh.h Trav); the loop condition is a warp
// vote, and whenever at least HL_REFILL_MIN lanes have finished their ray the warp takes that many new rays
// from the queue's atomic cursor in one transaction (ballot + popc prefix), so lanes do not idle until the
// warp's slowest ray is done.  Results are stored by ray index: the output does not depend on which lane
// traced which ray.  Q supplies load(i, ...) and done(i, hit).
#ifndef HL_REFILL_MIN
#define HL_REFILL_MIN 8
#endif
template <class Q>
__device__ __forceinline__ void trace_queue(const SceneView& s, const Q& q, uint32_t count, uint32_t* fetch, uint32_t flags, TravStack& st)
{
    const uint32_t lane = threadIdx.x & 31u;
    Trav           t;
    const TravStart start = trav_start(s);
    trav_begin(start, t, st, false, mk3(0.0f), 0.0f, mk3(0.0f), 0.0f, flags);
    uint32_t mine      = 0xFFFFFFFFu; // index of the ray this lane is tracing
    bool     exhausted = false;       // warp-uniform: the cursor ran past the end of the queue
    for (;;)
    {
        bool busy = trav_busy(t, st);
        if (!busy && mine != 0xFFFFFFFFu) q.done(mine, t.best), mine = 0xFFFFFFFFu;
        uint32_t bm = __ballot_sync(0xFFFFFFFFu, busy);
        if (!exhausted && 32 - __popc(bm) >= HL_REFILL_MIN)
        {
            const uint32_t idle = ~bm;
            const uint32_t want = (uint32_t)__popc(idle);
            uint32_t       base = 0;
            if (lane == 0) base = atomicAdd(fetch, want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            exhausted = base + want >= count;
            const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
            if (!busy && i < count)
            {
                f3    o, d;
                float tmin, tmax;
                q.load(i, o, tmin, d, tmax);
                trav_begin(start, t, st, true, o, tmin, d, tmax, flags);
                mine = i, busy = true; // (a query that starts with nothing to do is retired on the next pass)
            }
            bm = __ballot_sync(0xFFFFFFFFu, busy);
        }
        if (bm == 0)
        {
            if (exhausted) break;
            continue; // fewer than HL_REFILL_MIN idle lanes cannot happen with bm == 0; kept for clarity
        }
        trav_step_warp(s, t, st, busy, bm);
    }
}

// ---- extend ----------------------------------------------------------------------------------------
struct ExtendQueue
{
    const float4* __restrict__ ray_o;
    const float4* __restrict__ ray_d;
    float4* __restrict__ hit_a;
    uint2* __restrict__ hit_b;
    float tmin, tmax;
    __device__ __forceinline__ void load(uint32_t i, f3& o, float& t0, f3& d, float& t1) const
    {
        const float4 o4 = ld4(ray_o + i), d4 = ld4(ray_d + i);
        o = mk3(o4.x, o4.y, o4.z), d = mk3(d4.x, d4.y, d4.z), t0 = tmin, t1 = tmax;
    }
    __device__ __forceinline__ void done(uint32_t i, const Hit& h) const
    {
        hit_a[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.primitive));
        hit_b[i] = make_uint2(h.instance, h.geometry);
    }
};
__global__ void HL_TRACE_BOUNDS k_extend(SceneView s, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const uint32_t* __restrict__ count_ptr,
                                                           uint32_t* fetch, float tmin, float tmax, uint32_t flags, float4* __restrict__ hit_a, uint2* __restrict__ hit_b)
{
    HL_TRACE_SHARED(st);
    ExtendQueue q;
    q.ray_o = ray_o, q.ray_d = ray_d, q.hit_a = hit_a, q.hit_b = hit_b, q.tmin = tmin, q.tmax = tmax;
    trace_queue(s, q, *count_ptr, fetch, flags, st);
}

// ---- shade -----------------------------------------------------------------------------------------
// warp-aggregated queue append: one atomicAdd per warp, slots handed out by ballot prefix
__device__ __forceinline__ uint32_t warp_append(bool pred, uint32_t* counter, uint32_t lane)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
    uint32_t       base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

__global__ void HL_SHADE_BOUNDS k_shade(SceneView s, ShadeParams prm, uint32_t depth, const uint32_t* __restrict__ count_ptr, const float4* __restrict__ ray_o,
                                                          const float4* __restrict__ ray_d, const float4* __restrict__ hit_a, const uint2* __restrict__ hit_b, float4* state_a,
                                                          float4* state_b, float4* next_o, float4* next_d, uint32_t* next_count, float4* sh_o, float4* sh_d, float4* sh_c,
                                                          uint32_t* sh_count)
{
    const uint32_t count  = *count_ptr;
    const uint32_t lane   = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < count; base += stride)
    {
        const uint32_t i      = base + lane;
        const bool     active = i < count;
        bool           want_shadow = false, want_next = false;
        ShadeResult    r;
        uint32_t       path = 0;
        if (active)
        {
            const float4 ha = hit_a[i];
            const uint2  hb = hit_b[i];
            const float4 d4 = ray_d[i];
            path            = __float_as_uint(ray_o[i].w);
            float4 sa = state_a[path], sb = state_b[path];
            const f3 T = mk3(sa.x, sa.y, sa.z);
            const f3 dir = mk3(d4.x, d4.y, d4.z);
            if (hb.x == HL_MISS)
            {
                const f3 e = shade_miss(s, depth, dir, T);
                sb.x += e.x, sb.y += e.y, sb.z += e.z;
                state_b[path] = sb;
            }
            else
            {
                Hit h;
                h.t = ha.x, h.u = ha.y, h.v = ha.z, h.primitive = __float_as_uint(ha.w), h.instance = hb.x, h.geometry = hb.y;
                Rng rng;
                rng.x = __float_as_uint(sa.w), rng.y = __float_as_uint(sb.w);
                shade_hit(s, prm, depth, dir, h, T, rng, r);
                want_shadow = r.has_shadow, want_next = r.continues;
                sb.x += r.emitted.x, sb.y += r.emitted.y, sb.z += r.emitted.z;
                sb.w = __uint_as_float(rng.y);
                if (want_next) sa.x = r.T.x, sa.y = r.T.y, sa.z = r.T.z;
                sa.w          = __uint_as_float(rng.x);
                state_a[path] = sa, state_b[path] = sb;
            }
        }
        const uint32_t si = warp_append(want_shadow, sh_count, lane);
        if (want_shadow)
        {
            sh_o[si] = make_float4(r.shadow_o.x, r.shadow_o.y, r.shadow_o.z, __uint_as_float(path));
            sh_d[si] = make_float4(r.shadow_d.x, r.shadow_d.y, r.shadow_d.z, r.shadow_tmax);
            sh_c[si] = make_float4(r.direct.x, r.direct.y, r.direct.z, 0.0f);
        }
        const uint32_t ni = warp_append(want_next, next_count, lane);
        if (want_next)
        {
            next_o[ni] = make_float4(r.next_o.x, r.next_o.y, r.next_o.z, __uint_as_float(path));
            next_d[ni] = make_float4(r.next_d.x, r.next_d.y, r.next_d.z, 0.0f);
        }
    }
}

// ---- connect ---------------------------------------------------------------------------------------
struct ConnectQueue
{
    const float4* __restrict__ sh_o;
    const float4* __restrict__ sh_d;
    const float4* __restrict__ sh_c;
    float4* state_b;
    float   tmin;
    __device__ __forceinline__ void load(uint32_t i, f3& o, float& t0, f3& d, float& t1) const
    {
        const float4 o4 = ld4(sh_o + i), d4 = ld4(sh_d + i);
        o = mk3(o4.x, o4.y, o4.z), d = mk3(d4.x, d4.y, d4.z), t0 = tmin, t1 = d4.w;
    }
    __device__ __forceinline__ void done(uint32_t i, const Hit& h) const
    {
        if (h.instance != HL_MISS) return; // occluded
        // the shadow miss shader ran: p_Visibility = true.  One shadow ray per path and bounce: no other
        // thread touches this path's radiance while the connect stage runs.
        const float4   c    = ld4(sh_c + i);
        const uint32_t path = __float_as_uint(ld4(sh_o + i).w);
        float4         sb   = state_b[path];
        sb.x += c.x, sb.y += c.y, sb.z += c.z;
        state_b[path] = sb;
    }
};
__global__ void HL_TRACE_BOUNDS k_connect(SceneView s, const float4* __restrict__ sh_o, const float4* __restrict__ sh_d, const float4* __restrict__ sh_c,
                                                            const uint32_t* __restrict__ count_ptr, uint32_t* fetch, float tmin, uint32_t flags, float4* state_b)
{
    HL_TRACE_SHARED(st);
    ConnectQueue q;
    q.sh_o = sh_o, q.sh_d = sh_d, q.sh_c = sh_c, q.state_b = state_b, q.tmin = tmin;
    trace_queue(s, q, *count_ptr, fetch, flags, st);
}

// ---- tail ------------------------------------------------------------------------------------------
// Russian roulette thins the queues quickly: past the first bounces a launch carries a few thousand rays and
// its duration is the latency of its slowest ray, not throughput.  When the extension queue of bounce
// `depth0` holds at most `threshold` rays, this kernel finishes those paths in one launch — extend, shade and
// connect in a per-path loop over the remaining bounces, same arithmetic and same accumulation order as the
// wavefront stages — and zeroes the queue so that the wavefront launches enqueued behind it are no-ops.
__global__ void __launch_bounds__(HL_TRACE_BLOCK) k_tail(SceneView s, ShadeParams prm, uint32_t depth0, uint32_t threshold, uint32_t* counters, const float4* __restrict__ ray_o,
                                                         const float4* __restrict__ ray_d, const float4* __restrict__ state_a, float4* state_b)
{
    HL_TRACE_SHARED(st);
    const uint32_t count = counters[CTR_EXT_COUNT + depth0];
    if (count == 0 || count > threshold) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t       n_ext = 0, n_sh = 0;
    for (;;)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counters + CTR_TAIL_FETCH, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= count) break;
        const uint32_t i     = base + lane;
        bool           alive = i < count;
        f3             o = mk3(0.0f), d = mk3(0.0f), T = mk3(0.0f), L = mk3(0.0f);
        Rng            rng;
        rng.x = rng.y = 0;
        uint32_t path = 0, depth = depth0;
        float    lw = 0.0f;
        if (alive)
        {
            const float4 o4 = ray_o[i], d4 = ray_d[i];
            path            = __float_as_uint(o4.w);
            const float4 sa = state_a[path], sb = state_b[path];
            o = mk3(o4.x, o4.y, o4.z), d = mk3(d4.x, d4.y, d4.z);
            T = mk3(sa.x, sa.y, sa.z), L = mk3(sb.x, sb.y, sb.z);
            rng.x = __float_as_uint(sa.w), rng.y = __float_as_uint(sb.w), lw = sb.w;
        }
        const bool mine = alive;
        while (__any_sync(0xFFFFFFFFu, alive))
        {
            Hit h;
            trace_ray(s, alive, o, depth == 0 ? 0.001f : 0.0001f, d, 10000.0f, depth == 0 ? 0u : HL_RAY_OPAQUE, h, st);
            ShadeResult r;
            r.has_shadow = false, r.continues = false;
            if (alive)
            {
                n_ext++;
                if (h.instance == HL_MISS)
                    L = L + shade_miss(s, depth, d, T);
                else
                {
                    shade_hit(s, prm, depth, d, h, T, rng, r);
                    L = L + r.emitted;
                }
            }
            Hit hs;
            trace_ray(s, r.has_shadow, r.shadow_o, 0.0001f, r.shadow_d, r.shadow_tmax, depth == 0 ? HL_RAY_TERMINATE : (HL_RAY_OPAQUE | HL_RAY_TERMINATE), hs, st);
            if (r.has_shadow)
            {
                n_sh++;
                if (hs.instance == HL_MISS) L = L + r.direct;
            }
            alive = alive && r.continues;
            if (alive) o = r.next_o, d = r.next_d, T = r.T, depth++;
        }
        if (mine) state_b[path] = make_float4(L.x, L.y, L.z, lw);
    }
    // ray counters: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) n_ext += __shfl_down_sync(0xFFFFFFFFu, n_ext, off), n_sh += __shfl_down_sync(0xFFFFFFFFu, n_sh, off);
    if (lane == 0 && (n_ext | n_sh)) atomicAdd(counters + CTR_TAIL_EXT, n_ext), atomicAdd(counters + CTR_TAIL_SH, n_sh);
    // last block out closes the queue: the wavefront launches behind this kernel see an empty bounce
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        if (atomicAdd(counters + CTR_TAIL_DONE, 1u) == gridDim.x - 1) counters[CTR_EXT_COUNT + depth0] = 0, counters[CTR_TAIL_DONE] = 0, counters[CTR_TAIL_FETCH] = 0;
    }
}

// ---- resolve ---------------------------------------------------------------------------------------
__global__ void k_resolve(FrameParams fp, const float4* __restrict__ state_b, float4* accum, int accum_mode, uint32_t* rgba8, int fused, float exposure, int op, float scale)
{
    const uint32_t n = fp.lw * fp.lh;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t W = fp.pc.launch_id_size[2], H = fp.pc.launch_id_size[3];
    uint32_t       lx, ly;
    launch_pixel(fp.lw, fp.tiled, i, lx, ly);
    const uint32_t px = fp.pc.launch_id_size[0] + lx, py = fp.pc.launch_id_size[1] + ly;
    const size_t   pix = (size_t)py * W + px;
    const float4   sb  = state_b[i];
    const float4   pv  = accum[pix];
    const f3       L = mk3(sb.x, sb.y, sb.z), prev = mk3(pv.x, pv.y, pv.z);
    const f3       c = accum_mode == HL_ACCUM_SUM ? accumulate_sum(L, prev) : accumulate_running_mean(L, prev, fp.pc.num_frames);
    accum[pix]       = make_float4(c.x, c.y, c.z, 1.0f);
    // scale = 1 for the running mean (x * 1.0f is exact); 1 / samples for a per-GPU sum image (HL_ACCUM_SUM preview)
    if (fused) rgba8[(size_t)(H - 1 - py) * W + px] = tone_map_rgba8(mk3(c.x * scale, c.y * scale, c.z * scale), exposure, op);
}

__global__ void k_totals(uint32_t* counters, unsigned long long* totals, uint32_t bounces)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        unsigned long long e = 0, s = 0;
        for (uint32_t b = 0; b < bounces; b++) e += counters[CTR_EXT_COUNT + b], s += counters[CTR_SH_COUNT + b];
        e += counters[CTR_TAIL_EXT], s += counters[CTR_TAIL_SH];
        totals[0] += e, totals[1] += s;
    }
}

// ---- film ------------------------------------------------------------------------------------------
__global__ void k_clear(float4* accum, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) accum[i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f); // renderer.cpp:212-223 clear colour
}
__global__ void k_tonemap(const float4* __restrict__ accum, uint32_t W, uint32_t H, float exposure, int op, float scale, uint32_t* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)W * H) return;
    const uint32_t r = (uint32_t)(i / W), x = (uint32_t)(i % W);
    const float4   a = accum[(size_t)(H - 1 - r) * W + x];
    out[i]           = tone_map_rgba8(mk3(a.x * scale, a.y * scale, a.z * scale), exposure, op);
}
struct SkyCoeffs
{
    float cf[40];
    float sun[3];
};
__global__ void k_sky(SkyCoeffs c, uint32_t size, float4* out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)6 * size * size;
    if (i >= n) return;
    const int      face = (int)(i / ((size_t)size * size));
    const uint32_t j = (uint32_t)((i / size) % size), x = (uint32_t)(i % size);
    const f3       v = hosek_wilkie_radiance(c.cf, cube_texel_direction(face, x, j, size), mk3(c.sun));
    out[i]           = make_float4(v.x, v.y, v.z, 1.0f);
}
__global__ void __launch_bounds__(HL_TRACE_BLOCK) k_trace_generic(SceneView s, const float* __restrict__ rays, uint32_t n, uint32_t flags, float* __restrict__ hits)
{
    HL_TRACE_SHARED(st);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x)
    {
        const uint32_t i   = base + lane;
        const bool     act = i < n;
        const float*   r   = rays + (size_t)(act ? i : 0) * 8;
        Hit            h;
        trace_ray(s, act, mk3(r[0], r[1], r[2]), r[3], mk3(r[4], r[5], r[6]), r[7], flags, h, st);
        if (!act) continue;
        float*     o   = hits + (size_t)i * 6;
        const bool hit = h.instance != HL_MISS;
        o[0] = hit ? h.t : hl_inf(), o[1] = hit ? h.u : 0.0f, o[2] = hit ? h.v : 0.0f;
        o[3] = __uint_as_float(h.instance), o[4] = __uint_as_float(h.geometry), o[5] = __uint_as_float(h.primitive);
    }
}

// Debug output buffers (Renderer::set_current_output_buffer, include/gfx/renderer.h:25-33): what the reference
// rasterises with debug_visualization.frag:144-161 — albedo, shading normal * 0.5 + 0.5, roughness, metallic,
// emissive of the surface seen through each pixel — evaluated here on the primary hits with the path's own
// surface fetch (path_trace_rchit.glsl:206-252 and debug_visualization.frag:74-135 are the same fetch_* functions).
__global__ void k_output_buffer(SceneView s, const float4* __restrict__ hit_a, const uint2* __restrict__ hit_b, uint32_t n, uint32_t W, uint32_t tiled, int which, float4* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        uint32_t x, y;
        launch_pixel(W, tiled, i, x, y);
        const float4 ha = hit_a[i];
        const uint2  hb = hit_b[i];
        Hit          h;
        h.t = ha.x, h.u = ha.y, h.v = ha.z, h.primitive = __float_as_uint(ha.w), h.instance = hb.x, h.geometry = hb.y;
        const f4 c = output_buffer_value(s, h, which);
        out[(size_t)y * W + x] = make_float4(c.x, c.y, c.z, c.w);
    }
}

// Ray debug view (hl_debug.h): one path per lane; segments appended with one atomic per segment, as the reference's
// shaders do (rchit:552, rmiss:44) — their order in the buffer is not defined there either.
struct DebugRayOut
{
    float4*   verts; // two float4 per vertex: position.xyz 1, colour.rgb 1 (DebugRayVertex, common.glsl:54-58)
    uint32_t* count;
    uint32_t  capacity;
    __device__ __forceinline__ uint32_t alloc2() { return atomicAdd(count, 2u); }
    __device__ __forceinline__ void     put(uint32_t k, f3 p, f3 c)
    {
        if (k < capacity) verts[2 * (size_t)k] = make_float4(p.x, p.y, p.z, 1.0f), verts[2 * (size_t)k + 1] = make_float4(c.x, c.y, c.z, 1.0f);
    }
};
__global__ void __launch_bounds__(HL_TRACE_BLOCK) k_debug_rays(SceneView s, hl_push_constants pc, uint32_t n, DebugRayOut out)
{
    HL_TRACE_SHARED(st);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x)
        debug_ray_path(s, pc, base + lane, base + lane < n, st, out);
}

// ---- host side -------------------------------------------------------------------------------------
static void slot_drop_graph(hl_wave_slot& w);
static void slot_alloc(hl_context_t* ctx, hl_wave_slot& w, size_t n)
{
    slot_drop_graph(w); // the buffers below may move
    w.state_a.alloc(n * 16), w.state_b.alloc(n * 16);
    for (int k = 0; k < 2; k++) w.ext_o[k].alloc(n * 16), w.ext_d[k].alloc(n * 16);
    w.hit_a.alloc(n * 16), w.hit_b.alloc(n * 8);
    w.sh_o.alloc(n * 16), w.sh_d.alloc(n * 16), w.sh_c.alloc(n * 16);
    w.rgba8.alloc(n * 4);
    HL_CUDA(cudaMemsetAsync(w.rgba8.p, 0, n * 4, ctx->stream));
    if (!w.counters.p)
    {
        w.counters.alloc(CTR_BYTES);
        HL_CUDA(cudaMemsetAsync(w.counters.p, 0, CTR_BYTES, ctx->stream));
    }
    if (!w.stream) HL_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    if (!w.resolved) HL_CUDA(cudaEventCreateWithFlags(&w.resolved, cudaEventDisableTiming));
    if (!w.copy_stream) HL_CUDA(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
    if (!w.image_ready) HL_CUDA(cudaEventCreateWithFlags(&w.image_ready, cudaEventDisableTiming));
    if (!w.copy_done) HL_CUDA(cudaEventCreateWithFlags(&w.copy_done, cudaEventDisableTiming));
    w.pending = false, w.copy_pending = false;
}
static void slot_drop_graph(hl_wave_slot& w)
{
    if (w.graph_exec) cudaGraphExecDestroy(w.graph_exec);
    w.graph_exec = nullptr;
}
static void slot_release(hl_wave_slot& w)
{
    slot_drop_graph(w);
    w.state_a.release(), w.state_b.release();
    for (int k = 0; k < 2; k++) w.ext_o[k].release(), w.ext_d[k].release();
    w.hit_a.release(), w.hit_b.release(), w.sh_o.release(), w.sh_d.release(), w.sh_c.release(), w.rgba8.release();
}

void wavefront_alloc(hl_context_t* ctx)
{
    const size_t n = (size_t)ctx->W * ctx->H;
    ctx->accum.alloc(n * 16), ctx->rgba8.alloc(n * 4);
    for (int k = 0; k < ctx->n_slots; k++) slot_alloc(ctx, ctx->slot[k], n);
    if (!ctx->main_ev) HL_CUDA(cudaEventCreateWithFlags(&ctx->main_ev, cudaEventDisableTiming));
    ctx->queue_capacity = n;
    film_clear(ctx);
}

// HL_OPT_FRAMES_IN_FLIGHT; the caller has joined the frames in flight with the main stream
void wavefront_set_slots(hl_context_t* ctx, int n_slots)
{
    const size_t n = (size_t)ctx->W * ctx->H;
    for (int k = ctx->n_slots; k < n_slots; k++) slot_alloc(ctx, ctx->slot[k], n);
    if (n_slots < ctx->n_slots)
    {
        HL_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = n_slots; k < ctx->n_slots; k++)
        {
            // keep the ray totals (hl_get_counters sums the slots in use)
            unsigned long long t[2], t0[2];
            HL_CUDA(cudaMemcpy(t, (char*)ctx->slot[k].counters.p + CTR_TOTALS_OFFSET, 16, cudaMemcpyDeviceToHost));
            HL_CUDA(cudaMemcpy(t0, (char*)ctx->slot[0].counters.p + CTR_TOTALS_OFFSET, 16, cudaMemcpyDeviceToHost));
            t0[0] += t[0], t0[1] += t[1], t[0] = t[1] = 0;
            HL_CUDA(cudaMemcpy((char*)ctx->slot[0].counters.p + CTR_TOTALS_OFFSET, t0, 16, cudaMemcpyHostToDevice));
            HL_CUDA(cudaMemcpy((char*)ctx->slot[k].counters.p + CTR_TOTALS_OFFSET, t, 16, cudaMemcpyHostToDevice));
            slot_release(ctx->slot[k]);
        }
    }
    ctx->n_slots = n_slots;
}

void wavefront_release(hl_context_t* ctx)
{
    ctx->accum.release(), ctx->rgba8.release();
    for (hl_wave_slot& w : ctx->slot) slot_release(w);
}

// every frame in flight happens-before whatever is enqueued on the main stream next
void wavefront_join(hl_context_t* ctx)
{
    for (hl_wave_slot& w : ctx->slot)
    {
        if (w.pending)
        {
            HL_CUDA(cudaEventRecord(w.resolved, w.stream));
            HL_CUDA(cudaStreamWaitEvent(ctx->stream, w.resolved, 0));
            w.pending = false;
        }
        if (w.copy_pending)
        {
            HL_CUDA(cudaStreamWaitEvent(ctx->stream, w.copy_done, 0));
            w.copy_pending = false;
        }
    }
}

void film_clear(hl_context_t* ctx)
{
    const size_t n = (size_t)ctx->W * ctx->H;
    k_clear<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->accum.as<float4>(), n);
    ctx->launches++;
    HL_CUDA(cudaMemsetAsync(ctx->rgba8.p, 0, n * 4, ctx->stream));
    for (hl_wave_slot& w : ctx->slot)
        if (w.rgba8.p) HL_CUDA(cudaMemsetAsync(w.rgba8.p, 0, n * 4, ctx->stream));
    ctx->rgba8_cur = ctx->rgba8.p;
    ctx->sum_samples = 0;
}

void film_tonemap(hl_context_t* ctx, float exposure, int op, float scale)
{
    const size_t n = (size_t)ctx->W * ctx->H;
    k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->accum.as<float4>(), ctx->W, ctx->H, exposure, op, scale, ctx->rgba8.as<uint32_t>());
    ctx->rgba8_cur = ctx->rgba8.p;
    ctx->launches++;
}

void sky_bake(hl_context_t* ctx, const float* coeffs40, const float* sun3, uint32_t size)
{
    SkyCoeffs c;
    memcpy(c.cf, coeffs40, sizeof(c.cf));
    memcpy(c.sun, sun3, sizeof(c.sun));
    const size_t n = (size_t)6 * size * size;
    ctx->env_faces.alloc(n * 16);
    ctx->env_size = size;
    k_sky<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(c, size, ctx->env_faces.as<float4>());
    ctx->launches++;
}

__global__ void k_env_pad(const float4* __restrict__ faces, uint32_t size, float4* __restrict__ out)
{
    const uint32_t P = size + 2;
    const size_t   i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)6 * P * P) return;
    const int face = (int)(i / ((size_t)P * P));
    const int iy = (int)((i / P) % P) - 1, ix = (int)(i % P) - 1;
    const f4  c = cube_pad_texel((const f4*)faces, (int)size, face, ix, iy);
    out[i]      = make_float4(c.x, c.y, c.z, c.w);
}
void env_pad(hl_context_t* ctx)
{
    const uint32_t size = ctx->env_size;
    if (size)
    {
        const size_t n = (size_t)6 * (size + 2) * (size + 2);
        ctx->env_padded.alloc(n * 16);
        k_env_pad<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->env_faces.as<float4>(), size, ctx->env_padded.as<float4>());
        ctx->launches++;
    }
    ctx->view.env.faces = ctx->env_padded.as<f4>(), ctx->view.env.size = size;
}

static void ensure_events(hl_context_t* ctx)
{
    if (ctx->ev_ready) return;
    for (auto& e : ctx->ev) HL_CUDA(cudaEventCreate(&e));
    ctx->ev_ready = true;
}

static void run_bounces(hl_context_t* ctx, hl_wave_slot& w, cudaStream_t st, const FrameParams& fp, uint32_t bounces, bool shade)
{
    uint32_t*    ctr  = w.counters.as<uint32_t>();
    const int    tgrid = ctx->sm_count * HL_TRACE_GRID_MULT; // persistent trace kernels
    const int    sgrid = ctx->sm_count * HL_SHADE_GRID_MULT;
    ShadeParams  prm;
    prm.num_lights = fp.pc.num_lights, prm.max_ray_bounces = fp.pc.max_ray_bounces, prm.shadow_ray_bias = fp.pc.shadow_ray_bias;
    const bool prof = ctx->profiling;
    for (uint32_t b = 0; b < bounces; b++)
    {
        const int cur = b & 1, nxt = cur ^ 1;
        // ray parameters per SURVEY A.5: primary tmin 0.001 flags 0; indirect tmin 0.0001 Opaque
        const float    ext_tmin  = b == 0 ? 0.001f : 0.0001f;
        const uint32_t ext_flags = b == 0 ? 0u : HL_RAY_OPAQUE;
        if (prof) HL_CUDA(cudaEventRecord(ctx->ev[2 + 5 * b + 0], st));
        if (shade && b >= ctx->tail_start && ctx->tail_threshold > 0)
        {
            // sparse late bounces: finish the surviving paths in one launch when the queue is small (see k_tail)
            k_tail<<<tgrid, HL_TRACE_BLOCK, 0, st>>>(ctx->view, prm, b, ctx->tail_threshold, ctr, w.ext_o[cur].as<float4>(), w.ext_d[cur].as<float4>(), w.state_a.as<float4>(),
                                                     w.state_b.as<float4>());
            ctx->launches++;
        }
        if (prof) HL_CUDA(cudaEventRecord(ctx->ev[2 + 5 * b + 1], st));
        k_extend<<<tgrid, HL_TRACE_BLOCK, 0, st>>>(ctx->view, w.ext_o[cur].as<float4>(), w.ext_d[cur].as<float4>(), ctr + CTR_EXT_COUNT + b, ctr + CTR_EXT_FETCH + b, ext_tmin,
                                                   10000.0f, ext_flags, w.hit_a.as<float4>(), w.hit_b.as<uint2>());
        ctx->launches++;
        if (!shade) break;
        if (prof) HL_CUDA(cudaEventRecord(ctx->ev[2 + 5 * b + 2], st));
        k_shade<<<sgrid, HL_SHADE_BLOCK, 0, st>>>(ctx->view, prm, b, ctr + CTR_EXT_COUNT + b, w.ext_o[cur].as<float4>(), w.ext_d[cur].as<float4>(), w.hit_a.as<float4>(),
                                                  w.hit_b.as<uint2>(), w.state_a.as<float4>(), w.state_b.as<float4>(), w.ext_o[nxt].as<float4>(), w.ext_d[nxt].as<float4>(),
                                                  ctr + CTR_EXT_COUNT + b + 1, w.sh_o.as<float4>(), w.sh_d.as<float4>(), w.sh_c.as<float4>(), ctr + CTR_SH_COUNT + b);
        ctx->launches++;
        if (prof) HL_CUDA(cudaEventRecord(ctx->ev[2 + 5 * b + 3], st));
        // shadow rays: depth 0 -> flags 0 (any-hit runs); deeper -> Opaque | TerminateOnFirstHit (rchit:286-290).
        // Visibility only asks whether ANY accepted intersection exists in (tmin, tmax) — acceptance is a
        // per-candidate test (alpha), independent of order — so the depth-0 query may also stop at its first
        // accepted hit: the result is identical to the reference's closest-hit + shadow.rchit sequence.
        const uint32_t sh_flags = b == 0 ? HL_RAY_TERMINATE : (HL_RAY_OPAQUE | HL_RAY_TERMINATE);
        k_connect<<<tgrid, HL_TRACE_BLOCK, 0, st>>>(ctx->view, w.sh_o.as<float4>(), w.sh_d.as<float4>(), w.sh_c.as<float4>(), ctr + CTR_SH_COUNT + b, ctr + CTR_SH_FETCH + b, 0.0001f,
                                                    sh_flags, w.state_b.as<float4>());
        ctx->launches++;
        if (prof) HL_CUDA(cudaEventRecord(ctx->ev[2 + 5 * b + 4], st));
    }
}

// The bounce loop through a CUDA graph: its ~30 launches take no per-frame argument (the push constants only reach the
// generate and resolve kernels), so one cudaGraphLaunch replaces them.  The host cost of a frame drops from ~40 API calls
// to ~10 — at 512 x 512 (BASELINE configs[0]) enqueuing a frame cost as much as rendering it.  Captured per slot on first
// use and again whenever anything the launches carry by value changes (scene view, integrator settings, extent).
static void replay_bounces(hl_context_t* ctx, hl_wave_slot& w, cudaStream_t st, const FrameParams& fp, uint32_t bounces)
{
    hl_wave_slot::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.view = ctx->view;
    key.num_lights = fp.pc.num_lights, key.max_ray_bounces = fp.pc.max_ray_bounces, key.bounces = bounces, key.tail_start = ctx->tail_start, key.tail_threshold = ctx->tail_threshold;
    key.W = ctx->W, key.H = ctx->H, key.shadow_ray_bias = fp.pc.shadow_ray_bias;
    if (!w.graph_exec || memcmp(&key, &w.graph_key, sizeof(key)) != 0)
    {
        slot_drop_graph(w);
        const uint64_t before = ctx->launches;
        cudaGraph_t    graph  = nullptr;
        HL_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try
        {
            run_bounces(ctx, w, st, fp, bounces, true);
        }
        catch (...)
        {
            cudaStreamEndCapture(st, &graph);
            if (graph) cudaGraphDestroy(graph);
            ctx->launches = before;
            throw;
        }
        HL_CUDA(cudaStreamEndCapture(st, &graph));
        w.graph_launches = (uint32_t)(ctx->launches - before);
        ctx->launches    = before;
        const cudaError_t e = cudaGraphInstantiate(&w.graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess)
        {
            w.graph_exec = nullptr;
            HL_CUDA(e);
        }
        memcpy(&w.graph_key, &key, sizeof(key));
    }
    HL_CUDA(cudaGraphLaunch(w.graph_exec, st));
    ctx->launches += w.graph_launches;
}

void wavefront_render_frame(hl_context_t* ctx, const hl_push_constants& pc, uint32_t lw, uint32_t lh, const ResolveOptions& opt)
{
    FrameParams fp;
    fp.pc = pc;
    fp.lw = lw, fp.lh = lh, fp.tiled = launch_is_tiled(lw, lh);
    const uint32_t n = lw * lh;
    if (n == 0) return;
    // max_ray_bounces = 0 still traces and shades the primary ray (rgen:205; the closest-hit shader only skips the indirect ray)
    const uint32_t bounces = std::max<uint32_t>(1u, std::min<uint32_t>(pc.max_ray_bounces, HL_MAX_BOUNCES));
    const bool     prof    = ctx->profiling;
    const bool     piped   = ctx->pipeline && ctx->n_slots > 1 && !prof;
    if (!piped) wavefront_join(ctx);
    const uint64_t ns   = (uint64_t)ctx->n_slots;
    hl_wave_slot& w     = ctx->slot[piped ? ctx->frame_seq % ns : 0];
    hl_wave_slot& other = ctx->slot[piped ? (ctx->frame_seq + ns - 1) % ns : 0]; // the previous frame's
    cudaStream_t  st    = piped ? w.stream : ctx->stream;
    if (piped)
    {
        // whatever the caller enqueued on the main stream (uploads, clears, table updates) precedes this frame
        HL_CUDA(cudaEventRecord(ctx->main_ev, ctx->stream));
        HL_CUDA(cudaStreamWaitEvent(st, ctx->main_ev, 0));
    }
    if (prof) ensure_events(ctx);
    uint32_t* ctr = w.counters.as<uint32_t>();
    HL_CUDA(cudaMemsetAsync(ctr, 0, CTR_U32_TOTAL * 4, st));
    if (prof) HL_CUDA(cudaEventRecord(ctx->ev[0], st));
    k_generate<<<(n + 255) / 256, 256, 0, st>>>(fp, w.state_a.as<float4>(), w.state_b.as<float4>(), w.ext_o[0].as<float4>(), w.ext_d[0].as<float4>(), ctr);
    ctx->launches++;
    if (piped && ctx->use_graphs)
        replay_bounces(ctx, w, st, fp, bounces);
    else
        run_bounces(ctx, w, st, fp, bounces, true);
    const size_t last = 2 + 5 * (size_t)bounces;
    if (prof) HL_CUDA(cudaEventRecord(ctx->ev[last], st));
    // progressive blends are applied in frame order: wait for the previous frame's resolve pass
    if (piped && other.pending) HL_CUDA(cudaStreamWaitEvent(st, other.resolved, 0));
    // ... and this slot's RGBA8 target may still be on its way to the host (read-back of the frame issued n_slots ago)
    if (w.copy_pending && opt.tone_map)
    {
        HL_CUDA(cudaStreamWaitEvent(st, w.copy_done, 0));
        w.copy_pending = false;
    }
    const bool full  = lw == ctx->W && lh == ctx->H;
    const bool fused = opt.tone_map && full;
    if (ctx->accum_mode == HL_ACCUM_SUM && full) ctx->sum_samples++;
    const float scale = ctx->accum_mode == HL_ACCUM_SUM ? 1.0f / (float)std::max<uint64_t>(ctx->sum_samples, 1) : 1.0f;
    k_resolve<<<(n + 255) / 256, 256, 0, st>>>(fp, w.state_b.as<float4>(), ctx->accum.as<float4>(), ctx->accum_mode, w.rgba8.as<uint32_t>(), fused ? 1 : 0, opt.exposure, opt.op, scale);
    if (opt.tone_map)
    {
        if (!fused) // a tile launch resolves only its own pixels: tone map the whole image, as the reference's full-screen pass does
        {
            const size_t px = (size_t)ctx->W * ctx->H;
            k_tonemap<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(ctx->accum.as<float4>(), ctx->W, ctx->H, opt.exposure, opt.op, 1.0f, w.rgba8.as<uint32_t>());
            ctx->launches++;
        }
        ctx->rgba8_cur = w.rgba8.p;
    }
    if (piped)
    {
        HL_CUDA(cudaEventRecord(w.resolved, st)); // the next frame's blend waits for this, not for the copy below
        w.pending = true;
    }
    if (opt.tone_map && opt.host)
    {
        const size_t bytes = (size_t)ctx->W * ctx->H * 4;
        if (piped)
        {
            HL_CUDA(cudaEventRecord(w.image_ready, st));
            HL_CUDA(cudaStreamWaitEvent(w.copy_stream, w.image_ready, 0));
            // copies of different slots run on different streams: two read-backs into the SAME host memory (a caller without
            // the ring of n_slots buffers the header asks for) are ordered, the later frame's image lands last
            for (hl_wave_slot& o : ctx->slot)
                if (&o != &w && o.copy_pending && o.copy_host == opt.host) HL_CUDA(cudaStreamWaitEvent(w.copy_stream, o.copy_done, 0));
            HL_CUDA(cudaMemcpyAsync(opt.host, w.rgba8.p, bytes, cudaMemcpyDeviceToHost, w.copy_stream));
            HL_CUDA(cudaEventRecord(w.copy_done, w.copy_stream));
            w.copy_pending = true, w.copy_host = opt.host;
        }
        else
            HL_CUDA(cudaMemcpyAsync(opt.host, w.rgba8.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    k_totals<<<1, 32, 0, st>>>(ctr, (unsigned long long*)((char*)w.counters.p + CTR_TOTALS_OFFSET), bounces);
    ctx->launches += 2;
    if (prof)
    {
        HL_CUDA(cudaEventRecord(ctx->ev[1], st));
        HL_CUDA(cudaEventSynchronize(ctx->ev[1]));
        hl_counters& c = ctx->last;
        c.ms_generate = c.ms_extend = c.ms_shade = c.ms_connect = c.ms_resolve = 0.0f;
        float ms;
        HL_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]));
        c.ms_generate = ms;
        std::vector<uint32_t> hc(CTR_U32_TOTAL);
        HL_CUDA(cudaMemcpy(hc.data(), ctr, CTR_U32_TOTAL * 4, cudaMemcpyDeviceToHost));
        ctx->bounce_prof_n = bounces, ctx->prof_tail_ext = hc[CTR_TAIL_EXT], ctx->prof_tail_sh = hc[CTR_TAIL_SH];
        for (uint32_t b = 0; b < bounces; b++)
        {
            hl_bounce_profile& bp = ctx->bounce_prof[b];
            bp.extension_rays = hc[CTR_EXT_COUNT + b], bp.shadow_rays = hc[CTR_SH_COUNT + b];
            HL_CUDA(cudaEventElapsedTime(&bp.ms_tail, ctx->ev[2 + 5 * b], ctx->ev[2 + 5 * b + 1]));
            HL_CUDA(cudaEventElapsedTime(&bp.ms_extend, ctx->ev[2 + 5 * b + 1], ctx->ev[2 + 5 * b + 2]));
            HL_CUDA(cudaEventElapsedTime(&bp.ms_shade, ctx->ev[2 + 5 * b + 2], ctx->ev[2 + 5 * b + 3]));
            HL_CUDA(cudaEventElapsedTime(&bp.ms_connect, ctx->ev[2 + 5 * b + 3], ctx->ev[2 + 5 * b + 4]));
            // hl_counters: the tail kernel's time is booked under the extend stage as before (it replaces the three stages of the
            // bounces it finishes); hl_get_bounce_profile separates it
            c.ms_extend += bp.ms_tail + bp.ms_extend, c.ms_shade += bp.ms_shade, c.ms_connect += bp.ms_connect;
        }
        HL_CUDA(cudaEventElapsedTime(&ms, ctx->ev[last], ctx->ev[1]));
        c.ms_resolve = ms;
        HL_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        c.ms_frame = ms;
    }
    ctx->frames++;
    ctx->frame_seq++;
}

// primary-ray closest hits into slot 0's hit buffers, on the main stream (callers join first)
void wavefront_primary_hits(hl_context_t* ctx, const hl_push_constants& pc)
{
    cudaStream_t  st = ctx->stream;
    hl_wave_slot& w  = ctx->slot[0];
    FrameParams   fp;
    fp.pc = pc;
    fp.pc.launch_id_size[0] = fp.pc.launch_id_size[1] = 0;
    fp.lw = ctx->W, fp.lh = ctx->H, fp.tiled = launch_is_tiled(ctx->W, ctx->H);
    const uint32_t n   = fp.lw * fp.lh;
    uint32_t*      ctr = w.counters.as<uint32_t>();
    HL_CUDA(cudaMemsetAsync(ctr, 0, CTR_U32_TOTAL * 4, st));
    k_generate<<<(n + 255) / 256, 256, 0, st>>>(fp, w.state_a.as<float4>(), w.state_b.as<float4>(), w.ext_o[0].as<float4>(), w.ext_d[0].as<float4>(), ctr);
    ctx->launches++;
    run_bounces(ctx, w, st, fp, 1, false);
}

void wavefront_output_buffer(hl_context_t* ctx, const hl_push_constants& pc, int which, float4* d_out)
{
    wavefront_primary_hits(ctx, pc);
    const uint32_t n = ctx->W * ctx->H;
    k_output_buffer<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->view, ctx->slot[0].hit_a.as<float4>(), ctx->slot[0].hit_b.as<uint2>(), n, ctx->W, launch_is_tiled(ctx->W, ctx->H), which, d_out);
    ctx->launches++;
}

void wavefront_debug_rays(hl_context_t* ctx, const hl_push_constants& pc, uint32_t n, float4* d_verts, uint32_t capacity, uint32_t* d_count)
{
    if (!n) return;
    DebugRayOut out;
    out.verts = d_verts, out.count = d_count, out.capacity = capacity;
    const uint32_t blocks = std::min<uint32_t>((n + HL_TRACE_BLOCK - 1) / HL_TRACE_BLOCK, (uint32_t)ctx->sm_count * 8u);
    k_debug_rays<<<blocks, HL_TRACE_BLOCK, 0, ctx->stream>>>(ctx->view, pc, n, out);
    ctx->launches++;
}

// traversals that ran out of stack since the last reset (hl_bvh.h note_stack_overflow)
uint64_t trav_overflow_count(hl_context_t* ctx, bool reset)
{
    unsigned long long n = 0;
    HL_CUDA(cudaMemcpyFromSymbolAsync(&n, g_trav_overflow, sizeof(n), 0, cudaMemcpyDeviceToHost, ctx->stream));
    HL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (reset && n)
    {
        const unsigned long long z = 0;
        HL_CUDA(cudaMemcpyToSymbolAsync(g_trav_overflow, &z, sizeof(z), 0, cudaMemcpyHostToDevice, ctx->stream));
    }
    return n;
}

// pixel (row-major index) of path i of a full-frame launch: hl_trace_primary_ids reorders the hit records with it
uint32_t wavefront_path_pixel(hl_context_t* ctx, uint32_t i)
{
    uint32_t x, y;
    launch_pixel(ctx->W, launch_is_tiled(ctx->W, ctx->H), i, x, y);
    return y * ctx->W + x;
}

void wavefront_trace_rays(hl_context_t* ctx, const float* d_rays, uint32_t n, uint32_t flags, void* d_hits)
{
    if (!n) return;
    k_trace_generic<<<ctx->sm_count * 8, HL_TRACE_BLOCK, 0, ctx->stream>>>(ctx->view, d_rays, n, flags, (float*)d_hits);
    ctx->launches++;
}
} // namespace hl
