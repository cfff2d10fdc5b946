// Three-axis PMFT histograms over the bonds of a device NeighborList: freud::pmft::PMFTXYZ, PMFTXYT and PMFTR12
// (freud/pmft/PMFTXYZ.cc:111-147, PMFTXYT.cc:77-101, PMFTR12.cc:91-113) -- and, built from the same parts, the bond
// orientational order diagram freud::environment::BondOrder (freud/environment/BondOrder.cc:100-153; end of file).
// One thread per bond, u32 bin counts in a block-shared histogram when it fits and in global memory otherwise, linear
// index (b0 * n1 + b1) * n2 + b2 (Histogram.h:327-351).
//
//   XYZ  the bond vector rotated by conj(q_i) and then by every equivalent orientation, all in the reference's float
//        operation order (VectorMath.h:810-818) -> three RegularAxis bins.  Pure float arithmetic: bit-exact counts.
//   XY   freud::pmft::PMFTXY (PMFTXY.cc:65-87): (x, y) = rotmat2(-theta_i) * (dx, dy) on two RegularAxis; the third
//        axis has one bin.
//   XYT  the same (x, y), and the angle t = modulusPositive(theta_j - atan2f(-dy, -dx), 2 pi).
//   R12  r = bond distance, t1 = modulusPositive(theta_j - atan2f(dy, dx), 2 pi),
//        t2 = modulusPositive(theta_i - atan2f(-dy, -dx), 2 pi).
//
// atan2f is the host libm's in the reference and CUDA's differs from it in the last place (both are within 2 ulp of
// the true value).  Everything between atan2f and the bin index is float arithmetic that host and device evaluate
// identically and that is monotone in the angle, so the kernel follows that chain from CUDA's atan2f and from the two
// values a margin (twice the sum of the error bounds) below and above it: the same bin three times means the bin
// cannot depend on whose atan2f it was.  Otherwise -- about 1 bond in 10^4 -- the bond goes to a list and is binned
// by the host with its libm (capi.cu): counts are bit-identical to the reference's, and no libm call runs per bond.
// cosf / sinf of the query particle's angle (rotmat2::fromAngle, VectorMath.h:912-921) get the same treatment: CUDA's
// sincosf, and the rotated coordinate must land in the same bin when moved by what a few last places of (cos, sin)
// can move it -- the host evaluates libm only for the bonds that do not (it used to for every particle).  (Double
// precision atan2 / sincos were tried first: 390 instructions per bond, the kernel issue-bound at 0.59 ms per frame.)
#include <algorithm>
#include <cstring>

#include "internal.h"
#include "pair_math.cuh"

namespace fgpu {

namespace {

constexpr float kTwoPi = 6.28318548202514648f; // (float) (2.0 * M_PI), Box.h:24
// 2 ulp (CUDA) + 2 ulp (libm) of atan2f at pi = 9.6e-7: the margin is twice that.  kEdgeGuard keeps the kernel away
// from the wrap of the modulus.
constexpr float kAngleMargin = 2.0e-6f;
constexpr float kEdgeGuard = 1.0e-5f;

// fmodf(a, 2 pi) for moderate |a| without the library's bit-serial loop: the remainder of a truncated division is a
// float, so |a| - q 2 pi comes out of one fused multiply-add exactly once q is the right integer (the quotient from
// one multiplication can be off by one).  The result carries the sign of a, like fmodf's.
__device__ __forceinline__ float fmod_two_pi(float a)
{
    float const m = fabsf(a);
    if (!(m < 1.0e5f))
    {
        return fmodf(a, kTwoPi); // huge, infinite or NaN arguments: the library's path
    }
    float q = truncf(m * 0.15915494f);
    float r = __fmaf_rn(-q, kTwoPi, m);
    if (r < 0.0f)
    {
        q -= 1.0f;
        r = __fmaf_rn(-q, kTwoPi, m);
    }
    else if (r >= kTwoPi)
    {
        q += 1.0f;
        r = __fmaf_rn(-q, kTwoPi, m);
    }
    return copysignf(r, a);
}

// util::modulusPositive(a, 2 pi) = fmodf(fmodf(a, 2 pi) + 2 pi, 2 pi), utils.h:29-32.  The sum lies in [0, 4 pi], so the
// outer fmodf is at most two exact subtractions (Sterbenz).
__device__ __forceinline__ float mod_two_pi(float a)
{
    float r = __fadd_rn(fmod_two_pi(a), kTwoPi);
    if (r >= kTwoPi)
    {
        r = __fsub_rn(r, kTwoPi);
    }
    if (r >= kTwoPi)
    {
        r = __fsub_rn(r, kTwoPi);
    }
    return r;
}

// Bin of t = modulusPositive(orientation - atan2f(y, x), 2 pi) on `axis` = RegularAxis(n, 0, 2 pi); *sure = false if a
// last-place difference in atan2f could change it.  The chain from d to the bin is float arithmetic that both sides
// evaluate identically and that is monotone in d between two wraps of the modulus, so it is followed at d - m, d and
// d + m (m = kAngleMargin covers the two atan2f): same bin at all three and no wrap in between -- t falls as d grows
// -- settle it.  Close to 0 and 2 pi, where the modulus wraps (and can round onto 2 pi), the host decides.
__device__ __forceinline__ int angle_bin(const AxisDev& axis, float orientation, float y, float x, bool* sure)
{
    float const d = atan2f(y, x);
    float const t = mod_two_pi(__fsub_rn(orientation, d));
    float const t_lo = mod_two_pi(__fsub_rn(orientation, d + kAngleMargin));
    float const t_hi = mod_two_pi(__fsub_rn(orientation, d - kAngleMargin));
    int const bin = axis_bin(axis, t);
    // a NaN orientation is never sure
    *sure = axis_bin(axis, t_lo) == bin && axis_bin(axis, t_hi) == bin && t_lo <= t && t <= t_hi && t > kEdgeGuard
        && t < kTwoPi - kEdgeGuard;
    return bin;
}

// Bin of `value` on a RegularAxis when `value` is only known to +-slack: the bin is a monotone function of the value
// (float arithmetic both sides evaluate identically), so the bins of value - slack and value + slack decide: equal ->
// settled (both outside: the bond is dropped for good), else the host decides.
__device__ __forceinline__ int slack_bin(const AxisDev& axis, float value, float slack, bool* sure)
{
    int const lo = axis_bin(axis, value - slack), hi = axis_bin(axis, value + slack);
    if (lo == hi)
    {
        return lo;
    }
    *sure = false;
    return max(axis_bin(axis, value), 0); // the host decides, also whether the bond is inside at all
}

// (x, y) = rotmat2::fromAngle(-theta) * (vx, vy), VectorMath.h:912-936, with CUDA's sincosf (2 ulp); the slack of the
// result covers its difference to libm's cosf / sinf (1 ulp each way: < 2e-7 (|vx| + |vy|)) and the roundings of
// the two products and the sum, which may fall differently for a neighbouring (cos, sin) (< 1.5 ulp of |v|)
__device__ __forceinline__ void rotate_xy(float theta, float vx, float vy, float& rx, float& ry, float& slack)
{
    float c, sn;
    sincosf(-theta, &sn, &c);
    rx = __fadd_rn(__fmul_rn(c, vx), __fmul_rn(-sn, vy));
    ry = __fadd_rn(__fmul_rn(sn, vx), __fmul_rn(c, vy));
    slack = 5.0e-7f * (fabsf(vx) + fabsf(vy));
}

// rotate(q, v), VectorMath.h:810-818: (s^2 - v.v) b + (2 s) (v x b) + (2 v.b) v, one rounding per operation
__device__ __forceinline__ void quat_rotate(float s, float qx, float qy, float qz, float& x, float& y, float& z)
{
    float const vv = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    float const a = __fsub_rn(__fmul_rn(s, s), vv);
    float const two_s = __fmul_rn(2.0f, s);
    float const cx = __fsub_rn(__fmul_rn(qy, z), __fmul_rn(qz, y));
    float const cy = __fsub_rn(__fmul_rn(qz, x), __fmul_rn(qx, z));
    float const cz = __fsub_rn(__fmul_rn(qx, y), __fmul_rn(qy, x));
    float const vb = __fadd_rn(__fadd_rn(__fmul_rn(qx, x), __fmul_rn(qy, y)), __fmul_rn(qz, z));
    float const two_vb = __fmul_rn(2.0f, vb);
    float const ox = __fadd_rn(__fadd_rn(__fmul_rn(x, a), __fmul_rn(cx, two_s)), __fmul_rn(qx, two_vb));
    float const oy = __fadd_rn(__fadd_rn(__fmul_rn(y, a), __fmul_rn(cy, two_s)), __fmul_rn(qy, two_vb));
    float const oz = __fadd_rn(__fadd_rn(__fmul_rn(z, a), __fmul_rn(cz, two_s)), __fmul_rn(qz, two_vb));
    x = ox;
    y = oy;
    z = oz;
}

// One bond of the histogram: i = query point, j = point, (vx, vy, vz) = bond vector, dist = its length (R12)
template<int KIND, typename Count>
__device__ __forceinline__ void pmft3_bond(const Pmft3Args& a, uint32_t i, uint32_t j, float vx, float vy, float vz,
                                           float dist, Count count)
{
    if (KIND == FGPU_PMFT_XYZ)
    {
        float4 const q = a.query_quats[i]; // (s, x, y, z); conj: (s, -v), VectorMath.h:765
        float x = vx, y = vy, z = vz;
        quat_rotate(q.x, -q.y, -q.z, -q.w, x, y, z);
        for (uint32_t e = 0; e < a.n_equiv; ++e)
        {
            float4 const eq = a.equiv_quats[e];
            float ex = x, ey = y, ez = z;
            quat_rotate(eq.x, eq.y, eq.z, eq.w, ex, ey, ez);
            count(axis_bin(a.a0, ex), axis_bin(a.a1, ey), axis_bin(a.a2, ez));
        }
        return;
    }
    bool sure = true;
    int b0, b1, b2;
    if (KIND == FGPU_PMFT_XYT || KIND == FGPU_PMFT_XY)
    {
        float rx, ry, slack;
        rotate_xy(a.query_orientations[i], vx, vy, rx, ry, slack);
        b0 = slack_bin(a.a0, rx, slack, &sure);
        b1 = slack_bin(a.a1, ry, slack, &sure);
        b2 = 0;
        if (KIND == FGPU_PMFT_XYT && b0 >= 0 && b1 >= 0)
        {
            bool sure_t = true;
            b2 = angle_bin(a.a2, a.orientations[j], -vy, -vx, &sure_t);
            sure = sure && sure_t;
        }
    }
    else
    {
        bool sure1 = true, sure2 = true;
        b0 = axis_bin(a.a0, dist);
        b1 = angle_bin(a.a1, a.orientations[j], vy, vx, &sure1);
        b2 = angle_bin(a.a2, a.query_orientations[i], -vy, -vx, &sure2);
        sure = sure1 && sure2;
    }
    if (b0 < 0 || (KIND != FGPU_PMFT_R12 && b1 < 0))
    {
        return; // surely outside an axis
    }
    if (sure)
    {
        count(b0, b1, b2);
    }
    else
    {
        uint32_t const slot = atomicAdd(a.deferred_count, 1U);
        if (slot < a.deferred_cap)
        {
            a.deferred[slot] = make_uint4(i, j, __float_as_uint(vx), __float_as_uint(vy));
            if (KIND == FGPU_PMFT_R12)
            {
                a.deferred_dist[slot] = dist;
            }
        }
    }
}

// DSMEM helpers (sm_90+ thread-block clusters): the shared-memory window of CTA `rank` of this cluster, and a
// fire-and-forget add into it.
__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add_cluster(const uint32_t* local_word, uint32_t rank, uint32_t value)
{
    uint32_t const local = (uint32_t) __cvta_generic_to_shared(local_word);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(value) : "memory");
}

// ROWS = false: one thread per bond of a NeighborList.  ROWS = true: the bonds are still in the search's bag (no
// NeighborList was built): a group of lanes per query row, over the row's records {bond vector, bits(point index)}.
// Where the histogram lives (a.use_shared): 1 = in the block's shared memory (up to 10 k bins); 0 = global memory,
// one L2 atomic per bond (the 65-90 k bin histograms of PMFTXYT / PMFTR12 / PMFTXYZ: 39 M atomics per frame at half
// of L2's atomic throughput, profiles/ncu_r1_v9_summary.md); 2 = CLUSTER: spread over the shared memories of a
// thread-block cluster -- CTA c of the cluster owns bins [c, c + 1) * a.slice and every CTA adds into the owner's
// slice through distributed shared memory (red.shared::cluster), flushed once per CTA at the end.  Measured on B200
// (8 CTAs x 45 KB, 4 clusters' worth of CTAs per SM): PMFTXYT 1.39 ms against 0.60 ms with global atomics, PMFTR12
// 1.78 against 0.73, PMFTXYZ 0.20 against 0.16 -- seven of eight adds cross the cluster at DSMEM latency and L2's
// atomic units outrun them.  Global atomics stay the default; the cluster path is an alternative the tests keep honest
// (fgpu_ctx_set_tuning "pmft_cluster").
template<int KIND, bool ROWS, bool CLUSTER = false> __global__ void __launch_bounds__(256) k_pmft3(Pmft3Args a)
{
    extern __shared__ uint32_t p3_hist[];
    uint32_t const n_bins = a.a0.bins * a.a1.bins * a.a2.bins;
    // CLUSTER: this CTA's slice
    uint32_t const my_first = CLUSTER ? cluster_rank() * a.slice : 0U;
    uint32_t const my_bins = CLUSTER ? (my_first < n_bins ? min(a.slice, n_bins - my_first) : 0U) : n_bins;
    if (CLUSTER || a.use_shared)
    {
        for (uint32_t b = threadIdx.x; b < my_bins; b += blockDim.x)
        {
            p3_hist[b] = 0;
        }
        if (CLUSTER)
        {
            cluster_barrier(); // every slice is clear before anybody adds into it
        }
        else
        {
            __syncthreads();
        }
    }
    uint32_t* const h = a.use_shared ? p3_hist : a.hist;
    auto count = [&](int b0, int b1, int b2) {
        if (b0 >= 0 && b1 >= 0 && b2 >= 0)
        {
            uint32_t const bin = ((uint32_t) b0 * a.a1.bins + (uint32_t) b1) * a.a2.bins + (uint32_t) b2;
            if (CLUSTER)
            {
                uint32_t const owner = bin / a.slice;
                red_add_cluster(p3_hist + (bin - owner * a.slice), owner, 1U);
            }
            else
            {
                atomicAdd(&h[bin], 1U);
            }
        }
    };
    if (ROWS)
    {
        // a.group (4, 8 or 32) lanes share a row: rows of a few dozen bonds would leave most of a whole warp idle
        uint32_t const g = a.group, sub = threadIdx.x & (g - 1U);
        uint32_t const groups = gridDim.x * blockDim.x / g;
        for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) / g; row < a.n_rows; row += groups)
        {
            uint32_t const n = a.row_counts[row], start = a.row_bag_start[row];
            for (uint32_t k = sub; k < n; k += g)
            {
                float4 const r = a.bag[start + k];
                float const dist = KIND == FGPU_PMFT_R12 ? __fsqrt_rn(dot_exact(r.x, r.y, r.z)) : 0.0f; // NeighborBond.h:41-44
                pmft3_bond<KIND>(a, row, __float_as_uint(r.w), r.x, r.y, r.z, dist, count);
            }
        }
    }
    else
    {
        for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < a.n_bonds; k += (uint64_t) gridDim.x * blockDim.x)
        {
            uint2 const ij = reinterpret_cast<const uint2*>(a.neighbors)[k];
            float const vz = KIND == FGPU_PMFT_XYZ ? a.vectors[3 * k + 2] : 0.0f;
            float const dist = KIND == FGPU_PMFT_R12 ? a.distances[k] : 0.0f;
            pmft3_bond<KIND>(a, ij.x, ij.y, a.vectors[3 * k], a.vectors[3 * k + 1], vz, dist, count);
        }
    }
    if (CLUSTER)
    {
        cluster_barrier(); // every add of every CTA of the cluster has landed; from here on a CTA reads only its own slice
        for (uint32_t b = threadIdx.x; b < my_bins; b += blockDim.x)
        {
            if (p3_hist[b] != 0)
            {
                atomicAdd(&a.hist[my_first + b], p3_hist[b]);
            }
        }
    }
    else if (a.use_shared)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x)
        {
            if (p3_hist[b] != 0)
            {
                atomicAdd(&a.hist[b], p3_hist[b]);
            }
        }
    }
}

// resident histogram += the histogram of one frame
__global__ void __launch_bounds__(256) k_add_hist(const uint32_t* __restrict__ frame, uint32_t n, uint32_t* __restrict__ hist)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && frame[i] != 0)
    {
        hist[i] += frame[i];
    }
}

// BondOrder::accumulate: the bond vector (or, mode oocd, the z director of the query particle) rotated as the mode
// asks, then theta = modulusPositive(atan2f(v.y, v.x), 2 pi) on RegularAxis(n_theta, 0, 2 pi) and
// phi = acosf(v.z / sqrt(v.v)) on RegularAxis(n_phi, 0, pi).  Both angles are bracketed like the PMFT angles: CUDA's
// atan2f / acosf of the reference's float arguments, bins accepted away from the bin edges, the rest left to the host.
template<typename Hist>
__device__ __forceinline__ void bond_order_bond(const BondOrderArgs& a, uint32_t i, uint32_t j, float vx, float vy, float vz,
                                                Hist* h)
{
    float x = vx, y = vy, z = vz;
    if (a.mode != FGPU_BOND_ORDER_BOD)
    {
        float4 const rq = a.orientations[j];      // ref_q = orientations[point], BondOrder.cc:108
        float4 const q = a.query_orientations[i]; // q = query_orientations[query point], :110
        if (a.mode == FGPU_BOND_ORDER_OOCD)
        {
            x = 0.0f;
            y = 0.0f;
            z = 1.0f;
            quat_rotate(q.x, q.y, q.z, q.w, x, y, z); // :127-130
        }
        quat_rotate(rq.x, -rq.y, -rq.z, -rq.w, x, y, z); // rotate(conj(ref_q), .), :116, :123, :133
        if (a.mode == FGPU_BOND_ORDER_OBCD)
        {
            quat_rotate(q.x, q.y, q.z, q.w, x, y, z); // :117
        }
    }
    // theta: the chain of angle_bin with orientation - d replaced by d itself (t grows with d here)
    float const d = atan2f(y, x);
    float const theta = mod_two_pi(d);
    int const bt0 = axis_bin(a.at, theta);
    float const th_lo = mod_two_pi(d - kAngleMargin), th_hi = mod_two_pi(d + kAngleMargin);
    bool sure = axis_bin(a.at, th_lo) == bt0 && axis_bin(a.at, th_hi) == bt0 && th_lo <= theta && theta <= th_hi
        && theta > kEdgeGuard && theta < kTwoPi - kEdgeGuard;
    // phi: the argument is float arithmetic (one division, one square root), acosf is libm's; the bin is a
    // monotone function of phi
    float const c = __fdiv_rn(z, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
    float const phi = acosf(c);
    int const bp0 = axis_bin(a.ap, phi);
    sure = sure && axis_bin(a.ap, phi - kAngleMargin) == bp0 && axis_bin(a.ap, phi + kAngleMargin) == bp0
        && phi > kEdgeGuard && phi < a.ap.r_max - kEdgeGuard;
    if (sure)
    {
        int const bt = axis_bin(a.at, theta), bp = axis_bin(a.ap, phi);
        if (bt >= 0 && bp >= 0)
        {
            atomicAdd(&h[(uint32_t) bt * a.ap.bins + (uint32_t) bp], 1U);
        }
    }
    else
    {
        uint32_t const slot = atomicAdd(a.deferred_count, 1U);
        if (slot < a.deferred_cap)
        {
            a.deferred[slot] = make_uint4(i, j, __float_as_uint(vx), __float_as_uint(vy));
            a.deferred_z[slot] = vz;
        }
    }
}

// ROWS = false: one thread per bond of a NeighborList; ROWS = true: a.group lanes per query row of the search's bag
template<bool ROWS> __global__ void __launch_bounds__(256) k_bond_order(BondOrderArgs a)
{
    extern __shared__ uint32_t bo_hist[];
    uint32_t const n_bins = a.at.bins * a.ap.bins;
    if (a.use_shared)
    {
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x)
        {
            bo_hist[b] = 0;
        }
        __syncthreads();
    }
    uint32_t* const h = a.use_shared ? bo_hist : a.hist;
    if (ROWS)
    {
        uint32_t const g = a.group, sub = threadIdx.x & (g - 1U);
        uint32_t const groups = gridDim.x * blockDim.x / g;
        for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) / g; row < a.n_rows; row += groups)
        {
            uint32_t const n = a.row_counts[row], start = a.row_bag_start[row];
            for (uint32_t k = sub; k < n; k += g)
            {
                float4 const r = a.bag[start + k];
                bond_order_bond(a, row, __float_as_uint(r.w), r.x, r.y, r.z, h);
            }
        }
    }
    else
    {
        for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < a.n_bonds; k += (uint64_t) gridDim.x * blockDim.x)
        {
            uint2 const ij = reinterpret_cast<const uint2*>(a.neighbors)[k];
            bond_order_bond(a, ij.x, ij.y, a.vectors[3 * k], a.vectors[3 * k + 1], a.vectors[3 * k + 2], h);
        }
    }
    if (a.use_shared)
    {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x)
        {
            if (bo_hist[b] != 0)
            {
                atomicAdd(&a.hist[b], bo_hist[b]);
            }
        }
    }
}

// the host's share: one count per listed bin (0xffffffff: the bond fell outside the axes)
__global__ void __launch_bounds__(256) k_add_bins(const uint32_t* __restrict__ bins, uint32_t n, uint32_t* __restrict__ hist)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        uint32_t const b = bins[i];
        if (b != 0xffffffffU)
        {
            atomicAdd(&hist[b], 1U);
        }
    }
}

} // namespace

template<int KIND, bool ROWS> void launch_pmft3_cluster(fgpu_ctx* ctx, const Pmft3Args& a, unsigned blocks, int cluster, size_t dyn)
{
    auto kern = k_pmft3<KIND, ROWS, true>;
    if (dyn > 48 * 1024)
    {
        FGPU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned) cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    FGPU_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
}

void launch_pmft3(fgpu_ctx* ctx, int kind, Pmft3Args a)
{
    bool const rows = a.bag != nullptr;
    if (rows ? a.n_rows == 0 : a.n_bonds == 0)
    {
        return;
    }
    uint32_t const n_bins = a.a0.bins * a.a1.bins * a.a2.bins;
    size_t const smem = (size_t) n_bins * sizeof(uint32_t);
    a.use_shared = smem <= 40 * 1024 ? 1 : 0;
    a.slice = n_bins;
    // beyond one block's shared memory: a cluster of 2, 4 or 8 CTAs holds the histogram in slices of at most 48 KB
    // where that is possible (several clusters per SM), else of up to 200 KB
    int cluster = 0;
    if (!a.use_shared && ctx->tune_pmft_cluster != 0)
    {
        for (int c = 2; c <= 8 && cluster == 0; c *= 2)
        {
            if ((smem + c - 1) / c <= 48 * 1024)
            {
                cluster = c;
            }
        }
        if (cluster == 0 && (smem + 7) / 8 <= 200 * 1024)
        {
            cluster = 8;
        }
    }
    size_t dyn = a.use_shared ? smem : 0;
    uint64_t const want = rows ? ((uint64_t) a.n_rows * a.group + 255) / 256 : (a.n_bonds + 255) / 256;
    unsigned blocks = (unsigned) std::min<uint64_t>(want, (uint64_t) ctx->sm_count * 8U);
    if (cluster != 0)
    {
        a.slice = (n_bins + (uint32_t) cluster - 1U) / (uint32_t) cluster;
        dyn = (size_t) a.slice * sizeof(uint32_t);
        a.use_shared = 2;
        unsigned const per_sm = (unsigned) std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (dyn + 1024)));
        blocks = (unsigned) std::min<uint64_t>(want, (uint64_t) ctx->sm_count * per_sm);
        blocks = std::max(1U, (blocks + (unsigned) cluster - 1U) / (unsigned) cluster) * (unsigned) cluster;
    }
    {
        KernelScope ks(ctx, rows ? "pmft3_rows" : "pmft3");
#define FGPU_PMFT3_LAUNCH(KIND)                                                                                  \
    if (cluster != 0)                                                                                            \
    {                                                                                                            \
        if (rows)                                                                                                \
            launch_pmft3_cluster<KIND, true>(ctx, a, blocks, cluster, dyn);                                      \
        else                                                                                                     \
            launch_pmft3_cluster<KIND, false>(ctx, a, blocks, cluster, dyn);                                     \
    }                                                                                                            \
    else if (rows)                                                                                               \
    {                                                                                                            \
        k_pmft3<KIND, true><<<blocks, 256, dyn, ctx->stream>>>(a);                                               \
    }                                                                                                            \
    else                                                                                                         \
    {                                                                                                            \
        k_pmft3<KIND, false><<<blocks, 256, dyn, ctx->stream>>>(a);                                              \
    }
        switch (kind)
        {
        case FGPU_PMFT_XYZ:
            FGPU_PMFT3_LAUNCH(FGPU_PMFT_XYZ)
            break;
        case FGPU_PMFT_XYT:
            FGPU_PMFT3_LAUNCH(FGPU_PMFT_XYT)
            break;
        case FGPU_PMFT_XY:
            FGPU_PMFT3_LAUNCH(FGPU_PMFT_XY)
            break;
        default:
            FGPU_PMFT3_LAUNCH(FGPU_PMFT_R12)
            break;
        }
#undef FGPU_PMFT3_LAUNCH
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_add_hist(fgpu_ctx* ctx, const uint32_t* frame, uint32_t n, uint32_t* hist)
{
    {
        KernelScope ks(ctx, "pmft_add_hist");
        k_add_hist<<<(n + 255) / 256, 256, 0, ctx->stream>>>(frame, n, hist);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_bond_order(fgpu_ctx* ctx, BondOrderArgs a)
{
    bool const rows = a.bag != nullptr;
    if (rows ? a.n_rows == 0 : a.n_bonds == 0)
    {
        return;
    }
    size_t const smem = (size_t) a.at.bins * a.ap.bins * sizeof(uint32_t);
    a.use_shared = smem <= 40 * 1024 ? 1 : 0;
    uint64_t const want = rows ? ((uint64_t) a.n_rows * a.group + 255) / 256 : (a.n_bonds + 255) / 256;
    unsigned const blocks = (unsigned) std::min<uint64_t>(want, (uint64_t) ctx->sm_count * 8U);
    {
        KernelScope ks(ctx, rows ? "bond_order_rows" : "bond_order");
        if (rows)
        {
            k_bond_order<true><<<blocks, 256, a.use_shared ? smem : 0, ctx->stream>>>(a);
        }
        else
        {
            k_bond_order<false><<<blocks, 256, a.use_shared ? smem : 0, ctx->stream>>>(a);
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_add_bins(fgpu_ctx* ctx, const uint32_t* bins, uint32_t n, uint32_t* hist)
{
    if (n == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "pmft_add_bins");
        k_add_bins<<<(n + 255) / 256, 256, 0, ctx->stream>>>(bins, n, hist);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

} // namespace fgpu
