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
// atan2f is the host libm's in the reference and CUDA's differs from it in the last place (both are within 2-3 ulp of
// the true value).  The angle bin only depends on those last places when t lies within a few float ulps of a bin
// edge, so the kernel evaluates CUDA's atan2f, follows the reference's float chain from it, and accepts the bin only
// if t is farther from every bin edge than any such difference can move it (kAngleMargin, several times the two error
// bounds plus the rounding of the chain).  The few bonds inside a margin (~3 in 10^4) are written to a list and
// binned by the host with its libm (capi.cu): counts are bit-identical to the reference's, and no libm call runs per
// bond.  cosf / sinf of the query particle's angle (rotmat2::fromAngle, VectorMath.h:912-921) get the same treatment:
// CUDA's sincosf, and the rotated coordinate must clear every bin edge by more than a few last places of (cos, sin)
// can move it -- the host evaluates libm only for the bonds that do not (it used to for every particle).  (Double
// precision atan2 / sincos were tried first: 390 instructions per bond, the kernel issue-bound at 0.59 ms per frame.)
#include "internal.h"
#include "pair_math.cuh"

namespace fgpu {

namespace {

constexpr float kTwoPi = 6.28318548202514648f; // (float) (2.0 * M_PI), Box.h:24
constexpr float kAngleMargin = 1.0e-5f;        // >> (2 + 3) ulp of the two atan2f at pi (1.2e-6) + the chain's roundings (1e-6)

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
// last-place difference in atan2f could change it.
__device__ __forceinline__ int angle_bin(const AxisDev& axis, float orientation, float y, float x, bool* sure)
{
    float const d = atan2f(y, x);
    float const t = mod_two_pi(__fsub_rn(orientation, d));
    float const u = __fmul_rn(t, axis.inv_width);
    float const frac = u - floorf(u);
    float const margin = kAngleMargin * axis.inv_width + 1.0e-6f * (u + 1.0f);
    // near 0 or 2 pi the wrap of the modulus sits on a bin edge too; a NaN orientation is never sure
    *sure = frac > margin && frac < 1.0f - margin && t > kAngleMargin && t < kTwoPi - kAngleMargin;
    return axis_bin(axis, t);
}

// Bin of `value` on a RegularAxis when `value` is only known to +-slack: *sure = false if some value within the slack
// falls into another bin (or on the other side of an end of the axis).
__device__ __forceinline__ int slack_bin(const AxisDev& axis, float value, float slack, bool* sure)
{
    float const u = (value - axis.r_min) * axis.inv_width;
    float const m = slack * axis.inv_width + 1.0e-6f * (fabsf(u) + 1.0f);
    float const n = (float) axis.bins;
    if (u < -m || u > n + m)
    {
        return -1; // outside for every value within the slack
    }
    float const frac = u - floorf(u);
    bool const clear = frac > m && frac < 1.0f - m && u > m && u < n - m;
    *sure = *sure && clear;
    int const bin = axis_bin(axis, value);
    return clear ? bin : max(bin, 0); // not clear: the host decides, also whether the bond is inside at all
}

// (x, y) = rotmat2::fromAngle(-theta) * (vx, vy), VectorMath.h:912-936, with CUDA's sincosf (2 ulp); the slack of the
// result covers its difference to libm's cosf / sinf (1 ulp) several times
__device__ __forceinline__ void rotate_xy(float theta, float vx, float vy, float& rx, float& ry, float& slack)
{
    float c, sn;
    sincosf(-theta, &sn, &c);
    rx = __fadd_rn(__fmul_rn(c, vx), __fmul_rn(-sn, vy));
    ry = __fadd_rn(__fmul_rn(sn, vx), __fmul_rn(c, vy));
    slack = 1.0e-6f * (fabsf(vx) + fabsf(vy)); // 3 ulp of (cos, sin) move a coordinate by < 2e-7 (|vx| + |vy|)
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

template<int KIND> __global__ void __launch_bounds__(256) k_pmft3(Pmft3Args a)
{
    extern __shared__ uint32_t p3_hist[];
    uint32_t const n_bins = a.a0.bins * a.a1.bins * a.a2.bins;
    if (a.use_shared)
    {
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x)
        {
            p3_hist[b] = 0;
        }
        __syncthreads();
    }
    uint32_t* const h = a.use_shared ? p3_hist : a.hist;
    auto count = [&](int b0, int b1, int b2) {
        if (b0 >= 0 && b1 >= 0 && b2 >= 0)
        {
            atomicAdd(&h[((uint32_t) b0 * a.a1.bins + (uint32_t) b1) * a.a2.bins + (uint32_t) b2], 1U);
        }
    };
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < a.n_bonds; k += (uint64_t) gridDim.x * blockDim.x)
    {
        uint2 const ij = reinterpret_cast<const uint2*>(a.neighbors)[k];
        float const vx = a.vectors[3 * k], vy = a.vectors[3 * k + 1];
        if (KIND == FGPU_PMFT_XYZ)
        {
            float4 const q = a.query_quats[ij.x]; // (s, x, y, z); conj: (s, -v), VectorMath.h:765
            float x = vx, y = vy, z = a.vectors[3 * k + 2];
            quat_rotate(q.x, -q.y, -q.z, -q.w, x, y, z);
            for (uint32_t e = 0; e < a.n_equiv; ++e)
            {
                float4 const eq = a.equiv_quats[e];
                float ex = x, ey = y, ez = z;
                quat_rotate(eq.x, eq.y, eq.z, eq.w, ex, ey, ez);
                count(axis_bin(a.a0, ex), axis_bin(a.a1, ey), axis_bin(a.a2, ez));
            }
            continue;
        }
        bool sure = true;
        int b0, b1, b2;
        if (KIND == FGPU_PMFT_XYT || KIND == FGPU_PMFT_XY)
        {
            float rx, ry, slack;
            rotate_xy(a.query_orientations[ij.x], vx, vy, rx, ry, slack);
            b0 = slack_bin(a.a0, rx, slack, &sure);
            b1 = slack_bin(a.a1, ry, slack, &sure);
            b2 = 0;
            if (KIND == FGPU_PMFT_XYT && b0 >= 0 && b1 >= 0)
            {
                bool sure_t = true;
                b2 = angle_bin(a.a2, a.orientations[ij.y], -vy, -vx, &sure_t);
                sure = sure && sure_t;
            }
        }
        else
        {
            bool sure1 = true, sure2 = true;
            b0 = axis_bin(a.a0, a.distances[k]);
            b1 = angle_bin(a.a1, a.orientations[ij.y], vy, vx, &sure1);
            b2 = angle_bin(a.a2, a.query_orientations[ij.x], -vy, -vx, &sure2);
            sure = sure1 && sure2;
        }
        if (b0 < 0 || (KIND != FGPU_PMFT_R12 && b1 < 0))
        {
            continue; // surely outside an axis
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
                a.deferred[slot] = make_uint4(ij.x, ij.y, __float_as_uint(vx), __float_as_uint(vy));
                if (KIND == FGPU_PMFT_R12)
                {
                    a.deferred_dist[slot] = a.distances[k];
                }
            }
        }
    }
    if (a.use_shared)
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

// BondOrder::accumulate: the bond vector (or, mode oocd, the z director of the query particle) rotated as the mode
// asks, then theta = modulusPositive(atan2f(v.y, v.x), 2 pi) on RegularAxis(n_theta, 0, 2 pi) and
// phi = acosf(v.z / sqrt(v.v)) on RegularAxis(n_phi, 0, pi).  Both angles are bracketed like the PMFT angles: CUDA's
// atan2f / acosf of the reference's float arguments, bins accepted away from the bin edges, the rest left to the host.
__global__ void __launch_bounds__(256) k_bond_order(BondOrderArgs a)
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
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < a.n_bonds; k += (uint64_t) gridDim.x * blockDim.x)
    {
        uint2 const ij = reinterpret_cast<const uint2*>(a.neighbors)[k];
        float const vx = a.vectors[3 * k], vy = a.vectors[3 * k + 1], vz = a.vectors[3 * k + 2];
        float x = vx, y = vy, z = vz;
        if (a.mode != FGPU_BOND_ORDER_BOD)
        {
            float4 const rq = a.orientations[ij.y];      // ref_q = orientations[point], BondOrder.cc:108
            float4 const q = a.query_orientations[ij.x]; // q = query_orientations[query point], :110
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
        // theta: the chain of angle_bin with orientation - d replaced by d itself
        float const d = atan2f(y, x);
        float const theta = mod_two_pi(d);
        float const ut = __fmul_rn(theta, a.at.inv_width);
        float const ft = ut - floorf(ut);
        float const mt = kAngleMargin * a.at.inv_width + 1.0e-6f * (ut + 1.0f);
        bool sure = ft > mt && ft < 1.0f - mt && theta > kAngleMargin && theta < kTwoPi - kAngleMargin;
        // phi: the argument is float arithmetic (one division, one square root), acosf is libm's
        float const c = __fdiv_rn(z, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
        float const phi = acosf(c);
        float const up = __fmul_rn(phi, a.ap.inv_width);
        float const fp = up - floorf(up);
        float const mp = kAngleMargin * a.ap.inv_width + 1.0e-6f * (up + 1.0f);
        sure = sure && fp > mp && fp < 1.0f - mp && phi > kAngleMargin && phi < a.ap.r_max - kAngleMargin;
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
                a.deferred[slot] = make_uint4(ij.x, ij.y, __float_as_uint(vx), __float_as_uint(vy));
                a.deferred_z[slot] = vz;
            }
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

// the host's share: one count per listed bin
__global__ void __launch_bounds__(256) k_add_bins(const uint32_t* __restrict__ bins, uint32_t n, uint32_t* __restrict__ hist)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        atomicAdd(&hist[bins[i]], 1U);
    }
}

} // namespace

void launch_pmft3(fgpu_ctx* ctx, int kind, Pmft3Args a)
{
    if (a.n_bonds == 0)
    {
        return;
    }
    size_t const smem = (size_t) a.a0.bins * a.a1.bins * a.a2.bins * sizeof(uint32_t);
    a.use_shared = smem <= 40 * 1024 ? 1 : 0;
    size_t const dyn = a.use_shared ? smem : 0;
    unsigned const blocks = (unsigned) std::min<uint64_t>((a.n_bonds + 255) / 256, (uint64_t) ctx->sm_count * 8U);
    {
        KernelScope ks(ctx, "pmft3");
        switch (kind)
        {
        case FGPU_PMFT_XYZ:
            k_pmft3<FGPU_PMFT_XYZ><<<blocks, 256, dyn, ctx->stream>>>(a);
            break;
        case FGPU_PMFT_XYT:
            k_pmft3<FGPU_PMFT_XYT><<<blocks, 256, dyn, ctx->stream>>>(a);
            break;
        case FGPU_PMFT_XY:
            k_pmft3<FGPU_PMFT_XY><<<blocks, 256, dyn, ctx->stream>>>(a);
            break;
        default:
            k_pmft3<FGPU_PMFT_R12><<<blocks, 256, dyn, ctx->stream>>>(a);
            break;
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_bond_order(fgpu_ctx* ctx, BondOrderArgs a)
{
    if (a.n_bonds == 0)
    {
        return;
    }
    size_t const smem = (size_t) a.at.bins * a.ap.bins * sizeof(uint32_t);
    a.use_shared = smem <= 40 * 1024 ? 1 : 0;
    unsigned const blocks = (unsigned) std::min<uint64_t>((a.n_bonds + 255) / 256, (uint64_t) ctx->sm_count * 8U);
    {
        KernelScope ks(ctx, "bond_order");
        k_bond_order<<<blocks, 256, a.use_shared ? smem : 0, ctx->stream>>>(a);
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
