// Steinhardt q_l: per-particle accumulation of spherical harmonics over the neighbour list.
//
// Replaces Steinhardt::baseCompute + normalizeSystem (freud/order/Steinhardt.cc:120-222, 291-327) and the
// fsph::PointSPHEvaluator<float> recurrence it calls (extern/fsph/src/spherical_harmonics.hpp:155-290,
// value iterator :78-93, Condon-Shortley sign Steinhardt.cc:47-49).
//
// One thread per particle.  For every bond the thread recomputes delta = Box::wrap(p_j - p_i) (Steinhardt.cc:155;
// always the WRAP arithmetic, whoever found the bond), takes cos(theta) = clamp(z / d) with the list's
// distance d and (cos, sin)(phi) from (x, y), runs the Jacobi recurrence for m = 0..l, rotates exp(i m phi)
// from m to m + 1 and accumulates Y_lm for m >= 0 in registers.  Negative m needs no accumulator: Y_{l,-m} = conj(Y_{l,m}) * (-1)^m term by term, and both
// conj and the sign commute exactly with float summation, so q_{l,-m} is derived at the end.
// Tolerance vs the reference: 1e-5 (the reference goes through libm's atan2f/acosf/sinf/cosf/exp, this kernel
// through algebraically identical component formulas; both are a few ulp from the exact value).
//
// Options (Steinhardt.cc:224-289, 329-359; Wigner3j.cc:22-57) run as follow-up kernels over the per-particle
// q_lm array left in device memory by the base kernel:
//   k_steinhardt_average  second-shell average: q_lm(i) summed over the neighbours of i in bond order, plus
//                         i itself, divided by (bonds + 1); averaged q_l; system sums of the averaged q_lm.
//   k_steinhardt_wl       third-order invariant: sum over the Wigner 3j table of q_{l m1} q_{l m2} q_{l m3}
//                         (real part), optionally scaled by (sqrt(4 pi / (2l+1)) / q_l)^3.
// The 3j coefficients are computed on the host (Racah's formula with an exact integer sum) in the reference's
// table order -- m1 = -l..l, m2 = max(-l-m1, -l)..min(l-m1, l) -- and used as float, like upstream.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

#include "internal.h"

namespace fgpu {

namespace {

constexpr int kSphLmax = 32;
// Y = P / sqrt(2 pi) upstream (a double division per value, spherical_harmonics.hpp:85-91).  The Jacobi recurrence is
// linear in its first column, so the factor is folded into c_jac0 on the host (float(jac0 / sqrt(2 pi))): one
// rounding in a different place, ~1e-7 relative, and no float<->double conversion per (bond, m) -- those conversions
// and the wrap's truncations kept the XU pipe 52 % busy, the most loaded unit of the kernel (profiles, k19 capture).
constexpr double kInvSqrt2Pi = 1.0 / 2.5066282746310002;
constexpr int kThreads = 128;
constexpr int kStageBonds = 1792; // bonds staged per pass by k_steinhardt_single: 5 floats each, 35 KB
constexpr int kStageBatch = 7;    // gathers a thread keeps in flight while staging
static_assert(kStageBonds % (kThreads * kStageBatch) == 0, "staging passes must tile the chunk");

// positions padded to 16 bytes in the original order: one aligned load (one sector) per gathered neighbour
__global__ void __launch_bounds__(256) k_pad_positions(const float* __restrict__ xyz, uint32_t n, float4* __restrict__ out)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        out[i] = make_float4(xyz[3 * (size_t) i], xyz[3 * (size_t) i + 1], xyz[3 * (size_t) i + 2], 0.0f);
    }
}

// recurrence prefactors for lmax, laid out as the reference does: [0, lmax*(lmax+1)) first kind,
// [lmax*(lmax+1), 2*lmax*(lmax+1)) second kind; index lmax*m + (l-1)
__constant__ float c_pref[2 * (kSphLmax + 1) * kSphLmax];
__constant__ float c_jac0[kSphLmax + 1]; // jacobi[m][0] = 1/sqrt(2) * prod_{k<=m} sqrt(1 + 1/(2k)), times 1/sqrt(2 pi)
__constant__ int c_ls[kSphLmax + 1];     // requested l values
__constant__ int c_l_slot[kSphLmax + 1]; // l -> index into the request list, or -1
__constant__ int c_acc_off[kSphLmax + 1]; // request index -> offset of its (l+1) complex accumulators
__constant__ int c_out_off[kSphLmax + 1]; // request index -> offset (in complex elements per particle sum) of qlm block

struct Angles
{
    float sphi, cphi; // sin / cos of the polar angle ("phi" in fsph, theta in freud)
    float caz, saz;   // cos / sin of the azimuth
};

// One bond as it comes out of memory: the neighbour's position, the list's distance and weight.  Loading the next
// bond before the current one is evaluated keeps two dependent gathers (index -> position) in flight per thread.
struct BondIn
{
    float px, py, pz, dist, w;
};

__device__ __forceinline__ BondIn load_bond(const SteinhardtArgs& a, uint32_t b)
{
    uint32_t const j = a.neighbors[2 * (size_t) b + 1];
    BondIn in;
    in.px = __ldg(a.xyz + 3 * (size_t) j);
    in.py = __ldg(a.xyz + 3 * (size_t) j + 1);
    in.pz = __ldg(a.xyz + 3 * (size_t) j + 2);
    in.dist = a.distances[b];
    in.w = a.weighted ? a.weights[b] : 1.0f;
    return in;
}

__device__ __forceinline__ Angles vector_angles(float dx, float dy, float dz, float dist);

__device__ __forceinline__ Angles bond_angles(const SteinhardtArgs& args, float rx0, float ry0, float rz0, float px,
                                              float py, float pz, float dist)
{
    float dx, dy, dz;
    wrap_quick(args.box, args.rcp_lx, args.rcp_ly, args.rcp_lz, __fsub_rn(px, rx0), __fsub_rn(py, ry0), __fsub_rn(pz, rz0), dx, dy,
               dz);
    return vector_angles(dx, dy, dz, dist);
}

// sines and cosines of the two angles of a bond vector of length dist
__device__ __forceinline__ Angles vector_angles(float dx, float dy, float dz, float dist)
{
    // The reference takes phi = atan2(y, x) and theta = acos(clamp(z / d)) (Steinhardt.cc:161-174) and the
    // evaluator immediately goes back to sin/cos of both (spherical_harmonics.hpp:239-244, :272-281).  Here the
    // sines and cosines come straight from the components -- cos(theta) = clamp(z / d), sin(theta) =
    // sqrt(1 - cos^2), (cos, sin)(phi) = (x, y) / hypot(x, y) -- which agrees with the libm round trip to a few
    // ulp (the documented tolerance is 1e-5) and costs no transcendental at all.
    float ct = fmaxf(-1.0f, fminf(__fdiv_rn(dz, dist), 1.0f));
    if (dist == 0.0f)
    {
        ct = 1.0f; // theta = 0, Steinhardt.cc:171-174
    }
    float const rho_sq = dx * dx + dy * dy;
    float const inv_rho = rho_sq > 0.0f ? rsqrtf(rho_sq) : 0.0f;
    Angles a;
    a.cphi = ct;
    a.sphi = sqrtf(fmaxf(0.0f, (1.0f - ct) * (1.0f + ct)));
    a.caz = rho_sq > 0.0f ? dx * inv_rho : 1.0f; // atan2(0, 0) = 0
    a.saz = rho_sq > 0.0f ? dy * inv_rho : 0.0f;
    return a;
}

// Sum of v over the block, written to *dst by thread 0; every thread of the block must call it.
__device__ __forceinline__ void block_sum_to(double v, double* dst)
{
    __shared__ double s_w[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        v += __shfl_down_sync(0xffffffffU, v, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        s_w[threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w)
        {
            t += s_w[w];
        }
        *dst = t;
    }
    __syncthreads();
}

// One bond's Y_lm, m = 0..L, added to the particle's accumulators (the loop body of Steinhardt::baseCompute,
// Steinhardt.cc:141-193, with fsph's evaluator unrolled into registers)
template<int L> __device__ __forceinline__ void accumulate_ylm(const Angles& ang, float w, float* re, float* im)
{
    float sinpow = 1.0f;
    float c = 1.0f, s = 0.0f; // exp(i m phi) by rotation, m = 0
#pragma unroll
    for (int m = 0; m <= L; ++m)
    {
        // Jacobi recurrence in l' = l - m up to L - m (spherical_harmonics.hpp:246-270)
        float j_prev = 0.0f, j_cur = c_jac0[m];
#pragma unroll
        for (int lp = 1; lp <= L - m; ++lp)
        {
            float next = ang.cphi * c_pref[L * m + (lp - 1)] * j_cur;
            if (lp >= 2)
            {
                next += c_pref[L * (L + 1) + L * m + (lp - 1)] * j_prev;
            }
            j_prev = j_cur;
            j_cur = next;
        }
        float const legendre = sinpow * j_cur;                      // :272-281
        float const amp = legendre; // / sqrt(2 pi) is folded into c_jac0
        float const phase = (m & 1) ? -1.0f : 1.0f; // Steinhardt.cc:47-49
        re[m] += w * (phase * (amp * c));           // exp(i m theta), :239-244
        im[m] += w * (phase * (amp * s));
        sinpow *= ang.sphi;
        float const cn = c * ang.caz - s * ang.saz;
        s = s * ang.caz + c * ang.saz;
        c = cn;
    }
}

// normalise, q_l, outputs and the block's partial sums of the system q_lm; every thread of the block must call it
template<int L>
__device__ __forceinline__ void finish_particle(const SteinhardtArgs& a, uint32_t i, bool active, float* re, float* im,
                                                float total_weight)
{
    // normalise, q_l, outputs (Steinhardt.cc:195-220)
    float const nf = (float) (4.0 * 3.14159265358979323846 / (2 * L + 1));
    float sum = 0.0f;
    double sys_re[L + 1], sys_im[L + 1];
#pragma unroll
    for (int m = 0; m <= L; ++m)
    {
        re[m] = re[m] / total_weight;
        im[m] = im[m] / total_weight;
        sys_re[m] = active ? (double) re[m] : 0.0;
        sys_im[m] = active ? (double) im[m] : 0.0;
    }
    if (active)
    {
        // m = 0..L, then -1..-L, as upstream orders them
#pragma unroll
        for (int m = 0; m <= L; ++m)
        {
            sum += re[m] * re[m] + im[m] * im[m];
        }
#pragma unroll
        for (int m = 1; m <= L; ++m)
        {
            sum += re[m] * re[m] + im[m] * im[m];
        }
        a.ql[i] = sqrtf(sum * nf);
        if (a.qlm != nullptr)
        {
            float2* out = reinterpret_cast<float2*>(a.qlm) + (size_t) i * (2 * L + 1);
#pragma unroll
            for (int m = 0; m <= L; ++m)
            {
                out[m] = make_float2(re[m], im[m]);
            }
#pragma unroll
            for (int m = 1; m <= L; ++m)
            {
                float const phase = (m & 1) ? -1.0f : 1.0f;
                out[L + m] = make_float2(phase * re[m], -(phase * im[m])); // conj without the sign
            }
        }
    }
    // System q_lm partial sums in fp64: warp shuffle, block sum in shared memory, one row of partials per block.
    // (One atomicAdd per warp and component on the 2(l+1) accumulators serialises in L2: 31 k warps x 14 hot
    // addresses cost ~0.25 ms at C3, more than the rest of the kernel.  k_sum_partials adds the rows up, in a
    // fixed order.)
    if (a.sys_partials != nullptr)
    {
        __shared__ double s_part[kThreads / 32][2 * (L + 1)];
#pragma unroll
        for (int m = 0; m <= L; ++m)
        {
            double vr = sys_re[m], vi = sys_im[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                vr += __shfl_down_sync(0xffffffffU, vr, o);
                vi += __shfl_down_sync(0xffffffffU, vi, o);
            }
            if ((threadIdx.x & 31) == 0)
            {
                s_part[threadIdx.x >> 5][2 * m] = vr;
                s_part[threadIdx.x >> 5][2 * m + 1] = vi;
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * (L + 1))
        {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w)
            {
                v += s_part[w][threadIdx.x];
            }
            a.sys_partials[(size_t) blockIdx.x * (2 * (L + 1)) + threadIdx.x] = v;
        }
    }
}

// ---- single l, everything unrolled into registers ---------------------------------------------------
template<int L> __global__ void __launch_bounds__(kThreads) k_steinhardt_single(SteinhardtArgs a)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    float re[L + 1], im[L + 1];
#pragma unroll
    for (int m = 0; m <= L; ++m)
    {
        re[m] = 0.0f;
        im[m] = 0.0f;
    }
    float total_weight = 0.0f;
    bool const active = i < a.n;
    // The bonds of the block's particles are one contiguous range of the list.  The block stages them chunk by
    // chunk in shared memory -- coalesced reads of the index and distance arrays, one independent gather of the
    // neighbour's position per thread -- and only then walks its rows: the per-row loop is a chain of dependent
    // loads otherwise (index -> position), twelve deep, with nothing to overlap it.
    __shared__ float s_px[kStageBonds], s_py[kStageBonds], s_pz[kStageBonds], s_dist[kStageBonds], s_w[kStageBonds];
    uint32_t const i0 = blockIdx.x * blockDim.x, i1 = min(i0 + blockDim.x, a.n);
    uint32_t const block_beg = a.row_start[i0], block_end = a.row_start[i1];
    uint32_t const beg = active ? a.row_start[i] : 0U, end = active ? a.row_start[i + 1] : 0U;
    float rx0 = 0.0f, ry0 = 0.0f, rz0 = 0.0f;
    if (active)
    {
        size_t const pi = (size_t) i + a.row_offset;
        rx0 = a.xyz[3 * pi];
        ry0 = a.xyz[3 * pi + 1];
        rz0 = a.xyz[3 * pi + 2];
    }
    for (uint32_t c0 = block_beg; c0 < block_end; c0 += kStageBonds)
    {
        uint32_t const c1 = min(c0 + (uint32_t) kStageBonds, block_end);
        __syncthreads();
        // seven bonds per thread at a time: all indices first, then all gathers (16-byte padded positions, one
        // sector each), then the stores -- the hardware issues in order, so a gather that is consumed right
        // away would leave one chain in flight per warp
#pragma unroll
        for (int h = 0; h < kStageBonds / kThreads / kStageBatch; ++h)
        {
            uint32_t jj[kStageBatch];
            float4 pp[kStageBatch];
#pragma unroll
            for (int k = 0; k < kStageBatch; ++k)
            {
                uint32_t const b = c0 + (uint32_t) (h * kStageBatch + k) * kThreads + threadIdx.x;
                jj[k] = b < c1 ? a.neighbors[2 * (size_t) b + 1] : 0U;
            }
#pragma unroll
            for (int k = 0; k < kStageBatch; ++k)
            {
                pp[k] = __ldg(a.xyz4 + jj[k]);
            }
#pragma unroll
            for (int k = 0; k < kStageBatch; ++k)
            {
                uint32_t const b = c0 + (uint32_t) (h * kStageBatch + k) * kThreads + threadIdx.x;
                if (b < c1)
                {
                    s_px[b - c0] = pp[k].x;
                    s_py[b - c0] = pp[k].y;
                    s_pz[b - c0] = pp[k].z;
                    s_dist[b - c0] = a.distances[b];
                    s_w[b - c0] = a.weighted ? a.weights[b] : 1.0f;
                }
            }
        }
        __syncthreads();
        uint32_t const lo = max(beg, c0), hi = min(end, c1);
        for (uint32_t b = lo; b < hi; ++b)
        {
            float const w = s_w[b - c0];
            Angles const ang = bond_angles(a, rx0, ry0, rz0, s_px[b - c0], s_py[b - c0], s_pz[b - c0], s_dist[b - c0]);
            accumulate_ylm<L>(ang, w, re, im);
            total_weight += w;
        }
    }
    finish_particle<L>(a, i, active, re, im, total_weight);
}

// ---- k nearest neighbours -> Y_lm in one kernel (BASELINE.json configs[2]) ------------------------------------------
// Steinhardt(l).compute(system, neighbors = {num_neighbors: k}) needs no NeighborList: the window search has left
// every row's hits in the bag (16-byte records {bond vector, point index}, knn2.cu), and all this compute wants from
// them is WHICH k points are nearest.  One thread per row: it streams its row of the bag, keeps the k smallest keys
// (bits(r_sq) << 32 | point index -- r_sq >= +0, so the bit pattern orders like the value, and ties at the k-th place
// fall to the point index as in k_knn_select and the oracle) in a sorted array that lives in registers (branch-free
// insertion: K'[i] = min(K[i], max(x, K[i-1]))), then gathers the k positions and runs the Y_lm recurrence over
// Box::wrap(p_j - p_i) exactly as k_steinhardt_single does -- the reference evaluates Y_lm on the wrapped difference
// (Steinhardt.cc:155), whose float32 round trip moves a component by a few 1e-6, so the bag's own (IMAGE-arithmetic)
// vector would drift from the reference by more than the 1e-5 this path promises.
// Gone: k_knn_select (a warp per row ranking, ordering and writing a 28 B/bond NeighborList: 336 MB at 1 M particles),
// the row-offset scan, and k_steinhardt_single's staging of that list.  A first fused version kept the warp-per-row
// selection and handed the vectors to a thread per row through shared memory: 437 us against 337 + 180 us for the two
// kernels it replaced -- the selection leaves a third of the lanes idle and pays ~150 instructions per row.
constexpr int kFusedMaxK = 16; // neighbours per row this route serves

struct KnnYlmArgs
{
    const float4* bag;
    const float4* bag2;        // rows searched again with a wider window (top bit of tmp_start)
    const uint32_t* tmp_start; // per row: offset of its hits in the bag
    const uint32_t* hits;      // per row: number of hits in the window
    uint32_t k;
};

template<int L, int KMAX> __global__ void __launch_bounds__(kThreads) k_knn_ylm(SteinhardtArgs a, KnnYlmArgs s)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    bool const active = i < a.n;
    uint32_t const n = active ? __ldg(s.hits + i) : 0U;
    uint32_t const ts = active ? __ldg(s.tmp_start + i) : 0U;
    const float4* __restrict__ const bag = (ts & kSecondBag) != 0 ? s.bag2 + (ts & ~kSecondBag) : s.bag + ts;
    unsigned long long key[KMAX];
#pragma unroll
    for (int q = 0; q < KMAX; ++q)
    {
        key[q] = ~0ULL;
    }
    for (uint32_t h = 0; h < n; ++h)
    {
        float4 const v = bag[h];
        unsigned long long const x
            = ((unsigned long long) __float_as_uint(dot_exact(v.x, v.y, v.z)) << 32) | __float_as_uint(v.w);
#pragma unroll
        for (int q = KMAX - 1; q >= 1; --q)
        {
            unsigned long long const up = key[q - 1] > x ? key[q - 1] : x;
            key[q] = key[q] < up ? key[q] : up;
        }
        key[0] = key[0] < x ? key[0] : x;
    }
    uint32_t const kept = min(min(n, s.k), (uint32_t) KMAX);
    float re[L + 1], im[L + 1];
#pragma unroll
    for (int m = 0; m <= L; ++m)
    {
        re[m] = 0.0f;
        im[m] = 0.0f;
    }
    float total_weight = 0.0f;
    float rx0 = 0.0f, ry0 = 0.0f, rz0 = 0.0f;
    if (active)
    {
        rx0 = a.xyz[3 * (size_t) i];
        ry0 = a.xyz[3 * (size_t) i + 1];
        rz0 = a.xyz[3 * (size_t) i + 2];
    }
    // The kept points' Y_lm in a ROLLED loop: unrolled over the KMAX slots the kernel was 76 KB of code (l = 6, k = 12)
    // and its warps, each at its own place, waited for instruction fetches more than for anything else (ncu:
    // stalled_no_instruction 5.1 per issue).  The sorted keys go to the thread's column of a shared-memory table so
    // that the loop can index them; the next point's position is requested while the current one is evaluated.
    __shared__ unsigned long long s_key[KMAX][kThreads];
#pragma unroll
    for (int q = 0; q < KMAX; ++q)
    {
        s_key[q][threadIdx.x] = key[q];
    }
    float4 p_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (kept != 0)
    {
        p_next = __ldg(a.xyz4 + (uint32_t) (key[0] & 0xffffffffULL));
    }
#pragma unroll 1
    for (uint32_t q = 0; q < kept; ++q)
    {
        unsigned long long const kq = s_key[q][threadIdx.x];
        float4 const p = p_next;
        if (q + 1 < kept)
        {
            p_next = __ldg(a.xyz4 + (uint32_t) (s_key[q + 1][threadIdx.x] & 0xffffffffULL));
        }
        float const dist = __fsqrt_rn(__uint_as_float((uint32_t) (kq >> 32))); // NeighborBond.h:41-44
        Angles const ang = bond_angles(a, rx0, ry0, rz0, p.x, p.y, p.z, dist);
        accumulate_ylm<L>(ang, 1.0f, re, im);
        total_weight += 1.0f;
    }
    finish_particle<L>(a, i, active, re, im, total_weight);
}

// Column sums of the per-block partials: one block per column, fixed summation order (bitwise reproducible).  1024
// threads: a column of 7 814 partials (C3) is eight independent loads per thread, all in flight at once -- with 256
// threads the kernel was thirty dependent rounds of DRAM latency on fourteen SMs (18 us).
constexpr int kSumThreads = 1024;
__global__ void __launch_bounds__(kSumThreads) k_sum_partials(const double* __restrict__ partials, uint32_t n_rows,
                                                              uint32_t width, double* __restrict__ out)
{
    __shared__ double s_sum[kSumThreads];
    uint32_t const col = blockIdx.x;
    double v = 0.0;
    for (uint32_t r = threadIdx.x; r < n_rows; r += blockDim.x)
    {
        v += partials[(size_t) r * width + col];
    }
    s_sum[threadIdx.x] = v;
    __syncthreads();
    for (int o = kSumThreads / 2; o > 0; o >>= 1)
    {
        if ((int) threadIdx.x < o)
        {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        out[col] += s_sum[0];
    }
}

// ---- generic: any set of l values up to kSphLmax, accumulators in local memory ------------------------
constexpr int kMaxAcc = 2 * 160; // floats; sum over requested l of (l + 1) complex numbers

__global__ void __launch_bounds__(kThreads) k_steinhardt_generic(SteinhardtArgs a, int lmax, int n_ls, int n_acc,
                                                                 int tot_m)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    bool const active = i < a.n; // idle threads stay for the block sums at the end
    float acc[kMaxAcc];
    for (int k = 0; k < 2 * n_acc; ++k)
    {
        acc[k] = 0.0f;
    }
    float total_weight = 0.0f;
    uint32_t const beg = active ? a.row_start[i] : 0U, end = active ? a.row_start[i + 1] : 0U;
    size_t const ii = active ? (size_t) i + a.row_offset : 0;
    float const rx0 = a.xyz[3 * ii], ry0 = a.xyz[3 * ii + 1], rz0 = a.xyz[3 * ii + 2];
    BondIn next = beg < end ? load_bond(a, beg) : BondIn {0, 0, 0, 0, 0};
    for (uint32_t b = beg; b < end; ++b)
    {
        BondIn const cur = next;
        if (b + 1 < end)
        {
            next = load_bond(a, b + 1);
        }
        float const w = cur.w;
        Angles const ang = bond_angles(a, rx0, ry0, rz0, cur.px, cur.py, cur.pz, cur.dist);
        float sinpow = 1.0f;
        float c = 1.0f, s = 0.0f;
        for (int m = 0; m <= lmax; ++m)
        {
            float const phase = (m & 1) ? -1.0f : 1.0f;
            float j_prev = 0.0f, j_cur = c_jac0[m];
            for (int lp = 0; lp <= lmax - m; ++lp)
            {
                if (lp >= 1)
                {
                    float next = ang.cphi * c_pref[lmax * m + (lp - 1)] * j_cur;
                    if (lp >= 2)
                    {
                        next += c_pref[lmax * (lmax + 1) + lmax * m + (lp - 1)] * j_prev;
                    }
                    j_prev = j_cur;
                    j_cur = next;
                }
                int const slot = c_l_slot[lp + m];
                if (slot >= 0)
                {
                    float const legendre = sinpow * j_cur;
                    float const amp = legendre; // / sqrt(2 pi) is folded into c_jac0
                    int const o = 2 * (c_acc_off[slot] + m);
                    acc[o] += w * (phase * (amp * c));
                    acc[o + 1] += w * (phase * (amp * s));
                }
            }
            sinpow *= ang.sphi;
            float const cn = c * ang.caz - s * ang.saz;
            s = s * ang.caz + c * ang.saz;
            c = cn;
        }
        total_weight += w;
    }
    for (int r = 0; r < n_ls; ++r)
    {
        int const l = c_ls[r];
        float const nf = (float) (4.0 * 3.14159265358979323846 / (2 * l + 1));
        float* q = acc + 2 * c_acc_off[r];
        float sum = 0.0f;
        for (int m = 0; m <= l; ++m)
        {
            q[2 * m] = q[2 * m] / total_weight;
            q[2 * m + 1] = q[2 * m + 1] / total_weight;
            sum += q[2 * m] * q[2 * m] + q[2 * m + 1] * q[2 * m + 1];
        }
        for (int m = 1; m <= l; ++m)
        {
            sum += q[2 * m] * q[2 * m] + q[2 * m + 1] * q[2 * m + 1];
        }
        if (active)
        {
            a.ql[(size_t) i * n_ls + r] = sqrtf(sum * nf);
        }
        if (active && a.qlm != nullptr)
        {
            // per-l blocks are concatenated: block r starts at n * c_out_off[r] complex elements
            float2* out = reinterpret_cast<float2*>(a.qlm) + (size_t) a.n * c_out_off[r] + (size_t) i * (2 * l + 1);
            for (int m = 0; m <= l; ++m)
            {
                out[m] = make_float2(q[2 * m], q[2 * m + 1]);
            }
            for (int m = 1; m <= l; ++m)
            {
                float const phase = (m & 1) ? -1.0f : 1.0f;
                out[l + m] = make_float2(phase * q[2 * m], -(phase * q[2 * m + 1]));
            }
        }
        if (a.sys_partials != nullptr)
        {
            // sys_qlm holds (2l+1) complex per l at complex offset c_out_off[r]; only m >= 0 is accumulated,
            // the host derives m < 0.  Block sums, one row of 2 tot_m partials per block (see k_sum_partials).
            double* const row = a.sys_partials + (size_t) blockIdx.x * (2 * tot_m);
            for (int m = 0; m <= l; ++m)
            {
                block_sum_to(active ? (double) q[2 * m] : 0.0, row + 2 * (c_out_off[r] + m));
                block_sum_to(active ? (double) q[2 * m + 1] : 0.0, row + 2 * (c_out_off[r] + m) + 1);
            }
        }
    }
}

// ---- second-shell average (Steinhardt::computeAve, Steinhardt.cc:224-289) ----------------------------------
// One thread per particle.  Only m >= 0 is accumulated: q_{l,-m} = (-1)^m conj(q_{l,m}) holds term by term and
// both conj and the sign commute exactly with float summation and division.
__global__ void __launch_bounds__(kThreads) k_steinhardt_average(SteinhardtAveArgs a, int n_ls)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    bool const active = i < a.n; // idle threads stay for the block sums
    uint32_t const beg = active ? a.row_start[i] : 0U, end = active ? a.row_start[i + 1] : 0U;
    float const count = (float) (end - beg + 1U); // neighborcount starts at 1, Steinhardt.cc:244
    int const tot_m = c_out_off[n_ls - 1] + 2 * c_ls[n_ls - 1] + 1;
    for (int r = 0; r < n_ls; ++r)
    {
        int const l = c_ls[r];
        int const nm = 2 * l + 1;
        const float2* __restrict__ const block = reinterpret_cast<const float2*>(a.qlm) + (size_t) a.n * c_out_off[r];
        float re[kSphLmax + 1], im[kSphLmax + 1];
        for (int m = 0; m <= l; ++m)
        {
            re[m] = 0.0f;
            im[m] = 0.0f;
        }
        for (uint32_t b = beg; b < end; ++b)
        {
            uint32_t const j = a.neighbors[2 * (size_t) b + 1];
            const float2* __restrict__ const src = block + (size_t) j * nm;
            for (int m = 0; m <= l; ++m)
            {
                float2 const v = src[m];
                re[m] += v.x;
                im[m] += v.y;
            }
        }
        size_t const ii = active ? i : 0;
        const float2* __restrict__ const own = block + ii * nm;
        float2* const out = reinterpret_cast<float2*>(a.qlm_ave) + (size_t) a.n * c_out_off[r] + ii * nm;
        float sum = 0.0f;
        for (int m = 0; m <= l; ++m)
        {
            float2 const v = own[m];
            re[m] = (re[m] + v.x) / count;
            im[m] = (im[m] + v.y) / count;
            if (active)
            {
                out[m] = make_float2(re[m], im[m]);
            }
            sum += re[m] * re[m] + im[m] * im[m];
            if (a.sys_partials != nullptr)
            {
                // block sums instead of 2(l+1) hot atomics per particle (see k_steinhardt_single)
                double* const row = a.sys_partials + (size_t) blockIdx.x * (2 * tot_m);
                block_sum_to(active ? (double) re[m] : 0.0, row + 2 * (c_out_off[r] + m));
                block_sum_to(active ? (double) im[m] : 0.0, row + 2 * (c_out_off[r] + m) + 1);
            }
        }
        if (active)
        {
            for (int m = 1; m <= l; ++m)
            {
                float const phase = (m & 1) ? -1.0f : 1.0f;
                out[l + m] = make_float2(phase * re[m], -(phase * im[m]));
                sum += re[m] * re[m] + im[m] * im[m];
            }
            float const nf = (float) (4.0 * 3.14159265358979323846 / nm);
            a.ql_ave[(size_t) i * n_ls + r] = sqrtf(sum * nf);
        }
    }
}

// ---- third-order invariant (Steinhardt::aggregatewl + reduceWigner3j, Steinhardt.cc:329-359, Wigner3j.cc:22-57)
__global__ void __launch_bounds__(kThreads) k_steinhardt_wl(SteinhardtWlArgs a, int n_ls)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n)
    {
        return;
    }
    for (int r = 0; r < n_ls; ++r)
    {
        int const l = c_ls[r];
        int const nm = 2 * l + 1;
        const float2* __restrict__ const src
            = reinterpret_cast<const float2*>(a.qlm) + (size_t) a.n * c_out_off[r] + (size_t) i * nm;
        float2 q[2 * kSphLmax + 1];
        for (int k = 0; k < nm; ++k)
        {
            q[k] = src[k];
        }
        const float* __restrict__ const w3j = a.w3j + a.w3j_off[r];
        float result = 0.0f;
        int counter = 0;
        for (int m1 = -l; m1 <= l; ++m1)
        {
            float2 const s1 = q[m1 < 0 ? l - m1 : m1];
            int const lo = max(-l - m1, -l), hi = min(l - m1, l);
            for (int m2 = lo; m2 <= hi; ++m2)
            {
                int const m3 = -m1 - m2;
                float2 const s2 = q[m2 < 0 ? l - m2 : m2], s3 = q[m3 < 0 ? l - m3 : m3];
                float const w = __ldg(w3j + counter);
                // (w * s1) * s2 * s3, left to right, real part
                float const ar = w * s1.x, ai = w * s1.y;
                float const br = ar * s2.x - ai * s2.y, bi = ar * s2.y + ai * s2.x;
                result += br * s3.x - bi * s3.y;
                ++counter;
            }
        }
        if (a.normalize)
        {
            float const nf = (float) (4.0 * 3.14159265358979323846 / nm);
            float const nrm = sqrtf(nf) / a.ql[(size_t) i * n_ls + r];
            result *= nrm * nrm * nrm;
        }
        a.wl[(size_t) i * n_ls + r] = result;
    }
}

void upload_tables(fgpu_ctx* ctx, int lmax, const std::vector<uint32_t>& ls)
{
    // evaluatePrefactors, spherical_harmonics.hpp:216-237 (double expression, stored to float)
    std::vector<float> pref(2 * (kSphLmax + 1) * kSphLmax, 0.0f);
    unsigned const L = (unsigned) lmax;
    unsigned const f1 = L * (L + 1);
    for (unsigned m = 0; m < L + 1; ++m)
    {
        for (unsigned l = 1; l < L + 1; ++l)
        {
            pref[L * m + (l - 1)] = (float) (2 * std::sqrt(1 + (m - 0.5) / l) * std::sqrt(1 - (m - 0.5) / (l + 2 * m)));
        }
    }
    for (unsigned m = 0; m < L + 1; ++m)
    {
        if (L > 0)
        {
            pref[f1 + L * m] = 0.0f;
        }
        for (unsigned l = 2; l < L + 1; ++l)
        {
            pref[f1 + L * m + (l - 1)] = (float) (-std::sqrt(1.0 + 4.0 / (2 * l + 2 * m - 3)) * std::sqrt(1 - 1.0 / l)
                                                  * std::sqrt(1.0 - 1.0 / (l + 2 * m)));
        }
    }
    // compute_jacobis first column, spherical_harmonics.hpp:250-256
    float jac0[kSphLmax + 1]; // first column of the recurrence as upstream stores it (float), scaled at the end
    jac0[0] = (float) (1 / std::sqrt(2.0));
    for (unsigned m = 1; m < L + 1; ++m)
    {
        jac0[m] = (float) (jac0[m - 1] * std::sqrt(1 + 1.0 / 2 / m));
    }
    for (unsigned m = L + 1; m < kSphLmax + 1; ++m)
    {
        jac0[m] = 0.0f;
    }
    for (unsigned m = 0; m < L + 1; ++m)
    {
        jac0[m] = (float) ((double) jac0[m] * kInvSqrt2Pi);
    }
    int h_ls[kSphLmax + 1] = {0}, h_slot[kSphLmax + 1], h_acc[kSphLmax + 1] = {0}, h_out[kSphLmax + 1] = {0};
    for (int& s : h_slot)
    {
        s = -1;
    }
    int acc = 0, out = 0;
    for (size_t r = 0; r < ls.size(); ++r)
    {
        h_ls[r] = (int) ls[r];
        h_slot[ls[r]] = (int) r;
        h_acc[r] = acc;
        h_out[r] = out;
        acc += (int) ls[r] + 1;
        out += 2 * (int) ls[r] + 1;
    }
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_pref, pref.data(), pref.size() * sizeof(float), 0,
                                            cudaMemcpyHostToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_jac0, jac0, sizeof(jac0), 0, cudaMemcpyHostToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_ls, h_ls, sizeof(h_ls), 0, cudaMemcpyHostToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_l_slot, h_slot, sizeof(h_slot), 0, cudaMemcpyHostToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_acc_off, h_acc, sizeof(h_acc), 0, cudaMemcpyHostToDevice, ctx->stream));
    FGPU_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_out_off, h_out, sizeof(h_out), 0, cudaMemcpyHostToDevice, ctx->stream));
    // the host arrays above die at return; the copies must have been consumed by then
    FGPU_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

template<int L> void launch_single(fgpu_ctx* ctx, const SteinhardtArgs& a_in)
{
    SteinhardtArgs a = a_in;
    unsigned const blocks = (a.n + kThreads - 1) / kThreads;
    uint32_t const width = 2 * (L + 1);
    if (a.sys_qlm != nullptr)
    {
        ctx->st_partials.reserve((size_t) blocks * width);
        a.sys_partials = ctx->st_partials.ptr;
    }
    k_steinhardt_single<L><<<blocks, kThreads, 0, ctx->stream>>>(a);
    if (a.sys_qlm != nullptr)
    {
        k_sum_partials<<<width, kSumThreads, 0, ctx->stream>>>(a.sys_partials, blocks, width, a.sys_qlm);
    }
}

template<int L> void launch_fused(fgpu_ctx* ctx, SteinhardtArgs a, const KnnYlmArgs& s)
{
    unsigned const blocks = (a.n + kThreads - 1) / kThreads;
    uint32_t const width = 2 * (L + 1);
    if (a.sys_qlm != nullptr)
    {
        ctx->st_partials.reserve((size_t) blocks * width);
        a.sys_partials = ctx->st_partials.ptr;
    }
    if (s.k <= 12)
    {
        k_knn_ylm<L, 12><<<blocks, kThreads, 0, ctx->stream>>>(a, s);
    }
    else
    {
        k_knn_ylm<L, kFusedMaxK><<<blocks, kThreads, 0, ctx->stream>>>(a, s);
    }
    if (a.sys_qlm != nullptr)
    {
        k_sum_partials<<<width, kSumThreads, 0, ctx->stream>>>(a.sys_partials, blocks, width, a.sys_qlm);
    }
}

} // namespace

void ensure_steinhardt_tables(fgpu_ctx* ctx, const std::vector<uint32_t>& ls, int lmax)
{
    // the tables live in __constant__ memory, one copy per device: upload only when the l list changes
    static std::mutex mtx;
    static std::map<int, std::vector<uint32_t>> resident;
    std::lock_guard<std::mutex> lock(mtx);
    auto it = resident.find(ctx->device);
    if (it == resident.end() || it->second != ls)
    {
        upload_tables(ctx, lmax, ls);
        resident[ctx->device] = ls;
    }
}

bool knn_ylm_supported(const std::vector<uint32_t>& ls, uint32_t k)
{
    if (ls.size() != 1 || k == 0 || k > (uint32_t) kFusedMaxK)
    {
        return false;
    }
    uint32_t const l = ls[0];
    return l == 2 || l == 4 || l == 6 || l == 8 || l == 10 || l == 12;
}

void launch_knn_ylm(fgpu_ctx* ctx, const SteinhardtArgs& a, uint32_t l, const KnnSelectArgs& src)
{
    ensure_steinhardt_tables(ctx, std::vector<uint32_t> {l}, (int) l);
    if (a.n == 0)
    {
        return;
    }
    KnnYlmArgs s;
    s.bag = src.bag;
    s.bag2 = src.bag2;
    s.tmp_start = src.tmp_start;
    s.hits = src.hits;
    s.k = src.k;
    KernelScope ks(ctx, "knn_ylm");
    switch (l)
    {
    case 2: launch_fused<2>(ctx, a, s); break;
    case 4: launch_fused<4>(ctx, a, s); break;
    case 6: launch_fused<6>(ctx, a, s); break;
    case 8: launch_fused<8>(ctx, a, s); break;
    case 10: launch_fused<10>(ctx, a, s); break;
    case 12: launch_fused<12>(ctx, a, s); break;
    default: throw Error(FGPU_EINVALID, "knn_ylm: unsupported l");
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_steinhardt(fgpu_ctx* ctx, const SteinhardtArgs& a, const std::vector<uint32_t>& ls)
{
    int lmax = 0, n_acc = 0, tot_m = 0;
    for (uint32_t l : ls)
    {
        if (l > (uint32_t) kSphLmax)
        {
            throw Error(FGPU_EINVALID, "Steinhardt: l larger than 32 is not supported by the device path");
        }
        lmax = std::max(lmax, (int) l);
        n_acc += (int) l + 1;
        tot_m += 2 * (int) l + 1;
    }
    if (ls.empty() || (int) ls.size() > kSphLmax || 2 * n_acc > kMaxAcc)
    {
        throw Error(FGPU_EINVALID, "Steinhardt: unsupported list of l values");
    }
    ensure_steinhardt_tables(ctx, ls, lmax);
    if (a.n == 0)
    {
        return;
    }
    KernelScope ks(ctx, "steinhardt");
    bool single = ls.size() == 1;
    if (single)
    {
        switch (ls[0])
        {
        case 2: launch_single<2>(ctx, a); break;
        case 4: launch_single<4>(ctx, a); break;
        case 6: launch_single<6>(ctx, a); break;
        case 8: launch_single<8>(ctx, a); break;
        case 10: launch_single<10>(ctx, a); break;
        case 12: launch_single<12>(ctx, a); break;
        default: single = false; break;
        }
    }
    if (!single)
    {
        SteinhardtArgs g = a;
        unsigned const blocks = (a.n + kThreads - 1) / kThreads;
        uint32_t const width = 2 * (uint32_t) tot_m;
        if (a.sys_qlm != nullptr)
        {
            ctx->st_partials.reserve((size_t) blocks * width);
            g.sys_partials = ctx->st_partials.ptr;
            FGPU_CUDA_CHECK(cudaMemsetAsync(g.sys_partials, 0, (size_t) blocks * width * sizeof(double), ctx->stream));
        }
        k_steinhardt_generic<<<blocks, kThreads, 0, ctx->stream>>>(g, lmax, (int) ls.size(), n_acc, tot_m);
        if (a.sys_qlm != nullptr)
        {
            k_sum_partials<<<width, kSumThreads, 0, ctx->stream>>>(g.sys_partials, blocks, width, a.sys_qlm);
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_pad_positions(fgpu_ctx* ctx, const float* xyz, uint32_t n, float4* out)
{
    if (n == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "pad_positions");
        k_pad_positions<<<(n + 255) / 256, 256, 0, ctx->stream>>>(xyz, n, out);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_steinhardt_average(fgpu_ctx* ctx, const SteinhardtAveArgs& a, int n_ls, uint32_t tot_m)
{
    if (a.n == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "steinhardt_average");
        SteinhardtAveArgs g = a;
        unsigned const blocks = (a.n + kThreads - 1) / kThreads;
        uint32_t const width = 2 * tot_m;
        if (a.sys_qlm != nullptr)
        {
            ctx->st_partials.reserve((size_t) blocks * width);
            g.sys_partials = ctx->st_partials.ptr;
            FGPU_CUDA_CHECK(cudaMemsetAsync(g.sys_partials, 0, (size_t) blocks * width * sizeof(double), ctx->stream));
        }
        k_steinhardt_average<<<blocks, kThreads, 0, ctx->stream>>>(g, n_ls);
        if (a.sys_qlm != nullptr)
        {
            k_sum_partials<<<width, kSumThreads, 0, ctx->stream>>>(g.sys_partials, blocks, width, a.sys_qlm);
        }
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

void launch_steinhardt_wl(fgpu_ctx* ctx, const SteinhardtWlArgs& a, int n_ls)
{
    if (a.n == 0)
    {
        return;
    }
    {
        KernelScope ks(ctx, "steinhardt_wl");
        k_steinhardt_wl<<<(a.n + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(a, n_ls);
    }
    FGPU_CUDA_CHECK(cudaGetLastError());
}

// Wigner 3j symbols (l l l; m1 m2 m3) in the order reduceWigner3j walks its table (Wigner3j.cc:43-55).
// Racah's formula with the alternating sum written as a sum of products of three binomial coefficients,
//   (l l l; m1 m2 m3) = (-1)^m3 sqrt( prod (l +- m_i)! / ((3l+1)! (l!)^3) ) sum_k (-1)^k C(l,k) C(l,l-m1-k) C(l,l+m2-k),
// which is exact in 128-bit integers for l <= 32; the prefactor is evaluated in long double.
std::vector<float> wigner3j_table(uint32_t l_)
{
    int const l = (int) l_;
    std::vector<long double> fact(3 * l + 2);
    fact[0] = 1.0L;
    for (int k = 1; k < 3 * l + 2; ++k)
    {
        fact[k] = fact[k - 1] * (long double) k;
    }
    std::vector<std::vector<__int128>> binom(l + 1, std::vector<__int128>(l + 1, 0));
    for (int n = 0; n <= l; ++n)
    {
        binom[n][0] = 1;
        for (int k = 1; k <= n; ++k)
        {
            binom[n][k] = binom[n - 1][k - 1] + (k <= n - 1 ? binom[n - 1][k] : 0);
        }
    }
    auto C = [&](int n, int k) -> __int128 { return (k < 0 || k > n) ? (__int128) 0 : binom[n][k]; };
    std::vector<float> table;
    for (int m1 = -l; m1 <= l; ++m1)
    {
        for (int m2 = std::max(-l - m1, -l); m2 <= std::min(l - m1, l); ++m2)
        {
            int const m3 = -m1 - m2;
            __int128 sum = 0;
            for (int k = 0; k <= l; ++k)
            {
                __int128 const term = C(l, k) * C(l, l - m1 - k) * C(l, l + m2 - k);
                sum += (k & 1) ? -term : term;
            }
            long double const pref = std::sqrt(fact[l + m1] * fact[l - m1] * fact[l + m2] * fact[l - m2] * fact[l + m3]
                                               * fact[l - m3] / (fact[3 * l + 1] * fact[l] * fact[l] * fact[l]));
            long double const sign = (m3 & 1) ? -1.0L : 1.0L;
            double const value = (double) (sign * pref * (long double) sum);
            table.push_back((float) value); // upstream stores doubles and uses float(w), Wigner3j.cc:52
        }
    }
    return table;
}

} // namespace fgpu
