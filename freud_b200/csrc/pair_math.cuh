// Exact float32 pair arithmetic of the reference, for device code.
//
// The reference's x86-64 build is un-fused IEEE float32, so every operation that
// feeds a value compared against r_max^2, binned, or written to a NeighborList goes through the
// round-to-nearest intrinsics (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn), which nvcc never
// contracts into FMAs, in exactly the reference's operation order:
//   Box::makeFractional  freud/box/Box.h:243-255
//   Box::makeAbsolute    freud/box/Box.h:212-222
//   Box::wrap            freud/box/Box.h:307-329 with util::modulusPositive freud/util/utils.h:29-32
//   dot(vec3, vec3)      freud/util/VectorMath.h:270-273   ((x*x + y*y) + z*z)
//   image vectors        freud/locality/NeighborQuery.h:496-564
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fgpu {

struct BoxDev
{
    float Lx, Ly, Lz;    // Lz == 0 in 2-D (Box.h:102-106)
    float xy, xz, yz;    // tilt factors
    float lox, loy, loz; // -(L * 0.5f)   (Box.h:113-114)
    float t_xz;          // xz - yz * xy, the constant sub-expression of makeFractional (Box.h:246)
    // lattice vectors (Box.h:503-518); c == 0 in 2-D (NeighborQuery.h:533-537)
    float ax, bx, by, cx, cy, cz;
    int is2d;
};

// fmodf(a, 1.0f) is exactly a - trunc(a) for every finite a (the subtraction is exact); the sign of a
// zero result differs from fmodf only for a == -0.0f / negative integers, and modulusPositive adds 1.0f
// right after, which erases it.
__device__ __forceinline__ float modulus_positive_one(float a)
{
    float const t = __fsub_rn(a, truncf(a));
    float const u = __fadd_rn(t, 1.0f);
    return __fsub_rn(u, truncf(u));
}

// The same for f in (-1, 2) whose intermediate sum stays below 2: truncation is a compare (FSET), not an FRND on
// the XU pipe.
__device__ __forceinline__ float modulus_positive_one_small(float f)
{
    float const t = __fsub_rn(f, f >= 1.0f ? 1.0f : 0.0f);
    float const u = __fadd_rn(t, 1.0f);
    return __fsub_rn(u, u >= 1.0f ? 1.0f : 0.0f);
}

__device__ __forceinline__ float dot_exact(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// a / L with y = RN(1 / L) rounded on the host: q0 = RN(a y), one fused correction makes the quotient faithful,
// the second one makes it RN(a / L) (Markstein; the residuals a - L q are exact in an FMA) as long as nothing
// underflows.  Five issue slots instead of the ~15 of __fdiv_rn.  Users state why their operands are in range.
__device__ __forceinline__ float div_by_const(float a, float L, float y)
{
    float const q0 = __fmul_rn(a, y);
    float const r0 = __fmaf_rn(-q0, L, a);
    float const q1 = __fmaf_rn(r0, y, q0);
    float const r1 = __fmaf_rn(-q1, L, a);
    return __fmaf_rn(r1, y, q1);
}

// Box::wrap(v) with the three divisions done by div_by_const: the same result as wrap_exact except when a
// fractional coordinate is so small that a residual underflows (|v - lo| below ~1e-30 L).  For consumers with a
// tolerance (Steinhardt's 1e-5), not for bit-exact bond lists.
__device__ __forceinline__ void wrap_quick(const BoxDev& b, float ylx, float yly, float ylz, float vx, float vy,
                                           float vz, float& rx, float& ry, float& rz)
{
    float dx = __fsub_rn(vx, b.lox);
    float dy = __fsub_rn(vy, b.loy);
    float const dz = __fsub_rn(vz, b.loz);
    dx = __fsub_rn(dx, __fadd_rn(__fmul_rn(b.t_xz, vz), __fmul_rn(b.xy, vy)));
    dy = __fsub_rn(dy, __fmul_rn(b.yz, vz));
    float fx = div_by_const(dx, b.Lx, ylx);
    float fy = div_by_const(dy, b.Ly, yly);
    float fz = b.is2d ? 0.0f : div_by_const(dz, b.Lz, ylz);
    // two points of the box are less than one box apart: f in (-1, 2), where fmodf(fmodf(f, 1) + 1, 1) needs
    // compares only (bit-identical to the truncating form there); anything else takes the general form
    if (fabsf(fx - 0.5f) < 1.4f && fabsf(fy - 0.5f) < 1.4f && fabsf(fz - 0.5f) < 1.4f)
    {
        fx = modulus_positive_one_small(fx);
        fy = modulus_positive_one_small(fy);
        fz = modulus_positive_one_small(fz);
    }
    else
    {
        fx = modulus_positive_one(fx);
        fy = modulus_positive_one(fy);
        fz = modulus_positive_one(fz);
    }
    float x = __fadd_rn(b.lox, __fmul_rn(fx, b.Lx));
    float y = __fadd_rn(b.loy, __fmul_rn(fy, b.Ly));
    float z = __fadd_rn(b.loz, __fmul_rn(fz, b.Lz));
    x = __fadd_rn(x, __fadd_rn(__fmul_rn(b.xy, y), __fmul_rn(b.xz, z)));
    y = __fadd_rn(y, __fmul_rn(b.yz, z));
    rx = x;
    ry = y;
    rz = b.is2d ? 0.0f : z;
}

__device__ __forceinline__ bool in_window2(float r_sq, float r_max_sq, float r_min_sq)
{
    return r_sq < r_max_sq && r_sq >= r_min_sq; // LinkCell.cc:525, AABBQuery.cc:129
}

// Division by a box length: div_by_const (pair_math.cuh) replaces __fdiv_rn (x86 divss upstream, Box.h:248-250).
// Exact here because stage 1 bounds |a| / L away from 0 (no underflow in the residuals).
// modulus_positive_one_small (pair_math.cuh): util::modulusPositive(f, 1) (freud/util/utils.h:29-32) for f in (-1, 2).
// Box::wrap(v) (freud/box/Box.h:307-329) for displacements whose fractional coordinates are within
// (-1/2 - 0.35, 1/2 + 0.35) + {-1, 0, 1}; bit-identical to wrap_exact on that domain.
template<bool TRI>
__device__ __forceinline__ void wrap_fast(const BoxDev& b, float ylx, float yly, float ylz, float vx, float vy,
                                          float vz, float& rx, float& ry, float& rz)
{
    float dx = __fsub_rn(vx, b.lox);
    float dy = __fsub_rn(vy, b.loy);
    float const dz = __fsub_rn(vz, b.loz);
    if (TRI)
    {
        dx = __fsub_rn(dx, __fadd_rn(__fmul_rn(b.t_xz, vz), __fmul_rn(b.xy, vy)));
        dy = __fsub_rn(dy, __fmul_rn(b.yz, vz));
    }
    float fx = div_by_const(dx, b.Lx, ylx);
    float fy = div_by_const(dy, b.Ly, yly);
    float fz = b.is2d ? 0.0f : div_by_const(dz, b.Lz, ylz);
    fx = modulus_positive_one_small(fx);
    fy = modulus_positive_one_small(fy);
    fz = modulus_positive_one_small(fz);
    float x = __fadd_rn(b.lox, __fmul_rn(fx, b.Lx));
    float y = __fadd_rn(b.loy, __fmul_rn(fy, b.Ly));
    float z = __fadd_rn(b.loz, __fmul_rn(fz, b.Lz));
    if (TRI)
    {
        x = __fadd_rn(x, __fadd_rn(__fmul_rn(b.xy, y), __fmul_rn(b.xz, z)));
        y = __fadd_rn(y, __fmul_rn(b.yz, z));
    }
    if (b.is2d)
    {
        z = 0.0f;
    }
    rx = x;
    ry = y;
    rz = z;
}

// r = Box::wrap(v), all axes periodic.
__device__ __forceinline__ void wrap_exact(const BoxDev& b, float vx, float vy, float vz, float& rx, float& ry,
                                           float& rz)
{
    // makeFractional
    float dx = __fsub_rn(vx, b.lox);
    float dy = __fsub_rn(vy, b.loy);
    float const dz = __fsub_rn(vz, b.loz);
    dx = __fsub_rn(dx, __fadd_rn(__fmul_rn(b.t_xz, vz), __fmul_rn(b.xy, vy)));
    dy = __fsub_rn(dy, __fmul_rn(b.yz, vz));
    float fx = __fdiv_rn(dx, b.Lx);
    float fy = __fdiv_rn(dy, b.Ly);
    float fz = b.is2d ? 0.0f : __fdiv_rn(dz, b.Lz); // 2-D: 0/0 -> NaN upstream, then forced to 0
    // modulusPositive(f, 1)
    fx = modulus_positive_one(fx);
    fy = modulus_positive_one(fy);
    fz = modulus_positive_one(fz);
    // makeAbsolute
    float x = __fadd_rn(b.lox, __fmul_rn(fx, b.Lx));
    float y = __fadd_rn(b.loy, __fmul_rn(fy, b.Ly));
    float z = __fadd_rn(b.loz, __fmul_rn(fz, b.Lz));
    x = __fadd_rn(x, __fadd_rn(__fmul_rn(b.xy, y), __fmul_rn(b.xz, z)));
    y = __fadd_rn(y, __fmul_rn(b.yz, z));
    if (b.is2d)
    {
        z = 0.0f;
    }
    rx = x;
    ry = y;
    rz = z;
}

// image vector float(i)*a + float(j)*b + float(k)*c, component-wise, left to right
__device__ __forceinline__ void image_vector(const BoxDev& b, int i, int j, int k, float& ix, float& iy, float& iz)
{
    float const fi = (float) i, fj = (float) j, fk = (float) k;
    ix = __fadd_rn(__fadd_rn(__fmul_rn(fi, b.ax), __fmul_rn(fj, b.bx)), __fmul_rn(fk, b.cx));
    iy = __fadd_rn(__fadd_rn(__fmul_rn(fi, 0.0f), __fmul_rn(fj, b.by)), __fmul_rn(fk, b.cy));
    iz = __fadd_rn(__fadd_rn(__fmul_rn(fi, 0.0f), __fmul_rn(fj, 0.0f)), __fmul_rn(fk, b.cz));
}

// CellQuery's ghost displacement for a point seen across w = (wx, wy, wz) boundaries (CellQuery.h:246-262):
// shift = 0, then += +-a, += +-b, += +-c for the non-zero components, with a = (Lx, 0, 0), b = (Ly xy, Ly, 0),
// c = (Lz xz, Lz yz, Lz) (Box.h:503-518) and component-wise float adds.
__device__ __forceinline__ void ghost_shift(const BoxDev& b, int wx, int wy, int wz, float& sx, float& sy, float& sz)
{
    sx = sy = sz = 0.0f;
    if (wx != 0)
    {
        float const sg = wx > 0 ? 1.0f : -1.0f;
        sx = __fadd_rn(sx, sg * b.ax);
        sy = __fadd_rn(sy, sg * 0.0f);
        sz = __fadd_rn(sz, sg * 0.0f);
    }
    if (wy != 0)
    {
        float const sg = wy > 0 ? 1.0f : -1.0f;
        sx = __fadd_rn(sx, sg * b.bx);
        sy = __fadd_rn(sy, sg * b.by);
        sz = __fadd_rn(sz, sg * 0.0f);
    }
    if (wz != 0)
    {
        float const sg = wz > 0 ? 1.0f : -1.0f;
        sx = __fadd_rn(sx, sg * b.cx);
        sy = __fadd_rn(sy, sg * b.cy);
        sz = __fadd_rn(sz, sg * b.cz);
    }
}

// Bond vector of candidate p and query q when the QUERY is taken in image k (the candidate sits across w = -k
// boundaries).  IMAGE: r = p - (q + image_k), AABBQuery.cc:93,125.  GHOST: r = (p + shift_w) - q,
// CellQuery.cc:107 + CellIterator.h:167.  (WRAP never gets here: it wraps the difference.)
template<int FLAVOUR>
__device__ __forceinline__ void image_pair(const BoxDev& b, float px, float py, float pz, float qx, float qy, float qz,
                                           int kx, int ky, int kz, float& rx, float& ry, float& rz)
{
    if (FLAVOUR == FGPU_FLAVOUR_GHOST)
    {
        float sx, sy, sz;
        ghost_shift(b, -kx, -ky, -kz, sx, sy, sz);
        bool const real = kx == 0 && ky == 0 && kz == 0; // a real point is stored as it is, CellQuery.cc:121
        rx = __fsub_rn(real ? px : __fadd_rn(px, sx), qx);
        ry = __fsub_rn(real ? py : __fadd_rn(py, sy), qy);
        rz = __fsub_rn(real ? pz : __fadd_rn(pz, sz), qz);
    }
    else
    {
        float ix, iy, iz;
        image_vector(b, kx, ky, kz, ix, iy, iz);
        rx = __fsub_rn(px, __fadd_rn(qx, ix));
        ry = __fsub_rn(py, __fadd_rn(qy, iy));
        rz = __fsub_rn(pz, __fadd_rn(qz, iz));
    }
}

// Fractional coordinates for CELL ASSIGNMENT only (candidate generation is conservative);
// any consistent arithmetic works, the cell width carries a margin for its rounding.
__device__ __forceinline__ void fractional_for_cells(const BoxDev& b, float vx, float vy, float vz, float& fx,
                                                     float& fy, float& fz)
{
    float dx = vx - b.lox;
    float dy = vy - b.loy;
    float const dz = vz - b.loz;
    dx -= b.t_xz * vz + b.xy * vy;
    dy -= b.yz * vz;
    fx = dx / b.Lx;
    fy = dy / b.Ly;
    fz = b.is2d ? 0.0f : dz / b.Lz;
}

// Cell coordinate of a point: frac - floor(frac) scaled to the grid; integer image offset = floor(frac).
__device__ __forceinline__ void cell_coords(const BoxDev& b, int dx, int dy, int dz, float x, float y, float z,
                                            int& cx, int& cy, int& cz, int& nx, int& ny, int& nz)
{
    float fx, fy, fz;
    fractional_for_cells(b, x, y, z, fx, fy, fz);
    float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    // NaN / far-away guard: such points cannot be neighbours of anything in range
    if (!(fabsf(flx) < 500.0f))
    {
        flx = 0.0f;
    }
    if (!(fabsf(fly) < 500.0f))
    {
        fly = 0.0f;
    }
    if (!(fabsf(flz) < 500.0f))
    {
        flz = 0.0f;
    }
    cx = min(max((int) ((fx - flx) * (float) dx), 0), dx - 1);
    cy = min(max((int) ((fy - fly) * (float) dy), 0), dy - 1);
    cz = min(max((int) ((fz - flz) * (float) dz), 0), dz - 1);
    nx = (int) flx;
    ny = (int) fly;
    nz = (int) flz;
}

// integer image offsets of a point, 10 bits per axis, biased by 512
__device__ __forceinline__ int pack_shift(int nx, int ny, int nz)
{
    return ((nx + 512) & 1023) | (((ny + 512) & 1023) << 10) | (((nz + 512) & 1023) << 20);
}

__device__ __forceinline__ void unpack_shift(int packed, int& nx, int& ny, int& nz)
{
    nx = (packed & 1023) - 512;
    ny = ((packed >> 10) & 1023) - 512;
    nz = ((packed >> 20) & 1023) - 512;
}

// Per axis: the cells a query in home cell c must visit and, for each, how many times the periodic
// boundary was crossed to reach it (w in {-1, 0, 1}; 2 == ambiguous: the axis has fewer than 3 cells, so
// every image has to be tried on it).
struct AxisSlots
{
    int n;
    int cell[3];
    int w[3];
};

__device__ __forceinline__ void make_slots(int dim, int ambiguous, int c, AxisSlots& s)
{
    if (dim >= 3)
    {
        s.n = 3;
#pragma unroll
        for (int o = -1; o <= 1; ++o)
        {
            int t = c + o;
            int w = 0;
            if (t < 0)
            {
                t += dim;
                w = -1;
            }
            else if (t >= dim)
            {
                t -= dim;
                w = 1;
            }
            s.cell[o + 1] = t;
            s.w[o + 1] = w;
        }
    }
    else
    {
        s.n = dim;
        for (int t = 0; t < 3; ++t)
        {
            s.cell[t] = t < dim ? t : 0;
            s.w[t] = ambiguous ? 2 : 0;
        }
    }
}

// RegularAxis::bin, freud/util/Histogram.h:152-174.  Returns -1 for the overflow bin.
struct AxisDev
{
    float r_min, r_max, inv_width;
    uint32_t bins;
};

__device__ __forceinline__ int axis_bin(const AxisDev& a, float value)
{
    if (value < a.r_min || value >= a.r_max)
    {
        return -1;
    }
    float const val = __fmul_rn(__fsub_rn(value, a.r_min), a.inv_width);
    int bin = __float2int_rz(val); // truncation == _mm_cvtt_ss2si
    if ((uint32_t) bin == a.bins)
    {
        bin -= 1;
    }
    return bin;
}

} // namespace fgpu
