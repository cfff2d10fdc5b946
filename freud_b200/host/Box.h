// freud::box::Box, reduced to what the neighbour-query path consumes (freud/box/Box.h:44-606):
// lengths, tilt factors, dimensionality, periodicity, volume and nearest-plane distances.  The per-pair
// minimum-image arithmetic (makeFractional / wrap / makeAbsolute) lives in the CUDA kernels
// (csrc/pair_math.cuh); this value type only describes the box to them.
#pragma once
#include <cmath>
#include <stdexcept>

#include "VectorMath.h"

namespace freud { namespace box {

class Box
{
public:
    Box() = default;
    // same argument order as upstream (module-box.cc:20)
    Box(float Lx, float Ly, float Lz, float xy = 0, float xz = 0, float yz = 0, bool is2D = false)
        : m_Lx(Lx), m_Ly(Ly), m_Lz(is2D ? 0.0F : Lz), m_xy(xy), m_xz(xz), m_yz(yz), m_2d(is2D)
    {}

    float getLx() const { return m_Lx; }
    float getLy() const { return m_Ly; }
    float getLz() const { return m_Lz; }
    float getTiltFactorXY() const { return m_xy; }
    float getTiltFactorXZ() const { return m_xz; }
    float getTiltFactorYZ() const { return m_yz; }
    bool is2D() const { return m_2d; }

    vec3<bool> getPeriodic() const { return m_periodic; }
    void setPeriodic(bool x, bool y, bool z) { m_periodic = vec3<bool>(x, y, z); }

    // Box.h:143-150: area in 2-D
    float getVolume() const
    {
        volatile float a = m_Lx * m_Ly;
        if (m_2d)
        {
            return a;
        }
        volatile float v = a * m_Lz;
        return v;
    }

    // Box.h:489-497
    vec3<float> getNearestPlaneDistance() const
    {
        volatile float t0 = m_xy * m_yz;
        volatile float t = t0 - m_xz;
        volatile float a0 = m_xy * m_xy;
        volatile float a1 = 1.0F + a0;
        volatile float a2 = t * t;
        volatile float a3 = a1 + a2;
        volatile float c0 = m_yz * m_yz;
        volatile float c1 = 1.0F + c0;
        return vec3<float>(m_Lx / std::sqrt((float) a3), m_Ly / std::sqrt((float) c1), m_Lz);
    }

    // Lattice vectors (Box.h:503-518); the third one vanishes in 2-D
    vec3<float> getLatticeVector(unsigned int i) const
    {
        if (i == 0)
        {
            return vec3<float>(m_Lx, 0.0F, 0.0F);
        }
        if (i == 1)
        {
            volatile float bx = m_Ly * m_xy;
            return vec3<float>(bx, m_Ly, 0.0F);
        }
        if (i == 2 && !m_2d)
        {
            volatile float cx = m_Lz * m_xz, cy = m_Lz * m_yz;
            return vec3<float>(cx, cy, m_Lz);
        }
        throw std::out_of_range("Box lattice vector index requested does not exist.");
    }

    // Host restatement of the per-vector box arithmetic (Box.h:212-255, 307-329, freud/util/utils.h:29-32), one float32
    // rounding per operation in the reference's order (this file is compiled with -ffp-contract=off; the volatile
    // temporaries keep every intermediate in float).  The GPU kernels hold the production copy (csrc/pair_math.cuh);
    // this one serves the host-only corners of the API: NeighborList's all-pairs constructor and CellQuery's grid
    // introspection.
    vec3<float> makeFractional(const vec3<float>& v) const
    {
        volatile float lox = -(m_Lx * 0.5F), loy = -(m_Ly * 0.5F), loz = -(m_Lz * 0.5F);
        volatile float dx = v.x - lox, dy = v.y - loy, dz = v.z - loz;
        volatile float t0 = m_yz * m_xy;
        volatile float txz = m_xz - t0;
        volatile float a = txz * v.z, b = m_xy * v.y;
        volatile float ab = a + b;
        dx = dx - ab;
        volatile float c = m_yz * v.z;
        dy = dy - c;
        volatile float fx = dx / m_Lx, fy = dy / m_Ly, fz = m_2d ? 0.0F : dz / m_Lz;
        return vec3<float>(fx, fy, fz);
    }
    vec3<float> makeAbsolute(const vec3<float>& f) const
    {
        volatile float lox = -(m_Lx * 0.5F), loy = -(m_Ly * 0.5F), loz = -(m_Lz * 0.5F);
        volatile float px = f.x * m_Lx, py = f.y * m_Ly, pz = f.z * m_Lz;
        volatile float x = lox + px, y = loy + py, z = loz + pz;
        volatile float a = m_xy * y, b = m_xz * z;
        volatile float ab = a + b;
        x = x + ab;
        volatile float c = m_yz * z;
        y = y + c;
        return vec3<float>(x, y, m_2d ? 0.0F : (float) z);
    }
    vec3<float> wrap(const vec3<float>& v) const
    {
        vec3<float> f = makeFractional(v);
        auto mod1 = [](float a) {
            volatile float t = std::fmod(a, 1.0F);
            volatile float u = t + 1.0F;
            volatile float w = std::fmod((float) u, 1.0F);
            return (float) w;
        };
        f.x = mod1(f.x);
        f.y = mod1(f.y);
        f.z = m_2d ? 0.0F : mod1(f.z);
        return makeAbsolute(f);
    }

    bool operator==(const Box& o) const
    {
        return m_Lx == o.m_Lx && m_Ly == o.m_Ly && m_Lz == o.m_Lz && m_xy == o.m_xy && m_xz == o.m_xz
            && m_yz == o.m_yz && m_2d == o.m_2d;
    }

    // {Lx, Ly, Lz, xy, xz, yz} as the C ABI takes it
    void toArray6(float out[6]) const
    {
        out[0] = m_Lx;
        out[1] = m_Ly;
        out[2] = m_Lz;
        out[3] = m_xy;
        out[4] = m_xz;
        out[5] = m_yz;
    }

private:
    float m_Lx {0}, m_Ly {0}, m_Lz {0}, m_xy {0}, m_xz {0}, m_yz {0};
    bool m_2d {false};
    vec3<bool> m_periodic {true, true, true};
};

}} // namespace freud::box
