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
