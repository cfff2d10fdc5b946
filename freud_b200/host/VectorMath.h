// vec3<T>: the 12-byte POD the reference reinterprets (N, 3) float32 arrays as
// (freud/util/VectorMath.h:27-57, freud/locality/export-NeighborQuery.cc:30-31).  Only the layout matters
// on this path: all pair arithmetic runs on the GPU.
#pragma once

template<typename T> struct vec3
{
    T x {}, y {}, z {};
    vec3() = default;
    vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(vec3<float>) == 12, "vec3<float> must alias three packed floats");

// quat<T>: scalar part first, then the vector part -- the 16-byte POD the reference reinterprets (N, 4) float32 arrays
// as (freud/util/VectorMath.h:588-640).  Layout only, like vec3.
template<typename T> struct quat
{
    T s {1};
    vec3<T> v;
    quat() = default;
    quat(T s_, const vec3<T>& v_) : s(s_), v(v_) {}
};
static_assert(sizeof(quat<float>) == 16, "quat<float> must alias four packed floats");
