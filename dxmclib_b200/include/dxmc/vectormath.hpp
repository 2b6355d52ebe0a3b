// vectormath.hpp — 3-vector helpers with the reference's names and operation order
// (API of reference include/dxmc/vectormath.hpp:37-221). Host side only; the device versions
// live in csrc/physics.cuh.
#pragma once
#include "dxmc/floating.hpp"
#include <cmath>
#include <cstdint>
#include <type_traits>

namespace dxmc::vectormath {

template <typename T>
concept Index = std::is_integral_v<T> && !std::is_same_v<bool, T>;

template <Floating T>
inline T dot(const T a[3], const T b[3]) noexcept { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <Floating T>
inline T lenght_sqr(T v[3]) noexcept { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }

template <Floating T>
inline T lenght(T v[3]) noexcept { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

template <Floating T>
inline T lenght(const T v[3]) noexcept { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

template <Floating T>
inline void normalize(T v[3]) noexcept
{
    const T norm = T { 1 } / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; ++i)
        v[i] *= norm;
}

template <Floating T>
inline void cross(const T a[3], const T b[3], T out[3]) noexcept
{
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

// cross product of the two halves of a direction-cosine sextet
template <Floating T>
inline void cross(const T c[6], T out[3]) noexcept { cross(c, c + 3, out); }

// Rodrigues rotation of v about a unit axis
template <Floating T>
inline void rotate(T v[3], const T axis[3], const T angle) noexcept
{
    const T s = std::sin(angle);
    const T c = std::cos(angle);
    const T m = (T { 1 } - c) * dot(v, axis);
    const T r[3] = {
        c * v[0] + m * axis[0] + s * (axis[1] * v[2] - axis[2] * v[1]),
        c * v[1] + m * axis[1] + s * (-axis[0] * v[2] + axis[2] * v[0]),
        c * v[2] + m * axis[2] + s * (axis[0] * v[1] - axis[1] * v[0])
    };
    v[0] = r[0];
    v[1] = r[1];
    v[2] = r[2];
}

template <Floating T>
inline void projectToPlane(T v[3], const T normal[3]) noexcept
{
    const T d = dot(v, normal);
    for (int i = 0; i < 3; ++i)
        v[i] = v[i] - d * normal[i];
}

// Kahan's numerically stable angle between two vectors
template <Floating T>
inline T angleBetween(const T a[3], const T b[3]) noexcept
{
    const T d[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
    const T la = lenght(a), lb = lenght(b), lc = lenght(d);
    const T u = lb >= lc ? lc - (la - lb) : lb - (la - lc);
    const T nom = ((la - lb) + lc) * u;
    const T den = (la + (lb + lc)) * ((la - lc) + lb);
    return T { 2 } * std::atan(std::sqrt(nom / den));
}

template <Floating T>
inline T angleBetweenOnPlane(T a[3], T b[3], T normal[3]) noexcept
{
    normalize(a);
    normalize(b);
    normalize(normal);
    T c[3];
    cross(a, b, c);
    return std::atan2(dot(c, normal), dot(a, b));
}

template <Index U, Floating T>
inline U argmin3(const T v[3]) noexcept
{
    const T x = std::abs(v[0]), y = std::abs(v[1]), z = std::abs(v[2]);
    return x <= y ? (x <= z ? 0 : 2) : (y <= z ? 1 : 2);
}

template <Index U, Floating T>
inline U argmax3(const T v[3]) noexcept
{
    const T x = std::abs(v[0]), y = std::abs(v[1]), z = std::abs(v[2]);
    return x >= y ? (x >= z ? 0 : 2) : (y >= z ? 1 : 2);
}

// columns b1,b2,b3 times v
template <Floating T>
inline void changeBasis(const T b1[3], const T b2[3], const T b3[3], const T v[3], T out[3]) noexcept
{
    for (int i = 0; i < 3; ++i)
        out[i] = b1[i] * v[0] + b2[i] * v[1] + b3[i] * v[2];
}
template <Floating T>
inline void changeBasis(const T b1[3], const T b2[3], const T b3[3], T v[3]) noexcept
{
    T r[3];
    changeBasis(b1, b2, b3, v, r);
    v[0] = r[0];
    v[1] = r[1];
    v[2] = r[2];
}

// rows b1,b2,b3 times v (inverse of an orthonormal basis change)
template <Floating T>
inline void changeBasisInverse(const T b1[3], const T b2[3], const T b3[3], const T v[3], T out[3]) noexcept
{
    out[0] = b1[0] * v[0] + b1[1] * v[1] + b1[2] * v[2];
    out[1] = b2[0] * v[0] + b2[1] * v[1] + b2[2] * v[2];
    out[2] = b3[0] * v[0] + b3[1] * v[1] + b3[2] * v[2];
}
template <Floating T>
inline void changeBasisInverse(const T b1[3], const T b2[3], const T b3[3], T v[3]) noexcept
{
    T r[3];
    changeBasisInverse(b1, b2, b3, v, r);
    v[0] = r[0];
    v[1] = r[1];
    v[2] = r[2];
}

// Deflect v by polar angle theta, azimuth phi. The helper axis v x e_min is deliberately NOT
// normalised — the reference does not either, and step lengths scale with |dir| afterwards.
template <Floating T>
inline void peturb(T v[3], const T theta, const T phi) noexcept
{
    T k[3] = { 0, 0, 0 };
    k[argmin3<std::uint_fast32_t, T>(v)] = T { 1 };
    T ortho[3];
    cross(v, k, ortho);
    rotate(ortho, v, phi);
    const T s = std::sin(theta);
    const T c = std::cos(theta);
    for (int i = 0; i < 3; ++i)
        v[i] = v[i] * c + ortho[i] * s;
}
}
