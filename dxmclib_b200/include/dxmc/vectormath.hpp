// vectormath.hpp — 3-vector helpers with the reference's names and operation order
// (API of reference include/dxmc/vectormath.hpp:37-221). Host side only; the device versions
// live in csrc/physics.cuh. Every expression keeps the reference's evaluation order: the host
// tables and exposure frames built with these must come out bit-identical.
#pragma once
#include "dxmc/types.hpp"
#include <cmath>
#include <cstdint>
#include <type_traits>

namespace dxmc::vectormath {

template <typename T>
concept Index = std::is_integral_v<T> && !std::is_same_v<bool, T>;

namespace detail {
    template <Floating T>
    inline void store3(T dst[3], const T src[3]) noexcept
    {
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
    }
    // index of the component that is smallest (Smallest = true) or largest in magnitude; ties go to the lower axis
    template <bool Smallest, Floating T>
    inline int extremeAxis(const T v[3]) noexcept
    {
        const T m[3] = { std::abs(v[0]), std::abs(v[1]), std::abs(v[2]) };
        auto before = [](T a, T b) { return Smallest ? a <= b : a >= b; };
        if (before(m[0], m[1]))
            return before(m[0], m[2]) ? 0 : 2;
        return before(m[1], m[2]) ? 1 : 2;
    }
}

// ---- products and norms
template <Floating T>
inline T dot(const T a[3], const T b[3]) noexcept { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

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

template <Floating T>
inline T lenght_sqr(T v[3]) noexcept { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
template <Floating T>
inline T lenght(const T v[3]) noexcept { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
template <Floating T>
inline T lenght(T v[3]) noexcept { return lenght(static_cast<const T*>(v)); }

template <Floating T>
inline void normalize(T v[3]) noexcept
{
    const T scale = T { 1 } / lenght(static_cast<const T*>(v));
    v[0] *= scale;
    v[1] *= scale;
    v[2] *= scale;
}

template <Index U, Floating T>
inline U argmin3(const T v[3]) noexcept { return static_cast<U>(detail::extremeAxis<true>(v)); }
template <Index U, Floating T>
inline U argmax3(const T v[3]) noexcept { return static_cast<U>(detail::extremeAxis<false>(v)); }

// ---- rotations and scattering
// Rodrigues rotation of v about a unit axis
template <Floating T>
inline void rotate(T v[3], const T axis[3], const T angle) noexcept
{
    const T s = std::sin(angle);
    const T c = std::cos(angle);
    const T m = (T { 1 } - c) * dot(v, axis);
    const T turned[3] = {
        c * v[0] + m * axis[0] + s * (axis[1] * v[2] - axis[2] * v[1]),
        c * v[1] + m * axis[1] + s * (-axis[0] * v[2] + axis[2] * v[0]),
        c * v[2] + m * axis[2] + s * (axis[0] * v[1] - axis[1] * v[0])
    };
    detail::store3(v, turned);
}

// Deflect v by polar angle theta, azimuth phi. The helper axis v x e_min is deliberately NOT
// normalised — the reference does not either, and step lengths scale with |dir| afterwards.
template <Floating T>
inline void peturb(T v[3], const T theta, const T phi) noexcept
{
    T unit[3] = { 0, 0, 0 };
    unit[detail::extremeAxis<true>(v)] = T { 1 };
    T ortho[3];
    cross(v, unit, ortho);
    rotate(ortho, v, phi);
    const T s = std::sin(theta);
    const T c = std::cos(theta);
    const T bent[3] = { v[0] * c + ortho[0] * s, v[1] * c + ortho[1] * s, v[2] * c + ortho[2] * s };
    detail::store3(v, bent);
}

template <Floating T>
inline void projectToPlane(T v[3], const T normal[3]) noexcept
{
    const T along = dot(v, normal);
    const T inPlane[3] = { v[0] - along * normal[0], v[1] - along * normal[1], v[2] - along * normal[2] };
    detail::store3(v, inPlane);
}

// ---- frames: columns b1,b2,b3 times v, and rows b1,b2,b3 times v (the inverse for an orthonormal basis)
template <Floating T>
inline void changeBasis(const T b1[3], const T b2[3], const T b3[3], const T v[3], T out[3]) noexcept
{
    out[0] = b1[0] * v[0] + b2[0] * v[1] + b3[0] * v[2];
    out[1] = b1[1] * v[0] + b2[1] * v[1] + b3[1] * v[2];
    out[2] = b1[2] * v[0] + b2[2] * v[1] + b3[2] * v[2];
}
template <Floating T>
inline void changeBasisInverse(const T b1[3], const T b2[3], const T b3[3], const T v[3], T out[3]) noexcept
{
    out[0] = dot(b1, v);
    out[1] = dot(b2, v);
    out[2] = dot(b3, v);
}
template <Floating T>
inline void changeBasis(const T b1[3], const T b2[3], const T b3[3], T v[3]) noexcept
{
    T image[3];
    changeBasis(b1, b2, b3, static_cast<const T*>(v), image);
    detail::store3(v, image);
}
template <Floating T>
inline void changeBasisInverse(const T b1[3], const T b2[3], const T b3[3], T v[3]) noexcept
{
    T image[3];
    changeBasisInverse(b1, b2, b3, static_cast<const T*>(v), image);
    detail::store3(v, image);
}

// ---- angles
// Kahan's numerically stable angle between two vectors
template <Floating T>
inline T angleBetween(const T a[3], const T b[3]) noexcept
{
    const T d[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
    const T la = lenght(a), lb = lenght(b), lc = lenght(static_cast<const T*>(d));
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
}
