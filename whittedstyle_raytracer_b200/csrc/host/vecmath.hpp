// vecmath.hpp — float3 helpers whose rounding order follows the reference's
// Vector3f (include/Vector.hpp:59-139): every product and sum is a separate
// correctly-rounded float operation, dot = (x*x' + y*y') + z*z',
// normalized() multiplies by 1/mag and returns the input when mag == 0.
#pragma once
#include <cmath>

namespace wrt {

struct V2 {
    float x = -1.f, y = -1.f;             // Vector.hpp:43-46: default is (-1,-1) = "no texture"
    V2() = default;
    V2(float a, float b) : x(a), y(b) {}
};

struct V3 {
    float x = 0.f, y = 0.f, z = 0.f;
    V3() = default;
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(float c) const { return V3(x * c, y * c, z * c); }
    V3 operator/(float c) const { return V3(x / c, y / c, z / c); }
    float dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    float norm() const { return sqrtf(x * x + y * y + z * z); }
};

inline V3 operator*(float c, const V3& v) { return V3(v.x * c, v.y * c, v.z * c); }

inline V3 normalized(const V3& v) {        // Vector.hpp:127-134
    float mag = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (mag > 0) {
        float mag_inv = 1 / mag;
        return V3(v.x * mag_inv, v.y * mag_inv, v.z * mag_inv);
    }
    return v;
}

inline V3 cross(const V3& a, const V3& b) { // Vector.hpp:137-139
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

inline bool float_equal(float x, float y) { // global.hpp:92-94
    return fabsf(x - y) < 0.00001f;
}

} // namespace wrt
