/*
 * Minimal value types with the names the NTRTsim modelling API exposes in its signatures
 * (btVector3 / btQuaternion / btMatrix3x3 / btTransform, btScalar = double as in the reference build:
 * src/inc.CMakeBullet.txt:47-50).  Written for this repo; Bullet is not a dependency.  Only the members
 * the host-side builder path uses are provided (SURVEY.md Appendix C lists that surface).
 */
#ifndef TGB_HOST_BT_VECTOR3_H
#define TGB_HOST_BT_VECTOR3_H

#include <cfloat>
#include <cmath>
#include <ostream>

typedef double btScalar;
#define SIMD_EPSILON DBL_EPSILON
#define SIMD_PI 3.1415926535897932384626433832795029
#define SIMD_HALF_PI (SIMD_PI * 0.5)

class btVector3 {
public:
    btVector3() { v_[0] = v_[1] = v_[2] = v_[3] = 0.0; }
    btVector3(btScalar x, btScalar y, btScalar z) { v_[0] = x; v_[1] = y; v_[2] = z; v_[3] = 0.0; }

    btScalar x() const { return v_[0]; }
    btScalar y() const { return v_[1]; }
    btScalar z() const { return v_[2]; }
    btScalar w() const { return v_[3]; }
    btScalar getX() const { return v_[0]; }
    btScalar getY() const { return v_[1]; }
    btScalar getZ() const { return v_[2]; }
    void setX(btScalar a) { v_[0] = a; }
    void setY(btScalar a) { v_[1] = a; }
    void setZ(btScalar a) { v_[2] = a; }
    void setValue(btScalar a, btScalar b, btScalar c) { v_[0] = a; v_[1] = b; v_[2] = c; v_[3] = 0.0; }
    /* indexing (v[i]) and std::set<btVector3> in user code both go through these conversions */
    operator const btScalar*() const { return v_; }
    operator btScalar*() { return v_; }

    btVector3& operator+=(const btVector3& o) { v_[0] += o.v_[0]; v_[1] += o.v_[1]; v_[2] += o.v_[2]; return *this; }
    btVector3& operator-=(const btVector3& o) { v_[0] -= o.v_[0]; v_[1] -= o.v_[1]; v_[2] -= o.v_[2]; return *this; }
    btVector3& operator*=(btScalar s) { v_[0] *= s; v_[1] *= s; v_[2] *= s; return *this; }
    btVector3& operator/=(btScalar s) { return *this *= (btScalar(1.0) / s); }
    bool operator==(const btVector3& o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2] && v_[3] == o.v_[3]; }
    bool operator!=(const btVector3& o) const { return !(*this == o); }

    btScalar dot(const btVector3& o) const { return v_[0] * o.v_[0] + v_[1] * o.v_[1] + v_[2] * o.v_[2]; }
    btVector3 cross(const btVector3& o) const {
        return btVector3(v_[1] * o.v_[2] - v_[2] * o.v_[1], v_[2] * o.v_[0] - v_[0] * o.v_[2], v_[0] * o.v_[1] - v_[1] * o.v_[0]);
    }
    btScalar length2() const { return dot(*this); }
    btScalar length() const { return std::sqrt(length2()); }
    btScalar distance(const btVector3& o) const;
    btVector3& normalize() { return *this /= length(); }
    btVector3 normalized() const { btVector3 r(*this); r.normalize(); return r; }
    bool isZero() const { return v_[0] == 0.0 && v_[1] == 0.0 && v_[2] == 0.0; }
    bool fuzzyZero() const { return length2() < SIMD_EPSILON; }
    btVector3 absolute() const { return btVector3(std::fabs(v_[0]), std::fabs(v_[1]), std::fabs(v_[2])); }
    /* rotate about a unit axis */
    btVector3 rotate(const btVector3& axis, btScalar angle) const;
    btVector3 dot3(const btVector3& a, const btVector3& b, const btVector3& c) const { return btVector3(dot(a), dot(b), dot(c)); }

private:
    btScalar v_[4];
};

inline btVector3 operator+(const btVector3& a, const btVector3& b) { return btVector3(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }
inline btVector3 operator-(const btVector3& a, const btVector3& b) { return btVector3(a.x() - b.x(), a.y() - b.y(), a.z() - b.z()); }
inline btVector3 operator-(const btVector3& a) { return btVector3(-a.x(), -a.y(), -a.z()); }
inline btVector3 operator*(const btVector3& a, btScalar s) { return btVector3(a.x() * s, a.y() * s, a.z() * s); }
inline btVector3 operator*(btScalar s, const btVector3& a) { return a * s; }
inline btVector3 operator*(const btVector3& a, const btVector3& b) { return btVector3(a.x() * b.x(), a.y() * b.y(), a.z() * b.z()); }
inline btVector3 operator/(const btVector3& a, btScalar s) { return a * (btScalar(1.0) / s); }
inline btScalar btVector3::distance(const btVector3& o) const { return (o - *this).length(); }
inline btVector3 btVector3::rotate(const btVector3& axis, btScalar angle) const {
    btVector3 o = axis * axis.dot(*this);
    btVector3 px = *this - o;
    btVector3 py = axis.cross(*this);
    return o + px * std::cos(angle) + py * std::sin(angle);
}
inline std::ostream& operator<<(std::ostream& os, const btVector3& v) { return os << "(" << v.x() << ", " << v.y() << ", " << v.z() << ")"; }

#endif
