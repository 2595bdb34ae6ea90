/* Quaternion value type (x,y,z,w) for the host-side modelling API; see btVector3.h for scope. */
#ifndef TGB_HOST_BT_QUATERNION_H
#define TGB_HOST_BT_QUATERNION_H

#include "btVector3.h"

class btQuaternion {
public:
    btQuaternion() : x_(0), y_(0), z_(0), w_(1) {}
    btQuaternion(btScalar x, btScalar y, btScalar z, btScalar w) : x_(x), y_(y), z_(z), w_(w) {}
    btQuaternion(const btVector3& axis, btScalar angle) { setRotation(axis, angle); }

    static btQuaternion getIdentity() { return btQuaternion(0, 0, 0, 1); }
    btScalar x() const { return x_; }
    btScalar y() const { return y_; }
    btScalar z() const { return z_; }
    btScalar w() const { return w_; }
    btScalar getX() const { return x_; }
    btScalar getY() const { return y_; }
    btScalar getZ() const { return z_; }
    btScalar getW() const { return w_; }
    void setValue(btScalar x, btScalar y, btScalar z, btScalar w) { x_ = x; y_ = y; z_ = z; w_ = w; }

    void setRotation(const btVector3& axis, btScalar angle) {
        btScalar d = axis.length();
        btScalar s = std::sin(angle * btScalar(0.5)) / d;
        setValue(axis.x() * s, axis.y() * s, axis.z() * s, std::cos(angle * btScalar(0.5)));
    }
    /* yaw about Y, pitch about X, roll about Z */
    void setEuler(btScalar yaw, btScalar pitch, btScalar roll) {
        btScalar hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
        btScalar cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
        setValue(cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, sr * cp * cy - cr * sp * sy, cr * cp * cy + sr * sp * sy);
    }
    btScalar length2() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
    btScalar length() const { return std::sqrt(length2()); }
    btQuaternion& operator*=(btScalar s) { x_ *= s; y_ *= s; z_ *= s; w_ *= s; return *this; }
    btQuaternion& operator/=(btScalar s) { return *this *= (btScalar(1.0) / s); }
    btQuaternion& normalize() { return *this /= length(); }
    btQuaternion inverse() const { return btQuaternion(-x_, -y_, -z_, w_); }
    btScalar getAngle() const { return btScalar(2.0) * std::acos(w_); }
    btVector3 getAxis() const {
        btScalar s2 = btScalar(1.0) - w_ * w_;
        if (s2 < btScalar(10.0) * SIMD_EPSILON) return btVector3(1.0, 0.0, 0.0);
        btScalar s = btScalar(1.0) / std::sqrt(s2);
        return btVector3(x_ * s, y_ * s, z_ * s);
    }

private:
    btScalar x_, y_, z_, w_;
};

inline btQuaternion operator*(const btQuaternion& a, const btQuaternion& b) {
    return btQuaternion(a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                        a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                        a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x(),
                        a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z());
}
inline btQuaternion operator-(const btQuaternion& q) { return btQuaternion(-q.x(), -q.y(), -q.z(), -q.w()); }

#endif
