/* 3x3 rotation matrix value type for the host-side modelling API; see btVector3.h for scope. */
#ifndef TGB_HOST_BT_MATRIX3X3_H
#define TGB_HOST_BT_MATRIX3X3_H

#include "btQuaternion.h"

class btMatrix3x3 {
public:
    btMatrix3x3() {}
    explicit btMatrix3x3(const btQuaternion& q) { setRotation(q); }
    btMatrix3x3(btScalar xx, btScalar xy, btScalar xz, btScalar yx, btScalar yy, btScalar yz, btScalar zx, btScalar zy, btScalar zz) {
        setValue(xx, xy, xz, yx, yy, yz, zx, zy, zz);
    }
    void setValue(btScalar xx, btScalar xy, btScalar xz, btScalar yx, btScalar yy, btScalar yz, btScalar zx, btScalar zy, btScalar zz) {
        r_[0].setValue(xx, xy, xz); r_[1].setValue(yx, yy, yz); r_[2].setValue(zx, zy, zz);
    }
    void setIdentity() { setValue(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    static btMatrix3x3 getIdentity() { btMatrix3x3 m; m.setIdentity(); return m; }
    const btVector3& operator[](int i) const { return r_[i]; }
    btVector3& operator[](int i) { return r_[i]; }
    const btVector3& getRow(int i) const { return r_[i]; }
    btVector3 getColumn(int j) const { return btVector3(r_[0][j], r_[1][j], r_[2][j]); }

    void setRotation(const btQuaternion& q) {
        btScalar d = q.length2();
        btScalar s = btScalar(2.0) / d;
        btScalar xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
        btScalar wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
        btScalar xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
        btScalar yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
        setValue(btScalar(1.0) - (yy + zz), xy - wz, xz + wy, xy + wz, btScalar(1.0) - (xx + zz), yz - wx, xz - wy, yz + wx,
                 btScalar(1.0) - (xx + yy));
    }
    void getRotation(btQuaternion& q) const {
        btScalar trace = r_[0].x() + r_[1].y() + r_[2].z();
        btScalar t[4];
        if (trace > btScalar(0.0)) {
            btScalar s = std::sqrt(trace + btScalar(1.0));
            t[3] = s * btScalar(0.5);
            s = btScalar(0.5) / s;
            t[0] = (r_[2].y() - r_[1].z()) * s;
            t[1] = (r_[0].z() - r_[2].x()) * s;
            t[2] = (r_[1].x() - r_[0].y()) * s;
        } else {
            int i = r_[0].x() < r_[1].y() ? (r_[1].y() < r_[2].z() ? 2 : 1) : (r_[0].x() < r_[2].z() ? 2 : 0);
            int j = (i + 1) % 3, k = (i + 2) % 3;
            btScalar s = std::sqrt(r_[i][i] - r_[j][j] - r_[k][k] + btScalar(1.0));
            t[i] = s * btScalar(0.5);
            s = btScalar(0.5) / s;
            t[3] = (r_[k][j] - r_[j][k]) * s;
            t[j] = (r_[j][i] + r_[i][j]) * s;
            t[k] = (r_[k][i] + r_[i][k]) * s;
        }
        q.setValue(t[0], t[1], t[2], t[3]);
    }
    /* yaw-pitch-roll about Z, Y, X */
    void getEulerYPR(btScalar& yaw, btScalar& pitch, btScalar& roll) const {
        yaw = std::atan2(r_[1].x(), r_[0].x());
        pitch = std::asin(-r_[2].x());
        roll = std::atan2(r_[2].y(), r_[2].z());
        if (std::fabs(pitch) == SIMD_HALF_PI) {
            if (yaw > 0) yaw -= SIMD_PI; else yaw += SIMD_PI;
            if (roll > 0) roll -= SIMD_PI; else roll += SIMD_PI;
        }
    }
    btMatrix3x3 transpose() const {
        return btMatrix3x3(r_[0].x(), r_[1].x(), r_[2].x(), r_[0].y(), r_[1].y(), r_[2].y(), r_[0].z(), r_[1].z(), r_[2].z());
    }
    btMatrix3x3 inverse() const { return transpose(); } /* rotations only */
    btMatrix3x3 absolute() const {
        return btMatrix3x3(std::fabs(r_[0].x()), std::fabs(r_[0].y()), std::fabs(r_[0].z()), std::fabs(r_[1].x()), std::fabs(r_[1].y()),
                           std::fabs(r_[1].z()), std::fabs(r_[2].x()), std::fabs(r_[2].y()), std::fabs(r_[2].z()));
    }
    btMatrix3x3 scaled(const btVector3& s) const {
        return btMatrix3x3(r_[0].x() * s.x(), r_[0].y() * s.y(), r_[0].z() * s.z(), r_[1].x() * s.x(), r_[1].y() * s.y(), r_[1].z() * s.z(),
                           r_[2].x() * s.x(), r_[2].y() * s.y(), r_[2].z() * s.z());
    }

private:
    btVector3 r_[3];
};

inline btVector3 operator*(const btMatrix3x3& m, const btVector3& v) { return btVector3(m[0].dot(v), m[1].dot(v), m[2].dot(v)); }
inline btVector3 operator*(const btVector3& v, const btMatrix3x3& m) {
    return btVector3(m.getColumn(0).dot(v), m.getColumn(1).dot(v), m.getColumn(2).dot(v));
}
inline btMatrix3x3 operator*(const btMatrix3x3& a, const btMatrix3x3& b) {
    btMatrix3x3 bt = b.transpose();
    return btMatrix3x3(a[0].dot(bt[0]), a[0].dot(bt[1]), a[0].dot(bt[2]), a[1].dot(bt[0]), a[1].dot(bt[1]), a[1].dot(bt[2]), a[2].dot(bt[0]),
                       a[2].dot(bt[1]), a[2].dot(bt[2]));
}

#endif
