/* Rigid transform value type for the host-side modelling API; see btVector3.h for scope. */
#ifndef TGB_HOST_BT_TRANSFORM_H
#define TGB_HOST_BT_TRANSFORM_H

#include "btMatrix3x3.h"

class btTransform {
public:
    btTransform() {}
    btTransform(const btMatrix3x3& b, const btVector3& o) : basis_(b), origin_(o) {}
    explicit btTransform(const btQuaternion& q, const btVector3& o = btVector3(0, 0, 0)) : basis_(q), origin_(o) {}
    void setIdentity() { basis_.setIdentity(); origin_.setValue(0, 0, 0); }
    static btTransform getIdentity() { btTransform t; t.setIdentity(); return t; }
    void setOrigin(const btVector3& o) { origin_ = o; }
    const btVector3& getOrigin() const { return origin_; }
    btVector3& getOrigin() { return origin_; }
    void setRotation(const btQuaternion& q) { basis_.setRotation(q); }
    btQuaternion getRotation() const { btQuaternion q; basis_.getRotation(q); return q; }
    const btMatrix3x3& getBasis() const { return basis_; }
    btMatrix3x3& getBasis() { return basis_; }
    void setBasis(const btMatrix3x3& b) { basis_ = b; }
    btVector3 operator()(const btVector3& x) const { return x.dot3(basis_[0], basis_[1], basis_[2]) + origin_; }
    btVector3 operator*(const btVector3& x) const { return (*this)(x); }
    btTransform operator*(const btTransform& t) const { return btTransform(basis_ * t.basis_, (*this)(t.origin_)); }
    btTransform inverse() const {
        btMatrix3x3 inv = basis_.transpose();
        return btTransform(inv, inv * -origin_);
    }

private:
    btMatrix3x3 basis_;
    btVector3 origin_;
};

#endif
