/* Shim (oracle/ref) standing in for src/tgcreator/tgUtil.h while src/core/tgBulletContactSpringCable.cpp is compiled in place: the
   reference's header pulls in the whole creator layer and Bullet's shape classes for its stream operators.  The contact cable uses
   one function of it, tgUtil::getTransform(start, end), to place the cylinders of its ghost shape (updateCollisionObject,
   tgBulletContactSpringCable.cpp:717-727) -- a shape nothing reads in this harness (the manifolds are filled by ref_harness.cpp), so
   the transform only needs its origin. */
#ifndef REF_SHIM_TG_UTIL_H
#define REF_SHIM_TG_UTIL_H
#include "LinearMath/btTransform.h"
class tgUtil {
public:
    static btTransform getTransform(const btVector3& start, const btVector3& end) {
        btTransform t;
        t.setIdentity();
        t.setOrigin((start + end) * btScalar(0.5));
        return t;
    }
};
#endif
